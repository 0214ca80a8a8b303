"""Small driver for profiling: a few decoder-row launches of the chosen engine/mode on a 262,144-row batch."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
dec = synth.make_decoder("B").cuda(); lat = synth.make_latent().cuda()
g = torch.Generator().manual_seed(11)
pts = ((torch.rand(n, 3, generator=g) - 0.5) * 1.2).cuda()
for _ in range(4):
    if mode == "fwd":
        pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine=engine)
    else:
        pkg.decode_sdf_gradient(dec, lat, pts, clamp_dist=0.1, engine=engine)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    if mode == "fwd":
        pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine=engine)
    else:
        pkg.decode_sdf_gradient(dec, lat, pts, clamp_dist=0.1, engine=engine)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
F = 3146752 * (1 if mode == "fwd" else 2)
print("%s %s n=%d: %.3f ms/launch  useful %.1f TFLOP/s  issued(x3) %.1f TFLOP/s" % (engine, mode, n, ms, n * F / ms / 1e9, 3 * n * F / ms / 1e9))
