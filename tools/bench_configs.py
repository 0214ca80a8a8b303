"""Timing of the other BASELINE.json configurations (not the headline bench): config 3 (224x224, march 100, buffer 3,
fwd+bwd), config 4 (24 views 256x256 fwd+bwd, looped and as one fused march; see also tools/bench_views.py), config 2 variants (512x512 forward only, pyramid/trivial)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev); lat0 = synth.make_latent().to(dev)


def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def step(ren, R, T, kind, grad=True):
    l = lat0.detach().requires_grad_(grad)
    out = ren.render(l, R, T, ray_marching_type=kind, no_grad=not grad)
    if grad: bench.loss_of(out).backward()


R, T = synth.front_camera(); R, T = R.to(dev), T.to(dev)
for name, hw, ms_, bs in (("config2 512x512 s50 b5", 512, 50, 5), ("config3 224x224 s100 b3", 224, 100, 3), ("256x256 s50 b5", 256, 50, 5)):
    ren = pkg.SDFRenderer(dec, synth.intrinsic(hw, hw), img_hw=(hw, hw), march_step=ms_, buffer_size=bs)
    for kind in ("recursive", "pyramid_recursive", "trivial"):
        for grad in (False, True):
            ms = timeit(lambda: step(ren, R, T, kind, grad))
            print("%-26s %-18s %-8s %8.2f ms  %8.3f M rays/s" % (name, kind, "fwd+bwd" if grad else "fwd", ms, hw * hw / ms / 1e3), flush=True)
views = [(r.to(dev), t.to(dev)) for r, t in synth.ring_cameras(24, 25.0, 2.5)]
ren = pkg.SDFRenderer(dec, synth.intrinsic(256, 256, 1.2 * 2.5 / 1.6), img_hw=(256, 256))
def multi():
    l = lat0.detach().requires_grad_(True)
    tot = sum(bench.loss_of(ren.render(l, r, t, ray_marching_type="recursive")) for r, t in views)
    tot.backward()
ms = timeit(multi, n=3)
print("config4 24 views 256x256 fwd+bwd (looped): %.1f ms  %.3f M rays/s" % (ms, 24 * 65536 / ms / 1e3))
Rs, Ts = torch.stack([r for r, _ in views]), torch.stack([t for _, t in views])
def fused():
    l = lat0.detach().requires_grad_(True)
    bench.loss_of(ren.render_views(l, Rs, Ts, ray_marching_type="recursive")).backward()
ms = timeit(fused, n=3)
print("config4 24 views 256x256 fwd+bwd (one fused march): %.1f ms  %.3f M rays/s" % (ms, 24 * 65536 / ms / 1e3))
