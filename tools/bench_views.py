"""BASELINE.json config 4 (24 views of 256x256 on the ring, one shape, summed loss, one backward over the latent):
views rendered one after the other vs. `render_views` (back to back without host syncs / on several streams /
one fused march over all views)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev); lat0 = synth.make_latent().to(dev)
views = synth.ring_cameras(24, 25.0, 2.5)
Rs = torch.stack([r for r, _ in views]).to(dev); Ts = torch.stack([t for _, t in views]).to(dev)
ren = pkg.SDFRenderer(dec, synth.intrinsic(256, 256, 1.2 * 2.5 / 1.6), img_hw=(256, 256))


def timeit(fn, n=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for kind in ("recursive", "pyramid_recursive"):
    def looped():
        l = lat0.detach().requires_grad_(True)
        sum(bench.loss_of(ren.render(l, Rs[v], Ts[v], ray_marching_type=kind)) for v in range(24)).backward()
    ms = timeit(looped)
    print("%-18s looped          : %7.1f ms  %6.3f M rays/s" % (kind, ms, 24 * 65536 / ms / 1e3), flush=True)
    for ns in (1, 4):
        def batched():
            l = lat0.detach().requires_grad_(True)
            bench.loss_of(ren.render_views(l, Rs, Ts, fused=False, n_streams=ns, ray_marching_type=kind)).backward()
        ms = timeit(batched)
        print("%-18s back to back x%d : %7.1f ms  %6.3f M rays/s" % (kind, ns, ms, 24 * 65536 / ms / 1e3), flush=True)
    def fused():
        l = lat0.detach().requires_grad_(True)
        bench.loss_of(ren.render_views(l, Rs, Ts, ray_marching_type=kind)).backward()
    ms = timeit(fused)
    print("%-18s fused march     : %7.1f ms  %6.3f M rays/s" % (kind, ms, 24 * 65536 / ms / 1e3), flush=True)
