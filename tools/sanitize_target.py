#!/usr/bin/env python
"""Small end-to-end workload for `compute-sanitizer --tool memcheck` (and racecheck on the elementwise kernels): a 48x48
render with all three marching variants, forward + backward, on the tensor-core engine with two-tier precision, the
two-view warp loss and a decode_color call -- every kernel of the library runs at least once."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
warp = importlib.import_module("dist-renderer_b200.renderer_warp")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev)
hw = (48, 48)
K = synth.intrinsic(*hw, focal_scale=1.2 * 2.5 / 1.6)
(R1, T1), (R2, T2) = synth.lookat_camera(20.0, 25.0, 2.5), synth.lookat_camera(35.0, 25.0, 2.5)
R1, T1, R2, T2 = R1.to(dev), T1.to(dev), R2.to(dev), T2.to(dev)
ren = pkg.SDFRenderer(dec, K, img_hw=hw, march_step=30, buffer_size=5, engine="tc")
for kind in ("recursive", "pyramid_recursive", "trivial"):
    lat = synth.make_latent().to(dev).requires_grad_(True)
    Rg, Tg = R1.clone().requires_grad_(True), T1.clone().requires_grad_(True)
    out = ren.render(lat, Rg, Tg, ray_marching_type=kind)
    (torch.where(out[2].bool(), out[0], torch.zeros_like(out[0])).sum() + out[3].sum() + out[1].sum()).backward()
    print(kind, "hits", int(out[2].sum()), "|g_latent|", float(lat.grad.norm()))
rw = warp.SDFRenderer_warp(dec, K, img_hw=hw, march_step=30)
g = torch.Generator().manual_seed(9)
img1, img2 = torch.rand(*hw, 3, generator=g).to(dev), torch.rand(*hw, 3, generator=g).to(dev)
lat = synth.make_latent().to(dev).requires_grad_(True)
o = rw.render_warp(lat, R1, T1, R2, T2, img1, img2)
o[0].backward()
print("warp loss", float(o[0]))
col = synth.make_color_decoder().to(dev)
rgb = pkg.decode_color(col, (0.1 * torch.randn(1, 8)).to(dev), synth.make_latent().to(dev), (torch.rand(300, 3, device=dev) - 0.5))
print("colour", tuple(rgb.shape))
torch.cuda.synchronize()
print("done")
