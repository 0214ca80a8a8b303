"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the REAL reference (via oracle/ref_shim.py).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Each fixture holds the outputs of the unmodified reference `SDFRenderer.render` (depth, normal, mask, min_sdf),
the gradients of tests/cases.scalar_loss w.r.t. latent / R / T, and a checksum of the seeded decoder weights
the recipe regenerates.  `decoder_points.npz` pins decode_sdf / decode_sdf_gradient on random points.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
import cases  # noqa: E402


def ref_decoder(dec):
    _, _, RefDecoder = ref_shim.load()
    ref = RefDecoder(dec.latent_size, **cases.synth.STANDARD_SPEC).eval()
    ref.load_state_dict(dec.state_dict())
    return ref


def color_fixture():
    """tests/golden/color_24.npz from the reference's SDFRenderer_color.render (core/sdfrenderer/renderer_rgb.py:70)."""
    _, _, RefDecoder = ref_shim.load()
    Color = ref_shim.load_color()
    dec, col = cases.decoder("B"), cases.synth.make_color_decoder()
    spec = dict(cases.synth.STANDARD_SPEC, dims=list(col.dims[1:-1]))
    ref_col = RefDecoder(col.latent_size, last_dim=3, **spec).eval()
    ref_col.load_state_dict(col.state_dict())
    hw, K, R, T, cc, lights, energies = cases.color_case()
    ren = Color(ref_decoder(dec), ref_col, K, img_hw=hw, use_gpu=False)
    lat = cases.synth.make_latent().requires_grad_(True)
    ccg = cc.clone().requires_grad_(True)
    out = ren.render(ccg, lat, R, T, lighting_locations=lights, lighting_energies=energies)
    (out[2].sum() + out[0][out[3].bool()].sum()).backward()
    plain = ren.render(cc, cases.synth.make_latent(), R, T, no_grad=True)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "color_24.npz"), depth=out[0].detach().numpy(),
                        normal=out[1].detach().numpy(), color=out[2].detach().numpy(), mask=out[3].numpy(),
                        min_sdf=out[4].detach().numpy(), color_unlit=plain[2].numpy(), g_latent=lat.grad.numpy(),
                        g_color=ccg.grad.numpy(), weights_checksum=cases.weights_checksum(col))
    print("color_24 hits", int(out[3].sum()), "max rgb", float(out[2].abs().max()))


def main():
    if "--color-only" in sys.argv:      # adds the colour fixture without rewriting the others
        os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
        return color_fixture()
    Rmod, DU, _ = ref_shim.load()
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    lat0 = cases.synth.make_latent()
    for name, cs in cases.CASES.items():
        dec = cases.decoder(cs["decoder"])
        ref = ref_decoder(dec)
        K, R, T = cases.camera(cs["cam"], cs["hw"])
        ren = Rmod.SDFRenderer(ref, K, img_hw=cs["hw"], march_step=cs["march_step"], buffer_size=cs["buffer_size"],
                               use_gpu=False)
        lat = lat0.clone().requires_grad_(True)
        Rg, Tg = R.clone().requires_grad_(True), T.clone().requires_grad_(True)
        out = ren.render(lat, Rg, Tg, ray_marching_type=cs["kind"])
        cases.scalar_loss(out).backward()
        Zdepth, zmask, zmin = ren.render_depth(lat0, R, T, ray_marching_type=cs["kind"], no_grad=True)
        np.savez_compressed(
            os.path.join(cases.GOLDEN_DIR, name + ".npz"),
            depth=out[0].detach().numpy(), normal=out[1].detach().numpy(), mask=out[2].numpy(),
            min_sdf=out[3].detach().numpy(), g_latent=lat.grad.numpy(), g_R=Rg.grad.numpy(), g_T=Tg.grad.numpy(),
            Zdepth_nograd=Zdepth.numpy(), weights_checksum=cases.weights_checksum(dec))
        print(name, "hits", int(out[2].sum()), "of", out[2].numel())
    # decoder-level fixture
    dec = cases.decoder("B")
    ref = ref_decoder(dec)
    g = torch.Generator().manual_seed(7)
    pts = (torch.rand(3000, 3, generator=g) - 0.5) * 1.6
    sdf = DU.decode_sdf(ref, lat0, pts, clamp_dist=None).detach()
    p = pts.clone().requires_grad_(True)
    grad = DU.decode_sdf_gradient(ref, lat0, p, clamp_dist=0.1).detach()
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "decoder_points.npz"), points=pts.numpy(), sdf=sdf.numpy(),
                        grad=grad.numpy(), weights_checksum=cases.weights_checksum(dec))
    print("decoder_points", float(sdf.min()), float(sdf.max()))
    # two-view warp fixture (reference SDFRenderer_warp.render_warp, core/sdfrenderer/renderer_warp.py:103)
    Warp = ref_shim.load_warp()
    hw, K, (R1, T1), (R2, T2), img1, img2 = cases.warp_case()
    rw = Warp(ref, K, img_hw=hw, use_gpu=False)
    lat = lat0.clone().requires_grad_(True)
    out = rw.render_warp(lat, R1, T1, R2, T2, img1, img2)
    out[0].backward()
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "warp_40.npz"), loss=float(out[0]), g_latent=lat.grad.numpy(),
                        mask1=out[3].numpy(), mask2=out[4].numpy(), min_sdf1=out[5].detach().numpy(),
                        normal1=out[7].detach().numpy(), depth1=out[8].detach().numpy(),
                        weights_checksum=cases.weights_checksum(dec))
    print("warp_40 loss", float(out[0]))
    color_fixture()


if __name__ == "__main__":
    main()
