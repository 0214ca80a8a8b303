"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the REAL reference (via oracle/ref_shim.py).

Run in the build container (needs /root/reference):   python oracle/make_golden.py            (small cases)
                                                       python oracle/make_golden.py --flags    (gradient flags)
                                                       python oracle/make_golden.py --big      (BASELINE sizes; ~20 min)
                                                       python oracle/make_golden.py --chamfer  (eval_func.py outputs)
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


def _grads(ts):
    return [t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape), np.float32) for t in ts]


def big_fixtures(names=None):
    """tests/golden/big_*.npz: the BASELINE.json configurations at their own sizes (cases.BIG_CASES), rendered by the
    real reference (fp32) -- outputs, gradients of cases.scalar_loss w.r.t. latent / R / T -- plus the deviation of the
    fp64 twin (oracle/sdf_oracle.py in float64, itself pinned bit for bit to the reference in fp32) from them: the noise
    floor BASELINE.md section 3 asks to be printed beside every parity number."""
    import time
    import gpu_util as gu
    Rmod, _, _ = ref_shim.load()
    lat0 = cases.synth.make_latent()
    for name, cs in cases.BIG_CASES.items():
        if names and name not in names:
            continue
        t0 = time.time()
        dec = cases.decoder(cs["decoder"])
        K, R, T = cases.camera(cs["cam"], cs["hw"])
        ren = Rmod.SDFRenderer(ref_decoder(dec), K, img_hw=cs["hw"], march_step=cs["march_step"],
                               buffer_size=cs["buffer_size"], use_gpu=False)
        lat, Rg, Tg = lat0.clone().requires_grad_(True), R.clone().requires_grad_(True), T.clone().requires_grad_(True)
        out = ren.render(lat, Rg, Tg, ray_marching_type=cs["kind"])
        cases.scalar_loss(out).backward()
        ref = [o.detach() for o in out]
        gref = [torch.from_numpy(g) for g in _grads((lat, Rg, Tg))]
        t1 = time.time()
        o64, g64 = gu.run_oracle(cs, dtype=torch.float64)
        floor = gu.measure([o.float() if o.dtype == torch.float64 else o for o in o64], ref, [g.float() for g in g64], gref)
        floor = {k: (-1.0 if v is None else float(v)) for k, v in floor.items()}
        np.savez_compressed(
            os.path.join(cases.GOLDEN_DIR, "big_" + name + ".npz"),
            depth=ref[0].numpy(), normal=ref[1].numpy(), mask=ref[2].numpy(), min_sdf=ref[3].numpy(),
            g_latent=gref[0].numpy(), g_R=gref[1].numpy(), g_T=gref[2].numpy(),
            floor_keys=np.array(sorted(floor)), floor_vals=np.array([floor[k] for k in sorted(floor)]),
            weights_checksum=cases.weights_checksum(dec))
        print("big_" + name, "hits", int(ref[2].sum()), "of", ref[2].numel(), "ref %.0fs fp64 %.0fs" % (t1 - t0, time.time() - t1),
              "fp64 floor:", {k: ("%.3g" % v) for k, v in floor.items()}, flush=True)


def flag_fixtures():
    """tests/golden/flag_*.npz: render() of the real reference under each gradient flag (renderer.py:943-957) with the
    gradients that survive it (zeros where autograd gives None), and silhouette_48.npz: the (valid_mask, min_sdf)
    pair of render_depth (renderer.py:878) with the gradients of min_sdf.sum()."""
    Rmod, _, _ = ref_shim.load()
    lat0 = cases.synth.make_latent()
    for name, cs in cases.FLAG_CASES.items():
        dec = cases.decoder(cs["decoder"])
        K, R, T = cases.camera(cs["cam"], cs["hw"])
        ren = Rmod.SDFRenderer(ref_decoder(dec), K, img_hw=cs["hw"], march_step=cs["march_step"],
                               buffer_size=cs["buffer_size"], use_gpu=False)
        lat, Rg, Tg = lat0.clone().requires_grad_(True), R.clone().requires_grad_(True), T.clone().requires_grad_(True)
        out = ren.render(lat, Rg, Tg, ray_marching_type=cs["kind"], **cs["flags"])
        cases.scalar_loss(out).backward()
        g = _grads((lat, Rg, Tg))
        np.savez_compressed(os.path.join(cases.GOLDEN_DIR, name + ".npz"), depth=out[0].detach().numpy(),
                            normal=out[1].detach().numpy(), mask=out[2].numpy(), min_sdf=out[3].detach().numpy(),
                            g_latent=g[0], g_R=g[1], g_T=g[2], weights_checksum=cases.weights_checksum(dec))
        print(name, "hits", int(out[2].sum()), "|g|", [float(np.abs(x).sum()) for x in g])
    cs = cases._FLAG_BASE
    dec = cases.decoder(cs["decoder"])
    K, R, T = cases.camera(cs["cam"], cs["hw"])
    ren = Rmod.SDFRenderer(ref_decoder(dec), K, img_hw=cs["hw"], march_step=cs["march_step"],
                           buffer_size=cs["buffer_size"], use_gpu=False)
    lat, Rg, Tg = lat0.clone().requires_grad_(True), R.clone().requires_grad_(True), T.clone().requires_grad_(True)
    _, vm, ms = ren.render_depth(lat, Rg, Tg, ray_marching_type=cs["kind"])
    ms.sum().backward()
    g = _grads((lat, Rg, Tg))
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "silhouette_48.npz"), mask=vm.reshape(cs["hw"]).numpy(),
                        min_sdf=ms.detach().reshape(cs["hw"]).numpy(), g_latent=g[0], g_R=g[1], g_T=g[2],
                        weights_checksum=cases.weights_checksum(dec))
    print("silhouette_48 hits", int(vm.sum()))


def chamfer_fixture():
    """Outputs of the reference's core/evaluation/eval_func.py on the seeded point sets of tests/test_mesh_cpu.py."""
    EF = ref_shim.load_eval_func()
    rng = np.random.default_rng(7)
    a = rng.standard_normal((3000, 3)) * 0.3
    b = rng.standard_normal((2500, 3)) * 0.3 + 0.05
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "chamfer.npz"), a=a, b=b, sq=EF.compute_chamfer_distance(a, b),
                        lin=EF.compute_chamfer_distance(a, b, use_square_dist=False),
                        sep=np.array(EF.compute_chamfer_distance_separate(a, b)))
    print("chamfer", EF.compute_chamfer_distance(a, b))


def main():
    if "--chamfer" in sys.argv:
        os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
        return chamfer_fixture()
    if "--big" in sys.argv:              # minutes of CPU per case: written separately from the small fixtures
        os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
        return big_fixtures([a for a in sys.argv[1:] if not a.startswith("--")])
    if "--flags" in sys.argv:
        os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
        return flag_fixtures()
    if "--color-only" in sys.argv:      # adds the colour fixture without rewriting the others
        os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
        return color_fixture()
    Rmod, DU, _ = ref_shim.load()
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    lat0 = cases.synth.make_latent()
    only = sys.argv[sys.argv.index("--only") + 1:] if "--only" in sys.argv else None   # add fixtures without rewriting the rest
    for name, cs in cases.CASES.items():
        if only is not None and name not in only:
            continue
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
        more = {}
        if name.startswith("earlybreak"):
            # no ray converges in these cases, so render() hides Zdepth: pin render_depth itself -- the raw Zdepth of the
            # sphere-hit rays and the gradients of their sum, which run through ALL buffer_size selected samples
            l2, R2, T2 = lat0.clone().requires_grad_(True), R.clone().requires_grad_(True), T.clone().requires_grad_(True)
            Zd, _, _ = ren.render_depth(l2, R2, T2, ray_marching_type=cs["kind"])
            Zd[Zd < 1e10].sum().backward()
            more = dict(rd_Zdepth=Zd.detach().numpy(), rd_g_latent=l2.grad.numpy(), rd_g_R=R2.grad.numpy(),
                        rd_g_T=T2.grad.numpy())
        np.savez_compressed(
            os.path.join(cases.GOLDEN_DIR, name + ".npz"), **more,
            depth=out[0].detach().numpy(), normal=out[1].detach().numpy(), mask=out[2].numpy(),
            min_sdf=out[3].detach().numpy(), g_latent=lat.grad.numpy(), g_R=Rg.grad.numpy(), g_T=Tg.grad.numpy(),
            Zdepth_nograd=Zdepth.numpy(), weights_checksum=cases.weights_checksum(dec))
        print(name, "hits", int(out[2].sum()), "of", out[2].numel())
    if only is not None:
        return
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
