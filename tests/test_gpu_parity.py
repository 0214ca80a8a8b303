"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (fp32 path): depth / min_sdf rel-L2 < 1e-4 on the mask intersection; mask XOR <= 2 pixels; normal
rel-L2 < 1e-4 after excluding <= 0.1 % ReLU-flip outlier pixels; gradients rel-L2 < 2e-3.
"""
import os

import numpy as np
import pytest
import torch

import cases
import gpu_util as gu

pytestmark = pytest.mark.gpu
pkg = cases.pkg
RENDER_CASES = sorted(cases.CASES)


@pytest.fixture(scope="session", autouse=True)
def _built():
    import importlib
    importlib.import_module("dist-renderer_b200.build").build()


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_decoder_points_golden(engine):
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "decoder_points.npz"))
    dec = gu.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    pts = torch.from_numpy(gold["points"]).cuda()
    sdf = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, engine=engine)
    assert sdf.shape == (pts.shape[0], 1)
    assert gu.rel(sdf, gold["sdf"]) < 1e-5
    assert float((sdf.cpu() - torch.from_numpy(gold["sdf"])).abs().max()) < 2e-6
    grad = pkg.decode_sdf_gradient(dec, lat, pts, clamp_dist=0.1, engine=engine)
    err = (grad.cpu() - torch.from_numpy(gold["grad"])).norm(dim=1)
    assert int((err > 1e-3).sum()) <= 3          # ReLU-boundary flips
    keep = err <= 1e-3
    assert gu.rel(grad.cpu()[keep], torch.from_numpy(gold["grad"])[keep]) < 1e-5


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_decode_sdf_autograd(engine):
    """decode_sdf is differentiable w.r.t. latent and points like the reference's (decoder_utils.py:53)."""
    from oracle.sdf_oracle import decode_sdf as o_decode
    dec_c, dec_g = cases.decoder("B"), gu.gpu_decoder("B")
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(500, 3, generator=g) - 0.5) * 1.2
    w = torch.randn(500, 1, generator=g)
    lat_c = cases.synth.make_latent().requires_grad_(True)
    p_c = pts.clone().requires_grad_(True)
    (o_decode(dec_c, lat_c, p_c, clamp_dist=0.1) * w).sum().backward()
    lat_g = cases.synth.make_latent().cuda().requires_grad_(True)
    p_g = pts.cuda().requires_grad_(True)
    (pkg.decode_sdf(dec_g, lat_g, p_g, clamp_dist=0.1, engine=engine) * w.cuda()).sum().backward()
    assert gu.rel(lat_g.grad, lat_c.grad) < 1e-4
    assert gu.rel(p_g.grad, p_c.grad) < 1e-4


def test_tc_operand_range_margin_and_fallback():
    """The split-fp16 operands carry activations in units of 32, i.e. up to ~2000 before fp16 overflows.  Scale the
    weight-norm gains of the synthetic decoder so that its hidden activations reach trained-network magnitudes (1e2..1e3)
    and beyond: within the range the tensor-core engine stays at fp32 noise from the exact engine; beyond it, the
    prepare-time self-check disables it LOUDLY (warning; 'auto' resolves to the fp32 engine, engine='tc' raises) instead of
    returning non-finite values.  Prints the margin table."""
    import copy
    import warnings
    g = torch.Generator().manual_seed(2)
    pts = ((torch.rand(20000, 3, generator=g) - 0.5) * 1.6).cuda()
    lat = cases.synth.make_latent().cuda()
    rows = []
    for gain in (1.0, 1.6, 2.2, 3.0, 4.0, 5.5):
        dec = copy.deepcopy(cases.decoder("B")).cuda()
        with torch.no_grad():
            for l in range(1, 8):
                getattr(dec, "lin%d" % l).weight_g.mul_(gain)
            last = dec.lin8       # keep the sdf itself in range (un-saturated tanh), so that its error means something
            (last.weight_g if hasattr(last, "weight_g") else last.weight).mul_(gain ** -7)
        acts = []
        hooks = [getattr(dec, "lin%d" % l).register_forward_hook(lambda m, i, o: acts.append(float(o.abs().max()))) for l in range(8)]
        with torch.no_grad():
            dec._inference_torch(torch.cat([lat.expand(4096, -1), pts[:4096]], 1))
        for h in hooks:
            h.remove()
        amax = max(acts)
        ref = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine="simt")
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            try:
                out = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine="tc")
                err, state = float((out - ref).abs().max()), "tc"
            except NotImplementedError:
                err, state = float("nan"), "tc refused"
            auto = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine="auto")
        assert bool(torch.isfinite(auto).all()) and float((auto - ref).abs().max()) < 5e-5       # 'auto' is always safe
        refused = state == "tc refused" or len(wlist) >= 1      # the first call after a failed self-check warns and falls back
        rows.append((gain, amax, "fp32 (tc refused)" if refused else "tc", err, len(wlist)))
        if amax < 1500.0:
            assert not refused and err < 2e-5, rows[-1]
        if refused:
            assert amax > 1500.0, rows[-1]
    print("\n gain  max|activation|  engine      max|tc - fp32|  warnings")
    for r in rows:
        print(" %4.1f  %14.1f  %-18s  %14.3g  %d" % r)
    assert any(r[1] > 100.0 and r[2] == "tc" for r in rows)          # trained-network-like magnitudes are covered
    assert any(r[2] != "tc" for r in rows)                           # and the overflow case is exercised


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_decode_color_runs_on_the_fused_engines(engine):
    """next-3: the three-output colour network (decoder_utils.py:94-112) on the CUDA engines -- both codes folded into the
    per-render biases, hidden layers shared, three dot-product epilogues -- against the module's eager PyTorch layers:
    values and the gradients w.r.t. shape code, colour code and points; and it is the library that runs (launch count)."""
    import copy
    import importlib
    lib = importlib.import_module("dist-renderer_b200._abi").lib()
    col = copy.deepcopy(cases.synth.make_color_decoder()).cuda()
    g = torch.Generator().manual_seed(21)
    pts = ((torch.rand(700, 3, generator=g) - 0.5) * 1.2).cuda()
    w = torch.randn(700, 3, generator=g).cuda()
    shape0, cc0 = cases.synth.make_latent().cuda(), (0.1 * torch.randn(1, 8, generator=g)).cuda()
    sa, ca, pa = shape0.clone().requires_grad_(True), cc0.clone().requires_grad_(True), pts.clone().requires_grad_(True)
    pkg.decode_color(col, ca, sa, pa, engine=engine)          # engine preparation outside the counted region
    n0 = lib.dist_launch_count()
    rgb = pkg.decode_color(col, ca, sa, pa, engine=engine)
    assert lib.dist_launch_count() - n0 >= 3 and rgb.shape == (700, 3)
    (rgb * w).sum().backward()
    sb, cb, pb = shape0.clone().requires_grad_(True), cc0.clone().requires_grad_(True), pts.clone().requires_grad_(True)
    ref = col._inference_torch(torch.cat([sb.expand(700, -1), cb.expand(700, -1), pb], 1))
    (ref * w).sum().backward()
    assert gu.rel(rgb, ref) < 1e-5 and float((rgb - ref).abs().max()) < 5e-6
    assert gu.rel(sa.grad, sb.grad) < 1e-4 and gu.rel(ca.grad, cb.grad) < 1e-4 and gu.rel(pa.grad, pb.grad) < 1e-4


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_matches_oracle(name, engine):
    cs = cases.CASES[name]
    out, g, ren = gu.run_gpu(cs, engine=engine)
    ref, gref = gu.run_oracle(cs)
    res = gu.compare(out, ref, g, gref)
    print(name, res, "rows", int(ren.rows_evaluated.item()))


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_matches_golden(name, engine):
    """Directly against the outputs of the unmodified reference stored in tests/golden."""
    cs = cases.CASES[name]
    gold = np.load(os.path.join(cases.GOLDEN_DIR, name + ".npz"))
    out, g, _ = gu.run_gpu(cs, engine=engine)
    ref = [torch.from_numpy(gold[k]) for k in ("depth", "normal", "mask", "min_sdf")]
    gref = [torch.from_numpy(gold[k]) for k in ("g_latent", "g_R", "g_T")]
    gu.compare(out, ref, g, gref)


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_render_normal_isolated(engine):
    """render_normal fed with the ORACLE's Zdepth / mask (identical hit points): strict bar, <= 0.1 % outliers."""
    from oracle.sdf_oracle import OracleSDFRenderer
    cs = cases.CASES["c1_recursive_64"]
    K, R, T = cases.camera(cs["cam"], cs["hw"])
    lat = cases.synth.make_latent()
    ora = OracleSDFRenderer(cases.decoder("B"), K, img_hw=cs["hw"], march_step=50, buffer_size=5)
    Z, m, _ = ora.render_depth(lat, R, T, no_grad=True)
    n_ref = ora.render_normal(lat, R, T, Z, m, no_grad=True)
    ren = pkg.SDFRenderer(gu.gpu_decoder("B"), K, img_hw=cs["hw"], march_step=50, buffer_size=5, engine=engine)
    n_gpu = ren.render_normal(lat.cuda(), R.cuda(), T.cuda(), Z.cuda(), m.cuda())
    assert n_gpu.shape == (3, 64 * 64)
    assert float(n_gpu.cpu()[:, ~m].abs().max()) == 0.0
    r, n_out, _ = gu.normal_error(n_gpu.cpu().t(), n_ref.t(), m, outlier_frac=0.001)
    assert r < 1e-5 and n_out <= 2, (r, n_out)


def test_render_depth_no_grad_matches_golden():
    """no_grad render: the value-neutral (z-a)+a roundings are skipped exactly as renderer.py:410-411."""
    cs = cases.CASES["c1_recursive_64"]
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "c1_recursive_64.npz"))
    dec = gu.gpu_decoder("B")
    K, R, T = cases.camera(cs["cam"], cs["hw"])
    ren = pkg.SDFRenderer(dec, K, img_hw=cs["hw"], march_step=50, buffer_size=5, engine="simt")
    Z, m, s = ren.render_depth(cases.synth.make_latent().cuda(), R.cuda(), T.cuda(), no_grad=True)
    assert Z.dtype == torch.float32 and m.dtype == torch.bool and Z.shape == (64 * 64,)
    zg = torch.from_numpy(gold["Zdepth_nograd"])
    hit = zg < 1e10
    assert bool(((Z.cpu() < 1e10) == hit).all())
    assert gu.rel(Z.cpu()[hit], zg[hit]) < 1e-5


def test_tc_engine_matches_simt_at_scale():
    """1 M random rows: tensor-core engine (split-fp16, truncation-compensated) vs the exact-fp32 SIMT engine."""
    dec = gu.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    g = torch.Generator().manual_seed(5)
    pts = ((torch.rand(1000003, 3, generator=g) - 0.5) * 1.6).cuda()
    a = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, engine="simt")
    b = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, engine="tc")
    d = (a - b).abs()
    assert not bool(torch.isnan(b).any())
    assert float(d.max()) < 3e-6 and float(d.mean()) < 3e-7, (float(d.max()), float(d.mean()))
    # engine 'auto' resolves to the tensor-core engine for the standard spec on sm_100
    c = pkg.decode_sdf(dec, lat, pts[:1000], clamp_dist=None)
    assert torch.equal(c, b[:1000])


def test_tc_gradient_modes_match_simt_at_scale():
    """Input-gradient and backward-replay kernels of the tensor-core engine vs the fp32 SIMT engine, 200 K rows."""
    dec = gu.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    g = torch.Generator().manual_seed(8)
    n = 200003
    pts = ((torch.rand(n, 3, generator=g) - 0.5) * 1.4).cuda()
    ga = pkg.decode_sdf_gradient(dec, lat, pts, clamp_dist=0.1, engine="simt")
    gb = pkg.decode_sdf_gradient(dec, lat, pts, clamp_dist=0.1, engine="tc")
    err = (ga - gb).norm(dim=1)
    flips = int((err > 1e-3).sum())
    assert flips <= max(3, n // 2000), flips                   # ReLU-boundary flips only
    keep = err <= 1e-3
    assert gu.rel(gb[keep], ga[keep]) < 2e-6
    # positive row weights: a randomly signed sum over 200 K rows cancels to ~1/450 of its terms, so that a handful
    # of ReLU-boundary rows (present in ANY two fp32 evaluations: the fp32 oracle itself sits 2e-4..6e-4 from its fp64
    # twin on such a sum) would dominate the comparison
    w = torch.rand(n, 1, generator=g).cuda() + 0.1
    res = {}
    for eng in ("simt", "tc"):
        l = lat.clone().requires_grad_(True)
        p = pts.clone().requires_grad_(True)
        (pkg.decode_sdf(dec, l, p, clamp_dist=0.1, engine=eng) * w).sum().backward()
        res[eng] = (l.grad.clone(), p.grad.clone())
    assert gu.rel(res["tc"][0], res["simt"][0]) < 3e-4          # d/dlatent: sums over 200 K rows incl. flipped ones
    e2 = (res["tc"][1] - res["simt"][1]).norm(dim=1)
    k2 = e2 <= 1e-3 * w.abs().reshape(-1).clamp(min=1e-3)
    assert int((~k2).sum()) <= max(3, n // 2000)
    assert gu.rel(res["tc"][1][k2], res["simt"][1][k2]) < 2e-6


def test_tc_engine_ragged_counts_and_clamp():
    dec = gu.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    g = torch.Generator().manual_seed(6)
    for n in (1, 63, 64, 65, 127, 128, 129, 200, 9473):
        pts = ((torch.rand(n, 3, generator=g) - 0.5) * 1.6).cuda()
        a = pkg.decode_sdf(dec, lat, pts, clamp_dist=0.1, engine="simt")
        b = pkg.decode_sdf(dec, lat, pts, clamp_dist=0.1, engine="tc")
        assert float((a - b).abs().max()) < 3e-6, n
        assert float(b.abs().max()) <= 0.1 + 1e-7


def test_small_generic_network():
    """Non-standard shape (width 64, 4 hidden layers, latent 16, latent_in=2) through the generic SIMT engine."""
    dec_c = cases.synth.make_decoder("B", latent_size=16, width=64, depth=4, latent_in=2, seed=5)
    import copy
    dec_g = copy.deepcopy(dec_c).cuda()
    cs = dict(decoder=None, hw=(24, 24), cam=("front", 1.6), march_step=30, buffer_size=3, kind="recursive")
    out, g, _ = gu.run_gpu(cs, engine="simt", dec=dec_g)
    ref, gref = gu.run_oracle(cs, dec=dec_c)
    gu.compare(out, ref, g, gref)


@pytest.mark.parametrize("shape", [dict(latent_size=64, width=256, depth=6, latent_in=3),
                                   dict(latent_size=32, width=512, depth=4, latent_in=2),
                                   dict(latent_size=128, width=384, depth=5, latent_in=None)])
def test_tc_engine_other_network_shapes(shape):
    """The tensor-core engine on non-standard DeepSDF shapes (other widths / depths / latent sizes / no latent_in)."""
    import copy
    dec_c = cases.synth.make_decoder("B", seed=11, **shape)
    dec_g = copy.deepcopy(dec_c).cuda()
    lat = cases.synth.make_latent(shape["latent_size"])
    g = torch.Generator().manual_seed(12)
    pts = ((torch.rand(20000, 3, generator=g) - 0.5) * 1.4)
    a = pkg.decode_sdf(dec_g, lat.cuda(), pts.cuda(), clamp_dist=None, engine="simt")
    b = pkg.decode_sdf(dec_g, lat.cuda(), pts.cuda(), clamp_dist=None, engine="tc")
    assert float((a - b).abs().max()) < 3e-6
    ga = pkg.decode_sdf_gradient(dec_g, lat.cuda(), pts.cuda(), engine="simt")
    gb = pkg.decode_sdf_gradient(dec_g, lat.cuda(), pts.cuda(), engine="tc")
    err = (ga - gb).norm(dim=1)
    assert int((err > 1e-3).sum()) <= 10 and gu.rel(gb[err <= 1e-3], ga[err <= 1e-3]) < 5e-6
    cs = dict(decoder=None, hw=(24, 24), cam=("front", 1.6), march_step=30, buffer_size=3, kind="pyramid_recursive")
    out, gr, _ = gu.run_gpu(cs, engine="tc", dec=dec_g)
    ref, gref = gu.run_oracle(cs, dec=dec_c)
    gu.compare(out, ref, gr, gref)


def test_row_band_sharding_equals_full():
    """Rendering interleaved row bands (multi-GPU ray-tile sharding) reproduces the full image exactly."""
    cs = cases.CASES["ragged_37x53"]
    dec = gu.gpu_decoder("B")
    K, R, T = cases.camera(cs["cam"], cs["hw"])
    lat = cases.synth.make_latent().cuda()
    full = pkg.SDFRenderer(dec, K, img_hw=cs["hw"], march_step=cs["march_step"], engine="simt").render(
        lat, R.cuda(), T.cuda(), ray_marching_type="recursive", no_grad=True)
    H = cs["hw"][0]
    for r in range(3):
        n_rows = len(range(r, H, 3))
        part = pkg.SDFRenderer(dec, K, img_hw=cs["hw"], march_step=cs["march_step"], engine="simt",
                               rows=(r, 3, n_rows)).render(lat, R.cuda(), T.cuda(), ray_marching_type="recursive",
                                                           no_grad=True)
        for a, b in zip(part, full):
            assert torch.equal(a, b[r::3])


def test_forward_sampling_matches_reference_formula():
    """render(num_forward_sampling=k): sdf + offset at points pushed inside along the ray (renderer.py:912-941),
    checked against the oracle's decoder on the same points."""
    from oracle.sdf_oracle import decode_sdf as o_decode
    cs = cases.CASES["c1_recursive_64"]
    dec_g, dec_c = gu.gpu_decoder("B"), cases.decoder("B")
    K, R, T = cases.camera(cs["cam"], cs["hw"])
    lat = cases.synth.make_latent()
    ren = pkg.SDFRenderer(dec_g, K, img_hw=cs["hw"])
    out = ren.render(lat.cuda(), R.cuda(), T.cuda(), ray_marching_type="recursive", no_grad=True, num_forward_sampling=3)
    assert len(out) == 5 and out[4].shape == (64, 64, 3)
    Z, m, _ = ren.render_depth(lat.cuda(), R.cuda(), T.cuda(), no_grad=True)
    cam_pos = ren.get_camera_location(R.cuda(), T.cuda())
    rays = ren.get_camera_rays(R.cuda())
    for i in range(3):
        grid = 0.5 * 0.1 * (i + 1) / 3
        pts = ren.generate_point_samples(cam_pos, rays[:, m], Z[m] + grid).detach().t().cpu()
        ref = o_decode(dec_c, lat, pts, clamp_dist=None).squeeze(-1).detach() + grid
        got = out[4].reshape(-1, 3)[m.cpu(), i].cpu()
        assert float((got - ref).abs().max()) < 5e-6
    assert float(out[4].reshape(-1, 3)[~m.cpu()].abs().max()) == 0.0


def test_tc_self_check_downgrades_out_of_range_decoder():
    """Activations far beyond the fp16 operand range: the tensor-core self-check must trip, 'auto' must run on the
    fp32 engine (same results as engine='simt'), and an explicit engine='tc' must raise."""
    import copy
    import warnings
    dec = copy.deepcopy(cases.decoder("B"))
    with torch.no_grad():
        dec.lin1.weight_g.mul_(3e4)
    dec = dec.cuda()
    lat = cases.synth.make_latent().cuda()
    pts = ((torch.rand(3000, 3) - 0.5) * 1.2).cuda()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a = pkg.decode_sdf(dec, lat, pts, clamp_dist=None)
    assert any("tensor-core engine disabled" in str(x.message) for x in w)
    b = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, engine="simt")
    assert torch.equal(a, b) and not bool(torch.isnan(a).any())
    with pytest.raises(NotImplementedError):
        pkg.decode_sdf(dec, lat, pts, clamp_dist=None, engine="tc")


def test_decoder_inference_fused_path():
    """Decoder.inference under no_grad with a shared latent runs on the fused engines and equals the torch layers."""
    dec = gu.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    pts = ((torch.rand(5000, 3) - 0.5) * 1.2).cuda()
    x = torch.cat([lat.expand(5000, -1), pts], 1)
    ref = dec._inference_torch(x).detach()
    n0 = cases.pkg._abi.lib().dist_launch_count() if hasattr(cases.pkg, "_abi") else None
    with torch.no_grad():
        got = dec.inference(x)
    assert got.shape == (5000, 1) and float((got - ref).abs().max()) < 3e-6
    x2 = x.clone()
    x2[7, 0] += 0.5                                     # per-row latents: generic path, still correct
    with torch.no_grad():
        assert float((dec.inference(x2) - dec._inference_torch(x2)).abs().max()) == 0.0
    assert dec.inference(x).grad_fn is not None         # grad mode: plain PyTorch graph (weights may be trained)


def test_use_transform_false_and_custom_matrix():
    """use_transform=False (points stay in the world frame, normals are still mapped by transform_matrix:
    renderer.py:219-220 vs :902) and a non-default transform_matrix, vs the oracle."""
    from oracle.sdf_oracle import OracleSDFRenderer
    dec_c, dec_g = cases.decoder("B"), gu.gpu_decoder("B")
    K, R, T = cases.camera(("lookat", 70.0, 35.0, 1.8, 1.1), (40, 40))
    lat = cases.synth.make_latent()
    Mrot = np.array([[0., 1., 0.], [0., 0., 1.], [1., 0., 0.]])
    for tm, ut in ((None, False), (Mrot, True)):
        ora = OracleSDFRenderer(dec_c, K, img_hw=(40, 40), transform_matrix=tm)
        ren = pkg.SDFRenderer(dec_g, K, img_hw=(40, 40), transform_matrix=tm)
        ref = ora.render(lat, R, T, ray_marching_type="recursive", no_grad=True, use_transform=ut)
        out = ren.render(lat.cuda(), R.cuda(), T.cuda(), ray_marching_type="recursive", no_grad=True, use_transform=ut)
        gu.compare([t.cpu() for t in out], [t.detach() for t in ref])


def test_use_depth2normal_matches_oracle():
    """renderer.py:972-975: normals from central differences of the rendered depth (the function itself is pinned to the
    reference bit for bit on CPU, tests/test_oracle.py); here the wiring and the end-to-end agreement with the oracle."""
    from oracle.sdf_oracle import OracleSDFRenderer
    hw = (40, 40)
    K, R, T = cases.camera(("front", 1.6), hw)
    lat = cases.synth.make_latent()
    ora = OracleSDFRenderer(cases.decoder("B"), K, img_hw=hw, use_depth2normal=True)
    ren = pkg.SDFRenderer(gu.gpu_decoder("B"), K, img_hw=hw, use_depth2normal=True)
    ref = ora.render(lat, R, T, ray_marching_type="recursive", no_grad=True)
    out = [t.cpu() for t in ren.render(lat.cuda(), R.cuda(), T.cuda(), ray_marching_type="recursive", no_grad=True)]
    assert out[0].shape == hw and out[1].shape == hw + (3,) and out[2].dtype == torch.uint8
    mg, mo = out[2].bool(), ref[2].bool()
    assert int((mg != mo).sum()) <= 2
    assert float(out[0].min()) == 0.0 and float(out[0].max()) < 1e5        # background zeroed in place, as upstream
    both = mg & mo
    assert gu.rel(out[0][both], ref[0][both]) < 1e-5
    # compare normals where the 3x3 neighbourhood has the same silhouette in both (a flipped neighbour changes the stencil)
    agree = (mg == mo).float()[None, None]
    safe = torch.nn.functional.avg_pool2d(agree, 3, stride=1, padding=1)[0, 0] > 0.999
    assert int(safe.sum()) > 0.8 * hw[0] * hw[1]
    assert float((out[1] - ref[1])[safe].abs().max()) < 2e-3
    assert float(out[1][~mg].abs().max()) == 0.0


def test_api_errors():
    dec = gu.gpu_decoder("B")
    K, R, T = cases.camera(("front", 1.6), (16, 16))
    ren = pkg.SDFRenderer(dec, K, img_hw=(16, 16))
    lat = cases.synth.make_latent().cuda()
    with pytest.raises(NotImplementedError):
        pkg.SDFRenderer(dec, K, img_hw=(16, 16), scale_list=[2, 1], march_step_list=[3, -1]).render(lat, R.cuda(), T.cuda())
    with pytest.raises(ValueError):
        ren.render_depth(lat, R.cuda(), T.cuda(), ray_marching_type="bogus")
    with pytest.raises(RuntimeError):
        pkg.SDFRenderer(dec, K, img_hw=(16, 16), use_gpu=False)
    with pytest.raises(ValueError):
        pkg.SDFRenderer(cases.decoder("B"), K, img_hw=(16, 16))   # CPU decoder
    K2 = K.copy()
    K2[0, 2] += 4000.0
    with pytest.raises(ValueError, match="No valid depth"):
        pkg.SDFRenderer(dec, K2, img_hw=(16, 16)).render_depth(lat, R.cuda(), torch.tensor([0., 0., 50.]).cuda())
