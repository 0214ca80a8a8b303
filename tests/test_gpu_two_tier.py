"""GPU: two-tier precision of the march rows on the tensor-core engine (dist_march_t.screen, csrc/mlp_tc.cu).

Rows whose sdf is safely beyond the march's clamp are evaluated with ONE fp16 tensor-core pass; everything the
reference's results depend on (rows inside the clamp band, each ray's smallest |sdf|) is evaluated with the three
split-precision passes.  So a screened render must equal the unscreened one, and the kernel-level contract of
dist_decoder_forward_tiers must hold row by row.
"""
import importlib

import pytest
import torch

import cases
import gpu_util as gu

pytestmark = pytest.mark.gpu
pkg = cases.pkg
synth = cases.synth
abi = importlib.import_module("dist-renderer_b200._abi")
tc = importlib.import_module("dist-renderer_b200.tc")
plan_mod = importlib.import_module("dist-renderer_b200.plan")


def _tiers(plan, lat, pts, n_screen, n_exact, offset, thresh):
    """dist_decoder_forward_tiers on rows [0, n_screen) (one pass first) and [offset, offset + n_exact) (three passes)."""
    lib, st = abi.lib(), torch.cuda.current_stream().cuda_stream
    tc.prepare(plan)
    b0, bl, _ = plan.fold(lat, st)
    bl_tc = bl * tc.S_ACT if bl is not None else None       # keep alive: the descriptor holds raw addresses
    net = plan.c_net(b0, bl, bl_tc)
    n = pts.shape[0]
    sdf = torch.full((n,), float("nan"), device="cuda")
    seg = torch.full(((n + 63) // 64,), 7, device="cuda", dtype=torch.uint8)
    cnt = torch.zeros(2, device="cuda", dtype=torch.int64)
    abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n_screen, n_exact, offset, float(thresh), abi.ptr(sdf),
                                             abi.ptr(seg), abi.ptr(cnt), st))
    torch.cuda.synchronize()
    del b0, bl, bl_tc
    return sdf, seg, cnt


@pytest.mark.parametrize("n", [128 * 300 + 37, 128 * 74 * 3, 100])
def test_tiers_kernel_contract(n):
    dec = gu.gpu_decoder("B")
    plan = plan_mod.plan_for(dec)
    lat = synth.make_latent().cuda()
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(n, 3, generator=g) - 0.5) * 1.6
    pts = pts[pts.norm(dim=1).argsort()].contiguous().cuda()     # radius order: coherent tiles, some all far, some mixed
    exact = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine="tc").reshape(-1)
    tiles = (n + 127) // 128
    thresh = 0.102
    # (a) every row in the one-pass segment
    sdf, seg, cnt = _tiers(plan, lat, pts, n, 0, 0, thresh)
    assert int(seg.max()) <= 1
    row_approx = seg.repeat_interleave(64)[:n].bool()
    assert bool(torch.equal(sdf[~row_approx], exact[~row_approx]))          # full-precision rows: bit-identical
    if bool(row_approx.any()):
        assert float(exact[row_approx].abs().min()) > 0.1                   # one-pass values only beyond the clamp
        assert float((sdf[row_approx] - exact[row_approx]).abs().max()) < tc.SCREEN_MARGIN / 2
    pad = (-n) % 128
    far_tile = torch.cat([exact.abs() > thresh + tc.SCREEN_MARGIN, torch.ones(pad, dtype=torch.bool, device="cuda")]).reshape(-1, 128).all(1)
    seg_tile = torch.cat([seg, seg.new_ones(2 * tiles - seg.numel())]).reshape(-1, 2).bool().all(1)
    assert bool(seg_tile[far_tile].all())                                    # clearly far tiles keep their one-pass values
    assert int(cnt[0]) == tiles and int(cnt[1]) == int((~seg_tile).sum())    # failed tiles are redone, once
    # (b) every row in the full-precision segment (placed at an offset): three passes directly, nothing flagged
    off = 128 * 5
    pts_b = torch.cat([torch.zeros(off, 3, device="cuda"), pts])
    sdf, seg, cnt = _tiers(plan, lat, pts_b, 0, n, off, thresh)
    assert bool(torch.equal(sdf[off:], exact)) and int(seg[off // 64:].max()) == 0 and cnt.tolist() == [0, tiles]
    assert bool(torch.isnan(sdf[:off]).all())                                # rows outside both segments are untouched
    # (c) both segments: the first n1 rows screened, the rest (stored behind a gap) at full precision
    n1 = (n // 3) // 128 * 128 + min(17, n // 4)
    off = (n1 + 127) // 128 * 128 + 256
    pts_c = torch.cat([pts[:n1], torch.zeros(off - n1, 3, device="cuda"), pts[n1:]])
    sdf, seg, cnt = _tiers(plan, lat, pts_c, n1, n - n1, off, thresh)
    assert bool(torch.equal(sdf[off:], exact[n1:])) and int(seg[off // 64:].max()) == 0
    a1 = seg[: (n1 + 63) // 64].repeat_interleave(64)[:n1].bool()
    assert bool(torch.equal(sdf[:n1][~a1], exact[:n1][~a1]))
    if bool(a1.any()):
        assert float(exact[:n1][a1].abs().min()) > 0.1
    assert int(cnt[0]) == (n1 + 127) // 128 and int(cnt[1]) >= (n - n1 + 127) // 128


@pytest.mark.parametrize("kind", ["recursive", "pyramid_recursive", "trivial"])
@pytest.mark.parametrize("hw,cam", [((160, 160), ("front", 1.6)), ((96, 130), ("lookat", 40.0, 25.0, 2.5, 1.2 * 2.5 / 1.6))])
def test_screened_render_equals_full_precision_render(kind, hw, cam):
    """Same trajectories, same selected samples, same maps: the one-pass values never reach an output."""
    dec = gu.gpu_decoder("B")
    K, R, T = cases.camera(cam, hw)
    lat0, R, T = synth.make_latent().cuda(), R.cuda(), T.cuda()
    outs, grads, counters = [], [], []
    for screen in (True, False):
        ren = pkg.SDFRenderer(dec, K, img_hw=hw, march_step=50, buffer_size=5, engine="tc", screen=screen)
        lat = lat0.clone().requires_grad_(True)
        Rg, Tg = R.clone().requires_grad_(True), T.clone().requires_grad_(True)
        out = ren.render(lat, Rg, Tg, ray_marching_type=kind)
        cases.scalar_loss(out).backward()
        outs.append([o.detach() for o in out])
        grads.append((lat.grad, Rg.grad, Tg.grad))
        counters.append(ren.tile_counters.tolist())
    assert counters[0][0] > 0 and counters[1][0] == 0, counters       # the screened run did use one-pass tiles
    print(kind, hw, "tile programs [1-pass, 3-pass]: screened", counters[0], "full", counters[1])
    for name, a, b in zip(("depth", "normal", "mask", "min_sdf"), outs[0], outs[1]):
        assert torch.equal(a, b), (name, int((a != b).sum()), float((a.float() - b.float()).abs().max()))
    for a, b in zip(grads[0], grads[1]):
        assert gu.rel(a, b) < 1e-5


def test_screen_can_be_disabled_and_is_off_for_simt():
    dec = gu.gpu_decoder("B")
    K, (R, T) = synth.intrinsic(64, 64), synth.front_camera()
    lat, R, T = synth.make_latent().cuda(), R.cuda(), T.cuda()
    r0 = pkg.SDFRenderer(dec, K, img_hw=(64, 64), engine="simt")
    r0.render(lat, R, T, ray_marching_type="recursive", no_grad=True)
    assert r0.tile_counters.tolist() == [0, 0]
    r1 = pkg.SDFRenderer(dec, K, img_hw=(64, 64), engine="tc", screen=False)
    r1.render(lat, R, T, ray_marching_type="recursive", no_grad=True)
    assert r1.tile_counters[0].item() == 0 and r1.tile_counters[1].item() > 0


def test_mask_cache_kernels():
    """Mask cache (csrc/mlp_tc.cu): a forward launch records the ReLU sign bits of every hidden layer; the backward replay
    from those bits (mode 3: transposed chain only) equals the full replay (mode 2: forward + transposed chain)."""
    lib, st = abi.lib(), torch.cuda.current_stream().cuda_stream
    dec = gu.gpu_decoder("B")
    plan = plan_mod.plan_for(dec)
    tc.prepare(plan)
    lat = synth.make_latent().cuda()
    g = torch.Generator().manual_seed(8)
    n = 128 * 37 + 19
    pts = ((torch.rand(n, 3, generator=g) - 0.5) * 1.4).cuda()
    coef = torch.randn(n, generator=g).cuda()
    b0, bl, _ = plan.fold(lat, st)
    bl_tc = bl * tc.S_ACT
    net = plan.c_net(b0, bl, bl_tc)
    words, cap, base = 16 * (plan.n_layers - 1), 128 * 64, 128 * 3
    masks = torch.zeros(words * cap, device="cuda", dtype=torch.int32)
    sdf = torch.empty(n, device="cuda")
    abi.check(lib.dist_decoder_forward_masks(net, abi.ptr(pts), n, abi.ptr(sdf), abi.ptr(masks), cap, base, st))
    exact = pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine="tc").reshape(-1)
    assert torch.equal(sdf, exact)
    # the recorded bits against the module's eager layers (a pre-activation within rounding of zero may differ)
    acts = []
    hooks = [getattr(dec, "lin%d" % l).register_forward_hook(lambda m, i, o: acts.append(o)) for l in range(plan.n_layers - 1)]
    with torch.no_grad():
        dec._inference_torch(torch.cat([lat.expand(n, -1), pts], 1))
    for h in hooks:
        h.remove()
    mv = masks.view(words, cap)[:, base:base + n]
    bad = tot = 0
    for l, a in enumerate(acts):
        width = a.shape[1]
        bits = (a > 0).to(torch.int64)
        for kb in range((width + 31) // 32):
            blk = bits[:, 32 * kb:32 * kb + 32]
            word = (blk << torch.arange(blk.shape[1], device="cuda")).sum(1)
            got = mv[l * 16 + kb].to(torch.int64) & 0xFFFFFFFF
            valid = (1 << blk.shape[1]) - 1
            diff = (got ^ word) & valid
            bad += int(sum(((diff >> j) & 1).sum() for j in range(blk.shape[1])))
            tot += blk.numel()
    assert bad <= max(4, 2e-5 * tot), (bad, tot)
    # backward from the cache == full replay
    d_ref, a0_ref, al_ref = torch.empty(n, 3, device="cuda"), torch.zeros(plan.bias[0].numel(), device="cuda"), \
        torch.zeros(plan.bias[plan.latent_in].numel(), device="cuda")
    abi.check(lib.dist_decoder_backward(net, abi.ENGINE_TC, abi.ptr(pts), abi.ptr(coef), None, n, None, 0.0, abi.ptr(d_ref),
                                        abi.ptr(a0_ref), abi.ptr(al_ref), st))
    slots = (base + torch.arange(n, device="cuda")).to(torch.int32)
    d_m, a0_m, al_m = torch.empty(n, 3, device="cuda"), torch.zeros_like(a0_ref), torch.zeros_like(al_ref)
    abi.check(lib.dist_decoder_backward_masked(net, abi.ptr(slots), abi.ptr(sdf), abi.ptr(coef), n, 0.0, abi.ptr(masks), cap,
                                               abi.ptr(d_m), abi.ptr(a0_m), abi.ptr(al_m), st))
    torch.cuda.synchronize()
    assert torch.equal(d_m, d_ref)
    assert gu.rel(a0_m, a0_ref) < 1e-5 and gu.rel(al_m, al_ref) < 1e-5
    del b0, bl, bl_tc


@pytest.mark.parametrize("kind", ["recursive", "pyramid_recursive"])
def test_mask_cache_backward_equals_full_replay(kind):
    """render() + backward with the mask cache (default) and without: identical maps, gradients equal to rounding of the
    atomics, and the cached path really carries most of the replay rows."""
    dec = gu.gpu_decoder("B")
    hw = (128, 128)
    K, R, T = cases.camera(("lookat", 40.0, 25.0, 2.5, 1.2 * 2.5 / 1.6), hw)
    res = []
    for mc in (True, False):
        ren = pkg.SDFRenderer(dec, K, img_hw=hw, march_step=50, buffer_size=5, engine="tc", mask_cache=mc)
        lat = synth.make_latent().cuda().requires_grad_(True)
        Rg, Tg = R.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
        out = ren.render(lat, Rg, Tg, ray_marching_type=kind)
        cases.scalar_loss(out).backward()
        torch.cuda.synchronize()
        res.append(([o.detach() for o in out], (lat.grad, Rg.grad, Tg.grad), int(ren._scr["bm_cnt"].item()) if mc else 0,
                    int(ren._scr["b_cnt"].item())))
    (o1, g1, cached, rest), (o2, g2, _, full) = res
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)
    for a, b in zip(g1, g2):
        assert gu.rel(a, b) < 1e-5
    print(kind, "replay rows from the cache", cached, "full replay", rest, "(without the cache:", full, ")")
    assert cached + rest == full and cached > 0.8 * full
