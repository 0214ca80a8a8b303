"""Tensor-core engine operand preparation (split-fp16 weight tiles for csrc/mlp_tc.cu).

Precision scheme (single-pass TF32/BF16 breaks the 5e-5 convergence test): every fp32 operand x is
represented as hi + lo with hi = fp16(s*x), lo = fp16(s*x - hi) for a power-of-two scale s, and each logical GEMM
is issued as three fp16 tensor-core passes  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  with fp32 accumulation in TMEM
(the dropped lo*lo term is 2^-22 relative).  fp16 (11-bit significand) x 2 carries 22 bits -- the precision of the
"TF32x2" row of SURVEY Table P -- at the fp16/bf16 MMA rate (2x the TF32 rate) and half the operand bytes.

Blob layout (must match csrc/mlp_tc.cu): for each tensor-core layer, for each 256-wide N half h, for each 32-wide K
chunk kc, for each CTA r of the pair: a 16 KB stage = [hi 8 KB][lo 8 KB], each
[4 K-groups][128 n-rows][8 k] fp16 holding  W[n = 256 h + 128 r + row][k = 32 kc + 8 g + e] * sW  -- the no-swizzle
K-major "panel" image the UMMA shared-memory descriptor reads (LBO = 2048 B between K-groups, SBO = 128 B between
8-row groups), so one TMA box copy lands a stage without any reshuffling.
The transposed chain (input-gradient / backward) uses the same layout built from W^T.
"""
import math

import torch

from . import _abi

SCREEN_MARGIN = 2e-3  # two-tier precision: a one-pass sdf beyond clamp + margin is "far"; one-pass error must be < margin / 2
S_ACT = 32.0        # activation scale (power of two): post-ReLU activations up to ~2000 stay finite in fp16
S_GRAD = 256.0      # scale of the backward-chain operand (d sdf / d pre-activation, unit seed)


def supported(plan):
    """True when the tcgen05 engine covers this decoder shape on this device."""
    if not torch.cuda.is_available() or getattr(plan, "tc_unsafe", False):
        return False
    if _abi.lib().dist_device_supports_tc(plan.device.index if plan.device.index is not None else 0) != 1:
        return False
    n = plan.n_layers
    if n < 4 or n > 10:
        return False
    if plan.K[0] != 3:
        return False
    for l in range(n - 1):
        if plan.N[l] > 512 or plan.K[l + 1] > 512:      # A buffer / barrier sets cover 16 blocks of 32 features
            return False
        app = 3 if (l + 1 == plan.latent_in) else 0
        if plan.N[l] + app > 256 * ((plan.N[l] + 255) // 256):
            return False
    return True


def _pow2_scale(w, target_log2=14):
    m = float(w.abs().max())
    if m == 0.0 or not math.isfinite(m):
        return 1.0
    return 2.0 ** (target_log2 - 1 - math.floor(math.log2(m)))


def _split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi, lo


def _tiles(w, scale, c_trunc):
    """w: [N, K] fp32 (logical B operand: N output rows, K reduction).  Returns the stage blob for this layer as a
    flat fp16 tensor plus (k chunks of 32, n halves).

    Truncation pre-compensation: the tensor core adds each K=16 block into the fp32 accumulator with truncation
    (round toward zero), i.e. every one of the 3*J accumulation steps of a layer loses on average c*S_i of the
    running sum S_i.  The expected loss  -c * sum_i S_i = -c * sum_j 3(J-j) p_j  is a linear functional of the block
    products p_j, so scaling the weights of K-block j by (1 + 3c(J-j)) cancels it to first order (c is measured on
    the device by calibrate()).  What remains is zero-mean rounding noise of the size of an fp32 sequential sum's."""
    N, K = w.shape
    Kp, NH = ((K + 63) // 64) * 64, (N + 255) // 256
    wp = torch.zeros(NH * 256, Kp, device=w.device, dtype=torch.float64)
    J = Kp // 16
    comp = 1.0 + 3.0 * c_trunc * (J - torch.arange(Kp, device=w.device, dtype=torch.float64) // 16)
    wp[:N, :K] = w.double() * scale
    wp = (wp * comp[None, :]).float()
    hi, lo = _split(wp)
    kc = Kp // 32

    def panels(t):  # [NH*256, Kp] -> [kc, NH, 2(r), 4(g), 128(row), 8(e)]
        return t.reshape(NH, 2, 128, kc, 4, 8).permute(3, 0, 1, 4, 2, 5)
    both = torch.stack([panels(hi), panels(lo)], 3)          # [kc, NH, r, hi/lo, g, row, e]
    both = both.permute(1, 0, 2, 3, 4, 5, 6)                  # stage order of the kernel: N-half outer, K block inner
    return both.contiguous().reshape(-1), kc, NH


_C_TRUNC = {}       # device index -> measured per-accumulation truncation loss of the tcgen05 fp32 accumulator (_tiles)


def calibrate(plan, n_points=8192):
    """Measures the truncation constant c on this device: the final-sdf deviation of the tensor-core engine from the
    exact-fp32 SIMT engine is linear in the compensation constant, so two probes (c = 0 and c = c1) give its root."""
    dev_key = plan.device.index if plan.device.index is not None else torch.cuda.current_device()
    if dev_key in _C_TRUNC:
        return _C_TRUNC[dev_key]
    lib = _abi.lib()
    st = torch.cuda.current_stream(plan.device).cuda_stream
    g = torch.Generator().manual_seed(1234)
    pts = ((torch.rand(n_points, 3, generator=g) - 0.5) * 1.2).to(plan.device)
    lat = torch.zeros(plan.latent_size, device=plan.device) if plan.latent_size > 0 else None
    b0, bl, _ = plan.fold(lat, st)
    bl_tc = bl * S_ACT if bl is not None else None
    ref = torch.empty(n_points, device=plan.device)
    saved = plan.tc
    plan.tc = None
    _abi.check(lib.dist_decoder_forward(plan.c_net(b0, bl), _abi.ENGINE_SIMT, _abi.ptr(pts), n_points, None, 0.0,
                                        _abi.ptr(ref), st))

    def probe(c):
        plan.tc = _build(plan, c)
        out = torch.empty(n_points, device=plan.device)
        _abi.check(lib.dist_decoder_forward(plan.c_net(b0, bl, bl_tc), _abi.ENGINE_TC, _abi.ptr(pts), n_points, None,
                                            0.0, _abi.ptr(out), st))
        return float((out.double() - ref.double()).mean())
    c1 = 4.0e-8
    e0, e1 = probe(0.0), probe(c1)
    plan.tc = saved
    if not (math.isfinite(e0) and math.isfinite(e1)):
        # this decoder overflows the fp16 operands: nothing was measured, so nothing is cached for the device
        raise NotImplementedError("tensor-core calibration probes are not finite for this decoder")
    c = c1 * e0 / (e0 - e1) if abs(e0 - e1) > 1e-12 else 0.0
    if not (0.0 <= c <= 4.0e-7):      # outside anything physical: do not compensate
        c = 0.0
    _C_TRUNC[dev_key] = c
    return c


def prepare(plan):
    """Builds (once per weight version) the device blobs for the forward and the transposed chain."""
    if plan.tc is not None:
        return plan.tc
    if not supported(plan):
        raise NotImplementedError("the tensor-core engine does not cover this decoder shape / device")
    with torch.cuda.device(plan.device):
        dev_key = plan.device.index if plan.device.index is not None else torch.cuda.current_device()
        if dev_key not in _C_TRUNC:
            # the truncation constant is a property of the device, measured once -- but only through a decoder whose
            # operands provably fit fp16: self-check the uncompensated operands first (the bias being calibrated away is
            # ~5e-6, far below the self-check tolerance), then calibrate
            plan.tc = _build(plan, 0.0)
            _self_check(plan)
            try:
                calibrate(plan)
            except NotImplementedError:
                _mark_unsafe(plan, float("nan"))
        plan.tc = _build(plan, _C_TRUNC[dev_key])
        _self_check(plan)
        plan.tc["screen"] = _screen_check(plan)
    return plan.tc


def _self_check(plan, n_points=4096, tol=2e-5):
    """Guards the fp16 operand range: the split-fp16 engine must reproduce the exact-fp32 engine on a sample of this
    decoder's own rows (activations beyond ~2000 or non-finite values would overflow the fp16 operands).  On failure
    the plan is marked unsafe: 'auto' then resolves to the fp32 engine and engine='tc' raises."""
    lib = _abi.lib()
    st = torch.cuda.current_stream(plan.device).cuda_stream
    g = torch.Generator().manual_seed(4321)
    pts = ((torch.rand(n_points, 3, generator=g) - 0.5) * 2.0).to(plan.device)
    lat = (0.1 * torch.randn(plan.latent_size, generator=g)).to(plan.device) if plan.latent_size > 0 else None
    b0, bl, _ = plan.fold(lat, st)
    bl_tc = bl * S_ACT if bl is not None else None
    ref = torch.empty(n_points, device=plan.device)
    out = torch.empty(n_points, device=plan.device)
    net = plan.c_net(b0, bl, bl_tc)
    _abi.check(lib.dist_decoder_forward(net, _abi.ENGINE_SIMT, _abi.ptr(pts), n_points, None, 0.0, _abi.ptr(ref), st))
    _abi.check(lib.dist_decoder_forward(net, _abi.ENGINE_TC, _abi.ptr(pts), n_points, None, 0.0, _abi.ptr(out), st))
    err = float((out - ref).abs().max())
    if not (err < tol):
        _mark_unsafe(plan, err)


def _screen_check(plan, n_points=8192):
    """Two-tier precision of the march (dist_march_t.screen): rows whose sdf is safely beyond the clamp are evaluated with
    ONE fp16 pass.  Measures, on a sample of this decoder's own rows, how far the one-pass value is from the
    three-pass one; the scheme is enabled for the plan only if the worst deviation is below SCREEN_MARGIN / 4 (the
    kernels assume an error bound of SCREEN_MARGIN / 2).  Returns {'margin', 'err'} or None (screening off, warned)."""
    lib = _abi.lib()
    st = torch.cuda.current_stream(plan.device).cuda_stream
    g = torch.Generator().manual_seed(977)
    pts = ((torch.rand(n_points, 3, generator=g) - 0.5) * 2.0).to(plan.device)
    lat = (0.1 * torch.randn(plan.latent_size, generator=g)).to(plan.device) if plan.latent_size > 0 else None
    b0, bl, _ = plan.fold(lat, st)
    bl_tc = bl * S_ACT if bl is not None else None      # (kept alive: the descriptor only holds its address)
    net = plan.c_net(b0, bl, bl_tc)
    exact = torch.empty(n_points, device=plan.device)
    one = torch.empty(n_points, device=plan.device)
    seg = torch.zeros((n_points + 63) // 64, device=plan.device, dtype=torch.uint8)
    _abi.check(lib.dist_decoder_forward(net, _abi.ENGINE_TC, _abi.ptr(pts), n_points, None, 0.0, _abi.ptr(exact), st))
    # all rows in the one-pass segment; threshold -1: no |sdf| is <= -1, so every half-tile keeps its one-pass values
    _abi.check(lib.dist_decoder_forward_tiers(net, _abi.ptr(pts), n_points, 0, 0, -1.0, _abi.ptr(one), _abi.ptr(seg), None, st))
    err = float((one - exact).abs().max())
    ok = bool(seg.bool().all()) and err < SCREEN_MARGIN / 4
    if not ok:
        import warnings
        warnings.warn("dist-renderer_b200: two-tier precision disabled for this decoder (one-pass sdf deviates by %g, "
                      "limit %g): every march row runs at full split precision" % (err, SCREEN_MARGIN / 4))
        return None
    return {"margin": SCREEN_MARGIN, "err": err}


def _mark_unsafe(plan, err):
    import warnings
    plan.tc = None
    plan.tc_unsafe = True
    warnings.warn("dist-renderer_b200: tensor-core engine disabled for this decoder (max |tc - fp32| = %g on a "
                  "self-check sample: operand range exceeds fp16); using the fp32 engine" % err)
    raise NotImplementedError("tensor-core engine failed its self-check for this decoder")


def _build(plan, c_trunc):
    n = plan.n_layers
    blobs, meta = [], []
    stage = 0
    # forward: tensor-core layers are net layers 1 .. n-2
    for l in range(1, n - 1):
        w = plan.W[l][: plan.N[l], : plan.K[l]]
        s = _pow2_scale(w)
        b, kc, nh = _tiles(w, s, c_trunc)
        blobs.append(b)
        meta.append([kc, nh, stage, 1.0 / s])
        stage += kc * nh
    # transposed chain: gradient w.r.t. the input of net layer l (l = n-2 .. 1): B operand = W_l^T  ([K_l, N_l])
    for l in range(n - 2, 0, -1):
        w = plan.W[l][: plan.N[l], : plan.K[l]].t().contiguous()
        s = _pow2_scale(w)
        b, kc, nh = _tiles(w, s, c_trunc)
        blobs.append(b)
        meta.append([kc, nh, stage, 1.0 / s])
        stage += kc * nh
    blob = torch.cat(blobs).contiguous()
    assert blob.numel() * 2 == stage * 2 * 16384
    # meta as float tensor rows: kc, nh, stage_base, inv_scale  (read on the host side of the C ABI only)
    import ctypes
    inv = (ctypes.c_float * len(meta))(*[m[3] for m in meta])   # HOST array read by the launch code of the C ABI
    bias_s = [b * S_ACT for b in plan.bias]   # activations are carried in units of S_ACT (ReLU is positively homogeneous)
    return {"blob": blob, "inv_scale": inv, "meta": meta, "n_fwd": n - 2, "stages": stage, "c_trunc": c_trunc,
            "bias_s": bias_s}
