"""Host-side operand preparation of the tensor-core engine (dist-renderer_b200/tc.py), checked on CPU: the stage blob
must be the exact shared-memory image csrc/mlp_tc.cu expects (UMMA no-swizzle K-major panels, N-half outer / K chunk
inner stage order, [hi 8 KB][lo 8 KB] per CTA), and hi + lo must carry the fp32 weight to ~2^-22."""
import importlib
import math

import torch

tc = importlib.import_module("dist-renderer_b200.tc")


def _decode(blob, N, K):
    """Independent reading of the blob, following the kernel's addressing: stage s = h * kc + c (N-half h, K chunk c of
    32); CTA r of the pair holds rows 256 h + 128 r + row; inside a CTA's 16 KB: hi then lo, each [4 K-groups][128 rows][8]."""
    Kp, NH = ((K + 63) // 64) * 64, (N + 255) // 256
    kc = Kp // 32
    t = blob.reshape(NH, kc, 2, 2, 4, 128, 8).float()          # [h][c][r][hi/lo][g][row][e]
    hi = torch.zeros(NH * 256, Kp)
    lo = torch.zeros(NH * 256, Kp)
    for h in range(NH):
        for c in range(kc):
            for r in range(2):
                for g in range(4):
                    rows = slice(256 * h + 128 * r, 256 * h + 128 * r + 128)
                    cols = slice(32 * c + 8 * g, 32 * c + 8 * g + 8)
                    hi[rows, cols] = t[h, c, r, 0, g]
                    lo[rows, cols] = t[h, c, r, 1, g]
    return hi, lo, kc, NH


def test_pow2_scale_puts_the_largest_weight_below_fp16_max():
    g = torch.Generator().manual_seed(0)
    for scale in (1e-3, 0.07, 1.0, 37.0):
        w = torch.randn(64, 48, generator=g) * scale
        s = tc._pow2_scale(w)
        assert math.log2(s) == round(math.log2(s))
        m = float(w.abs().max()) * s
        assert 2.0 ** 12 <= m < 2.0 ** 14 and m < 65504.0
    assert tc._pow2_scale(torch.zeros(4, 4)) == 1.0


def test_stage_blob_is_the_kernel_smem_image():
    g = torch.Generator().manual_seed(1)
    for (N, K) in [(512, 512), (253, 512), (512, 256), (512, 259), (300, 70)]:
        w = torch.randn(N, K, generator=g) * (math.sqrt(2.0) / math.sqrt(N))
        s = tc._pow2_scale(w)
        blob, kc, nh = tc._tiles(w, s, 0.0)
        assert blob.dtype == torch.float16 and blob.numel() * 2 == kc * nh * 2 * 16384
        hi, lo, kc2, nh2 = _decode(blob, N, K)
        assert (kc, nh) == (kc2, nh2) == (((K + 63) // 64) * 2, (N + 255) // 256)
        ws = w.double() * s
        rec = hi.double() + lo.double()
        assert float((rec[:N, :K] - ws).abs().max()) <= float(ws.abs().max()) * 2.0 ** -21
        assert float(rec[N:].abs().max() if rec[N:].numel() else 0.0) == 0.0 and float(rec[:, K:].abs().max() if rec[:, K:].numel() else 0.0) == 0.0
        assert torch.equal(hi[:N, :K], (w * s).half().float())                  # hi is the fp16 rounding of the scaled weight


def test_truncation_compensation_scales_k_blocks():
    """Weights of K-block j (16 columns) are multiplied by 1 + 3 c (J - j): the first block sees all 3 J accumulation
    steps of the layer, the last block only its own three."""
    w = torch.ones(256, 128)
    c = 1e-3                                   # exaggerated so that fp16 resolves it
    blob, kc, nh = tc._tiles(w, 1024.0, c)
    hi, lo, _, _ = _decode(blob, 256, 128)
    rec = (hi.double() + lo.double()) / 1024.0
    J = 128 // 16
    for j in range(J):
        expect = 1.0 + 3.0 * c * (J - j)
        assert float((rec[:, 16 * j:16 * j + 16] - expect).abs().max()) < 1e-6


def test_three_pass_split_product_matches_fp32_gemm():
    """A_hi W_hi + A_lo W_hi + A_hi W_lo with exact accumulation reproduces A W^T to ~2^-21 (the dropped lo*lo term)."""
    g = torch.Generator().manual_seed(2)
    N, K = 512, 512
    w = torch.randn(N, K, generator=g) * (math.sqrt(2.0) / math.sqrt(N))
    a = torch.relu(torch.randn(64, K, generator=g)) * tc.S_ACT
    s = tc._pow2_scale(w)
    blob, _, _ = tc._tiles(w, s, 0.0)
    w_hi, w_lo, _, _ = _decode(blob, N, K)
    a_hi = a.half().float()
    a_lo = (a - a_hi).half().float()
    acc = a_hi.double() @ w_hi.double().t() + a_lo.double() @ w_hi.double().t() + a_hi.double() @ w_lo.double().t()
    ref = (a.double() / tc.S_ACT) @ w.double().t()
    got = acc / (tc.S_ACT * s)
    assert float((got - ref).abs().max()) < 2e-6 * float(ref.abs().max())
