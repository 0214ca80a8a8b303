#!/usr/bin/env python
"""Decoder-kernel microbenchmark of the two precision tiers (csrc/mlp_tc.cu): one launch over 262,144 rows that are all
far from the surface, evaluated (a) with the three split-precision passes, (b) with the one-pass attempt that every
half-tile passes, (c) with the one-pass attempt that every tile FAILS (threshold 2: cost of a failed screen + redo)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
abi = importlib.import_module("dist-renderer_b200._abi"); tc = importlib.import_module("dist-renderer_b200.tc")
plan_mod = importlib.import_module("dist-renderer_b200.plan")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev); lat = synth.make_latent().to(dev)
plan = plan_mod.plan_for(dec); tc.prepare(plan)
lib, st = abi.lib(), torch.cuda.current_stream().cuda_stream
b0, bl, _ = plan.fold(lat, st); bl_tc = bl * tc.S_ACT
net = plan.c_net(b0, bl, bl_tc)
F = 2 * sum(k * n for k, n in zip(plan.K, plan.N))


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for n in (262144, 9472, 1024):
    g = torch.Generator().manual_seed(3)
    d = torch.randn(n, 3, generator=g); pts = (d / d.norm(dim=1, keepdim=True) * 0.95).to(dev)   # on a sphere of radius 0.95: sdf ~ 0.45
    sdf = torch.empty(n, device=dev); tiles = (n + 127) // 128
    mode0 = torch.zeros(tiles, device=dev, dtype=torch.uint8); seg = torch.zeros(2 * tiles, device=dev, dtype=torch.uint8)
    cnt = torch.zeros(2, device=dev, dtype=torch.int64)
    t3 = timeit(lambda: abi.check(lib.dist_decoder_forward(net, abi.ENGINE_TC, abi.ptr(pts), n, None, 0.0, abi.ptr(sdf), st)))
    t1 = timeit(lambda: abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n, 0, 0, 0.102, abi.ptr(sdf), abi.ptr(seg), abi.ptr(cnt), st)))
    ok = bool(seg[: (n + 63) // 64].bool().all())
    tf = timeit(lambda: abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n, 0, 0, 2.0, abi.ptr(sdf), abi.ptr(seg), abi.ptr(cnt), st)))
    print("%7d rows: three passes %.3f ms (%.0f useful TFLOP/s) | one pass, all pass %.3f ms (x%.2f, all flagged %s) | one pass, all fail + redo %.3f ms (x%.2f)"
          % (n, t3, n * F / t3 / 1e9, t1, t3 / t1, ok, tf, tf / t3), flush=True)
