#!/usr/bin/env python
"""Per-layer timeline of one tile of mlp_tc_kernel<0> in both precision tiers.  Needs a library built with
DIST_EXTRA_NVCC_FLAGS=-DDIST_TC_TIMELINE and DIST_TC_DEBUG=4 in the environment (the launch code prints to stderr)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
synth = importlib.import_module("dist-renderer_b200.synth")
abi = importlib.import_module("dist-renderer_b200._abi"); tc = importlib.import_module("dist-renderer_b200.tc")
plan_mod = importlib.import_module("dist-renderer_b200.plan")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev); lat = synth.make_latent().to(dev)
plan = plan_mod.plan_for(dec); tc.prepare(plan)
lib, st = abi.lib(), torch.cuda.current_stream().cuda_stream
b0, bl, _ = plan.fold(lat, st); bl_tc = bl * tc.S_ACT
net = plan.c_net(b0, bl, bl_tc)
n = 9472 * 3
g = torch.Generator().manual_seed(3)
d = torch.randn(n, 3, generator=g); pts = (d / d.norm(dim=1, keepdim=True) * 0.95).to(dev)
sdf = torch.empty(n, device=dev); tiles = (n + 127) // 128
mode0 = torch.zeros(tiles, device=dev, dtype=torch.uint8); seg = torch.zeros(2 * tiles, device=dev, dtype=torch.uint8)
for rep in range(2):
    sys.stderr.write("==== three passes (rep %d)\n" % rep); sys.stderr.flush()
    abi.check(lib.dist_decoder_forward(net, abi.ENGINE_TC, abi.ptr(pts), n, None, 0.0, abi.ptr(sdf), st))
    torch.cuda.synchronize()
    sys.stderr.write("==== one pass, all half-tiles pass (rep %d)\n" % rep); sys.stderr.flush()
    abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n, 0, 0, 0.102, abi.ptr(sdf), abi.ptr(seg), None, st))
    torch.cuda.synchronize()
