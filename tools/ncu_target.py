#!/usr/bin/env python
"""Target for `ncu --profile-from-start off --set full`: after engine preparation, one launch each of
 (1) mlp_tc_kernel<0>, 262,144 rows, three split-precision passes         (the round-1 comparison point)
 (2) mlp_tc_kernel<0>, 262,144 rows in the one-pass segment, pair mode    (all half-tiles pass)
 (3) mlp_tc_kernel<0>, a dense march step's mix: 60 % one-pass rows + 40 % full-precision rows
 (4) mlp_tc_kernel<1>, 65,536 rows forward + input gradient (normals)
inside a cudaProfilerStart/Stop window."""
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
n = 262144
g = torch.Generator().manual_seed(3)
d = torch.randn(2 * n, 3, generator=g); pts = (d / d.norm(dim=1, keepdim=True) * 0.95).to(dev)
sdf = torch.empty(2 * n, device=dev); seg = torch.zeros(2 * n // 64, device=dev, dtype=torch.uint8)
gp = ((torch.rand(65536, 3, generator=g) - 0.5) * 1.2).to(dev)
for _ in range(2):   # warm-up outside the window
    abi.check(lib.dist_decoder_forward(net, abi.ENGINE_TC, abi.ptr(pts), n, None, 0.0, abi.ptr(sdf), st))
    abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n, 0, 0, 0.102, abi.ptr(sdf), abi.ptr(seg), None, st))
    pkg.decode_sdf_gradient(dec, lat, gp, engine="tc")
torch.cuda.synchronize()
torch.cuda.profiler.start()
abi.check(lib.dist_decoder_forward(net, abi.ENGINE_TC, abi.ptr(pts), n, None, 0.0, abi.ptr(sdf), st))
abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n, 0, 0, 0.102, abi.ptr(sdf), abi.ptr(seg), None, st))
n1 = int(0.6 * n) // 128 * 128
abi.check(lib.dist_decoder_forward_tiers(net, abi.ptr(pts), n1, n - n1, n, 0.102, abi.ptr(sdf), abi.ptr(seg), None, st))
pkg.decode_sdf_gradient(dec, lat, gp, engine="tc")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
