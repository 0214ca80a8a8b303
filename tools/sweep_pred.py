#!/usr/bin/env python
"""Sweep of the far/near prediction thresholds of the two-tier march (dist_march_t.screen_tpred / screen_ext_margin):
forward 512x512 and 224x224 renders, time and tile-program counters (a mispredicted tile shows up as an extra three-pass
program on top of its one-pass attempt)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
dev = torch.device("cuda")
dec = synth.make_decoder("B").to(dev); lat = synth.make_latent().to(dev)


def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for hw, steps, cam in ((512, 50, synth.front_camera()), (224, 100, synth.lookat_camera(40.0, 25.0, 2.5))):
    R, T = cam[0].to(dev), cam[1].to(dev)
    K = synth.intrinsic(hw, hw) if hw == 512 else synth.intrinsic(hw, hw, 1.2 * 2.5 / 1.6)
    base = None
    for tp, ext in ((0.30, 0.04), (0.30, 0.02), (0.27, 0.03), (0.25, 0.02), (0.22, 0.02), (0.20, 0.01), (0.18, 0.01), (9.0, 9.0)):
        ren = pkg.SDFRenderer(dec, K, img_hw=(hw, hw), march_step=steps, buffer_size=5, screen_tpred=tp, screen_ext_margin=ext)
        f = lambda: ren.render(lat, R, T, ray_marching_type="recursive", no_grad=True)
        ms = timeit(f)
        ren.reset_row_counter(); out = f(); torch.cuda.synchronize()
        c = ren.tile_counters.tolist()
        if base is None: base = out
        same = all(torch.equal(a, b) for a, b in zip(out, base))
        print("%4d^2  tpred %.2f ext %.2f : %7.2f ms   one-pass tiles %6d  three-pass %6d   maps identical to the first setting: %s"
              % (hw, tp, ext, ms, c[0], c[1], same), flush=True)
