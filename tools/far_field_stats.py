#!/usr/bin/env python
"""Measurement behind the two-tier precision march: which share of the decoder rows of a render
could be evaluated with ONE fp16 tensor-core pass instead of the three split-precision passes?

A march row whose |sdf| is safely beyond the clamp (0.1) advances by exactly ratio * clamp whatever its last bits are
(renderer.py:548-551), so only rows inside the band need fp32-level values.  This tool replays the 'recursive' march
of the BASELINE configurations in PyTorch on top of decode_sdf (CUDA engines), in compaction order, and reports per
step and in total:
  rows            decoder rows of the step (live rays)
  far             rows with |sdf| > clamp + margin                       (row-granular upper bound of the saving)
  far_tiles       rows that sit in 128-row tiles (consecutive list entries) whose rows are ALL far
  pred_far_ok     rows in tiles predicted far from the PREVIOUS step (all rows had |sdf_prev| > T_pred) that are far
  pred_far_miss   rows in tiles predicted far that contain a near row   (screen fails, tile is redone at 3 passes)
Cost model (in 3-pass row units): screened tile that passes 1/3, that fails 1/3 + 1, unscreened 1.

  python tools/far_field_stats.py [--out profiles/r2_far_field_stats.txt]
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("dist-renderer_b200")
synth = importlib.import_module("dist-renderer_b200.synth")

CLAMP, RATIO, THRESH = 0.1, 1.5, 5e-5


def march_stats(name, hw, cam, steps, margin, t_pred, engine, lines):
    gpu = torch.cuda.is_available()
    dev = torch.device("cuda" if gpu else "cpu")       # without a GPU: the module's eager layers (a preview at small sizes)
    dec = synth.make_decoder("B").to(dev)
    H, W = hw
    if cam[0] == "front":
        K, (R, T) = synth.intrinsic(H, W), synth.front_camera(cam[1])
    else:
        K, (R, T) = synth.intrinsic(H, W, focal_scale=cam[4]), synth.lookat_camera(cam[1], cam[2], cam[3])
    lat, R, T = synth.make_latent().to(dev), R.to(dev), T.to(dev)
    Kinv = torch.from_numpy(np.linalg.inv(K)).float().to(dev)
    M = torch.tensor([[1., 0., 0.], [0., 0., -1.], [0., 1., 0.]], device=dev)     # default transform_matrix (renderer.py:44-48)
    Y, X = torch.meshgrid(torch.arange(H, device=dev).float(), torch.arange(W, device=dev).float(), indexing="ij")
    homo = Kinv @ torch.stack([X.reshape(-1), Y.reshape(-1), torch.ones(H * W, device=dev)], 0)
    c = (-R.t() @ T[:, None]).squeeze(1)
    rays = R.t() @ homo
    rays = rays / (rays.norm(dim=0, keepdim=True) + 1e-12)

    def query(pts):
        if gpu:
            return pkg.decode_sdf(dec, lat, pts, clamp_dist=None, no_grad=True, engine=engine).reshape(-1)
        with torch.no_grad():
            return dec._inference_torch(torch.cat([lat.expand(pts.shape[0], -1), pts], 1)).reshape(-1)
    ptq = (c[:, None] * rays).sum(0)
    dist = (c[:, None] - ptq[None] * rays).norm(dim=0)
    hit = dist <= 1.0
    chord = 2 * torch.sqrt(torch.clamp(1 - dist * dist, min=0))
    cd = c.norm()
    entry = torch.sqrt(torch.clamp(cd * cd - dist * dist, min=0)) - chord / 2 if float(cd) >= 1.0 else torch.zeros_like(dist)
    ex = entry + chord
    idx = torch.nonzero(hit & (entry < ex)).reshape(-1)          # live list in pixel order (= compaction order)
    z = torch.zeros(H * W, device=dev)
    prev = torch.full((H * W,), 0.0, device=dev)                  # |sdf| of the previous step (0: no prediction at step 0)
    tot = dict(rows=0, far=0, far_tiles=0, ok=0, miss=0, pred_near=0)
    lines.append("== %s: %dx%d, %d steps, margin %.0e, T_pred %.2f, engine %s" % (name, H, W, steps, margin, t_pred, engine))
    lines.append("%4s %9s %9s %9s %9s %9s" % ("step", "rows", "far", "far_tiles", "pred_ok", "pred_miss"))
    for s in range(steps):
        n = idx.numel()
        if n == 0:
            break
        pts = M.t() @ (c[:, None] + rays[:, idx] * (entry[idx] + z[idx])[None])
        sdf = query(pts.t().contiguous())
        far = sdf.abs() > CLAMP + margin
        pad = (-n) % 128
        tile_far = torch.cat([far, far.new_ones(pad)]).reshape(-1, 128).all(1)
        pred = torch.cat([prev[idx] > t_pred, far.new_ones(pad)]).reshape(-1, 128).all(1) if s > 0 else torch.ones_like(tile_far)
        rows_t = torch.cat([far.new_ones(n), far.new_zeros(pad)]).reshape(-1, 128).sum(1)
        st = dict(rows=n, far=int(far.sum()), far_tiles=int(rows_t[tile_far].sum()), ok=int(rows_t[pred & tile_far].sum()),
                  miss=int(rows_t[pred & ~tile_far].sum()), pred_near=int(rows_t[~pred].sum()))
        for k in tot:
            tot[k] += st[k]
        if s < 16 or s % 10 == 0:
            lines.append("%4d %9d %9d %9d %9d %9d" % (s, st["rows"], st["far"], st["far_tiles"], st["ok"], st["miss"]))
        z[idx] = z[idx] + torch.clamp(sdf, -CLAMP, CLAMP) * RATIO
        prev[idx] = sdf.abs()
        keep = (z[idx] + entry[idx] < ex[idx]) & (sdf.abs() >= THRESH)
        idx = idx[keep]
    r = float(tot["rows"])
    cost_row = (tot["far"] / 3 + (tot["rows"] - tot["far"]) * (1 + 1 / 3)) / r           # verdict's scheme: screen all, redo near rows
    cost_tile = (tot["far_tiles"] / 3 + (tot["rows"] - tot["far_tiles"]) * (1 + 1 / 3)) / r
    cost_pred = (tot["ok"] / 3 + tot["miss"] * (1 + 1 / 3) + tot["pred_near"]) / r
    lines.append("total rows %d (%.1f per ray): far %.1f %%, in all-far tiles %.1f %%, predicted+far %.1f %%, predicted but near %.1f %%"
                 % (tot["rows"], r / (H * W), 100 * tot["far"] / r, 100 * tot["far_tiles"] / r, 100 * tot["ok"] / r, 100 * tot["miss"] / r))
    lines.append("march MMA cost vs 3 passes everywhere: row-granular screen+redo %.3f, tile-granular screen+redo %.3f, "
                 "tile-granular with prediction %.3f" % (cost_row, cost_tile, cost_pred))
    lines.append("")
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--margin", type=float, default=2e-3)
    ap.add_argument("--scale", type=float, default=1.0, help="image side multiplier (CPU previews)")
    args = ap.parse_args()
    sz = lambda n: int(round(n * args.scale))
    lines = []
    ring = ("lookat", 40.0, 25.0, 2.5, 1.2 * 2.5 / 1.6)
    for t_pred in (0.25, 0.3):
        march_stats("config 2 / bench (512x512 front)", (sz(512), sz(512)), ("front", 1.6), 50, args.margin, t_pred, args.engine, lines)
        march_stats("config 3 (224x224 look-at)", (sz(224), sz(224)), ring, 100, args.margin, t_pred, args.engine, lines)
    march_stats("config 4 (256x256 ring view)", (sz(256), sz(256)), ("lookat", 45.0, 25.0, 2.5, 1.2 * 2.5 / 1.6), 100, args.margin, 0.25,
                args.engine, lines)
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
