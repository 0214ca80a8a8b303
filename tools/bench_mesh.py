"""Stage timings of the device `latent_vec_to_points` + chamfer path at the reference's defaults
(N = 256, 30 000 samples), CUDA events, with the achieved HBM rate of the marching-cubes passes and scipy's KD-tree chamfer
(what the reference runs, core/evaluation/eval_func.py) timed on the host beside the brute-force kernel.
    python tools/bench_mesh.py [N]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("dist-renderer_b200")
ev, synth = pkg.evaluation, importlib.import_module("dist-renderer_b200.synth")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dec = synth.make_decoder("B").cuda()
    lat = synth.make_latent().cuda()
    t_grid, (vol, n_fine) = timed(lambda: ev.sdf_grid_speedup(dec, lat, N=N), 3)
    t_full, _ = timed(lambda: ev.sdf_grid(dec, lat, N=N), 2)
    vs = 2.0 / (N - 1)
    t_mc, (v, f) = timed(lambda: ev.marching_cubes(vol, 0.0, [vs] * 3, [-1.0] * 3))
    t_s, pts = timed(lambda: ev.sample_surface(v, f, 30000))
    pts2 = ev.sample_surface(v, f, 30000) + 0.01
    t_c, cd = timed(lambda: ev.compute_chamfer_distance(pts, pts2))
    M = N ** 3
    # marching cubes algorithmic traffic: count pass reads the grid (4 B) and writes scan (8 B) + mask (1 B); the scan reads and
    # writes 8 B twice; the emit pass reads grid + scan + mask (13 B) -- 58 B per grid point, outputs negligible
    print("N = %d: %d fine voxels of %d (%.1f %%), %d vertices, %d triangles" % (N, n_fine, M, 100.0 * n_fine / M, len(v), len(f)))
    print("sdf grid, coarse-to-fine   %8.2f ms   (%.1f M decoder rows/s)" % (t_grid, ((N // 2) ** 3 + n_fine) / t_grid / 1e3))
    print("sdf grid, full resolution  %8.2f ms   (%.1f M decoder rows/s)" % (t_full, M / t_full / 1e3))
    print("marching cubes             %8.2f ms   (%.0f GB/s over 58 B/grid point, one 8-byte host sync included)" % (t_mc, 58.0 * M / t_mc / 1e6))
    print("surface sampling (30 K)    %8.2f ms" % t_s)
    print("chamfer 30 K x 30 K        %8.2f ms   (%.2f G distance evaluations/s)" % (t_c, 2 * 30000.0 * 30000 / t_c / 1e6))
    a, b = pts.double().cpu().numpy(), pts2.double().cpu().numpy()
    from scipy.spatial import cKDTree
    t0 = time.perf_counter()
    d1 = cKDTree(a).query(b)[0]
    d2 = cKDTree(b).query(a)[0]
    ref = float(np.mean(d1 ** 2) + np.mean(d2 ** 2))
    t_cpu = (time.perf_counter() - t0) * 1e3
    print("chamfer, scipy cKDTree     %8.2f ms   (host, 1 thread; value %.6e vs %.6e)" % (t_cpu, ref, cd))


if __name__ == "__main__":
    main()
