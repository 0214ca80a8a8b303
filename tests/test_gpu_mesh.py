"""GPU: csrc/mesh.cu (marching cubes, surface sampling, nearest neighbours) through the C-ABI against oracle/mesh_oracle.py,
and the device `latent_vec_to_points` / chamfer path of dist-renderer_b200/evaluation.py."""
import importlib
import os

import numpy as np
import pytest
import torch

import cases
import gpu_util
import mesh_cases
from oracle import mesh_oracle as O

pytestmark = pytest.mark.gpu
ev = importlib.import_module("dist-renderer_b200.evaluation")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chamfer.npz")


def _mc(cs):
    vol = cs["vol"]()
    v, f = ev.marching_cubes(torch.from_numpy(vol).cuda(), cs["level"], cs["spacing"], cs["origin"])
    return vol, v.cpu().numpy(), f.cpu().numpy()


@pytest.mark.parametrize("name", sorted(mesh_cases.VOLUMES))
def test_marching_cubes_bit_exact_vs_oracle(name):
    """Same vertices (bit for bit, same order) and same faces (same order, same winding) as the oracle."""
    cs = mesh_cases.VOLUMES[name]
    vol, v, f = _mc(cs)
    ov, of = O.marching_cubes(vol, cs["level"], cs["spacing"], cs["origin"])
    assert v.shape == ov.shape and f.shape == of.shape and len(f) > 0
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(f, of)


def test_marching_cubes_level_outside_range_and_flat_volume():
    vol = torch.from_numpy(mesh_cases.sphere(17)).cuda()
    with pytest.raises(ValueError, match="data range"):
        ev.marching_cubes(vol, 5.0)
    assert ev.convert_sdf_samples_to_mesh(vol + 10.0, [-1, -1, -1], 0.1) is None
    v, f = ev.marching_cubes(torch.zeros(4, 5, 6, device="cuda"), 0.0)       # level in range, nothing below it
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(ValueError):
        ev.marching_cubes(vol.cpu(), 0.0)
    with pytest.raises(cases.pkg._abi.DistError):
        ev.marching_cubes(torch.zeros(1, 5, 6, device="cuda") - 1 + torch.arange(6, device="cuda"), 0.5)


def test_marching_cubes_256_cubed_properties():
    """BASELINE-size grid (256^3, the default N of latent_vec_to_points): closed, oriented, genus 0, right volume."""
    N, r = 256, 0.6
    g = torch.linspace(-1, 1, N, device="cuda")
    vol = (g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2).sqrt() - r
    v, f = ev.marching_cubes(vol, 0.0, [2 / (N - 1)] * 3, [-1, -1, -1])
    f64 = f.long()
    e = torch.cat([f64[:, [0, 1]], f64[:, [1, 2]], f64[:, [2, 0]]], 0)
    key = e[:, 0] * (len(v) + 1) + e[:, 1]
    rkey = e[:, 1] * (len(v) + 1) + e[:, 0]
    assert key.unique().numel() == key.numel()                                  # no directed edge twice
    assert torch.equal(key.sort().values, rkey.sort().values)                   # every edge has its reverse: closed + oriented
    n_edges = key.numel() // 2
    assert len(v) - n_edges + len(f) == 2
    assert f64.unique().numel() == len(v)
    a, b, c = (v[f64[:, k]].double() for k in range(3))
    volume = float((a * torch.linalg.cross(b, c)).sum() / 6)
    assert abs(volume - 4 / 3 * np.pi * r ** 3) / (4 / 3 * np.pi * r ** 3) < 2e-4
    assert float((v.norm(dim=1) - r).abs().max()) < 2e-5


def test_surface_sample_vs_oracle_same_uniforms():
    cs = mesh_cases.VOLUMES["torus40"]
    vol, v, f = _mc(cs)
    n = 100000
    pts, fidx, u = ev.sample_surface(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), n,
                                     generator=torch.Generator(device="cuda").manual_seed(5), return_index=True)
    op, ofi = O.sample_surface(v, f, u.cpu().numpy())
    same = fidx.cpu().numpy() == ofi
    assert same.mean() > 1 - 1e-4                      # fp64 prefix sums in a different order: a pick on a boundary may flip
    assert np.array_equal(pts.cpu().numpy()[same].view(np.uint32), op[same].view(np.uint32))
    # and the points are on their triangles' surface: the torus sdf at the samples is ~0
    p = pts.double().cpu().numpy()
    q = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2) - 0.55
    assert np.abs(np.sqrt(q * q + p[:, 2] ** 2) - 0.22).max() < 5e-3              # chord sag of a 40^3 grid


def test_surface_sample_rejects_empty_mesh():
    with pytest.raises(ValueError):
        ev.sample_surface(torch.zeros(0, 3, device="cuda"), torch.zeros(0, 3, dtype=torch.int32, device="cuda"), 10)


@pytest.mark.parametrize("n_ref,n_query", [(1, 7), (1000, 1), (8193, 3001), (30000, 30000)])
def test_nearest_sqdist_vs_kdtree(n_ref, n_query):
    rng = np.random.default_rng(n_ref + n_query)
    ref = (rng.standard_normal((n_ref, 3)) * 0.4).astype(np.float32)
    qry = (rng.standard_normal((n_query, 3)) * 0.4).astype(np.float32)
    d2, idx = ev.nearest_sqdist(torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda(), return_index=True)
    od2, oi = O.nearest_sqdist(ref, qry)
    d2, idx = d2.cpu().numpy(), idx.cpu().numpy()
    assert np.allclose(d2, od2, rtol=2e-6, atol=1e-12)
    # the chosen neighbour is a nearest one (ties aside): its exact distance equals the tree's
    chosen = ((qry.astype(np.float64) - ref[idx].astype(np.float64)) ** 2).sum(1)
    assert np.allclose(chosen, od2, rtol=2e-6, atol=1e-12)


def test_chamfer_vs_reference_golden_and_oracle():
    g = np.load(GOLD)
    a, b = g["a"], g["b"]
    tol = dict(rtol=2e-6)                                  # inputs are rounded to fp32 on the device
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    assert np.isclose(ev.compute_chamfer_distance(a, b), O.compute_chamfer_distance(a32, b32), **tol)
    assert np.isclose(ev.compute_chamfer_distance(a, b), float(g["sq"]), rtol=1e-5)
    assert np.isclose(ev.compute_chamfer_distance(a, b, use_square_dist=False), float(g["lin"]), rtol=1e-5)
    assert np.allclose(ev.compute_chamfer_distance_separate(a, b), g["sep"], rtol=1e-5)
    ta, tb = torch.from_numpy(a32).cuda(), torch.from_numpy(b32).cuda()
    assert np.isclose(ev.compute_chamfer_distance(ta, tb), O.compute_chamfer_distance(a32, b32), **tol)


def test_latent_vec_to_points_end_to_end(tmp_path):
    """Decoder B (sphere-like): grid -> mesh -> samples on the device; the samples sit on the decoder's zero set, the two
    mesh creators agree, and the mesh of the same grid equals the oracle's."""
    dec = gpu_util.gpu_decoder("B")
    lat = cases.synth.make_latent().cuda()
    N = 64
    E = ev.Evaluator(dec)
    gen = torch.Generator(device="cuda").manual_seed(0)
    fname = str(tmp_path / "mesh.ply")
    pts = E.latent_vec_to_points(lat, N=N, num_points=20000, fname=fname, generator=gen, silent=True)
    assert isinstance(pts, np.ndarray) and pts.shape == (20000, 3) and pts.dtype == np.float64
    sdf = cases.pkg.decode_sdf(dec, lat, torch.from_numpy(pts).float().cuda(), no_grad=True).abs().max()
    assert float(sdf) < 0.5 * 2 / (N - 1)                                     # within half a voxel of the surface
    # .ply round trip
    raw = open(fname, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    nv = int([l for l in head.split(b"\n") if l.startswith(b"element vertex")][0].split()[-1])
    nf = int([l for l in head.split(b"\n") if l.startswith(b"element face")][0].split()[-1])
    assert len(body) == nv * 12 + nf * 13
    # original vs speedup creator: same surface
    p2 = E.latent_vec_to_points(lat, N=N, num_points=20000, meshcreator_type="original", generator=gen, as_numpy=False)
    cd = E.compute_chamfer_distance(torch.from_numpy(pts).float().cuda(), p2)
    assert cd < 2 * (0.03 ** 2)                                                # two 20 K-point samplings of one surface
    assert np.isclose(cd, O.compute_chamfer_distance(pts.astype(np.float32), p2.cpu().numpy()), rtol=1e-5)
    # the mesh of the device grid equals the oracle's mesh of that grid
    vol, _ = ev.sdf_grid_speedup(dec, lat, N=32)
    v, f = ev.marching_cubes(vol, 0.0, [2 / 31] * 3, [-1, -1, -1])
    ov, of = O.marching_cubes(vol.cpu().numpy(), 0.0, [2 / 31] * 3, [-1, -1, -1])
    assert np.array_equal(v.cpu().numpy().view(np.uint32), ov.view(np.uint32)) and np.array_equal(f.cpu().numpy(), of)
    r = O.mesh_report(ov, of)
    assert r["closed"] and r["oriented"] and r["euler"] == 2
    with pytest.raises(NotImplementedError):
        E.latent_vec_to_points(lat, N=N, meshcreator_type="fancy")


def test_latent_vec_to_points_none_without_surface():
    """Decoder A (random init) is positive everywhere: the reference returns None (create_mesh.py:163-168, transforms.py:24-25)."""
    dec = gpu_util.gpu_decoder("A")
    assert ev.latent_vec_to_points(dec, cases.synth.make_latent().cuda(), N=32, silent=True) is None
