"""Evaluation path `latent_vec_to_points` + chamfer distance on the device -- drop-in for `core/evaluation/`
(`create_mesh.py:16-175`, `transforms.py:8-32`, `eval_func.py:5-39`, `evaluator.py:8-21`).

The reference fills an N^3 grid through `decode_sdf` in 32^3-row batches with a host->device and device->host copy
per batch (`create_mesh.py:35-54`), runs scikit-image's marching cubes on a host copy, writes a .ply, reloads it with
trimesh to draw surface samples, and builds two scipy KD-trees per chamfer evaluation.  Here the grid is generated and
evaluated on the fused decoder engines, meshed (`marching_cubes`), sampled (`sample_surface`) and compared
(`compute_chamfer_distance`) by the kernels of `csrc/mesh.cu` without leaving the device; a .ply is written only when
the caller names one.  `sdf_grid_speedup` is the coarse-to-fine variant of `create_mesh_speedup` (`:110-142`): an (N/2)^3
pass classifies voxels as far-outside / far-inside / near-surface (|sdf| <= 1.5 coarse voxels) and only the near-surface
voxels are evaluated at full resolution.

Grid indexing follows the reference's *intent* (DeepSDF upstream): integer division in `get_samples`
(`create_mesh.py:23-24` uses `/`, which is true division on torch >= 1.6 and yields fractional indices there --
SURVEY.md Appendix D).  Marching-cubes conventions (scikit-image is absent here; see `mc_tables.py`): inside = sdf < level,
ambiguous faces cut off inside corners, normals point outwards.  There is no CPU path: every function needs the CUDA
library and CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from . import _abi
from .functional import decode_sdf


def get_samples(N, voxel_origin, voxel_size, transform=False, device=None):
    """(N^3, 3) grid coordinates, x slowest / z fastest (create_mesh.py:16-33)."""
    idx = torch.arange(0, N ** 3, device=device)
    ijk = torch.stack([(idx // N) // N % N, (idx // N) % N, idx % N], 1).float()
    pts = ijk * voxel_size + torch.tensor(voxel_origin, device=device, dtype=torch.float32)
    if transform:                                             # create_mesh.py:10-14
        pts = torch.stack([pts[:, 0], pts[:, 2], -pts[:, 1]], 1)
    return pts


def infer_samples(decoder, latent_vec, points, max_batch=2 ** 22, engine=None):
    """sdf (clamped to +-0.1 like decode_sdf's default) of `points` (M,3), device-resident (create_mesh.py:35-54)."""
    out = torch.empty(points.shape[0], device=points.device)
    for s in range(0, points.shape[0], max_batch):
        out[s:s + max_batch] = decode_sdf(decoder, latent_vec, points[s:s + max_batch], no_grad=True,
                                          engine=engine).squeeze(1)
    return out


def sdf_grid(decoder, latent_vec, N=256, transform=False, engine=None):
    """Full-resolution grid (create_mesh.py:56-79 up to the marching-cubes call)."""
    dev = next(decoder.parameters()).device
    pts = get_samples(N, [-1.0, -1.0, -1.0], 2.0 / (N - 1), transform=transform, device=dev)
    return infer_samples(decoder, latent_vec, pts, engine=engine).reshape(N, N, N)


def sdf_grid_speedup(decoder, latent_vec, N=256, transform=False, engine=None):
    """Coarse-to-fine grid (create_mesh.py:110-133): returns (sdf[N,N,N], number of voxels evaluated at full res)."""
    dev = next(decoder.parameters()).device
    Nh = int(N / 2)
    origin, vs, vs_half = [-1.0, -1.0, -1.0], 2.0 / (N - 1), 2.0 / (N / 2 - 1)
    half = infer_samples(decoder, latent_vec, get_samples(Nh, origin, vs_half, transform, dev), engine=engine)
    up = half.reshape(Nh, Nh, Nh).repeat_interleave(2, 0).repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1)
    relax = 1.5                                               # create_mesh.py:101-108
    near = up.abs() <= vs_half * relax
    sdf = torch.where(up > vs_half * relax, torch.full_like(up, 0.1), torch.full_like(up, -0.1))
    pts = get_samples(N, origin, vs, transform, dev)
    idx = torch.nonzero(near).reshape(-1)
    sdf[idx] = infer_samples(decoder, latent_vec, pts[idx], engine=engine)
    return sdf.reshape(N, N, N), int(idx.numel())


def _cuda(t, name, dtype):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise ValueError("%s must be a CUDA tensor (the mesh kernels have no CPU path)" % name)
    return t.detach().to(dtype).contiguous()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def marching_cubes(vol, level=0.0, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
    """(verts[nv,3] f32, faces[nt,3] i32) of the `level` iso-surface of vol[n0,n1,n2] -- both on vol's device.
    verts = origin + spacing * (fractional grid index), like `marching_cubes_lewiner(vol, level, spacing)` plus the
    origin shift of `create_mesh.py:173-176`.  Raises ValueError when `level` is outside [vol.min(), vol.max()], as
    scikit-image does (the reference turns that into `return False`, `create_mesh.py:163-168`)."""
    vol = _cuda(vol, "vol", torch.float32)
    if vol.dim() != 3:
        raise ValueError("vol must be (n0, n1, n2)")
    lo, hi = torch.aminmax(vol)
    if not (float(lo) <= level <= float(hi)):
        raise ValueError("Surface level must be within volume data range.")
    lib, dev = _abi.lib(), vol.device
    n0, n1, n2 = vol.shape
    M = vol.numel()
    with torch.cuda.device(dev):
        st = _stream(dev)
        scan = torch.empty(M, dtype=torch.int64, device=dev)
        mask = torch.empty(M, dtype=torch.uint8, device=dev)
        scratch = torch.empty(lib.dist_scan_scratch_elems(M), dtype=torch.int64, device=dev)
        totals = torch.empty(1, dtype=torch.int64, device=dev)
        _abi.check(lib.dist_mc_count(_abi.ptr(vol), n0, n1, n2, float(level), _abi.ptr(scan), _abi.ptr(mask), _abi.ptr(scratch),
                                     _abi.ptr(totals), st))
        packed = int(totals.item())                       # the one host sync: the output size is data dependent
        nv, nt = packed & 0xffffffff, packed >> 32
        if nv >= 2 ** 31 or nt >= 2 ** 31:
            raise ValueError("marching cubes: %d vertices / %d triangles exceed the int32 index range" % (nv, nt))
        verts = torch.empty(nv, 3, dtype=torch.float32, device=dev)
        faces = torch.empty(nt, 3, dtype=torch.int32, device=dev)
        if nv:
            o3 = (ctypes.c_float * 3)(*[float(x) for x in origin])
            s3 = (ctypes.c_float * 3)(*[float(x) for x in spacing])
            _abi.check(lib.dist_mc_emit(_abi.ptr(vol), n0, n1, n2, float(level), o3, s3, _abi.ptr(scan), _abi.ptr(mask),
                                        _abi.ptr(verts), _abi.ptr(faces), st))
    return verts, faces


def sample_surface(verts, faces, count, generator=None, return_index=False):
    """`count` points drawn uniformly over the mesh surface (the scheme of `trimesh.sample.sample_surface`,
    `transforms.py:8-11`): faces by area through a cumulative sum, a uniform point inside each.  Device tensors in and out."""
    verts = _cuda(verts, "verts", torch.float32)
    faces = _cuda(faces, "faces", torch.int32)
    nt = faces.shape[0]
    if nt == 0:
        raise ValueError("sample_surface: the mesh has no faces")
    lib, dev = _abi.lib(), verts.device
    with torch.cuda.device(dev):
        st = _stream(dev)
        cum = torch.empty(nt, dtype=torch.float64, device=dev)
        scratch = torch.empty(lib.dist_scan_scratch_elems(nt), dtype=torch.float64, device=dev)
        total = torch.empty(1, dtype=torch.float64, device=dev)
        _abi.check(lib.dist_tri_area_scan(_abi.ptr(verts), _abi.ptr(faces), nt, _abi.ptr(cum), _abi.ptr(scratch), _abi.ptr(total), st))
        u = torch.rand(count, 3, device=dev, generator=generator)
        pts = torch.empty(count, 3, dtype=torch.float32, device=dev)
        fidx = torch.empty(count, dtype=torch.int32, device=dev) if return_index else None
        _abi.check(lib.dist_surface_sample(_abi.ptr(verts), _abi.ptr(faces), nt, _abi.ptr(cum), _abi.ptr(total), _abi.ptr(u), count,
                                           _abi.ptr(pts), _abi.ptr(fidx), st))
    return (pts, fidx, u) if return_index else pts


def nearest_sqdist(ref, query, return_index=False):
    """Squared distance from every query point to its nearest point of `ref` (`KDTree(ref).query(query)`,
    `eval_func.py:10-11`); (M,) f32 on the device [+ indices]."""
    ref = _cuda(ref, "ref", torch.float32)
    query = _cuda(query, "query", torch.float32)
    if ref.dim() != 2 or ref.shape[1] != 3 or query.dim() != 2 or query.shape[1] != 3:
        raise ValueError("point sets must be (n, 3)")
    if ref.shape[0] == 0:
        raise ValueError("nearest_sqdist: empty reference set")
    lib, dev = _abi.lib(), ref.device
    n = query.shape[0]
    with torch.cuda.device(dev):
        best = torch.empty(n, dtype=torch.int64, device=dev)
        d2 = torch.empty(n, dtype=torch.float32, device=dev)
        idx = torch.empty(n, dtype=torch.int32, device=dev) if return_index else None
        _abi.check(lib.dist_nearest_sqdist(_abi.ptr(ref), ref.shape[0], _abi.ptr(query), n, _abi.ptr(best), _abi.ptr(d2),
                                           _abi.ptr(idx), _stream(dev)))
    return (d2, idx) if return_index else d2


def _points(p, dev):
    if isinstance(p, np.ndarray):
        p = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32))
    return p.to(device=dev, dtype=torch.float32)


def compute_chamfer_distance_separate(points_1, points_2, device=None):
    """(mean squared distance points_2 -> points_1, points_1 -> points_2), `eval_func.py:26-39`.  numpy arrays (as the
    reference passes them) are copied to `device` (default: the first tensor's device, else cuda:0)."""
    dev = device or next((p.device for p in (points_1, points_2) if torch.is_tensor(p) and p.is_cuda), torch.device("cuda", 0))
    p1, p2 = _points(points_1, dev), _points(points_2, dev)
    d21 = nearest_sqdist(p1, p2).double().mean()
    d12 = nearest_sqdist(p2, p1).double().mean()
    return float(d21), float(d12)


def compute_chamfer_distance(points_1, points_2, use_square_dist=True, device=None):
    """Symmetric chamfer distance, `eval_func.py:5-24`."""
    dev = device or next((p.device for p in (points_1, points_2) if torch.is_tensor(p) and p.is_cuda), torch.device("cuda", 0))
    p1, p2 = _points(points_1, dev), _points(points_2, dev)
    d21, d12 = nearest_sqdist(p1, p2).double(), nearest_sqdist(p2, p1).double()
    if not use_square_dist:
        d21, d12 = d21.sqrt(), d12.sqrt()
    return float(d21.mean() + d12.mean())


def write_ply(verts, faces, filename):
    """Binary little-endian .ply with float x/y/z vertices and int32 `vertex_indices` faces -- the elements
    `convert_sdf_samples_to_ply` describes (`create_mesh.py:180-198`)."""
    v = np.ascontiguousarray(verts.detach().cpu().numpy(), dtype="<f4")
    f = np.ascontiguousarray(faces.detach().cpu().numpy(), dtype="<i4")
    rec = np.empty(len(f), dtype=[("n", "u1"), ("idx", "<i4", (3,))])
    rec["n"], rec["idx"] = 3, f
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
              "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(v), len(f)))
    with open(filename, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(v.tobytes())
        fh.write(rec.tobytes())


def convert_sdf_samples_to_mesh(sdf, voxel_grid_origin, voxel_size):
    """(verts, faces) in the reference's mesh frame (`create_mesh.py:144-176`), or None when the grid has no zero crossing
    (the reference's `return False`)."""
    try:
        return marching_cubes(sdf, 0.0, [voxel_size] * 3, voxel_grid_origin)
    except ValueError as e:
        if "data range" not in str(e):
            raise
        return None


def create_mesh(decoder, latent_vec, filename=None, N=256, max_batch=2 ** 22, silent=True, transform=False, engine=None):
    """`create_mesh.py:56-79`: full-resolution grid -> mesh.  Returns (verts, faces) on the device or None (no surface);
    writes `filename + '.ply'` when a filename is given."""
    mesh = convert_sdf_samples_to_mesh(sdf_grid(decoder, latent_vec, N, transform, engine), [-1.0, -1.0, -1.0], 2.0 / (N - 1))
    if mesh is not None and filename is not None:
        write_ply(mesh[0], mesh[1], filename + ".ply")
    return mesh


def create_mesh_speedup(decoder, latent_vec, filename=None, N=256, max_batch=2 ** 22, silent=True, transform=False, engine=None):
    """`create_mesh.py:110-142`: coarse-to-fine grid -> mesh; same return convention as `create_mesh`."""
    vol, _ = sdf_grid_speedup(decoder, latent_vec, N=N, transform=transform, engine=engine)
    mesh = convert_sdf_samples_to_mesh(vol, [-1.0, -1.0, -1.0], 2.0 / (N - 1))
    if mesh is not None and filename is not None:
        write_ply(mesh[0], mesh[1], filename + ".ply")
    return mesh


def latent_vec_to_points(decoder, latent_vec, N=256, max_batch=2 ** 22, num_points=30000, silent=False, fname=None,
                         transform=False, meshcreator_type="speedup", generator=None, as_numpy=True, engine=None):
    """`transforms.py:13-32`: decode the shape code to `num_points` surface samples; None when the grid has no surface.
    Returns a float64 numpy array like trimesh's (as_numpy=False keeps the (num_points, 3) f32 tensor on the device, which
    `compute_chamfer_distance` takes as it is).  `fname`, if given, receives the mesh as a .ply."""
    if meshcreator_type == "original":
        mesh = create_mesh(decoder, latent_vec, None, N=N, transform=transform, engine=engine)
    elif meshcreator_type == "speedup":
        mesh = create_mesh_speedup(decoder, latent_vec, None, N=N, transform=transform, engine=engine)
    else:
        raise NotImplementedError
    if mesh is None:
        return None
    verts, faces = mesh
    if fname is not None:
        write_ply(verts, faces, fname)
    pts = sample_surface(verts, faces, num_points, generator=generator)
    return pts.double().cpu().numpy() if as_numpy else pts


class Evaluator(object):
    """`core/evaluation/evaluator.py:8-21`."""

    def __init__(self, decoder):
        self.decoder = decoder
        self.device = next(self.decoder.parameters()).device
        self.decoder.eval()

    def latent_vec_to_points(self, latent_vec, N=256, max_batch=2 ** 22, num_points=30000, silent=False, fname=None,
                             transform=False, meshcreator_type="speedup", **kw):
        return latent_vec_to_points(self.decoder, latent_vec=latent_vec, N=N, max_batch=max_batch, num_points=num_points,
                                    silent=silent, fname=fname, transform=transform, meshcreator_type=meshcreator_type, **kw)

    def compute_chamfer_distance(self, points_1, points_2, separate=False):
        if not separate:
            return compute_chamfer_distance(points_1, points_2, device=self.device)
        return compute_chamfer_distance_separate(points_1, points_2, device=self.device)
