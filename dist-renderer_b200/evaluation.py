"""Dense-grid SDF evaluation for mesh extraction -- drop-in for the sampling half of
`core/evaluation/create_mesh.py:16-142`.

The reference fills an N^3 grid through `decode_sdf` in 32^3-row batches with a host->device and device->host copy
per batch (`create_mesh.py:35-54`); here the whole grid is generated, evaluated (fused decoder engines) and kept on
the device.  `sdf_grid_speedup` is the coarse-to-fine variant of `create_mesh_speedup` (`:110-142`): an (N/2)^3 pass
classifies voxels as far-outside / far-inside / near-surface (|sdf| <= 1.5 coarse voxels) and only the near-surface
voxels are evaluated at full resolution.  The returned (N,N,N) tensor is exactly what the reference hands to
marching cubes; `create_mesh*` call `skimage.measure.marching_cubes` when scikit-image is installed.

Grid indexing follows the reference's *intent* (DeepSDF upstream): integer division in `get_samples`
(`create_mesh.py:23-24` uses `/`, which is true division on torch >= 1.6 and yields fractional indices there --
SURVEY.md Appendix D).
"""
import torch

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


def create_mesh_speedup(decoder, latent_vec, N=256, transform=False, engine=None):
    """(verts, faces) via marching cubes on the coarse-to-fine grid; needs scikit-image (create_mesh.py:144-175)."""
    try:
        from skimage import measure
    except ImportError as e:                                   # pragma: no cover
        raise ImportError("create_mesh_speedup needs scikit-image for marching cubes; use sdf_grid_speedup() for the "
                          "SDF volume") from e
    vol, _ = sdf_grid_speedup(decoder, latent_vec, N=N, transform=transform, engine=engine)
    vs = 2.0 / (N - 1)
    verts, faces, _, _ = measure.marching_cubes(vol.cpu().numpy(), level=0.0, spacing=[vs] * 3)
    return verts - 1.0, faces
