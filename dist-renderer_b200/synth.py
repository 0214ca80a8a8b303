"""Synthetic decoders, latents and cameras.

The reference's pretrained DeepSDF checkpoints, latents and datasets are FTP downloads that are not
obtainable offline, so every test and benchmark runs on seeded synthetic inputs:

* decoder "A" -- standard DeepSDF spec (8x512, latent_in=[4], weight_norm, CodeLength 256) with PyTorch's
  default init.  SDF ~ constant -> no ray ever converges: a fixed-work plumbing / throughput fixture.
* decoder "B" -- same module with a SAL/IGR-style geometric initialisation: a latent-perturbed sphere of
  radius ~0.5, so that rays hit, converge, graze and miss like on a real shape.
* cameras -- K = [[1.2W,0,W/2],[0,1.2W,H/2],[0,0,1]], R = I, T = (0,0,1.6) (every ray meets the unit sphere),
  and an OpenCV-convention look-at ring for the multi-view layout.
All generators are deterministic functions of a seed on the CPU generator, so the build container and the
GPU box produce bit-identical weights (same image, same torch build).
"""
import math

import numpy as np
import torch

from .decoder import Decoder

STANDARD_SPEC = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)),
                     latent_in=[4], xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)


def make_decoder(kind="B", latent_size=256, width=512, depth=8, latent_in=4, seed=0, latent_scale=0.05):
    """kind 'A' = default init, 'B' = geometric (sphere) init.  Returns an eval-mode Decoder on CPU."""
    spec = dict(STANDARD_SPEC)
    spec["dims"] = [width] * depth
    spec["dropout"] = list(range(depth))
    spec["norm_layers"] = list(range(depth))
    spec["latent_in"] = [latent_in] if latent_in is not None else []
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        dec = Decoder(latent_size, **spec)
    finally:
        torch.random.set_rng_state(state)
    if kind == "A":
        return dec.eval()
    assert kind == "B"
    L = latent_size
    last = dec.num_layers - 2
    with torch.no_grad():
        for l in range(dec.num_layers - 1):
            lin = dec.layer(l)
            wn = hasattr(lin, "weight_v")
            out_f, in_f = (lin.weight_v if wn else lin.weight).shape
            if l < last:
                v = torch.randn(out_f, in_f, generator=g) * (math.sqrt(2.0) / math.sqrt(out_f))
                lin.bias.zero_()
            else:
                v = torch.randn(out_f, in_f, generator=g) * 1e-4 + math.sqrt(math.pi) / math.sqrt(in_f)
                lin.bias.fill_(-0.5)
            if l == 0:
                v[:, :L] *= latent_scale
            if l in dec.latent_in:
                h = in_f - (L + 3)
                v[:, h:h + L] *= latent_scale
            if wn:
                lin.weight_v.copy_(v)
                lin.weight_g.copy_(v.norm(2, dim=1, keepdim=True))
            else:
                lin.weight.copy_(v)
    return dec.eval()


def make_color_decoder(latent_size=256, color_size=8, seed=5):
    """Colour network of `load_decoder(color_size=...)` (decoder_utils.py:15-24): latent = shape + colour code,
    dims[3] widened by color_size, three outputs; seeded default init (any smooth rgb field serves the tests)."""
    spec = dict(STANDARD_SPEC)
    spec["dims"] = list(spec["dims"])
    spec["dims"][3] = spec["dims"][3] + color_size
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        dec = Decoder(latent_size + color_size, last_dim=3, **spec)
    finally:
        torch.random.set_rng_state(state)
    return dec.eval()


def make_latent(latent_size=256, seed=1, std=0.1):
    g = torch.Generator().manual_seed(seed)
    return std * torch.randn(1, latent_size, generator=g)


def intrinsic(H, W, focal_scale=1.2):
    return np.array([[focal_scale * W, 0.0, W / 2.0], [0.0, focal_scale * W, H / 2.0], [0.0, 0.0, 1.0]])


def front_camera(dist=1.6):
    """R = I, T = (0,0,dist): camera on the -z axis of the world frame looking down +z."""
    return torch.eye(3), torch.tensor([0.0, 0.0, dist])


def lookat_camera(azimuth_deg, elevation_deg, dist):
    """OpenCV-convention (x right, y down, z forward) world->camera extrinsic looking at the origin.

    Camera centre c = dist*(cos az cos el, sin az cos el, sin el); returns (R[3,3], T[3]) with X_cam = R X + T,
    so the origin lands at (0,0,dist) in camera coordinates.
    """
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    c = np.array([dist * math.cos(az) * math.cos(el), dist * math.sin(az) * math.cos(el), dist * math.sin(el)])
    fwd = -c / np.linalg.norm(c)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 0)
    T = -R @ c
    return torch.from_numpy(R).float(), torch.from_numpy(T).float()


def ring_cameras(n_views=24, elevation_deg=25.0, dist=2.5):
    return [lookat_camera(360.0 / n_views * i, elevation_deg, dist) for i in range(n_views)]
