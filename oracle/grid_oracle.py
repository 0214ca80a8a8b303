"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the sampling half of core/evaluation/create_mesh.py
(get_samples :16-33 with integer division, infer_samples :35-54, check_valid :101-108, the grid of
create_mesh_speedup :110-133), pinned against the reference's own functions in tests/test_oracle.py."""
import numpy as np
import torch

from .sdf_oracle import decode_sdf


def get_samples(N, origin, vs, transform=False):
    idx = torch.arange(0, N ** 3)
    s = torch.zeros(N ** 3, 3)
    s[:, 2] = idx % N
    s[:, 1] = (idx // N) % N
    s[:, 0] = ((idx // N) // N) % N
    for k in range(3):
        s[:, k] = s[:, k] * vs + origin[k]
    if transform:
        s = torch.stack([s[:, 0], s[:, 2], -s[:, 1]], 1)
    return s


def infer(decoder, latent, pts):
    return decode_sdf(decoder, latent, pts).squeeze(1).detach()


def grid_speedup(decoder, latent, N, transform=False):
    Nh = int(N / 2)
    vs, vsh = 2.0 / (N - 1), 2.0 / (N / 2 - 1)
    half = infer(decoder, latent, get_samples(Nh, [-1, -1, -1], vsh, transform))
    up = half.reshape(Nh, Nh, Nh).repeat_interleave(2, 0).repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1)
    pos, neg, near = up > vsh * 1.5, -up > vsh * 1.5, up.abs() <= vsh * 1.5
    pts = get_samples(N, [-1, -1, -1], vs, transform)
    out = torch.zeros(N ** 3)
    out[pos], out[neg] = 0.1, -0.1
    out[near] = infer(decoder, latent, pts[near])
    return out.reshape(N, N, N), int(near.sum())
