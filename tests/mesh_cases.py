"""Volumes shared by the CPU (oracle) and GPU (kernel vs oracle) mesh tests."""
import numpy as np


def sphere(n=33, r=0.6, shape=None):
    shape = shape or (n, n, n)
    ax = [np.linspace(-1, 1, k, dtype=np.float32) for k in shape]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    return (np.sqrt(X * X + Y * Y + Z * Z) - np.float32(r)).astype(np.float32)


def torus(n=40, R=0.55, r=0.22):
    g = np.linspace(-1, 1, n, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    q = np.sqrt(X * X + Y * Y) - np.float32(R)
    return (np.sqrt(q * q + Z * Z) - np.float32(r)).astype(np.float32)


def noise(shape=(13, 10, 9), seed=0, border=False):
    """White noise: every cube case, ambiguous faces included, shows up; border=True closes the surface."""
    v = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    if border:
        v[0], v[-1], v[:, 0], v[:, -1], v[:, :, 0], v[:, :, -1] = 1, 1, 1, 1, 1, 1
    return v


def all_cases():
    """A 2x(2*256)x2-ish strip is not needed: place the 256 corner patterns in separated cubes of one volume."""
    v = np.ones((3 * 16, 3 * 16, 3), dtype=np.float32)
    for c in range(256):
        i, j = 3 * (c // 16), 3 * (c % 16)
        for k in range(8):
            if (c >> k) & 1:
                v[i + ((k >> 2) & 1), j + ((k >> 1) & 1), k & 1] = -1.0 - 0.01 * k
    return v


VOLUMES = {
    "sphere33": dict(vol=lambda: sphere(33), level=0.0, spacing=(2 / 32,) * 3, origin=(-1, -1, -1)),
    "sphere_ragged": dict(vol=lambda: sphere(shape=(20, 33, 17)), level=0.05, spacing=(0.1, 0.0625, 0.125), origin=(-1, -1, -1)),
    "torus40": dict(vol=lambda: torus(40), level=0.0, spacing=(2 / 39,) * 3, origin=(-1, -1, -1)),
    "noise": dict(vol=lambda: noise(), level=0.1, spacing=(1, 1, 1), origin=(0, 0, 0)),
    "noise_closed": dict(vol=lambda: noise((16, 15, 14), 3, True), level=0.0, spacing=(0.5, 1.0, 2.0), origin=(3, -2, 1)),
    "all_cases": dict(vol=all_cases, level=0.0, spacing=(1, 1, 1), origin=(0, 0, 0)),
    "two_cube": dict(vol=lambda: np.array([[[-1, 1], [1, 1]], [[1, 1], [1, -2]]], dtype=np.float32), level=0.0,
                     spacing=(1, 1, 1), origin=(0, 0, 0)),
}
