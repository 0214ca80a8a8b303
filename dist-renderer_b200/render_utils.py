"""depth2normal -- host-side mirror of `core/utils/render_utils.py:9-43` (the `use_depth2normal` branch of
`SDFRenderer.render`, renderer.py:972-975).

Normals from central differences of the depth map instead of the decoder gradient.  This is a handful of elementwise
ops on an (H, W) map -- five orders of magnitude less work than the march that produced the depth -- and is kept in
plain (device-agnostic) PyTorch: the result is differentiable w.r.t. the depth exactly as in the reference.
A CPU test pins it bit-for-bit to the reference's own function.
"""
import torch


def depth2normal(depth, f_pix_x, f_pix_y=None):
    """depth (H, W) -> normal (H, W, 3).  Like the reference, background pixels (depth > 1e5 or == 0) are set to 0 IN
    PLACE in `depth` (render_utils.py:24-25) -- callers of render(use_depth2normal=True) therefore see a depth map whose
    background is 0, not 1e11 -- and get a zero normal; the one-pixel image border has zero finite differences."""
    if f_pix_y is None:
        f_pix_y = f_pix_x
    h, w = depth.shape
    bg = (depth > 1e5) | (depth == 0)
    depth[bg] = 0.0
    dzdx, dzdy = torch.zeros_like(depth), torch.zeros_like(depth)
    if w > 2:
        dzdx[:, 1:w - 1] = depth[:, 2:] - depth[:, :w - 2]
    if h > 2:
        dzdy[1:h - 1, :] = depth[2:, :] - depth[:h - 2, :]
    dzdx = dzdx * f_pix_x / 2.0
    dzdy = dzdy * f_pix_y / 2.0
    normal = torch.stack([dzdx, dzdy, -torch.ones_like(dzdx)]).permute(1, 2, 0)
    normal = normal / (torch.norm(normal, p=2, dim=2) + 1e-12)[:, :, None]
    normal[bg] = 0.0
    return normal
