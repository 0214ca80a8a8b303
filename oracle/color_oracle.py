"""CPU restatement of the reference's colour renderer -- TEST INFRASTRUCTURE ONLY (see oracle/sdf_oracle.py).

Follows `core/sdfrenderer/renderer_rgb.py:12-125` (class SDFRenderer_color) and `core/utils/decoder_utils.py:94-112`
(decode_color).  Pinned bit-for-bit against the reference itself (oracle/ref_shim.py::load_color) in
tests/test_oracle.py::test_color_oracle_matches_live_reference and against tests/golden/color_24.npz.
"""
import torch

from .sdf_oracle import OracleSDFRenderer


def decode_color(decoder, color_code, shape_code, points, max_points=100000, no_grad=False):
    """decoder_utils.py:94-112: rows [shape code | colour code | xyz] -> rgb, in chunks of max_points."""
    n, out, start = points.shape[0], [], 0
    while True:
        end = min(start + max_points, n)
        rows = torch.cat([shape_code.expand(end - start, -1), color_code.expand(end - start, -1), points[start:end]], 1)
        c = decoder.inference(rows)
        out.append(c.detach() if no_grad else c)
        start = end
        if end == n:
            break
    return torch.cat(out, 0)


class OracleColorRenderer(OracleSDFRenderer):
    def __init__(self, decoder, decoder_color, intrinsic, **kw):      # renderer_rgb.py:13-18
        super().__init__(decoder, intrinsic, **kw)
        self.decoder_color = decoder_color.eval()

    def render_color(self, latent_color, latent, c, rays, Zdepth, mask, no_grad=False):   # renderer_rgb.py:20-38
        h, w = self.img_hw
        if int(mask.sum()) == 0:
            return torch.zeros(3, h * w, dtype=self.dtype)            # :27-28 (shape quirk of the reference kept)
        pts = self.points_on_rays(c, rays[:, mask], Zdepth[mask])     # :30, depth detached, decoder frame
        rgb = decode_color(self.decoder_color, latent_color, latent, pts.t(), no_grad=no_grad)
        idx = torch.nonzero(mask).reshape(-1)
        color = torch.zeros(h * w, 3, dtype=self.dtype).index_copy(0, idx, rgb).reshape(h, w, 3)   # :34 copy_index
        return color.detach() if no_grad else color

    def shading_maps(self, R, T, lights, Zdepth, Znormal, mask):     # renderer_rgb.py:40-68
        c, rays = self.camera_location(R, T), self.camera_rays(R)
        pts = self.points_on_rays(c, rays[:, mask], Zdepth[mask], inv_transform=False).t()      # (N, 3) world frame
        d = (lights[:, None, :] - pts[None, :, :]).permute(0, 2, 1)                               # (M, 3, N)
        d = d / torch.norm(d, p=2, dim=1)[:, None, :].repeat(1, 3, 1)
        zd = torch.bmm(R.unsqueeze(0).expand(d.shape[0], 3, 3), d).permute(0, 2, 1)              # (M, N, 3) camera frame
        s = (zd * Znormal[mask, :][None]).sum(2)                                                  # (M, N)
        maps = torch.zeros(lights.shape[0], Zdepth.shape[0], dtype=self.dtype)
        maps[:, mask] = s
        return maps

    def render(self, latent_color, latent, R, T, clamp_dist=0.1, no_grad=False, lighting_locations=None,
               lighting_energies=None):                                                           # renderer_rgb.py:70-125
        h, w = self.img_hw
        Zdepth, mask, min_map = self.render_depth(latent, R, T, clamp_dist=clamp_dist, no_grad=no_grad)
        normal = self.render_normal(latent, R, T, Zdepth, mask, clamp_dist=clamp_dist, no_grad=no_grad)
        Zn = torch.matmul(R, normal)
        Zn = torch.cat([Zn[:1] * (-1), Zn[1:]], 0)                                               # :93 flip x
        color = self.render_color(latent_color, latent, self.camera_location(R, T), self.camera_rays(R), Zdepth, mask,
                                  no_grad=no_grad)
        depth = torch.ones_like(Zdepth) * 1e11
        depth[mask] = Zdepth[mask].clone() * self.calib_map[mask]
        Zn = Zn.reshape(3, h, w).permute(1, 2, 0)
        out_mask = mask.reshape(h, w).to(torch.uint8)
        if lighting_locations is not None:
            e = torch.ones_like(lighting_locations[:, 0]) if lighting_energies is None else lighting_energies
            maps = self.shading_maps(R, T, lighting_locations, Zdepth.reshape(-1), Zn.reshape(-1, 3), mask.reshape(-1))
            shading = (maps * e[:, None].repeat(1, maps.shape[1])).sum(0).reshape(h, w)
            color = color * shading[:, :, None].repeat(1, 1, 3)
        return depth.reshape(h, w), Zn, color, out_mask, min_map.reshape(h, w)
