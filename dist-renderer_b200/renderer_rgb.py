"""SDFRenderer_color -- the colour / shading renderer on top of the CUDA sphere tracer.

Drop-in for `core/sdfrenderer/renderer_rgb.py:12-125` (the demo path): depth, normal and
silhouette come from `renderer.SDFRenderer` (tensor-core march, fused analytic normals); the colour network -- a
DeepSDF-style MLP with three outputs fed [shape code | colour code | xyz] -- is queried ONCE per hit pixel on the same
fused engines (`functional.decode_color`: both codes folded into the per-render biases, hidden layers shared, three
dot-product epilogues), and the optional point-light shading is a few elementwise ops on the hit pixels.  The colour
carries gradients to both codes as upstream.
"""
import torch

from .functional import decode_color
from .renderer import SDFRenderer


class SDFRenderer_color(SDFRenderer):
    def __init__(self, decoder, decoder_color, intrinsic, img_hw=None, march_step=50, buffer_size=5,
                 ray_marching_ratio=1.5, max_sample_dist=0.2, threshold=5e-5, use_gpu=True, is_eval=True, engine=None):
        # renderer_rgb.py:13-18
        super(SDFRenderer_color, self).__init__(decoder, intrinsic, img_hw=img_hw, march_step=march_step,
                                                buffer_size=buffer_size, ray_marching_ratio=ray_marching_ratio,
                                                max_sample_dist=max_sample_dist, threshold=threshold, use_gpu=use_gpu,
                                                is_eval=is_eval, engine=engine)
        self.decoder_color = decoder_color.eval() if is_eval else decoder_color

    def render_color(self, latent_color, latent, cam_pos, cam_rays, Zdepth, valid_mask, no_grad=False):
        """rgb map (H, W, 3), zero off the surface -- renderer_rgb.py:20-38.  The hit points are built from the detached
        depth (`has_zdepth_grad=False`), so the colour depends on the codes and the camera, not on the march."""
        h, w = self.img_hw
        valid_mask = valid_mask.bool()
        idx = torch.nonzero(valid_mask).reshape(-1)
        if idx.numel() == 0:
            # upstream returns the empty canvas reshaped to (3, H*W) here (renderer_rgb.py:27-28); kept as is
            return torch.zeros(3, h * w, device=self.device, dtype=torch.float32)
        points = self.generate_point_samples(cam_pos, cam_rays[:, idx], Zdepth[idx], has_zdepth_grad=False)
        rgb = decode_color(self.decoder_color, latent_color, latent, points.transpose(1, 0), no_grad=no_grad,
                           engine=self.engine)
        color = torch.zeros(h * w, 3, device=self.device, dtype=rgb.dtype).index_copy(0, idx, rgb).reshape(h, w, 3)
        return color.detach() if no_grad else color

    def compute_shading_maps(self, R, T, lighting_locations, Zdepth, Znormal, valid_mask):
        """Lambertian term n . l per light, (M, H*W), zero off the surface -- renderer_rgb.py:40-68.
        `Znormal` (H*W, 3) is in the camera frame, the light positions (M, 3) in the world frame."""
        valid_mask = valid_mask.bool()
        cam_pos, cam_rays = self.get_camera_location(R, T), self.get_camera_rays(R)
        points = self.generate_point_samples(cam_pos, cam_rays[:, valid_mask], Zdepth[valid_mask], inv_transform=False,
                                             has_zdepth_grad=False).transpose(1, 0)                  # (N, 3)
        directions = (lighting_locations[:, None, :] - points[None, :, :]).permute(0, 2, 1)         # (M, 3, N)
        directions = directions / torch.norm(directions, p=2, dim=1)[:, None, :].repeat(1, 3, 1)
        # upstream's bmm(R[None], directions) only runs for one light; the rotation is broadcast over the lights here
        Zdirections = torch.bmm(R.unsqueeze(0).expand(directions.shape[0], 3, 3), directions).permute(0, 2, 1)
        valid = (Zdirections * Znormal[valid_mask, :][None, :, :]).sum(2)                           # (M, N)
        maps = torch.zeros(lighting_locations.shape[0], Zdepth.shape[0], device=Zdepth.device, dtype=valid.dtype)
        maps[:, valid_mask] = valid
        return maps

    def render(self, latent_color, latent, R, T, clamp_dist=0.1, profile=False, no_grad=False, lighting_locations=None,
               lighting_energies=None):
        """(depth[H,W], Znormal[H,W,3], color[H,W,3], mask[H,W] uint8, min_sdf[H,W]) -- renderer_rgb.py:70-125."""
        h, w = self.img_hw
        Zdepth, valid_mask, min_sdf_sample = self.render_depth(latent, R, T, clamp_dist=clamp_dist, profile=profile,
                                                               no_grad=no_grad)
        normal = self.render_normal(latent, R, T, Zdepth, valid_mask, clamp_dist=clamp_dist, no_grad=no_grad)
        Znormal = torch.matmul(R, normal)
        Znormal = torch.cat([Znormal[:1] * (-1), Znormal[1:]], 0)                                    # :93
        color = self.render_color(latent_color, latent, self.get_camera_location(R, T), self.get_camera_rays(R), Zdepth,
                                  valid_mask, no_grad=no_grad)
        depth = torch.where(valid_mask, Zdepth * self.calib_map, torch.full_like(Zdepth, 1e11)).reshape(h, w)
        Znormal = Znormal.reshape(3, h, w).permute(1, 2, 0)
        mask8 = valid_mask.reshape(h, w).type(torch.uint8)
        min_sdf_sample = min_sdf_sample.reshape(h, w)
        if lighting_locations is None:
            return depth, Znormal, color, mask8, min_sdf_sample
        if lighting_energies is None:
            lighting_energies = torch.ones_like(lighting_locations[:, 0])
        maps = self.compute_shading_maps(R, T, lighting_locations, Zdepth.reshape(-1), Znormal.reshape(-1, 3),
                                         valid_mask.reshape(-1))
        shading = (maps * lighting_energies[:, None].repeat(1, maps.shape[1])).sum(0).reshape(h, w)
        color = color * shading[:, :, None].repeat(1, 1, 3)
        return depth, Znormal, color, mask8, min_sdf_sample
