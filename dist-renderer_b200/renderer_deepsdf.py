"""SDFRenderer_deepsdf -- SDF supervision samples around / in front of an observed depth map.

Drop-in for `core/sdfrenderer/renderer_deepsdf.py:10-66` (the fourth class the reference's `core.sdfrenderer` package
exports): back-projects a depth map (and its normals) into the decoder frame with the renderer's camera helpers and
queries the decoder a small distance on either side of the observed surface, or at random depths in front of it.
The decoder rows run on the fused CUDA engines through `functional.decode_sdf`; the geometry is a few elementwise ops.
"""
import torch

from . import functional
from .renderer import SDFRenderer


class SDFRenderer_deepsdf(SDFRenderer):
    def __init__(self, decoder, intrinsic, img_hw=None, march_step=50, buffer_size=5, ray_marching_ratio=1.5,
                 max_sample_dist=0.2, threshold=5e-5, use_gpu=True, is_eval=True, engine=None):
        # renderer_deepsdf.py:11-12
        super(SDFRenderer_deepsdf, self).__init__(decoder, intrinsic, img_hw=img_hw, march_step=march_step,
                                                  buffer_size=buffer_size, ray_marching_ratio=ray_marching_ratio,
                                                  max_sample_dist=max_sample_dist, threshold=threshold, use_gpu=use_gpu,
                                                  is_eval=is_eval, engine=engine)

    def _observed_points(self, RT, depth):
        """(valid pixel mask, camera position, rays of the valid pixels, their ray depths) of a depth map (H, W) whose
        background is 0 or >= 1e5; `depth` is z-depth, `calib_map` converts it to depth along the ray."""
        R, T = RT[:, :3], RT[:, 3]
        depth = depth.reshape(-1)
        valid = (depth < 1e5) & (depth > 0)
        zdepth = depth[valid] / self.calib_map[valid]
        return valid, self.get_camera_location(R, T), self.get_camera_rays(R)[:, valid], zdepth

    def get_samples(self, latent, RT, depth, normal, clamp_dist=0.1, eta=0.01, use_rand=True):
        """(sdf(p + eta n) - eta, sdf(p - eta n) + eta) at the observed surface points p with normals n: both are zero
        for a decoder that reproduces the observation (renderer_deepsdf.py:14-43).  eta is drawn per pixel in [0, eta)
        when `use_rand`."""
        valid, cam_pos, rays, zdepth = self._observed_points(RT, depth)
        n_cam = normal.reshape(-1, 3)[valid, :]
        points = self.generate_point_samples(cam_pos, rays, zdepth, has_zdepth_grad=False).transpose(1, 0)
        eta_map = (torch.rand_like(zdepth) if use_rand else torch.ones_like(zdepth)) * eta
        offset = self.inv_transform_points(n_cam.transpose(1, 0)).transpose(1, 0) * eta_map.unsqueeze(-1)
        pos = functional.decode_sdf(self.decoder, latent, points + offset, clamp_dist=clamp_dist).squeeze(-1) - eta_map
        neg = functional.decode_sdf(self.decoder, latent, points - offset, clamp_dist=clamp_dist).squeeze(-1) + eta_map
        return pos, neg

    def get_freespace_samples(self, latent, RT, depth, clamp_dist=0.1, number=1):
        """sdf at `number` random depths per valid pixel between the camera and the observed surface (free space, where
        the sdf should be positive) -- renderer_deepsdf.py:45-65."""
        valid, cam_pos, rays, zdepth = self._observed_points(RT, depth)
        samples = []
        for _ in range(number):
            z = zdepth * (torch.rand_like(zdepth) * 1.0)
            points = self.generate_point_samples(cam_pos, rays, z, has_zdepth_grad=False).transpose(1, 0)
            samples.append(functional.decode_sdf(self.decoder, latent, points, clamp_dist=clamp_dist).squeeze(-1))
        return torch.cat(samples, 0)
