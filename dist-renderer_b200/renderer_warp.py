"""SDFRenderer_warp -- two-view photometric warping on top of the CUDA sphere tracer.

Drop-in for `core/sdfrenderer/renderer_warp.py:13-144`: view 1 is rendered with depth
gradients, view 2 without; the hit points of view 1 are reprojected into view 2, filtered by a depth-consistency test
against view 2's rendered depth, and the colours of both images are compared at the corresponding pixels (L1).
All decoder work (ONE two-view march + one `render_normal`) runs on the engines of `renderer.SDFRenderer`; the
reprojection, the depth-consistency test, the bilinear sampling and the L1 loss are ONE kernel forward and one backward
(`dist_warp_loss_fwd / _bwd`, csrc/warp.cu) instead of a dozen elementwise PyTorch ops on compacted (3, N) tensors with
their boolean-mask host synchronisations; the loss carries gradients to `latent` (through view 1's depth), `R1`, `T1`,
`R2`, `T2` as in the reference.  The PyTorch formulation (`get_valid_points`, `compute_loss_color`) is kept as methods with
the reference's signatures for subclasses / callers that use them directly.

`grid_sample` is called with ``align_corners=True``: the reference was written for torch 1.1, whose default sampling
convention that is; the pixel normalisation `2 x / (W - 1) - 1` of `loss_utils.py:19-21`
only makes sense under it.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from . import _abi
from .renderer import SDFRenderer, _stream


class _WarpLossFn(torch.autograd.Function):
    """Sum of |img1 - sample(img2)| over the view-1 hit pixels whose reprojection into view 2 passes the depth test
    (renderer_warp.py:18-101), and the number of such pixels."""

    @staticmethod
    def forward(ctx, Zdepth1, R1, T1, R2, T2, ren, mask1, depth2, img1, img2, thres):
        with torch.cuda.device(ren.device):
            lib, st, dev, P = _abi.lib(), _stream(ren.device), ren.device, ren.P
            R1d, T1d = R1.detach().float().contiguous(), T1.detach().float().contiguous()
            R2d, T2d = R2.detach().float().contiguous(), T2.detach().float().contiguous()
            c1 = ren.get_camera_location(R1d, T1d).contiguous()
            cam = ren._c_camera(R1d, c1)
            K_host = (ctypes.c_float * 9)(*[float(v) for v in ren.intrinsic.reshape(-1)])
            Z1 = Zdepth1.detach().float().contiguous()
            m1 = mask1.to(torch.uint8).contiguous()
            loss_sum = torch.empty(1, device=dev)
            count = torch.empty(1, device=dev, dtype=torch.int32)
            keep = torch.empty(P, device=dev, dtype=torch.uint8)
            vis1, vis2 = torch.empty(P, 3, device=dev), torch.empty(P, 3, device=dev)
            _abi.check(lib.dist_warp_loss_fwd(cam, K_host, _abi.ptr(R2d), _abi.ptr(T2d), _abi.ptr(Z1), _abi.ptr(m1),
                                              _abi.ptr(depth2), _abi.ptr(img1), _abi.ptr(img2), float(thres),
                                              _abi.ptr(loss_sum), _abi.ptr(count), _abi.ptr(keep), _abi.ptr(vis1),
                                              _abi.ptr(vis2), st))
            ctx.ren = ren
            ctx.save_for_backward(Z1, R1d, T1d, R2d, T2d, keep, img1, img2)
            ctx.mark_non_differentiable(count, keep, vis1, vis2)
            return loss_sum, count, keep, vis1, vis2

    @staticmethod
    def backward(ctx, g, *_):
        Z1, R1d, T1d, R2d, T2d, keep, img1, img2 = ctx.saved_tensors
        ren = ctx.ren
        with torch.cuda.device(ren.device):
            lib, st, dev, P = _abi.lib(), _stream(ren.device), ren.device, ren.P
            c1 = ren.get_camera_location(R1d, T1d).contiguous()
            cam = ren._c_camera(R1d, c1)
            K_host = (ctypes.c_float * 9)(*[float(v) for v in ren.intrinsic.reshape(-1)])
            gs = g.detach().reshape(1).float().contiguous()
            dZ1, d_ray = torch.empty(P, device=dev), torch.empty(3, P, device=dev)
            d_c1, dR2, dT2 = torch.empty(3, device=dev), torch.empty(9, device=dev), torch.empty(3, device=dev)
            _abi.check(lib.dist_warp_loss_bwd(cam, K_host, _abi.ptr(R2d), _abi.ptr(T2d), _abi.ptr(Z1), _abi.ptr(keep),
                                              _abi.ptr(img1), _abi.ptr(img2), _abi.ptr(gs), _abi.ptr(dZ1), _abi.ptr(d_ray),
                                              _abi.ptr(d_c1), _abi.ptr(dR2), _abi.ptr(dT2), st))
            gR1 = gT1 = None
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                with torch.enable_grad():       # c1 = -R1^T T1, ray = normalize(R1^T K^-1 u): 12-float Jacobian on the host
                    Rg, Tg = R1d.clone().requires_grad_(True), T1d.clone().requires_grad_(True)
                    gR1, gT1 = torch.autograd.grad([ren.get_camera_location(Rg, Tg), ren.get_camera_rays(Rg)], [Rg, Tg],
                                                   [d_c1, d_ray], allow_unused=True)
            return (dZ1 if ctx.needs_input_grad[0] else None, gR1, gT1,
                    dR2.reshape(3, 3) if ctx.needs_input_grad[3] else None, dT2 if ctx.needs_input_grad[4] else None,
                    None, None, None, None, None, None)


def grid_sample_on_img(img, xy):
    """Bilinear sampling of img [B,C,H,W] at pixel coordinates xy [B,2,Ho,Wo]  (core/utils/loss_utils.py:9-25)."""
    H, W = img.shape[2], img.shape[3]
    gx = 2.0 * xy[:, 0] / max(W - 1, 1) - 1.0
    gy = 2.0 * xy[:, 1] / max(H - 1, 1) - 1.0
    return F.grid_sample(img, torch.stack([gx, gy], -1), align_corners=True)


class SDFRenderer_warp(SDFRenderer):
    def __init__(self, decoder, intrinsic, img_hw=None, march_step=50, buffer_size=5, ray_marching_ratio=1.5,
                 max_sample_dist=0.2, threshold=5e-5, use_gpu=True, is_eval=True, transform_matrix=None, engine=None):
        # renderer_warp.py:14-16
        super(SDFRenderer_warp, self).__init__(decoder, intrinsic, img_hw=img_hw, transform_matrix=transform_matrix,
                                               march_step=march_step, buffer_size=buffer_size,
                                               ray_marching_ratio=ray_marching_ratio, max_sample_dist=max_sample_dist,
                                               threshold=threshold, use_gpu=use_gpu, is_eval=is_eval, engine=engine)
        self.counter = 0

    # ------------------------------------------------------------------------------------------------------
    def valid_points_depth(self, xy_proj, Zdepth2, depth2_proj, thres_depth):
        """Squared difference between the reprojected depth and view 2's rendered depth (renderer_warp.py:62-72)."""
        h, w = self.img_hw
        depth2 = (Zdepth2 * self.calib_map).reshape(1, 1, h, w)
        sampled = grid_sample_on_img(depth2, xy_proj).reshape(-1)
        return (depth2_proj - sampled) ** 2 < thres_depth

    def get_valid_points(self, render_out1, render_out2, R1, T1, R2, T2, thres_depth, gt_mask=None):
        """renderer_warp.py:18-53: view-1 hit points (world frame, with depth gradient) projected into view 2."""
        Zdepth1, valid_mask1, _ = render_out1
        Zdepth2, _, _ = render_out2
        cam_pos1 = self.get_camera_location(R1, T1)
        rays1 = self.get_camera_rays(R1)[:, valid_mask1]
        pts = self.generate_point_samples(cam_pos1, rays1, Zdepth1[valid_mask1], inv_transform=False,
                                          has_zdepth_grad=True)
        xyz = torch.matmul(self.K, torch.matmul(R2, pts) + T2[:, None])
        xy = (xyz[:2, :] / xyz[2, :])[None, :, :, None]                      # [1, 2, N, 1]
        keep_mask = torch.ones(xyz.shape[1], dtype=torch.bool, device=xyz.device)  # the mask test is disabled upstream (:43)
        keep_depth = self.valid_points_depth(xy, Zdepth2, xyz[2, :], thres_depth)
        return xy[:, :, keep_depth, :], keep_mask, keep_depth

    def compute_loss_color(self, img1, img2, xy_proj, valid_mask1, valid_mask_index, valid_depth_index):
        """Mean L1 colour difference at corresponding pixels + the two masked colour maps (renderer_warp.py:74-101)."""
        h, w = self.img_hw
        img1 = img1.to(self.device)
        c1 = img1.reshape(h * w, 3)[valid_mask1][valid_mask_index][valid_depth_index]
        c2 = grid_sample_on_img(img2.to(self.device).permute(2, 0, 1)[None], xy_proj)   # [1, 3, n, 1]
        c2 = c2.reshape(3, -1).permute(1, 0)
        loss = torch.mean(torch.abs(c1 - c2))
        idx1 = torch.nonzero(valid_mask1).reshape(-1)[valid_mask_index][valid_depth_index]
        final = torch.zeros(h * w, dtype=torch.bool, device=self.device)
        final[idx1] = True
        final = final.reshape(h, w)
        vis1, vis2 = torch.zeros_like(img1), torch.zeros_like(img1)
        vis1[final] = c1
        vis2[final] = c2
        return loss, vis1, vis2

    def render_warp(self, latent, R1, T1, R2, T2, img1, img2, clamp_dist=0.1, profile=False, no_grad_normal=False,
                    thres_depth=0.001):
        """renderer_warp.py:103-144.  Returns (loss_color, color_valid_1, color_valid_2, valid_mask1, valid_mask2,
        min_sdf_sample1, min_sdf_sample2, Znormal1, depth1_rendered)."""
        h, w = self.img_hw
        # renderer_warp.py:108-109 renders the two views one after the other (the second with no_grad_depth=True); here
        # both are marched TOGETHER (dist_camera_t.n_views = 2, per-view depth-gradient flag): one compaction list, one
        # decoder launch per step, one latency-bound tail -- with the maps of two separate render_depth calls bit for bit
        pair = self._fused_child(2)
        Z, M, S = pair.render_depth(latent, torch.stack([R1, R2], 0), torch.stack([T1, T2], 0), clamp_dist=clamp_dist,
                                    profile=profile, no_grad_depth=[False, True], check_empty=False)
        P = h * w
        out1, out2 = (Z[:P], M[:P], S[:P]), (Z[P:], M[P:], S[P:])
        pair._raise_if_empty()
        Zdepth1, valid_mask1, min_sdf1 = out1
        Zdepth2, valid_mask2, min_sdf2 = out2
        min_sdf1, min_sdf2 = min_sdf1.reshape(h, w), min_sdf2.reshape(h, w)
        if int(valid_mask1.sum()) == 0:
            loss_color = torch.zeros((), device=self.device, requires_grad=True)
            vis1 = torch.zeros_like(img1).to(self.device)
            vis2 = torch.zeros_like(img1).to(self.device)
        else:
            # renderer_warp.py:125-127 (get_valid_points + compute_loss_color): one fused kernel
            depth2 = (Zdepth2.detach() * self.calib_map).contiguous()
            i1 = img1.to(self.device).float().reshape(P, 3).contiguous()
            i2 = img2.to(self.device).float().reshape(P, 3).contiguous()
            loss_sum, count, _keep, v1, v2 = _WarpLossFn.apply(Zdepth1, R1, T1, R2, T2, self, valid_mask1, depth2, i1, i2,
                                                               float(thres_depth))
            loss_color = (loss_sum / (3.0 * count.float())).reshape(())      # mean over (n, 3); NaN when nothing is kept
            vis1, vis2 = v1.reshape(h, w, 3).to(img1.dtype), v2.reshape(h, w, 3).to(img1.dtype)
        normal1 = self.render_normal(latent, R1, T1, Zdepth1, valid_mask1, no_grad=no_grad_normal, clamp_dist=clamp_dist)
        Zn = torch.matmul(R1, normal1)
        Zn = torch.cat([Zn[:1] * (-1), Zn[1:]], 0).reshape(3, h, w).permute(1, 2, 0)
        depth1 = torch.where(valid_mask1, Zdepth1 * self.calib_map, torch.zeros_like(Zdepth1)).reshape(h, w)
        return (loss_color, vis1, vis2, valid_mask1.reshape(h, w).type(torch.uint8),
                valid_mask2.reshape(h, w).type(torch.uint8), min_sdf1, min_sdf2, Zn, depth1)
