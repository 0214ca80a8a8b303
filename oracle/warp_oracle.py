"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's two-view warp renderer
(core/sdfrenderer/renderer_warp.py:18-144 with core/utils/loss_utils.py:9-25), on top of oracle/sdf_oracle.py.
Pinned against the reference itself (loaded by oracle/ref_shim.load_warp) in tests/test_oracle.py and through
tests/golden/warp_40.npz."""
import torch
import torch.nn.functional as F

from .sdf_oracle import OracleSDFRenderer


def sample_img(img, xy):                                          # loss_utils.py:9-25 (torch-1.1 convention)
    H, W = img.shape[2], img.shape[3]
    g = torch.stack([2.0 * xy[:, 0] / max(W - 1, 1) - 1.0, 2.0 * xy[:, 1] / max(H - 1, 1) - 1.0], -1)
    return F.grid_sample(img, g, align_corners=True)


class OracleWarpRenderer(OracleSDFRenderer):
    def render_warp(self, latent, R1, T1, R2, T2, img1, img2, clamp_dist=0.1, thres_depth=0.001):
        h, w = self.img_hw
        Z1, m1, s1 = self.render_depth(latent, R1, T1, clamp_dist=clamp_dist)                      # renderer_warp.py:108
        Z2, m2, s2 = self.render_depth(latent, R2, T2, clamp_dist=clamp_dist, no_grad_depth=True)  # :109
        if int(m1.sum()) == 0:
            loss = torch.zeros((), requires_grad=True)
        else:
            c1 = self.camera_location(R1, T1)                                                       # :22-28
            pts = self.camera_rays(R1)[:, m1] * Z1[m1][None, :] + c1[:, None]
            xyz = self.K @ (R2 @ pts + T2[:, None])                                                 # :31
            xy = (xyz[:2] / xyz[2])[None, :, :, None]
            d2 = sample_img((Z2 * self.calib_map).reshape(1, 1, h, w), xy).reshape(-1)              # :62-68
            keep = (xyz[2] - d2) ** 2 < thres_depth                                                 # :70-71
            a = img1.reshape(h * w, 3)[m1][keep]                                                    # :76-79
            b = sample_img(img2.permute(2, 0, 1)[None], xy[:, :, keep, :]).reshape(3, -1).t()       # :81-83
            loss = torch.mean(torch.abs(a - b))                                                     # :85
        n1 = self.render_normal(latent, R1, T1, Z1, m1, clamp_dist=clamp_dist)
        Zn = R1 @ n1
        Zn = torch.cat([Zn[:1] * (-1), Zn[1:]], 0).reshape(3, h, w).permute(1, 2, 0)
        depth1 = torch.where(m1, Z1 * self.calib_map, torch.zeros_like(Z1)).reshape(h, w)
        return (loss, m1.reshape(h, w).to(torch.uint8), m2.reshape(h, w).to(torch.uint8), s1.reshape(h, w),
                s2.reshape(h, w), Zn, depth1)
