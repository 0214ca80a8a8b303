"""TEST INFRASTRUCTURE ONLY -- CPU/GPU-agnostic restatement of the single-view loss the reference's optimisation loop puts on
the renderer's outputs (config 3, `run_single_shape.py` path):

    compute_all_loss                          core/inv_optimizer/loss_single.py:7-57
    compute_loss_mask / depth / normal        core/utils/loss_utils.py:59-179
    weighting                                 core/inv_optimizer/optimize_single.py:74-78, run_single_shape.py:93-98

for ground-truth maps of the rendered resolution (ratio 1: `downsize_img_tensor` is then the identity; the reference's own
call needs a one-line patch for uint8 masks on CPU).  Pinned to the reference in tests/test_oracle.py
(`test_loss_oracle_matches_live_reference`); the GPU suite uses it to check that the product's outputs -- uint8 mask,
1e11 background depth, zero background normals, autograd connectivity -- are what `compute_all_loss` expects.
Nothing in the product package imports this file.
"""
import torch

WEIGHTS = dict(w_depth=10.0, w_normal=5.0, w_mask_gt=1.0, w_mask_out=1.0, w_l2reg=1.0)   # run_single_shape.py:93-98


def _mean_or_zero(values, like_mask):
    return values.mean() if values.numel() else torch.zeros_like(like_mask).float().mean()


def loss_mask(min_sdf, mask, mask_gt, threshold=5e-5):     # loss_utils.py:59-103
    mask, mask_gt = mask.bool(), mask_gt.bool()
    false_gt = mask_gt & ~(mask & mask_gt)
    l_gt = _mean_or_zero(torch.max(min_sdf[false_gt] - threshold, torch.zeros_like(min_sdf[false_gt])), false_gt)
    false_out = mask & ~(mask & mask_gt)
    l_out = _mean_or_zero(torch.max(-min_sdf[false_out] + threshold, torch.zeros_like(min_sdf[false_out])), false_out)
    return l_gt, l_out


def loss_depth(depth, mask, depth_gt, mask_gt):            # loss_utils.py:105-132
    ok = mask.bool() & mask_gt.bool() & (depth_gt > 0) & (depth_gt < 1e5)
    return torch.abs(depth[ok] - depth_gt[ok]).mean() if bool(ok.any()) else torch.zeros_like(ok).float().mean()


def loss_normal(normal, mask, normal_gt, mask_gt):         # loss_utils.py:134-179
    ok = mask.bool() & mask_gt.bool() & (torch.norm(normal, p=2, dim=2) != 0)
    if not bool(ok.any()):
        return torch.zeros_like(ok).float().mean()
    a, b = normal[ok], normal_gt[ok]
    a = a.div(torch.norm(a, p=2, dim=1)[:, None].repeat(1, 3) + 1e-12)
    b = b.div(torch.norm(b, p=2, dim=1)[:, None].repeat(1, 3) + 1e-12)
    return (-(a * b).sum(1)).mean()


def compute_all_loss(renderer, latent, extrinsic, gt_pack, threshold=5e-5, ray_marching_type='pyramid_recursive'):
    """loss_single.py:7-57 with all three tasks on.  gt_pack = {depth (H,W), normal (H,W,3), silhouette (H,W) uint8}."""
    depth, normal, mask, min_sdf = renderer.render(latent, extrinsic[:, :3], extrinsic[:, 3],
                                                   ray_marching_type=ray_marching_type, no_grad_depth=False, no_grad_normal=False)
    pack = {}
    pack['mask_gt'], pack['mask_out'] = loss_mask(min_sdf, mask, gt_pack["silhouette"], threshold)
    pack['depth'] = loss_depth(depth, mask, gt_pack["depth"], gt_pack["silhouette"])
    pack['normal'] = loss_normal(normal, mask, gt_pack["normal"], gt_pack["silhouette"])
    pack['l2reg'] = torch.mean(latent.pow(2))
    return pack


def total(pack, w=WEIGHTS):                                # optimize_single.py:74-78
    return (w['w_depth'] * pack['depth'] + w['w_normal'] * pack['normal'] + w['w_mask_gt'] * pack['mask_gt'] +
            w['w_mask_out'] * pack['mask_out'] + w['w_l2reg'] * pack['l2reg'])
