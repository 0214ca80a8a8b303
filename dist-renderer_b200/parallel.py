"""Ray-tile sharding across GPUs (no counterpart in the reference).

Rays are independent, so an image is split into interleaved row bands -- rank r renders rows r, r+N, r+2N, ... with
its own SDFRenderer(rows=...) and replicated decoder weights (7 MB) -- which balances the hit density across ranks.
The only data-path collective is ONE all-gather of the packed per-rank outputs (depth, normal, mask, min_sdf, and
optionally a small vector such as the partial latent/camera gradients, summed locally after the gather).
"""
import torch
import torch.distributed as dist

from .renderer import SDFRenderer


def band(H, rank, world):
    """(row0, row_step, n_rows) of `rank`'s interleaved band of an H-row image."""
    return rank, world, len(range(rank, H, world))


class ShardedSDFRenderer(object):
    def __init__(self, decoder, intrinsic, img_hw, rank=None, world_size=None, group=None, **kw):
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        self.local = SDFRenderer(decoder, intrinsic, img_hw=img_hw, rows=band(self.img_hw[0], self.rank, self.world), **kw)

    def render(self, latent, R, T, **kw):
        """Local band of render(): (depth[n_rows,W], normal[n_rows,W,3], mask, min_sdf)."""
        return self.local.render(latent, R, T, **kw)

    def gather(self, outs, extra=None):
        return gather_bands(outs, self.img_hw, self.rank, self.world, extra=extra, group=self.group)


def pack_band(outs, max_rows, extra=None):
    """[6, max_rows, W] fp32 (depth, normal xyz, mask, min_sdf; bands shorter than max_rows are zero padded) + extra."""
    depth, normal, mask, min_sdf = outs
    n_rows, W = depth.shape
    body = torch.cat([depth.detach()[None], normal.detach().permute(2, 0, 1), mask.detach().float()[None],
                      min_sdf.detach()[None]], 0)
    if n_rows < max_rows:
        body = torch.cat([body, body.new_zeros(6, max_rows - n_rows, W)], 1)
    flat = body.reshape(-1)
    if extra is not None:
        flat = torch.cat([flat, extra.detach().reshape(-1).float()])
    return flat


def unpack_bands(gathered, img_hw, world, n_extra=0):
    """gathered: [world, 6*max_rows*W + n_extra] -> full-image outputs (+ the per-rank extras [world, n_extra]).
    Image row k*world + r is row k of rank r's band, so one permute interleaves all bands."""
    H, W = img_hw
    max_rows = (H + world - 1) // world
    body = gathered[:, :6 * max_rows * W].reshape(world, 6, max_rows, W)
    full = body.permute(1, 2, 0, 3).reshape(6, max_rows * world, W)[:, :H]
    extras = gathered[:, 6 * max_rows * W:] if n_extra else None
    return (full[0], full[1:4].permute(1, 2, 0), full[4].to(torch.uint8), full[5]), extras


def gather_bands(outs, img_hw, rank, world, extra=None, group=None):
    """One all-gather of the packed band; every rank returns the full image (and the stacked extras)."""
    H, W = img_hw
    max_rows = (H + world - 1) // world
    flat = pack_band(outs, max_rows, extra)
    if world == 1:
        gathered = flat[None]
    else:
        gathered = torch.empty(world * flat.numel(), device=flat.device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, flat, group=group)      # flat output: accepted by NCCL and gloo alike
        gathered = gathered.view(world, flat.numel())
    return unpack_bands(gathered, img_hw, world, 0 if extra is None else extra.numel())
