"""Ray-tile sharding across GPUs (no counterpart in the reference).

Rays are independent, so an image is split into interleaved bands of 4-row groups -- rank r renders row groups
r, r+N, r+2N, ... (rows 4r..4r+3, 4(r+N)..4(r+N)+3, ...) with its own SDFRenderer(rows=...) and replicated decoder
operands -- which balances the hit density across ranks AND keeps the 1/2- and 1/4-resolution levels of the
reference's default `pyramid_recursive` march local to a band (a 4x4 pixel block never straddles two ranks).
The only data-path collective is ONE all-gather per render of the packed per-rank outputs: 21 bytes per ray
(depth f32, min_sdf f32, normal 3 x f32, mask u8) + the band's view statistics (for the 'No valid depth' test, which
has to be taken over the whole image: every rank raises, or none) + optionally a small fp32 vector such as the partial
latent / camera gradients, summed locally after the gather.

Not reproduced across ranks: the early-break padding rule of renderer.py:562-567 looks at the number of march steps
the WHOLE image needed; each band uses its own count.  It only matters when every ray of a band finishes in fewer than
`buffer_size` steps.
"""
import torch
import torch.distributed as dist

from .renderer import SDFRenderer

GROUP = 4   # rows per interleaved group: the pyramid march pools 4x4 pixel blocks (renderer.py:13 scale_list=[4,2,1])


def band(H, rank, world, group=GROUP):
    """(row0, row_step, n_rows, row_group) of `rank`'s band of an H-row image: row groups rank, rank+world, ..."""
    n_groups = (H + group - 1) // group
    mine = range(rank, n_groups, world)
    n_rows = sum(min(group, H - g * group) for g in mine)
    return rank * group, world * group, n_rows, group


def _max_groups(H, world, group=GROUP):
    return ((H + group - 1) // group + world - 1) // world


def _layout(H, W, world, n_extra, group=GROUP):
    """Byte offsets of one rank's packed band: depth | min_sdf | normal | mask | (pad) | stat[4] i32 | extra f32."""
    n = _max_groups(H, world, group) * group * W        # pixels of the largest band; shorter bands are zero padded
    o_depth, o_min, o_nrm, o_mask = 0, 4 * n, 8 * n, 20 * n
    o_stat = (21 * n + 3) // 4 * 4
    o_extra = o_stat + 16
    return n, o_depth, o_min, o_nrm, o_mask, o_stat, o_extra, o_extra + 4 * n_extra


def pack_band(outs, img_hw, world, stat=None, extra=None, group=GROUP):
    """One uint8 buffer holding this rank's band (21 B per ray), its view statistics and the optional fp32 extras."""
    depth, normal, mask, min_sdf = outs
    H, W = img_hw
    n_extra = 0 if extra is None else extra.numel()
    n, o_depth, o_min, o_nrm, o_mask, o_stat, o_extra, total = _layout(H, W, world, n_extra, group)
    k = depth.numel()
    buf = torch.zeros(total, device=depth.device, dtype=torch.uint8)
    buf[o_depth:o_depth + 4 * k].view(torch.float32).copy_(depth.detach().reshape(-1))
    buf[o_min:o_min + 4 * k].view(torch.float32).copy_(min_sdf.detach().reshape(-1))
    buf[o_nrm:o_nrm + 12 * k].view(torch.float32).copy_(normal.detach().reshape(-1))
    buf[o_mask:o_mask + k].copy_(mask.detach().reshape(-1))
    if stat is not None:
        buf[o_stat:o_stat + 16].view(torch.int32).copy_(stat.reshape(-1)[:4])
    if n_extra:
        buf[o_extra:o_extra + 4 * n_extra].view(torch.float32).copy_(extra.detach().reshape(-1).float())
    return buf


def unpack_bands(gathered, img_hw, world, n_extra=0, group=GROUP):
    """gathered: [world, bytes] uint8 -> full-image outputs, per-rank stats [world, 4] int32, extras [world, n_extra].
    Image row (gl * world + r) * group + i is row gl * group + i of rank r's band: one permute per map."""
    H, W = img_hw
    n, o_depth, o_min, o_nrm, o_mask, o_stat, o_extra, total = _layout(H, W, world, n_extra, group)
    mg = _max_groups(H, world, group)

    def image(off, nbytes, dtype, tail):
        x = gathered[:, off:off + nbytes * n].contiguous().view(dtype).reshape((world, mg, group, W) + tail)
        return x.permute(1, 0, 2, 3, *range(4, 4 + len(tail))).reshape((mg * world * group, W) + tail)[:H]
    depth = image(o_depth, 4, torch.float32, ())
    min_sdf = image(o_min, 4, torch.float32, ())
    normal = image(o_nrm, 12, torch.float32, (3,))
    mask = image(o_mask, 1, torch.uint8, ())
    stats = gathered[:, o_stat:o_stat + 16].contiguous().view(torch.int32)
    extras = gathered[:, o_extra:o_extra + 4 * n_extra].contiguous().view(torch.float32) if n_extra else None
    return (depth, normal, mask, min_sdf), stats, extras


def gather_bands(outs, img_hw, rank, world, stat=None, extra=None, group=None, row_group=GROUP, events=None):
    """One all-gather of the packed band; every rank returns (full image maps, stats [world,4], extras [world,n])."""
    def mark(name):
        if events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.setdefault(name, []).append(ev)
    mark("pack0")
    flat = pack_band(outs, img_hw, world, stat, extra, row_group)
    mark("pack1")
    if world == 1:
        gathered = flat[None]
    else:
        gathered = torch.empty(world * flat.numel(), device=flat.device, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, flat, group=group)      # flat output: accepted by NCCL and gloo alike
        gathered = gathered.view(world, flat.numel())
    mark("gather1")
    res = unpack_bands(gathered, img_hw, world, 0 if extra is None else extra.numel(), row_group)
    mark("unpack1")
    return res


class ShardedSDFRenderer(object):
    """SDFRenderer over this rank's band + the all-gather.  `render` takes the arguments of SDFRenderer.render (default
    march: 'pyramid_recursive', as in the reference) and returns the LOCAL band; `gather` assembles full images."""

    def __init__(self, decoder, intrinsic, img_hw, rank=None, world_size=None, group=None, **kw):
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        rows = band(self.img_hw[0], self.rank, self.world)
        if rows[2] == 0:
            raise ValueError("rank %d of %d has no rows of a %d-row image (4-row groups)" % (self.rank, self.world, self.img_hw[0]))
        self.local = SDFRenderer(decoder, intrinsic, img_hw=img_hw, rows=rows, **kw)
        self.events = None      # set to {} to have gather() record CUDA events around pack / all-gather / unpack

    def render(self, latent, R, T, **kw):
        """Local band of render(): (depth[n_rows,W], normal[n_rows,W,3], mask, min_sdf).  The 'No valid depth' test
        (renderer.py:214) is deferred to gather(): a band without a live ray is legal as long as another band has one,
        and a rank raising alone would leave the others blocked in the collective."""
        kw.pop("check_empty", None)
        return self.local.render(latent, R, T, check_empty=False, **kw)

    def gather(self, outs, extra=None, check_empty=True):
        """(full-image maps, stacked extras or None).  Raises ValueError('No valid depth.') on EVERY rank when no ray of
        the whole image meets the unit sphere (reads 16 bytes per rank back, i.e. waits for the render)."""
        full, stats, extras = gather_bands(outs, self.img_hw, self.rank, self.world, stat=self.local._last_counts,
                                           extra=extra, group=self.group, events=self.events)
        if check_empty:
            st = stats.cpu()
            total = torch.stack([st[:, 0].sum(), st[:, 1].max(), st[:, 2].max(), st[:, 3].max()])[None]
            self.local._raise_if_empty(total)
        return full, extras
