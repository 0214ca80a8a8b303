"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sphere-tracing renderer.

Nothing in the product package imports this file.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg use it, as the checker or as the timed CPU baseline ("port").

It restates, in plain PyTorch tensor ops (the arithmetic of the reference *is* PyTorch fp32 -- nn.Linear / MKL
SGEMM, autograd, topk), the algorithm of B1ueber2y/DIST-Renderer for the path

    SDFRenderer.render / render_depth / render_normal      core/sdfrenderer/renderer.py:836-999
    ray_marching_{trivial,recursive,pyramid_recursive}      core/sdfrenderer/renderer.py:472-583, 713-805
    decode_sdf / decode_sdf_gradient                        core/utils/decoder_utils.py:53-92
    Decoder.inference (called through decoder.inference)    core/graph/deep_sdf_decoder.py:80-111

Pinning: the reference has no tests / golden vectors of its own.  This restatement is
pinned against the *reference itself*, executed here through oracle/ref_shim.py: tests/test_oracle_vs_reference.py
compares every output and gradient where /root/reference exists, and tests/golden/*.npz (written by
oracle/make_golden.py from the real reference) pin it on the GPU box where the reference tree is absent.

``dtype=torch.float64`` gives the fp64 twin used to measure each output's noise floor.
"""
import math

import numpy as np
import torch

MAX_POINTS = 100000  # decoder_utils.py:53 chunk size


# ----------------------------------------------------------------------------- decoder glue
def decode_sdf(decoder, latent, points, clamp_dist=0.1, no_grad=False):
    """decoder_utils.py:53-74 -- rows [latent | xyz] through decoder.inference, 100 K-row chunks, optional clamp.

    As in the reference, ``no_grad`` only detaches the result: the autograd graph is still built and dropped.
    """
    chunks = []
    for s in range(0, max(points.shape[0], 1), MAX_POINTS):
        pts = points[s:s + MAX_POINTS]
        x = pts if latent is None else torch.cat([latent.expand(pts.shape[0], -1), pts], 1)
        y = decoder.inference(x)
        chunks.append(y.detach() if no_grad else y)
    sdf = torch.cat(chunks, 0)
    if clamp_dist is not None:
        sdf = torch.clamp(sdf, -clamp_dist, clamp_dist)
    return sdf


def decode_sdf_gradient(decoder, latent, points, clamp_dist=0.1, no_grad=False):
    """decoder_utils.py:76-92 -- d clamp(sdf) / d xyz by autograd (create_graph=True).

    grad_outputs = ones_like(sdf): the reference passes ones_like(points) (N,3) for an (N,1) output, which
    torch >= 2 rejects; normalised normals are unaffected by the implied factor.
    """
    outs = []
    for s in range(0, max(points.shape[0], 1), MAX_POINTS):
        pts = points[s:s + MAX_POINTS]
        sdf = decode_sdf(decoder, latent, pts, clamp_dist=clamp_dist)
        g = torch.autograd.grad(sdf, pts, grad_outputs=torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
        outs.append(g.detach() if no_grad else g)
    return torch.cat(outs, 0)


# ----------------------------------------------------------------------------- renderer
def _unit(v):
    """renderer.py:171-178 -- column-normalise with eps added to the norm."""
    return v / (torch.norm(v, p=2, dim=0).expand_as(v) + 1e-12)


def _put(base, mask, src):
    """Out-of-place masked write (renderer.py:284-302 copy_index)."""
    idx = torch.nonzero(mask.reshape(-1)).reshape(-1)
    return base.index_copy(0, idx, src)


def depth2normal(depth, f_pix_x, f_pix_y=None):
    """Normal map (H, W, 3) from central differences of a depth map (H, W) -- core/utils/render_utils.py:9-43.
    Background (depth > 1e5 or == 0) is zeroed IN PLACE in `depth` (:24-25) and gets a zero normal (:42); the shifted
    copies leave a one-pixel border of zeros (:27-34)."""
    f_pix_y = f_pix_x if f_pix_y is None else f_pix_y
    h, w = depth.shape
    bg = (depth > 1e5) | (depth == 0)
    depth[bg] = 0.0
    left, right, up, down = (torch.zeros(h, w, dtype=depth.dtype) for _ in range(4))
    left[:, 1:w - 1] = depth[:, :w - 2].clone()
    right[:, 1:w - 1] = depth[:, 2:].clone()
    up[1:h - 1, :] = depth[:h - 2, :].clone()
    down[1:h - 1, :] = depth[2:, :].clone()
    dzdx = (right - left) * f_pix_x / 2.0                                            # :36
    dzdy = (down - up) * f_pix_y / 2.0                                               # :37
    n = torch.stack([dzdx, dzdy, -torch.ones_like(dzdx)]).permute(1, 2, 0)           # :39
    n = n / (torch.norm(n, p=2, dim=2) + 1e-12)[:, :, None]                          # :40-41
    n[bg] = 0.0
    return n


class OracleSDFRenderer(object):
    def __init__(self, decoder, intrinsic, img_hw=None, transform_matrix=None, march_step=50, buffer_size=5,
                 ray_marching_ratio=1.5, radius=1.0, threshold=5e-5, scale_list=(4, 2, 1),
                 march_step_list=(3, 3, -1), dtype=torch.float32, use_depth2normal=False):
        # renderer.py:13-59
        self.decoder = decoder
        self.dtype = dtype
        self.use_depth2normal = use_depth2normal
        self.march_step, self.buffer_size = march_step, buffer_size
        self.ratio, self.radius, self.threshold = ray_marching_ratio, radius, threshold
        self.scale_list, self.march_step_list = list(scale_list), list(march_step_list)
        intrinsic = np.asarray(intrinsic, dtype=np.float64)
        if img_hw is None:
            img_hw = (int(intrinsic[1, 2] * 2), int(intrinsic[0, 2] * 2))
        self.img_hw = tuple(int(v) for v in img_hw)
        h, w = self.img_hw
        self.K = torch.from_numpy(intrinsic).to(dtype)
        self.K_inv = torch.from_numpy(np.linalg.inv(intrinsic)).to(dtype)
        self.grid = self._pixel_grid(h, w)                          # (h, w, 2) as (x, y)
        self.homo_2d = self._homo(self.grid).reshape(-1, 3).t()     # (3, P)
        self.homo_calib = self.K_inv @ self.homo_2d
        self.calib_map = _unit(self.homo_calib)[2]
        if transform_matrix is None:
            transform_matrix = np.array([[1., 0., 0.], [0., 0., -1.], [0., 1., 0.]])
        self.M = torch.from_numpy(np.asarray(transform_matrix, dtype=np.float64)).to(dtype)

    # --- camera / geometry -------------------------------------------------------------------------------
    def _pixel_grid(self, h, w):
        ys, xs = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
        return torch.stack([xs, ys], 2).to(self.dtype)

    def _homo(self, xy):
        return torch.cat([xy, torch.ones(xy.shape[0], xy.shape[1], 1, dtype=self.dtype)], 2)

    def camera_location(self, R, T):                                # renderer.py:180-188
        return torch.matmul(-R.t(), T[:, None]).squeeze(1)

    def camera_rays(self, R, homo=None):                            # renderer.py:190-200
        return _unit(torch.matmul(R.t(), self.homo_calib if homo is None else homo))

    def points_on_rays(self, c, rays, depth, inv_transform=True):   # renderer.py:202-223
        depth = depth.detach()
        if depth.shape[0] == 0:
            raise ValueError('No valid depth.')
        pts = rays * depth[None, :] + c[:, None]
        if inv_transform:
            pts = torch.matmul(self.M.t(), pts)                     # renderer.py:119
        if not pts.requires_grad:
            pts.requires_grad = True
        return pts

    def distance_from_origin(self, c, rays):                        # renderer.py:225-239
        ptq = (c[:, None] * rays).sum(0)
        return torch.norm(c[:, None] - ptq[None, :] * rays, p=2, dim=0)

    def chord_from_distance(self, d):                               # renderer.py:241-252
        with torch.no_grad():
            v = self.radius ** 2 - d ** 2
            ok = v >= 0
            chord = torch.zeros_like(d)
            chord[ok] = 2 * torch.sqrt(v[ok])
        return chord

    def sphere_entry(self, c, rays):                                # renderer.py:254-273
        with torch.no_grad():
            d = self.distance_from_origin(c, rays)
            hit = d <= self.radius
            chord = self.chord_from_distance(d)
            cdist = torch.sqrt((c ** 2).sum())
            if bool(cdist < self.radius):
                entry = torch.zeros_like(d)
            else:
                e_hit = torch.sqrt(cdist ** 2 - d[hit] ** 2) - chord[hit] / 2.0
                entry = _put(torch.ones_like(d) * e_hit.max(), hit, e_hit)
        return entry, hit

    def sphere_exit(self, c, rays):                                 # renderer.py:275-282
        with torch.no_grad():
            entry, _ = self.sphere_entry(c, rays)
            return entry + self.chord_from_distance(self.distance_from_origin(c, rays))

    # --- marching ----------------------------------------------------------------------------------------
    def _query(self, latent, pts, no_grad):
        return decode_sdf(self.decoder, latent, pts.t(), clamp_dist=None, no_grad=no_grad).squeeze(-1)

    def march_trivial(self, c, rays, entry, hit, latent, steps, clamp_dist, no_grad, use_transform=True):
        # renderer.py:472-510
        rays_h, entry_h = rays[:, hit], entry[hit]
        z = torch.zeros_like(entry_h)
        zs, sdfs, pts_l = [], [], []
        for _ in range(steps):
            pts = self.points_on_rays(c, rays_h, entry_h + z, inv_transform=use_transform)
            sdf = self._query(latent, pts, no_grad).detach()
            pts_l.append(pts.t()[None])
            sdfs.append(sdf[None])
            z = z + torch.clamp(sdf, -clamp_dist, clamp_dist) * self.ratio
            zs.append(z[None])
        zs, sdfs, pts_l = torch.cat(zs, 0), torch.cat(sdfs, 0), torch.cat(pts_l, 0)
        ok = (zs[-1] + entry_h < self.sphere_exit(c, rays_h)) & (torch.abs(sdfs).min(0)[0] <= self.threshold) \
            & (sdfs[0] > self.threshold)
        return sdfs, zs, pts_l, ok

    def march_recursive(self, c, rays, entry, hit, latent, steps, clamp_dist, no_grad, use_transform=True,
                        first_query_check=True):
        # renderer.py:512-583
        rays_h, entry_h = rays[:, hit], entry[hit]
        exit_h = self.sphere_exit(c, rays_h)
        z = torch.zeros_like(entry_h)
        live = (z + entry_h < exit_h)
        zs, sdfs, pts_l = [], [], []
        for _ in range(steps):
            pts_now = self.points_on_rays(c, rays_h[:, live], entry_h[live] + z[live], inv_transform=use_transform)
            if no_grad:
                pts_now = pts_now.detach()
            sdf_now = self._query(latent, pts_now, no_grad)
            pts = torch.zeros_like(z)[:, None].repeat(1, 3)
            pts[live, :] = pts_now.t()
            if no_grad:
                pts = pts.detach()
            pts_l.append(pts[None])
            sdf = torch.zeros_like(z)
            sdf[live] = sdf_now.detach()
            z = z + torch.clamp(sdf, -clamp_dist, clamp_dist) * self.ratio
            zs.append(z[None])
            sdf[~live] = 1.0
            sdfs.append(sdf[None])
            live = live & (z + entry_h < exit_h) & (torch.abs(sdf) >= self.threshold)
            if int(live.sum()) == 0:
                while len(zs) < self.buffer_size:   # renderer.py:562-567 pad with copies of the last step
                    zs.append(z[None]); sdfs.append(sdf[None]); pts_l.append(pts[None])
                break
        zs, sdfs, pts_l = torch.cat(zs, 0), torch.cat(sdfs, 0), torch.cat(pts_l, 0)
        ok = (zs[-1] + entry_h < exit_h) & (torch.abs(sdfs).min(0)[0] <= self.threshold)
        if first_query_check:
            ok = ok & (sdfs[0] > self.threshold)
        return sdfs, zs, pts_l, ok

    def _coarser_level(self, grid, R):
        """renderer.py:604-666 with scale 2: returns (coarse grid, coarse rays, fine->coarse flat index)."""
        h, w = grid.shape[0], grid.shape[1]
        stride = grid[0, 1, 0] - grid[0, 0, 0]
        nh, nw = int(np.ceil(h / 2.0)), int(np.ceil(w / 2.0))
        coarse = (2.0 * stride) * self._pixel_grid(nh, nw) + ((2.0 * stride) - 1) / 2
        idx_grid = grid if stride == 1 else self._pixel_grid(h, w)
        imap = torch.ceil((idx_grid + 1) / 2.0) - 1
        imap = (imap[:, :, 0] + imap[:, :, 1] * nw).reshape(-1).long()
        homo = self.K_inv @ self._homo(coarse).reshape(-1, 3).t()
        return coarse, self.camera_rays(R, homo=homo), imap

    def march_pyramid(self, c, R, hit, latent, clamp_dist, no_grad, use_transform=True):
        # renderer.py:713-805 (split_type='raydepth' -> recalibration map forced to ones, :740-741)
        steps = list(self.march_step_list)
        if steps[-1] == -1:
            steps[-1] = self.march_step - sum(steps[:-1])
        assert self.scale_list[-1] == 1
        scales, steps = self.scale_list[::-1], steps[::-1]           # fine -> coarse
        grids, rays_l, imaps, hits = [self.grid], [self.camera_rays(R)], [None], [hit]
        for li in range(1, len(scales)):
            assert scales[li] / scales[li - 1] == 2
            g, r, im = self._coarser_level(grids[-1], R)
            pooled = torch.zeros(g.shape[0] * g.shape[1], dtype=torch.uint8).scatter_reduce(
                0, im, hits[-1].to(torch.uint8), "amax").bool()      # scatter_max, renderer.py:668-680
            grids.append(g); rays_l.append(r); imaps.append(im); hits.append(pooled)
        entry_coarse, _ = self.sphere_entry(c, rays_l[-1])
        entry_full, _ = self.sphere_entry(c, rays_l[0])
        h_sdf = h_z = h_pts = None
        for li in range(len(scales) - 1, -1, -1):
            start = entry_coarse if li == len(scales) - 1 else h_z[-1]
            if li != 0:
                s, zz, pp, _ = self.march_trivial(c, rays_l[li], start, hits[li], latent, steps[li], clamp_dist,
                                                  no_grad, use_transform)
                n = hits[li].shape[0]
                s_f = torch.ones(s.shape[0], n, dtype=self.dtype); s_f[:, hits[li]] = s
                z_f = torch.zeros(s.shape[0], n, dtype=self.dtype); z_f[:, hits[li]] = zz
                p_f = torch.zeros(s.shape[0], n, 3, dtype=self.dtype); p_f[:, hits[li], :] = pp
                z_f = z_f + start
                if h_sdf is not None:
                    s_f, p_f, z_f = torch.cat([h_sdf, s_f], 0), torch.cat([h_pts, p_f], 0), torch.cat([h_z, z_f], 0)
                im = imaps[li]
                h_sdf, h_pts, h_z = s_f[:, im], p_f[:, im, :], z_f[:, im] * 1.0
            else:
                s, zz, pp, ok = self.march_recursive(c, rays_l[0], start, hit, latent, steps[0], clamp_dist, no_grad,
                                                     use_transform, first_query_check=False)
                h_sdf = torch.cat([h_sdf[:, hit], s], 0)
                h_pts = torch.cat([h_pts[:, hit, :], pp], 0)
                h_z = torch.cat([h_z[:, hit], zz + start[hit]], 0)
        return h_sdf, h_z - entry_full[hit][None, :], h_pts, ok

    def march(self, c, R, entry, hit, latent, clamp_dist, no_grad, kind, use_transform=True):  # renderer.py:807-834
        if kind == 'trivial':
            return self.march_trivial(c, self.camera_rays(R), entry, hit, latent, self.march_step, clamp_dist,
                                      no_grad, use_transform)
        if kind == 'recursive':
            return self.march_recursive(c, self.camera_rays(R), entry, hit, latent, self.march_step, clamp_dist,
                                        no_grad, use_transform)
        if kind == 'pyramid_recursive':
            return self.march_pyramid(c, R, hit, latent, clamp_dist, no_grad, use_transform)
        raise ValueError('Error! Invalid type of ray marching: {}.'.format(kind))

    # --- selection + differentiable re-queries -------------------------------------------------------------
    @staticmethod
    def _gather(data, index):                                       # renderer.py:343-362
        k, n = index.shape[1], index.shape[0]
        flat = index.t().reshape(-1) * data.shape[1] + torch.arange(n).repeat(k)
        if data.dim() == 3:
            return data.reshape(-1, data.shape[-1])[flat].reshape(k, n, data.shape[-1]).clone()
        return data.reshape(-1)[flat].reshape(k, n).clone()

    def _topk_min_abs(self, sdfs, k):                               # renderer.py:316-319
        _, index = torch.topk(-torch.abs(sdfs).t(), k, dim=1)
        return self._gather(sdfs, index), index

    def render_depth(self, latent, R, T, clamp_dist=0.1, no_grad=False, no_grad_depth=False, no_grad_mask=False,
                     no_grad_camera=False, ray_marching_type='recursive', use_transform=True):
        # renderer.py:836-878
        if no_grad:
            no_grad_depth = no_grad_mask = no_grad_camera = True
        c = self.camera_location(R, T)
        rays = self.camera_rays(R)
        dist = self.distance_from_origin(c, rays)
        entry, hit = self.sphere_entry(c, rays)
        sdfs, zs, pts_l, ok = self.march(c, R, entry, hit, latent, clamp_dist, no_grad_camera, ray_marching_type,
                                         use_transform)
        # renderer.py:382-390 -- unclamped sdf at each ray's min-|sdf| sample, with grad
        _, i1 = self._topk_min_abs(sdfs, 1)
        min_sdf = decode_sdf(self.decoder, latent, self._gather(pts_l, i1)[0], clamp_dist=None,
                             no_grad=no_grad_mask).squeeze(-1)
        if no_grad_mask:
            min_sdf = min_sdf.detach()
        # renderer.py:392-420 -- depth estimate + value-neutral gradient carriers
        sel_sdf, ik = self._topk_min_abs(sdfs, self.buffer_size)
        sel_pts = self._gather(pts_l, ik)
        z = self._gather(zs, ik[:, [0]])[0]
        z = z + (1 - self.ratio) * torch.clamp(sel_sdf[0, :], -clamp_dist, clamp_dist)
        if not no_grad_depth:
            for i in range(self.buffer_size):
                s = decode_sdf(self.decoder, latent, sel_pts[i], clamp_dist=clamp_dist).squeeze(-1)
                z = z - s.detach() * self.ratio
                z = z + s * self.ratio
        # renderer.py:859-878 -- scatter to the image
        P = hit.shape[0]
        min_map = _put(torch.zeros(P, dtype=self.dtype), hit, min_sdf)
        min_map = _put(min_map, ~hit, dist[~hit] + self.threshold - self.radius)
        Zdepth = _put(torch.ones(P, dtype=self.dtype) * 1e11, hit, entry[hit] + z)
        mask = hit.clone()
        mask[hit.clone()] = ok
        if no_grad_depth:
            Zdepth = Zdepth.detach()
        return Zdepth, mask, min_map

    def render_normal(self, latent, R, T, Zdepth, valid_mask, clamp_dist=0.1, no_grad=False, normalize=True,
                      use_transform=True):
        # renderer.py:880-910
        c = self.camera_location(R, T)
        rays = self.camera_rays(R)
        P = valid_mask.shape[0]
        out = torch.zeros(P, 3, dtype=self.dtype)
        if int(valid_mask.sum()) == 0:
            return out.t()
        pts = self.points_on_rays(c, rays[:, valid_mask], Zdepth[valid_mask], inv_transform=use_transform)
        g = decode_sdf_gradient(self.decoder, latent, pts.t(), clamp_dist=clamp_dist, no_grad=no_grad).t()
        n = _unit(g) if normalize else g
        n = torch.matmul(self.M, n)                                  # renderer.py:97 (3x3 transform only)
        out = _put(out, valid_mask, n.t()).t()
        return out.detach() if no_grad else out

    def render(self, latent, R, T, clamp_dist=0.1, no_grad=False, no_grad_depth=False, no_grad_normal=False,
               no_grad_mask=False, no_grad_camera=False, normalize_normal=True, use_transform=True,
               ray_marching_type='pyramid_recursive'):
        # renderer.py:943-999
        if no_grad:
            no_grad_depth = no_grad_normal = no_grad_mask = no_grad_camera = True
        h, w = self.img_hw
        Zdepth, mask, min_map = self.render_depth(latent, R, T, clamp_dist=clamp_dist, no_grad=no_grad,
                                                  no_grad_depth=no_grad_depth, no_grad_mask=no_grad_mask,
                                                  no_grad_camera=no_grad_camera,
                                                  ray_marching_type=ray_marching_type, use_transform=use_transform)
        depth = torch.ones_like(Zdepth) * 1e11
        depth[mask] = Zdepth[mask].clone() * self.calib_map[mask]
        if self.use_depth2normal:      # renderer.py:972-975
            depth = depth.reshape(h, w)
            normal = depth2normal(depth, np.float32(self.K[0, 0].item()), np.float32(self.K[1, 1].item()))
            return depth, normal, mask.reshape(h, w).to(torch.uint8), min_map.reshape(h, w)
        normal = self.render_normal(latent, R, T, Zdepth, mask, clamp_dist=clamp_dist, no_grad=no_grad_normal,
                                    normalize=normalize_normal, use_transform=use_transform)
        normal = torch.matmul(R, normal)
        normal = torch.cat([normal[:1] * (-1), normal[1:]], 0)       # renderer.py:979 flip x
        return (depth.reshape(h, w), normal.reshape(3, h, w).permute(1, 2, 0), mask.reshape(h, w).to(torch.uint8),
                min_map.reshape(h, w))

    def render_silhouette(self, latent, R, T, **kw):
        """(mask, min_abs_query) of render_depth -- the pair the reference calls the silhouette (renderer.py:878)."""
        _, mask, min_map = self.render_depth(latent, R, T, **kw)
        h, w = self.img_hw
        return mask.reshape(h, w).to(torch.uint8), min_map.reshape(h, w)
