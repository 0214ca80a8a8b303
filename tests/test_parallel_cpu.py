"""CPU, gloo, world_size 2 and 3: the ray-tile sharding logic (band assignment, pack, ONE all-gather, unpack, gradient sum)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (puts the repo root on sys.path)

par = importlib.import_module("dist-renderer_b200.parallel")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_image(H, W):
    g = torch.Generator().manual_seed(3)
    depth = torch.rand(H, W, generator=g)
    normal = torch.rand(H, W, 3, generator=g)
    mask = (torch.rand(H, W, generator=g) > 0.5).to(torch.uint8)
    min_sdf = torch.rand(H, W, generator=g)
    return depth, normal, mask, min_sdf


def _band_rows(H, rank, world):
    """Image rows of `rank`'s band, spelled out independently of parallel.band: 4-row groups rank, rank+world, ..."""
    return [y for y in range(H) if (y // 4) % world == rank]


def _worker(rank, world, port, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _full_image(H, W)
    rows = _band_rows(H, rank, world)
    assert par.band(H, rank, world)[2] == len(rows)
    local = tuple(t[rows] for t in full)
    extra = torch.full((5,), float(rank + 1))
    stat = torch.tensor([[10 * rank + 1, rank, 7, 0]], dtype=torch.int32)
    outs, stats, extras = par.gather_bands(local, (H, W), rank, world, stat=stat, extra=extra)
    ok = all(torch.equal(a, b) for a, b in zip(outs, full))          # bit-exact, dtypes included (mask stays uint8)
    ok = ok and outs[2].dtype == torch.uint8 and extras.shape == (world, 5) and stats.shape == (world, 4)
    ok = ok and stats[:, 0].tolist() == [10 * r + 1 for r in range(world)]
    ok = ok and torch.equal(extras.sum(0), torch.full((5,), float(sum(range(1, world + 1)))))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


# even split (2 x 2 groups), partial last group (rows 16-17), an odd group count, three ranks with a ragged tail
@pytest.mark.parametrize("world,hw", [(2, (16, 7)), (2, (18, 5)), (2, (21, 3)), (3, (30, 4))])
def test_band_gather_multi_process(world, hw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, hw[0], hw[1], q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_band_single_process():
    H, W = 6, 4
    full = _full_image(H, W)
    outs, stats, extras = par.gather_bands(full, (H, W), 0, 1, extra=torch.ones(3))
    assert all(torch.equal(a, b) for a, b in zip(outs, full))
    assert extras.shape == (1, 3) and stats.shape == (1, 4)


def test_band_rows_match_renderer_row_mapping():
    """parallel.band's (row0, row_step, n_rows, row_group) means: local row l is image row
    row0 + (l // g) * row_step + l % g -- the mapping csrc/march.cu (global_row) and SDFRenderer._image_rows apply."""
    for H in (16, 18, 21, 30, 512, 37):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                row0, step, n_rows, g = par.band(H, rank, world)
                rows = [row0 + (l // g) * step + l % g for l in range(n_rows)]
                assert rows == _band_rows(H, rank, world)
                seen += rows
            assert sorted(seen) == list(range(H))
