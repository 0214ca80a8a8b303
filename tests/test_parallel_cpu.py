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


def _worker(rank, world, port, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _full_image(H, W)
    row0, step, n_rows = par.band(H, rank, world)
    assert n_rows == len(range(rank, H, world))
    local = tuple(t[row0::step] for t in full)
    extra = torch.full((5,), float(rank + 1))
    outs, extras = par.gather_bands(local, (H, W), rank, world, extra=extra)
    ok = all(torch.equal(a.float(), b.float()) for a, b in zip(outs, full))
    ok = ok and outs[2].dtype == torch.uint8 and extras.shape == (world, 5)
    ok = ok and torch.equal(extras.sum(0), torch.full((5,), float(sum(range(1, world + 1)))))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


# even split, ragged split (5 + 4 rows), three ranks with a ragged tail (4 + 3 + 3 rows)
@pytest.mark.parametrize("world,hw", [(2, (10, 7)), (2, (9, 5)), (3, (10, 4))])
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
    outs, extras = par.gather_bands(full, (H, W), 0, 1, extra=torch.ones(3))
    assert all(torch.equal(a.float(), b.float()) for a, b in zip(outs, full))
    assert extras.shape == (1, 3)
