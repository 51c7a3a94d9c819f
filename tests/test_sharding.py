"""CPU, 2 gloo ranks: the per-view shard + single all-gather reassembles every grid bit-identically to the
single-process layout (pure partition, no reduction)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from signerf_b200 import sharding as SH

V, H, W = 6, 5, 7


def _tile(grid: int, view: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 * grid + view)
    rgb, depth, cond = torch.rand(H, W, 3, generator=g), torch.rand(H, W, 1, generator=g), torch.rand(H, W, 1, generator=g)
    mask = torch.rand(H, W, 1, generator=g) > 0.5
    return SH.pack_tiles(rgb, depth, cond, mask)


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = SH.views_of_rank(V, world, rank)
    local = torch.stack([torch.stack([_tile(g, v) for v in mine]) for g in range(world)])
    full = SH.gather_grids(local, V, world, rank)
    expect = torch.stack([torch.stack([_tile(g, v) for v in range(V)]) for g in range(world)])
    q.put((rank, bool(torch.equal(full, expect)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_view_shard_all_gather_is_bit_identical(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (world, V, H, W, SH.PACK_CHANNELS) for _, _, shape in res)


def test_partition_helpers():
    assert SH.views_of_rank(16, 4, 1) == [1, 5, 9, 13]
    assert sorted(sum((SH.views_of_rank(16, 8, r) for r in range(8)), [])) == list(range(16))
    with pytest.raises(ValueError):
        SH.views_of_rank(16, 4, 4)
    t = _tile(0, 0)
    rgb, depth, cond, mask = SH.unpack_tiles(t)
    assert torch.equal(SH.pack_tiles(rgb, depth, cond, mask), t)
    one = SH.gather_grids(t[None, None], 1, 1, 0)
    assert torch.equal(one[0, 0], t)
    with pytest.raises(ValueError):
        SH.gather_grids(torch.zeros(2, 3, H, W, 6), 7, 2, 0)
