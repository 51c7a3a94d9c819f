"""2 GPUs: the fused NVLink peer-store exchange (sgn_scatter_tiles_peer over torch symmetric memory) delivers exactly
the tiles the NCCL all-gather path delivers.  Skipped on boxes with fewer than two GPUs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

V, H, W = 4, 24, 40


def _worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist
    from signerf_b200 import sharding as SH
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        mine = SH.views_of_rank(V, world, rank)
        parts = []
        for g in range(world):
            for v in mine:
                gen = torch.Generator().manual_seed(100 * g + v)
                parts.append((torch.rand(H, W, 3, generator=gen), torch.rand(H, W, 1, generator=gen),
                              torch.rand(H, W, 1, generator=gen), (torch.rand(H, W, 1, generator=gen) > 0.5)))
        rgb = torch.stack([p[0] for p in parts]).to(dev)
        depth = torch.stack([p[1] for p in parts]).to(dev)
        cond = torch.stack([p[2] for p in parts]).to(dev)
        mask = torch.stack([p[3] for p in parts]).to(dev).to(torch.uint8)
        packed = SH.pack_tiles(rgb, depth, cond, mask).view(world, len(mine), H, W, SH.PACK_CHANNELS)
        ref = SH.gather_grids(packed, V, world, rank)[rank]
        ex = SH.PeerTileExchange(world, rank, V, H, W, dev)
        ok = True
        for _ in range(3):   # repeated steps reuse the buffers: the leading barrier protects the readers
            got = ex.exchange(rgb, depth, cond, mask)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(got, ref))
        q.put((rank, ok, ""))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, f"{type(e).__name__}: {e}"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_store_exchange_equals_nccl_all_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
