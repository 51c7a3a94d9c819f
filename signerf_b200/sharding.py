"""Multi-GPU partition of the reference-sheet path.

Views are independent (reference datasetgenerator.py:517-519) and every dataset camera owns its own sheet (:331-338),
so there is no data-path reduction: with `world` ranks a step works on `world` grids; rank r renders the views
v = r (mod world) of EVERY grid, ONE all-gather of the packed tiles reassembles all sheets, and rank r keeps grid r
(paste + denoise).  The result is bit-identical to the single-process one (pure partition + gather)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

PACK_CHANNELS = 6  # rgb(3) | depth | condition | mask


def views_of_rank(num_views: int, world: int, rank: int) -> List[int]:
    """Round-robin view ownership: rank r renders views r, r + world, ..."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_views, world))


def pack_tiles(rgb: Tensor, depth: Tensor, cond: Tensor, mask: Tensor) -> Tensor:
    """[..,H,W,3],[..,H,W,1] x3 -> [..,H,W,6] fp32 (one buffer = one collective)."""
    return torch.cat([rgb, depth, cond, mask.to(rgb.dtype)], dim=-1).contiguous()


def unpack_tiles(t: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    return t[..., 0:3].contiguous(), t[..., 3:4].contiguous(), t[..., 4:5].contiguous(), t[..., 5:6].contiguous()


def gather_grids(packed_local: Tensor, num_views: int, world: int, rank: int, group: Optional[object] = None,
                 out: Optional[Tensor] = None) -> Tensor:
    """packed_local [G, V_local, H, W, 6] = this rank's views of all G grids (G == world in the benchmark)
    -> [G, num_views, H, W, 6] with view v of every grid taken from rank v % world.  One all_gather."""
    G, v_loc = packed_local.shape[0], packed_local.shape[1]
    if num_views % world != 0:
        raise ValueError("views must divide evenly over the ranks (pad the grid with repeated cameras otherwise)")
    if v_loc != num_views // world:
        raise ValueError(f"rank holds {v_loc} views per grid, expected {num_views // world}")
    if out is None:
        out = torch.empty((G, num_views) + tuple(packed_local.shape[2:]), dtype=packed_local.dtype, device=packed_local.device)
    if world == 1:
        out.copy_(packed_local)
        return out
    parts = [torch.empty_like(packed_local) for _ in range(world)]
    dist.all_gather(parts, packed_local.contiguous(), group=group)
    for r in range(world):
        out[:, r::world] = parts[r]
    return out


def all_gather_views(local: Tensor, num_views: int, world: int, rank: int, group: Optional[object] = None) -> Tensor:
    """local [len(views_of_rank(num_views, world, rank)), ...] -> [num_views, ...] on every rank, view v taken from rank
    v % world.  `num_views` need not divide by `world` (the reference sheet has rows * cols - 1 views): short ranks pad
    their shard to ceil(num_views / world) rows for the ONE all_gather and the padding is dropped afterwards."""
    mine = views_of_rank(num_views, world, rank)
    if local.shape[0] != len(mine):
        raise ValueError(f"rank {rank} holds {local.shape[0]} views, expected {len(mine)}")
    if world == 1:
        return local
    k = (num_views + world - 1) // world
    padded = local.new_zeros((k,) + tuple(local.shape[1:]))
    padded[:len(mine)] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = local.new_empty((num_views,) + tuple(local.shape[1:]))
    for r in range(world):
        n_r = len(range(r, num_views, world))
        out[r::world] = parts[r][:n_r]
    return out


def merge_frames(per_rank: List[List[Tuple[int, dict]]]) -> List[dict]:
    """[(global index, transforms.json frame)] lists of all ranks -> frames in index order (each index exactly once)."""
    flat = sorted((pair for part in per_rank for pair in part), key=lambda p: p[0])
    idx = [i for i, _ in flat]
    if idx != sorted(set(idx)):
        raise ValueError("two ranks produced the same frame index")
    return [f for _, f in flat]


class PeerTileExchange:
    """The same exchange without a collective: every rank owns a symmetric-memory tile buffer [num_views, H, W, 6] for
    ITS grid, and the producers store their packed tiles straight into the owner's buffer over NVLink
    (`sgn_scatter_tiles_peer`), bracketed by two device-side barriers of the symmetric-memory handle:
        barrier   - every rank has finished reading the tiles of the previous step
        scatter   - rank r writes its views of grid g into rank g's buffer        (all-to-all, 1/world of the all-gather bytes)
        barrier   - every store has landed
    Bit-identical to `gather_grids(...)[rank]` (a pure permutation of the same fp32 values)."""

    def __init__(self, world: int, rank: int, num_views: int, height: int, width: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        if num_views % world != 0:
            raise ValueError("views must divide evenly over the ranks")
        self.world, self.rank, self.num_views, self.h, self.w = world, rank, num_views, height, width
        self.tiles = symm_mem.empty((num_views, height, width, PACK_CHANNELS), dtype=torch.float32, device=device)
        grp = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.tiles, grp)
        self.peer_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)

    def exchange(self, rgb: Tensor, depth: Tensor, cond: Tensor, mask: Tensor) -> Tensor:
        """Inputs: this rank's views of all `world` grids, [world * v_loc, H, W, C] (grid-major).  Returns this rank's
        grid as packed tiles [num_views, H, W, 6] (valid until the next call)."""
        import ctypes as C
        from . import _lib
        v_loc = self.num_views // self.world
        if rgb.shape[0] != self.world * v_loc:
            raise ValueError(f"expected {self.world * v_loc} local views, got {rgb.shape[0]}")
        rgb, depth, cond = rgb.contiguous(), depth.contiguous(), cond.contiguous()
        mask = mask.to(torch.uint8).contiguous()
        self.handle.barrier(channel=0)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.load().sgn_scatter_tiles_peer(
                C.c_void_p(rgb.data_ptr()), C.c_void_p(depth.data_ptr()), C.c_void_p(cond.data_ptr()),
                C.c_void_p(mask.data_ptr()), self.world, v_loc, self.h, self.w, self.world, self.rank,
                C.c_void_p(self.peer_ptrs.data_ptr()), C.c_void_p(torch.cuda.current_stream(rgb.device).cuda_stream)))
        self.handle.barrier(channel=1)
        return self.tiles
