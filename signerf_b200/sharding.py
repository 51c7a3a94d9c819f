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
