"""Multi-GPU: one process per GPU, images sharded by contiguous row blocks.  The path has no cross-image
dependency (model.py:105-169) so there is NO data-path collective; the only communication is an optional
all-gather of the decoded token ids (int32 [B/G, L], ~53 KB per GPU at B=4096) and, for the reference's
`[B, S, C]` early-exit shape contract, a MAX all-reduce of the per-rank step count S (model.py:144)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_rows(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of `total` rows owned by `rank`; blocks differ by at most one row."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_ids(local_ids: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather per-rank id blocks [n_r, L] into [total, L] in global row order (ragged-safe)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ids
    world, rank = dist.get_world_size(), dist.get_rank()
    L = local_ids.shape[1]
    sizes = [shard_rows(total, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width, L), dtype=local_ids.dtype, device=local_ids.device)
    pad[: local_ids.shape[0]] = local_ids
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def global_steps(local_steps: int, device=None) -> int:
    """S of the whole (sharded) batch = max over ranks (every row must have produced EOS, model.py:144)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(local_steps)
    t = torch.tensor([int(local_steps)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
