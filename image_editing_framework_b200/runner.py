"""Image-sharded sweeps over the GPUs of one box (the reference's `*/test.py` loops serially over PIE-Bench images,
e.g. p2p/test.py:114-181, hard-wired to one device).

Edits are independent, so item `i` goes to rank `i mod world`; every rank owns a full pipeline replica and there is NO
collective on the hot path. The only communication is the final gather of the (small) per-item results on the host side.
Launch one process per GPU with torchrun; without an initialised process group this degrades to a plain serial loop.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch.distributed as dist


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin: image i -> rank i mod G."""
    return list(range(rank, n_items, world_size))


def run_sharded(work: Callable[[int], Any], n_items: int, gather: bool = True) -> Dict[int, Any]:
    """Run `work(i)` for this rank's share of range(n_items). With gather=True every rank returns the full
    {index: result} map (results must be picklable and small: latents, timings, file names)."""
    rank, ws = world()
    mine = {i: work(i) for i in shard_indices(n_items, rank, ws)}
    if ws == 1 or not gather:
        return mine
    parts: List[Optional[Dict[int, Any]]] = [None] * ws
    dist.all_gather_object(parts, mine)
    merged: Dict[int, Any] = {}
    for p in parts:
        merged.update(p)
    return dict(sorted(merged.items()))
