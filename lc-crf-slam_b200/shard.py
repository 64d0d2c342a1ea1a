"""Multi-GPU sharding of independent CRF problems.

Each frame's CRF is an independent unit (own features, lattices and buffers; the reference
constructs and destroys the object per frame, src/Tracking.cc:1920-1958), so problems are
sharded across ranks with NO data-path collective: one process per GPU, each rank runs the
engine on its shard.  torch.distributed is used only for the barrier / max-over-ranks timing
in bench.py.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_contiguous(weights: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split problems [0, n) into `world` contiguous ranges with balanced total weight (sum of N_b).
    Returns [(begin, end)] per rank; every problem appears in exactly one range; ranges may be empty
    when there are fewer problems than ranks."""
    n = len(weights)
    total = float(sum(weights))
    out, begin, acc = [], 0, 0.0
    for r in range(world):
        target = total * (r + 1) / world
        end = begin
        while end < n and (acc + weights[end] / 2.0 <= target or n - end <= 0):
            acc += weights[end]
            end += 1
        if r == world - 1:
            end = n
        # leave at least one problem for each remaining rank when possible
        remaining_ranks = world - r - 1
        if n - end < remaining_ranks and end - begin > 1:
            give = min(end - begin - 1, remaining_ranks - (n - end))
            for _ in range(give):
                end -= 1
                acc -= weights[end]
        out.append((begin, end))
        begin = end
    return out


def shard_round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Sequences -> GPUs round-robin (config C5: 8 sequences over 1/2/4/8 GPUs)."""
    return list(range(rank, n_items, world))
