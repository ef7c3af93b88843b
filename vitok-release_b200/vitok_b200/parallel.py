"""Host-side data-parallel helpers: how images are dealt to ranks.

The path shards by image with no data-path collective (the reference launches one process per GPU,
vitok/utils.py:50-60, and its DataLoader gives every rank its own images).  For fixed-size batches a contiguous split is
balanced; for NaFlex batches the cost of an image is roughly linear in its token count for the GEMMs plus quadratic for
attention, and -- because masked batches run on their valid tokens only (token packing) -- a rank's time follows the sum
over ITS images, so ragged batches are dealt by estimated cost (longest-processing-time greedy).
"""
from __future__ import annotations

from typing import List, Sequence


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of ``range(n_items)``: the first ``n_items % world`` ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def image_cost(tokens: int, width: int = 1024, attn_weight: float = 1.0) -> float:
    """Relative cost of one image: per layer 2 n D (4D + 3 Hf) ~ 24 n D^2 of GEMM FLOPs + 4 n^2 D of attention FLOPs."""
    return 24.0 * tokens * width * width + attn_weight * 4.0 * tokens * tokens * width


def shard_by_tokens(token_counts: Sequence[int], world: int, width: int = 1024) -> List[List[int]]:
    """Deal images to ``world`` ranks so that the estimated cost per rank is balanced (LPT greedy: most expensive image first,
    always to the least loaded rank).  Returns one list of image indices per rank; every index appears exactly once.  The
    result is deterministic, so every rank can compute it locally from the same size list -- no collective."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(token_counts)), key=lambda i: (-image_cost(token_counts[i], width), i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += image_cost(token_counts[i], width)
    for lst in out:
        lst.sort()
    return out


__all__ = ["shard_range", "shard_by_tokens", "image_cost"]
