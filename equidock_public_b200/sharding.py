"""Data-parallel sharding of protein pairs across ranks: pairs are independent units, so a batch is
cut into contiguous per-rank slices balanced by estimated cost, every rank runs the engine on its slice, and the
only communication is outside the data path (a barrier and a max-reduction of the timed region in bench.py; the
gradient all-reduce of training arrives with the backward kernels)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def pair_cost(n_lig: int, n_rec: int, e_lig: int, e_rec: int, n_layers: int = 8) -> float:
    """Per-pair forward cost model (FLOPs): edges, nodes and the N_l x N_r attention term."""
    return n_layers * (38.3e3 * (e_lig + e_rec) + 66e3 * (n_lig + n_rec) + 512.0 * n_lig * n_rec)


def shard_bounds(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) slices of the pair list, one per rank, with near-equal summed cost: pair i goes to the rank
    whose share of the cumulative cost contains the pair's midpoint.  Every pair belongs to exactly one rank; a rank
    is empty only when there are fewer pairs than ranks (or one pair dwarfs the rest)."""
    n = len(costs)
    total = float(sum(costs))
    owners, cum = [], 0.0
    for c in costs:
        mid = cum + 0.5 * c
        owners.append(min(world - 1, int(mid * world / total)) if total > 0 else 0)
        cum += c
    bounds, lo = [], 0
    for r in range(world):
        hi = lo
        while hi < n and owners[hi] <= r:
            hi += 1
        bounds.append((lo, hi))
        lo = hi
    return bounds


def my_shard(costs: Sequence[float], world: int, rank: int) -> Tuple[int, int]:
    return shard_bounds(costs, world)[rank]
