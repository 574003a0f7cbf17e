"""Multi-GPU sharding of independent resample jobs: no data-path collective.

Every image -- and every cascade chain of (image, output size) pairs, which must stay on one GPU because later
sizes are resampled from earlier results (imageflow_tool self_test.rs:184-198 export_4_sizes) -- is independent.
Uniform batches are split into contiguous blocks; mixed workloads are binned by greedy LPT on input pixels.
The only cross-rank step is the host-side reduction of per-rank pixel counts and the max of the elapsed times.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_contiguous(n_items: int, world: int, rank: int) -> range:
    """Contiguous block of a uniform batch for `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_lpt(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time binning: returns, per rank, the indices of the chains it owns.
    costs[i] = cost of chain i (input pixels of every resample in the chain)."""
    if world <= 0:
        raise ValueError("bad world")
    bins: List[List[int]] = [[] for _ in range(world)]
    load = [0.0] * world
    for i in sorted(range(len(costs)), key=lambda k: (-costs[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += costs[i]
    for b in bins:
        b.sort()
    return bins


def lpt_imbalance(costs: Sequence[float], bins: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(costs[i] for i in b) for b in bins]
    mean = sum(loads) / len(loads) if loads else 0.0
    return max(loads) / mean if mean > 0 else 1.0


def aggregate(local_units: float, local_ms: float, device=None) -> Tuple[float, float]:
    """(sum of units over ranks, max of elapsed ms over ranks) -- the only collective on this path, and it is
    host bookkeeping (one 2-element all_reduce per measurement), not part of the data path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_units), float(local_ms)
    s = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    m = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(s.item()), float(m.item())


def constrain_within(w: int, h: int, box: int) -> Tuple[int, int]:
    """Constrain::Within(box, box) sizing without upscaling (imageflow_riapi sizing as used by self_test.rs:184-198):
    the longer edge is limited to `box`, the aspect ratio kept, each edge rounded to the nearest pixel (>= 1)."""
    if w <= box and h <= box:
        return w, h
    if w >= h:
        return box, max(1, int(h * box / w + 0.5))
    return max(1, int(w * box / h + 0.5)), box


def export_4_sizes_chain(w: int, h: int) -> List[Tuple[Tuple[int, int], Tuple[int, int]]]:
    """The cascade of self_test.rs:184-198 as (source size, target size) resamples:
    src->1600, 1600->1200, 1200->400, 1600->800.  Steps that would not change the size are dropped
    (resample_when default: the node deletes itself, flow/nodes/scale_render.rs:113-115)."""
    s1600 = constrain_within(w, h, 1600)
    s1200 = constrain_within(*s1600, 1200)
    s400 = constrain_within(*s1200, 400)
    s800 = constrain_within(*s1600, 800)
    steps = [((w, h), s1600), (s1600, s1200), (s1200, s400), (s1600, s800)]
    return [(a, b) for a, b in steps if a != b]
