"""Box sharding of the chemistry advance over the GPUs of one node, and the single statistics reduction.

The reference distributes work in units of AMReX boxes: ``BoxArray::maxSize(max_grid_size)`` chops the
domain (PP/Testing/Exec/ReactEval/main.cpp:198-203; PeleC ``amr.max_grid_size``) and a
``DistributionMapping`` hands boxes to ranks, by default with the cost-weighted KNAPSACK strategy
(amrex/Src/Base/AMReX_DistributionMapping.H:48,138-146) whose weights PeleC takes from the previous
step's ``Work_Estimate`` (= the reactor's ``FC`` output, PC/Source/React.cpp:316-322).

Chemistry needs no exchange between cells, so there is NO collective on the data path: every rank reacts
its own boxes on its own GPU.  The only communication is one all-reduce of a 12-number statistics vector
(sums of the work counters, max of the step count and of the device time) at the end of a call/benchmark:
NCCL on the GPUs, gloo in the CPU tests (tests/test_sharding.py, world_size 2).
"""
from typing import Dict, List, Sequence, Tuple

BoxT = Tuple[Tuple[int, int, int], Tuple[int, int, int]]  # (lo, hi), inclusive, like amrex::Box

SUM_KEYS = ("cells", "cells_masked", "n_steps", "n_rhs", "n_jac", "n_setups", "n_newton", "n_errfail",
            "n_convfail", "n_failed")
MAX_KEYS = ("max_steps", "device_ms")


def chop_domain(ncell: Sequence[int], max_grid_size: int) -> List[BoxT]:
    """BoxArray(domain).maxSize(max_grid_size) = BoxList::maxSize (amrex/Src/Base/AMReX_BoxList.cpp:765-820), per
    direction: while the chunk and the length are both even, halve both (the pieces stay aligned to that power of two);
    cut the reduced length nlen into numblk = ceil(nlen/bs) pieces of nlen // numblk cells, the first nlen % numblk one cell
    longer; scale back.  Length 10 / chunk 4 -> 4, 4, 2; length 100 / chunk 32 -> 28, 24, 24, 24.  Boxes are returned in
    AMReX's order (x fastest)."""
    cuts = []
    for n in ncell:
        if n <= 0 or max_grid_size <= 0:
            raise ValueError("chop_domain: sizes must be positive")
        if n <= max_grid_size:
            cuts.append([(0, n - 1)])
            continue
        ratio, bs, nlen = 1, max_grid_size, n
        while bs % 2 == 0 and nlen % 2 == 0:
            ratio, bs, nlen = ratio * 2, bs // 2, nlen // 2
        numblk = (nlen + bs - 1) // bs
        sz = nlen // numblk
        extra = nlen - sz * numblk
        edges = []
        for b in range(numblk):
            lo = b * (sz + 1) * ratio if b < extra else (b * sz + extra) * ratio
            hi = lo + ((sz + 1) if b < extra else sz) * ratio - 1
            edges.append((lo, hi))
        cuts.append(edges)
    return [((x[0], y[0], z[0]), (x[1], y[1], z[1])) for z in cuts[2] for y in cuts[1] for x in cuts[0]]


def box_cells(b: BoxT) -> int:
    lo, hi = b
    return (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)


def knapsack(weights: Sequence[float], nranks: int) -> List[int]:
    """Greedy longest-processing-time knapsack: heaviest box first onto the least-loaded rank (ties: lowest
    rank, lowest box index) - deterministic on every rank, so no broadcast of the map is needed.
    Returns owner[box]."""
    if nranks <= 0:
        raise ValueError("knapsack: nranks must be positive")
    load = [0.0] * nranks
    owner = [0] * len(weights)
    for i in sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i)):
        r = min(range(nranks), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += float(weights[i])
    return owner


def my_boxes(boxes: Sequence[BoxT], rank: int, nranks: int, weights: Sequence[float] = None) -> List[int]:
    """Indices of the boxes `rank` reacts.  Weights default to the cell count (first step); afterwards pass
    the per-box sum of the previous call's FC (the reference's Work_Estimate)."""
    w = [box_cells(b) for b in boxes] if weights is None else list(weights)
    own = knapsack(w, nranks)
    return [i for i, r in enumerate(own) if r == rank]


def reduce_stats(local: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """The one collective of the path: all-reduce (sum for counters, max for max_steps / device_ms) of the
    statistics vector.  Without an initialised process group this is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    s = torch.tensor([float(local.get(k, 0)) for k in SUM_KEYS], dtype=torch.float64, device=device)
    m = torch.tensor([float(local.get(k, 0)) for k in MAX_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    out = {k: int(v) for k, v in zip(SUM_KEYS, s.tolist())}
    out.update({k: v for k, v in zip(MAX_KEYS, m.tolist())})
    out["max_steps"] = int(out["max_steps"])
    return out
