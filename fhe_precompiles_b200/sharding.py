"""Batch sharding across GPUs: every precompile call is independent, so a batch splits into contiguous
per-rank slices with no data-path collective.  torch.distributed is used for the barrier and
for the max-over-ranks of a timing, nothing else."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open [lo, hi) slice of `n` items owned by `rank` of `world`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Cost-weighted assignment of call indices to ranks (greedy longest-processing-time): used for mixed
    batches where ct x ct multiplies dominate (BASELINE config 4). Deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


# relative device cost of one call by precompile family (measured on B200, us): mul ct x ct dominates
OP_COST = {"mul_ctct": 2.8, "mul_ctpt": 0.45, "add_ctct": 0.06, "sub_ctct": 0.06, "add_ctpt": 0.05, "sub_ctpt": 0.05}


def call_cost(name: str) -> float:
    op = name.split("_", 1)[0]
    ctct = name.count("cipher") == 2
    return OP_COST.get(f"{op}_{'ctct' if ctct else 'ctpt'}", 1.0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """max of a per-rank scalar (elapsed time) over all ranks; identity when not distributed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank: int, world: int, seconds_max: float) -> float:
    """units processed by all ranks divided by the slowest rank's time (weak scaling)."""
    return units_per_rank * world / seconds_max
