"""
glass_b200.sharding -- how the path is split over the GPUs of one box (one process per GPU).

The per-shell chain (a_lm draw/combine -> synthesis -> transformation -> Poisson/positions)
has no exchange step: shell j depends on other shells only through the random normals
z_{j-ncorr..j}, which are a pure function of (seed, shell, index) and are regenerated locally.
So shells are dealt round-robin to the ranks and nothing crosses NVLink on the data path;
torch.distributed is used only for the barrier and the max-over-ranks timing.
"""

from __future__ import annotations


def shard_shells(nshells: int, rank: int, world: int) -> range:
    """Shell indices owned by ``rank`` (round-robin, so every rank's shells span the same
    redshift range and cost)."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return range(rank, nshells, world)


def neighbours_needed(shells, ncorr: int) -> set[int]:
    """Shells whose normal deviates a rank must (re)generate: its own and the ``ncorr``
    preceding ones of each (glass/fields.py:410-420)."""
    need: set[int] = set()
    for j in shells:
        need.update(range(max(0, j - ncorr), j + 1))
    return need


def max_over_ranks(seconds: float, device=None) -> float:
    """Device-time reduction used by bench.py: MAX over ranks (1 rank: identity)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
