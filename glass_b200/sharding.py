"""
glass_b200.sharding -- how the path is split over the GPUs of one box (one process per GPU).

The per-shell chain (a_lm draw/combine -> synthesis -> transformation -> Poisson/positions)
has no exchange step: shell j depends on other shells only through the random normals
z_{j-ncorr..j}, which are a pure function of (seed, shell, index) and are regenerated locally.
So every rank takes a contiguous block of shells and nothing crosses NVLink on the data path;
torch.distributed is used only for the barrier and the max-over-ranks timing.
"""

from __future__ import annotations


def shard_shells(nshells: int, rank: int, world: int, mode: str = "block") -> range:
    """Shell indices owned by ``rank``.

    ``"block"`` (default): contiguous blocks of ceil/floor(nshells / world) shells.  Every shell
    costs the same (one nside, one lmax), and a block re-uses the normal deviates of its own
    preceding shells, so a rank regenerates only the ``ncorr`` deviates before its block --
    dealt ``"round-robin"`` (``rank, rank + world, ...``) it regenerates ``ncorr`` per shell
    (measured at nside 4096, ncorr 3: +0.37 ms per regenerated shell) and walks ``world`` times
    more of the host-side iternorm recursion per produced shell."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    if mode == "round-robin":
        return range(rank, nshells, world)
    if mode != "block":
        raise ValueError("mode must be 'block' or 'round-robin'")
    base, extra = divmod(nshells, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def bind_to_local_cpus(device_index: int) -> list[int] | None:
    """Pin this process to the CPU cores NVML reports as local to the GPU (same NUMA node /
    PCIe root), BEFORE any pinned host buffer is allocated: page-locked staging memory is then
    first-touched on the GPU's own socket and the device->host copies of several ranks do not
    cross the inter-socket link.  Returns the cores, or None if NVML / affinity is unavailable
    (then nothing is changed)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


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


# ------------------------------------------------------------------------------------------
# m-split of one transform (SURVEY.md 8e axis 2): Legendre stage sharded by m, Fourier and
# pixel stages sharded by ring band, one all-to-all of the phase array in between.
# ------------------------------------------------------------------------------------------


def msplit_layout(nside: int, lmax: int, world: int) -> dict:
    """
    Ownership and buffer layout for the m-split transform.

    * rank r owns the m values ``r, r + world, ...`` (cost per m falls linearly with m, so
      round-robin balances the Legendre work) -- slot of m on its owner: ``m // world``,
      ``W = ceil((lmax+1)/world)`` slots per rank;
    * rank d owns a band of ring pairs ``[pair_lo[d], pair_hi[d])`` -- the northern rings and
      their southern mirrors, so the north/south symmetry of the Legendre stage stays local --
      with bands cut at equal estimated Fourier-stage cost;
    * ``rings[d]``: the rings of rank d in local-row order (north ascending, then mirrors);
      ``rowmap[ring]``: row of the ring in the send buffer ``[row][W]``, rows grouped by
      owner so the all-to-all splits are contiguous.
    """
    import numpy as np

    nring, npair = 4 * nside - 1, 2 * nside
    nphi = np.array([4 * (r + 1) if r + 1 < nside else 4 * nside for r in range(npair)], dtype=np.int64)
    # cost of a ring pair: its pixels, times 3 where the ring FFT length is not a power of two
    # (those rings go through the chirp-z path: four FFTs instead of one)
    h = nphi // 2
    pow2 = (h & (h - 1)) == 0
    cost = nphi * np.where(pow2, 1, 3) * np.where(np.arange(npair) == npair - 1, 1, 2)
    cum = np.concatenate([[0], np.cumsum(cost)])
    cuts = [int(np.searchsorted(cum, cum[-1] * d / world, side="left")) for d in range(world + 1)]
    cuts[0], cuts[-1] = 0, npair
    for d in range(1, world + 1):
        cuts[d] = max(cuts[d], cuts[d - 1])
    rings = []
    for d in range(world):
        lo, hi = cuts[d], cuts[d + 1]
        north = list(range(lo, hi))
        south = [nring - 1 - r for r in north if r != npair - 1]
        rings.append(north + south)
    rowmap = np.empty(nring, dtype=np.int32)
    row = 0
    for d in range(world):
        for r in rings[d]:
            rowmap[r] = row
            row += 1
    assert row == nring
    return {
        "world": world,
        "W": (lmax + 1 + world - 1) // world,
        "pair_lo": cuts[:-1],
        "pair_hi": cuts[1:],
        "rings": rings,
        "rows": [len(r) for r in rings],
        "rowmap": rowmap,
    }


def ring_start(nside: int, ring: int) -> int:
    """First pixel of ring index ``ring`` (0-based, RING scheme)."""
    i = ring + 1
    if i < nside:
        return 2 * i * (i - 1)
    if i <= 3 * nside:
        return 2 * nside * (nside - 1) + (i - nside) * 4 * nside
    ip = 4 * nside - i
    return 12 * nside * nside - 2 * ip * (ip + 1)


def owned_pixel_ranges(nside: int, layout: dict, rank: int) -> list[tuple[int, int]]:
    """Contiguous pixel ranges (north band, south band) written by ``rank``."""
    nring, npair = 4 * nside - 1, 2 * nside
    lo, hi = layout["pair_lo"][rank], layout["pair_hi"][rank]
    if hi <= lo:
        return []
    npix = 12 * nside * nside

    def end_of(ring):
        return ring_start(nside, ring + 1) if ring + 1 < nring else npix

    out = [(ring_start(nside, lo), end_of(hi - 1))]
    s_hi = nring - 1 - lo  # southernmost mirror
    s_lo = nring - 1 - (hi - 1)
    if hi - 1 == npair - 1:  # the equator has no mirror
        s_lo += 1
    if s_lo <= s_hi:
        out.append((ring_start(nside, s_lo), end_of(s_hi)))
    # merge if the two bands touch (the rank owning the equator)
    if len(out) == 2 and out[0][1] == out[1][0]:
        out = [(out[0][0], out[1][1])]
    return out
