"""
End-to-end run of the BASELINE.json configurations through the glass_b200 API (the user loop
of examples/2-advanced/stage_4_galaxies.ipynb cell 13), device-resident:

    matter = generate(lognormal_fields(shells), gls, nside, ncorr=3)
    for delta_i in matter:
        convergence.add_window(delta_i, shell_i); kappa_i = convergence.kappa
        gamma1, gamma2 = shear_from_convergence(kappa_i, lmax, discretized=False)
        for lon, lat, count in positions_from_delta(ngal_i, delta_i, bias):
            z = redshifts(count, shell_i); eps = ellipticity_intnorm(count, sigma_e)
            she = galaxy_shear(lon, lat, eps, kappa_i, gamma1, gamma2)

    python tools/run_config.py 1|2|3|4 [--shells S] [--niter K] [--lensing]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_config.py 4 --lensing            # N GPUs: contiguous blocks of shells per rank,
                                                   # multi-plane recurrence pipelined over the ranks
Prints per-stage device times (CUDA events) and totals as one JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glass_b200  # noqa: E402
from bench import synthetic_gls  # noqa: E402

CONFIGS = {1: (10, 128, 383, False), 2: (20, 1024, 2047, False), 3: (40, 2048, 4095, True), 4: (60, 4096, 8191, False)}


class MockCosmology:  # reference tests/fixtures/domain.py:36-97
    Omega_m0 = 0.3
    hubble_distance = 4.4e3

    def H_over_H0(self, z):  # noqa: N802
        return (self.Omega_m0 * (1 + z) ** 3 + 1 - self.Omega_m0) ** 0.5

    def transverse_comoving_distance(self, z, z2=None):
        if z2 is None:
            return self.hubble_distance * np.asarray(z) * 1_000
        return self.hubble_distance * (np.asarray(z2) - np.asarray(z)) * 1_000


class HostCatalogSink(glass_b200.user._FitsWriter):
    """The catalogue columns of every batch copied to HOST memory through the product's
    double-buffered pinned staging on a side stream (glass_b200.user: the copy of batch i overlaps
    the kernels of batch i+1); the rows are handed to ``consume`` (default: counted and dropped --
    what happens to a catalogue on the host is the user's business, the maps never leave HBM)."""

    def __init__(self, consume=None):
        super().__init__(fh=None)
        self.bytes = 0
        self.consume = consume

    def _append_host(self, host_cols: dict) -> None:
        n = len(next(iter(host_cols.values())))
        self.nrows += n
        self.bytes += sum(a.nbytes for a in host_cols.values())
        if self.consume is not None:
            self.consume(host_cols)

    def close(self) -> None:
        self._drain()


def run_chain(config: int, *, dev, rank: int = 0, world: int = 1, shells: int | None = None, niter: int = 3, ngal: float | None = None,
              lensing: bool = False, ncorr: int = 3, galaxies: bool = True, host_catalog: bool = False, batch: int = 1_000_000,
              group: int = 4, discretized: bool = True) -> dict:
    """One pass of the user loop of the module docstring over a BASELINE.json configuration,
    device-resident (maps stay in HBM); with ``host_catalog`` the galaxy columns (lon, lat, z and,
    with lensing, the sheared ellipticity) are copied to host memory inside the timed region.
    ``world`` > 1: the caller has initialised torch.distributed (NCCL); contiguous blocks of shells
    per rank, multi-plane recurrence pipelined over the ranks.  Returns the result dict (identical
    on every rank: wall = max over ranks, galaxies = sum).  ``group``: convergence planes lensed per
    call of shear_from_convergence (a stack of planes shares the Legendre recurrences of the refinement
    syntheses); ``discretized``: the reference's default, pixel windows from hp.pixwin."""
    S, nside, lmax, cfg_lensing = CONFIGS[config]
    S = shells or S
    lensing = lensing or cfg_lensing
    args = argparse.Namespace(config=config, niter=niter, ncorr=ncorr, no_galaxies=not galaxies)
    if world > 1:
        import torch.distributed as dist
    mine = list(glass_b200.sharding.shard_shells(S, rank, world))
    npix = 12 * nside * nside
    dz = 1.0 / (S + 1)
    shells = [glass_b200.RadialWindow(np.array([i, i + 1.0, i + 2.0]) * dz, np.array([0.0, 1.0, 0.0]), (i + 1.0) * dz) for i in range(S)]
    shells_dev = [glass_b200.RadialWindow(torch.as_tensor(w.za, device=dev), torch.as_tensor(w.wa, device=dev), w.zeff) for w in shells]
    gls = [torch.as_tensor(g).to(dev) for g in synthetic_gls(S, lmax, args.ncorr)]
    ngal = ngal if ngal is not None else 6.7335 / 60  # per arcmin^2 per shell: 1.0e9 galaxies over 60 shells
    stages = {k: 0.0 for k in ("generate", "multiplane", "shear_from_convergence", "positions", "redshifts", "ellipticity", "galaxy_shear")}
    sink = HostCatalogSink() if host_catalog else None

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        pend.append((name, a, b))
        return r

    pend = []
    conv = glass_b200.MultiPlaneConvergence(MockCosmology())
    matter = glass_b200.generate(glass_b200.lognormal_fields(shells), gls, nside, ncorr=args.ncorr, rng=42, shells=mine if world > 1 else None)
    ngal_tot = 0
    if world > 1:
        from glass_b200.dist import multi_plane_block

        # the hand-off below is the first NCCL point-to-point of each pair: open the channels
        # outside the timed region (communicator setup, not the path)
        tok = torch.zeros(1, device=dev)
        if rank > 0:
            dist.recv(tok, rank - 1)
        if rank + 1 < world:
            dist.send(tok, rank + 1)
        dist.barrier()
    if lensing and discretized:
        glass_b200.healpix.pixwin(nside, lmax=lmax, pol=True)  # table set-up (cached per nside), like the transform plans
    torch.cuda.synchronize()
    t0 = time.perf_counter()

    def per_shell(i, delta, kappa, shear=None):
        nonlocal ngal_tot
        g1 = g2 = None
        rng_i = np.random.default_rng([42, i])  # the shell's own generator: the same galaxies on any number of ranks
        if lensing:
            g1, g2 = shear
        it = iter(()) if args.no_galaxies else glass_b200.positions_from_delta(ngal, delta, 1.2, rng=rng_i, batch=batch)
        while True:
            try:
                lon, lat, cnt = timed("positions", lambda: next(it))
            except StopIteration:
                break
            ngal_tot += cnt
            z = timed("redshifts", lambda: glass_b200.redshifts(torch.as_tensor(cnt), shells_dev[i], rng=rng_i))
            cols = {"RA": lon, "DEC": lat, "Z": z}
            if lensing:
                eps = timed("ellipticity", lambda: glass_b200.ellipticity_intnorm(cnt, 0.27, rng=rng_i, xp=torch))
                she = timed("galaxy_shear", lambda: glass_b200.galaxy_shear(lon, lat, eps, kappa, g1, g2))
                cols["E1"], cols["E2"] = she.real, she.imag
            if sink is not None and cnt:
                sink.write(**cols)

    def shear_group(kappas_g):
        """kappa -> shear for up to four planes at once (the refinement syntheses of the planes share
        one Legendre recurrence); one plane: the plain call of the reference's loop."""
        if len(kappas_g) == 1:
            return [timed("shear_from_convergence", lambda: glass_b200.shear_from_convergence(kappas_g[0], lmax, discretized=discretized, niter=args.niter))]
        g1, g2 = timed("shear_from_convergence", lambda: glass_b200.shear_from_convergence(torch.stack(kappas_g), lmax, discretized=discretized, niter=args.niter))
        return [(g1[b], g2[b]) for b in range(len(kappas_g))]

    group = max(1, int(group))
    if world == 1:
        # the user loop in groups of `group` shells: the recurrence of the convergence is local, so a
        # group's planes are made, lensed together and populated before the next group is generated
        for a in range(0, S, group):
            idx = list(range(a, min(a + group, S)))
            deltas = [timed("generate", lambda: next(matter)) for _ in idx]
            kappas = [None] * len(idx)
            if lensing:
                for b, i in enumerate(idx):
                    timed("multiplane", lambda: conv.add_window(deltas[b], shells[i]))
                    kappas[b] = conv.kappa.clone() if len(idx) > 1 else conv.kappa  # the recurrence recycles its buffers
            shears = shear_group(kappas) if lensing else [None] * len(idx)
            for b, i in enumerate(idx):
                per_shell(i, deltas[b], kappas[b], shears[b])
    else:
        # 1. matter planes of the block (no communication)  2. multi-plane recurrence, pipelined
        # over the ranks (dist.multi_plane_block)  3. transforms and galaxies (no communication)
        # (copies: the generator yields views of its batches of eight maps, and a batch would stay allocated until the
        # last of its eight shells is finished)
        deltas = [timed("generate", lambda: next(matter).clone()) for _ in mine]
        matter.close()  # the generator's batch buffers and the transform's scratch go back before the planes are lensed
        glass_b200.healpix.release_scratch()
        torch.cuda.empty_cache()  # the library allocates its lensing workspace with cudaMalloc: give it what torch has cached
        conv2 = None
        if lensing:
            # the recurrence runs through the block TWICE: once right away, to hand its state to the next rank as early
            # as possible (1 ms per plane), and again group by group below from a snapshot of the state it started
            # from -- the block's 30 convergence planes (48 GB at nside 4096) are never held at the same time
            from glass_b200.dist import _MP_MAPS, _MP_SCALARS, recv_multi_plane_state, send_multi_plane_state

            like = deltas[0] if deltas else torch.empty(npix, dtype=torch.float64, device=dev)

            def advance():
                if rank > 0:
                    recv_multi_plane_state(conv, rank - 1, like=like)
                snap = ({k: getattr(conv, k) for k in _MP_SCALARS},
                        {k: (getattr(conv, k).clone() if getattr(conv, k) is not None else None) for k in _MP_MAPS})
                for b, i in enumerate(mine):
                    conv.add_window(deltas[b], shells[i])
                if rank + 1 < world:
                    send_multi_plane_state(conv, rank + 1, like=like)
                return snap

            snap = timed("multiplane", advance)
            conv2 = glass_b200.MultiPlaneConvergence(MockCosmology())
            for k, v in snap[0].items():
                setattr(conv2, k, v)
            for k, v in snap[1].items():
                conv2._set_state_map(k, v)
        torch.cuda.empty_cache()  # (the transforms' workspace is allocated by the library on first use)
        for a in range(0, len(mine), group):
            hi = min(a + group, len(mine))
            kappas = [None] * (hi - a)
            if lensing:
                for b in range(a, hi):
                    timed("multiplane", lambda: conv2.add_window(deltas[b], shells[mine[b]]))
                    kappas[b - a] = conv2.kappa.clone()
            shears = shear_group(kappas) if lensing else [None] * (hi - a)
            for b in range(a, hi):
                per_shell(mine[b], deltas[b], kappas[b - a], shears[b - a])
                deltas[b] = None  # this shell is finished: its map goes back to the allocator
            kappas = shears = None
    if sink is not None:
        sink.close()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    for name, a, b in pend:
        stages[name] += a.elapsed_time(b)
    host_bytes = sink.bytes if sink is not None else 0
    if world > 1:
        # whole-job numbers: wall = max over ranks, galaxies = sum, stage times = max over ranks
        t = torch.tensor([wall] + [stages[k] for k in stages], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
        for k, v in zip(stages, t[1:].tolist()):
            stages[k] = v
        g = torch.tensor([int(ngal_tot), int(host_bytes)], dtype=torch.int64, device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ngal_tot, host_bytes = int(g[0]), int(g[1])
    return {
        "config": args.config, "n_gpus": world, "group": group, "discretized": discretized, "shells": S, "ncorr": args.ncorr, "nside": nside, "lmax": lmax, "lensing": lensing, "niter": args.niter,
        "galaxies": int(ngal_tot), "wall_s": wall, "shells_per_s": S / wall, "galaxies_per_s": ngal_tot / wall,
        "catalog_d2h_bytes": host_bytes,
        "stage_ms_total": {k: round(v, 2) for k, v in stages.items()},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int, choices=[1, 2, 3, 4])
    ap.add_argument("--shells", type=int, default=None)
    ap.add_argument("--niter", type=int, default=3)
    ap.add_argument("--ngal", type=float, default=None, help="galaxies per arcmin^2 per shell")
    ap.add_argument("--lensing", action="store_true")
    ap.add_argument("--ncorr", type=int, default=3, help="correlated shells (59 = all 60 shells of config 4 fully correlated)")
    ap.add_argument("--no-galaxies", action="store_true")
    ap.add_argument("--host-catalog", action="store_true", help="copy the galaxy columns to host memory inside the timed region")
    ap.add_argument("--group", type=int, default=4, help="convergence planes per shear_from_convergence call")
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    out = run_chain(args.config, dev=dev, rank=rank, world=world, shells=args.shells, niter=args.niter, ngal=args.ngal, lensing=args.lensing,
                    ncorr=args.ncorr, galaxies=not args.no_galaxies, host_catalog=args.host_catalog, group=args.group)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
