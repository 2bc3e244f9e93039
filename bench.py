#!/usr/bin/env python
"""
bench.py -- lognormal HEALPix shells/sec at nside=4096 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = every rank draws, combines and synthesises SHELLS_PER_STEP correlated
lognormal matter shells (nside=4096, lmax=8191, ncorr=3; the shape of BASELINE.json
configs[3], which fits one B200) through glass_b200.generate():
iternorm (host) -> Philox a_lm draw + banded combine -> Legendre synthesis (FP64) ->
ring FFT with fused lognormal.  Shells are sharded across ranks (rank r takes shells
r, r+N, ...); the path has no exchange step, so there is no data-path collective and
scaling is weak.

Printed JSON line (rank 0): value = device-resident throughput (maps stay in HBM),
e2e = same through the public API with HOST buffers (NumPy gls in, NumPy maps out; the
device->host copy of every map is inside the timed region), roofline = the Legendre
kernel against the FP64 DFMA peak measured in this run, cpu_baseline = the CPU arm (GLASS's
NumPy steps + the SIMD/OpenMP synthesis of oracle/sht_fast.cpp standing in for healpy) timed on
the host cores (N=1 only).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDE, LMAX, NCORR = 4096, 8191, 3
SHELLS_PER_STEP = 8  # eight shells share one Legendre recurrence (INT8 tensor-core contraction, csrc/sht_ozaki.cu)
METRIC = "lognormal HEALPix shells/sec at nside=4096"
UNIT = "shells/s"


def synthetic_gls(nshell: int, lmax: int, ncorr: int):
    """g_l = 1e-2 (l+1)^-1.5 (g_0 = 0), cross 0.5^|i-j| g_l within ncorr."""
    l = np.arange(lmax + 1)
    g = 1e-2 * (l + 1.0) ** -1.5
    g[0] = 0.0
    gls = []
    for i in range(nshell):
        for j in range(i, -1, -1):
            gls.append(0.5 ** (i - j) * g if i - j <= ncorr else np.zeros(0))
    return gls


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # the busiest samples are the ones under load
        sm_load = sorted(sm)[len(sm) // 3 :] if sm else []
        return {
            "sm_mhz": float(np.median(sm_load)) if sm_load else None,
            "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------
# CPU arm (oracle port): the reference path for one lognormal shell on the host cores
# ------------------------------------------------------------------------------------------
def host_threads() -> int:
    """The cores this process may run on.  NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which is a launcher default and not the box."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuShells:
    """The reference path shell by shell on the CPU at one size: what glass.generate does per
    shell (glass/fields.py:404-425, grf/_transformations.py:83-89) with healpy.alm2map
    (glass/healpix.py:71) replaced by oracle/sht_fast.cpp.  The spectra, the iternorm weights
    (a (lmax+1) x (ncorr+1) array) and the normals of the NCORR earlier shells are prepared once;
    ``step`` then does ALL the work of one further shell and nothing else."""

    def __init__(self, nside: int, lmax: int, nthreads: int):
        from oracle import glass_ref as G

        self.nside, self.lmax, self.nthreads = nside, lmax, nthreads
        self.gls = synthetic_gls(NCORR + 1, lmax, NCORR)
        self.w = G.iternorm(G.cls2cov_rows(self.gls, lmax + 1, NCORR + 1, NCORR))[-1]  # every later shell has this row
        self.var = G.cltovar(self.gls[0])
        self.rng = np.random.default_rng(42)
        self.n = (lmax + 1) * (lmax + 2) // 2
        self.y = [self.rng.standard_normal((self.n, 2)) @ np.array([1, 1j]) for _ in range(NCORR)]

    def step(self) -> tuple[float, float]:
        """One shell; returns (seconds in GLASS's NumPy code, seconds in alm2map); their sum is
        the wall time of the shell."""
        from oracle import glass_ref as G
        from oracle import sht_c

        t0 = time.perf_counter()
        self.y.append(self.rng.standard_normal((self.n, 2)) @ np.array([1, 1j]))  # fields.py:407
        self.y = self.y[-(NCORR + 1):]  # fields.py:410-414
        alm = sum(G.multalm(z, self.w[:, i]) for i, z in enumerate(self.y))  # fields.py:420
        alm = G.glass_to_healpix_alm(alm)  # fields.py:422
        alm[: self.lmax + 1] = alm[: self.lmax + 1].real + alm[: self.lmax + 1].imag + 0j  # fields.py:425
        t1 = time.perf_counter()
        m = sht_c.alm2map_fast(alm, self.nside, self.lmax, nthreads=self.nthreads)
        t2 = time.perf_counter()
        m = G.lognormal(m, self.var, 1.0)
        t3 = time.perf_counter()
        self.checksum = float(m[0])
        return (t1 - t0) + (t3 - t2), t2 - t1


CPU_NOTE = (
    "CPU restatement of the reference path (healpy/libsharp2 unavailable offline): GLASS's NumPy steps on one "
    "thread as in the reference; alm2map = oracle/sht_fast.cpp (AVX-512 across rings, register-blocked "
    "recurrence, mlim ring skipping, OpenMP over m)"
)


def workload_config(nside: int, lmax: int, world: int) -> dict:
    """The ``config`` object of BOTH arms (the driver compares them)."""
    return {
        "workload": f"correlated lognormal matter shells through generate(), nside={nside} lmax={lmax} ncorr={NCORR}, "
        f"synthetic power-law C_l (BASELINE.json configs[3] shape; {SHELLS_PER_STEP} shells per GPU per step, shell-sharded over {world} rank(s))",
        "shells_per_step_per_rank": SHELLS_PER_STEP,
        "l2": "inputs larger than L2 (a_lm 0.54 GB, phases 2.1 GB, map 1.6 GB per shell)",
        "parallelism": f"shell-sharded x{world}, no data-path collective",
    }


def cpu_baseline(nside: int, lmax: int, nshells: int = 3) -> dict:
    """cpu_baseline leg of the B200 line (N = 1): ``nshells`` shells at the bench size itself
    (about 5 s each on 16 cores) after one untimed shell."""
    cores = host_threads()
    cpu = CpuShells(nside, lmax, cores)
    cpu.step()
    t = [cpu.step() for _ in range(nshells)]
    t_np, t_sht = (sum(x[i] for x in t) / nshells for i in (0, 1))
    sample = (
        f"{nshells} lognormal shells (alm draw + combine + re-order, alm2map, expm1) at nside={nside} lmax={lmax} after one untimed "
        f"shell; per shell {t_np:.2f} s NumPy (1 thread) + {t_sht:.2f} s alm2map on {cores} threads; nothing extrapolated"
    )
    return {"value": 1.0 / (t_np + t_sht), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE}


def extra_stage_rooflines(dev, hbm_peak: float, fp64_peak: float) -> dict:
    """Secondary measurements (N=1 only, a few seconds): the HBM-bound sampling / lensing
    kernels at nside=4096 and the other two FP64 transforms at nside=2048, each timed with CUDA
    events on the launching stream, against the algorithmic bytes / flops of SURVEY.md 8(d)."""
    import ctypes as C

    import torch

    import glass_b200
    from glass_b200 import _lib
    from glass_b200 import healpix as hp
    from glass_b200.points import ARCMIN2_SPHERE

    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream

    def ev(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    out = {}
    nside = 4096
    npix = 12 * nside * nside
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    delta = torch.expm1(0.5 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g) - 0.125)
    counts = torch.empty(npix, dtype=torch.int64, device=dev)
    off = torch.empty(npix + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.glb_points_workspace_bytes(npix)), dtype=torch.uint8, device=dev)
    scale = 0.083  # expected galaxies per pixel (1e9 galaxies over 60 shells at nside 4096)

    tot_d = torch.zeros(1, dtype=torch.int64, device=dev)

    def k67():
        _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, 1.2, scale, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                         counts.data_ptr(), off.data_ptr(), None, 0, tot_d.data_ptr(), ws.data_ptr(), st))

    t = ev(k67)
    out["points_counts+offsets (K6+K7, scan mode)"] = {"ms": t, "GB/s": npix * 32 / t / 1e6, "frac_hbm": npix * 32 / t / 1e6 / hbm_peak, "algorithmic_bytes": npix * 32,
                                                       "note": "per-pixel counts + exclusive offsets; positions_from_delta uses it above 1 galaxy per pixel"}
    tot = int(off[-1].item())
    lon = torch.empty(tot, dtype=torch.float64, device=dev)
    lat = torch.empty(tot, dtype=torch.float64, device=dev)

    def k8():
        _lib.check(lib.glb_points_fill(nside, counts.data_ptr(), off.data_ptr(), 0, npix, None, None, C.c_uint64(42), C.c_uint32(0),
                                       lon.data_ptr(), lat.data_ptr(), None, st))

    t = ev(k8)
    by = npix * 8 + tot * 16
    out["points_fill (K8, scan mode)"] = {"ms": t, "GB/s": by / t / 1e6, "frac_hbm": by / t / 1e6 / hbm_peak, "galaxies": tot, "algorithmic_bytes": by}
    # LIST mode (the default of positions_from_delta at this density): K6 emits the galaxy -> pixel list
    # (8 B per galaxy) instead of counts + offsets (16 B per pixel); K8 is one thread per galaxy.  The
    # algorithmic bytes are SURVEY.md 8(d)'s (32 B/pixel; 8 B/pixel + 16 B/galaxy) -- what a
    # pixel-array formulation has to move -- so the fractions can exceed 1; "moved" is what this one moves.
    gpix = torch.empty(tot + 1024, dtype=torch.int64, device=dev)

    def k67l():
        _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, 1.2, scale, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                         None, None, gpix.data_ptr(), gpix.numel(), tot_d.data_ptr(), ws.data_ptr(), st))

    t = ev(k67l)
    moved = npix * 8 + tot * 8
    out["points_counts+offsets (K6+K7)"] = {"ms": t, "GB/s": npix * 32 / t / 1e6, "frac_hbm": npix * 32 / t / 1e6 / hbm_peak, "algorithmic_bytes": npix * 32,
                                            "moved_bytes": moved, "moved_GB/s": moved / t / 1e6, "frac_hbm_moved": moved / t / 1e6 / hbm_peak, "mode": "list"}

    def k8l():
        _lib.check(lib.glb_points_fill_list(nside, gpix.data_ptr(), 0, tot, None, None, C.c_uint64(42), C.c_uint32(0), lon.data_ptr(), lat.data_ptr(), st))

    t = ev(k8l)
    moved = tot * 24
    out["points_fill (K8)"] = {"ms": t, "GB/s": by / t / 1e6, "frac_hbm": by / t / 1e6 / hbm_peak, "galaxies": tot, "algorithmic_bytes": by,
                               "moved_bytes": moved, "moved_GB/s": moved / t / 1e6, "frac_hbm_moved": moved / t / 1e6 / hbm_peak, "mode": "list"}
    k3 = torch.zeros(npix, dtype=torch.float64, device=dev)
    k2 = torch.rand(npix, dtype=torch.float64, device=dev, generator=g)
    t = ev(lambda: _lib.check(lib.glb_multiplane_update(k3.data_ptr(), k2.data_ptr(), delta.data_ptr(), 0.0, npix, 0.3, 0.01, st)))
    out["multiplane_update (K9)"] = {"ms": t, "GB/s": npix * 32 / t / 1e6, "frac_hbm": npix * 32 / t / 1e6 / hbm_peak, "algorithmic_bytes": npix * 32}
    eps = glass_b200.ellipticity_intnorm(tot, 0.27, rng=1, xp=torch)
    res = torch.empty(tot, dtype=torch.complex128, device=dev)
    # three DISTINCT maps: every gather is its own 32-byte sector (144 B of DRAM traffic per galaxy
    # against the 72 B the algorithm needs -- the sector granularity of random 8-byte reads)
    t = ev(lambda: _lib.check(lib.glb_galaxy_shear(nside, lon.data_ptr(), lat.data_ptr(), None, eps.data_ptr(), tot, k2.data_ptr(),
                                                   k3.data_ptr(), delta.data_ptr(), 1, res.data_ptr(), st)))
    out["galaxy_shear (K12)"] = {"ms": t, "GB/s": tot * 72 / t / 1e6, "frac_hbm": tot * 72 / t / 1e6 / hbm_peak, "algorithmic_bytes": tot * 72,
                                 "sector_bytes": tot * 144,
                                 "note": "ncu (profiles/r01_ncu_summary_v4.txt): 243 B of DRAM reads per galaxy (three sparse gathers at 0.083 "
                                 "galaxies per pixel pull 64 B each) at 87 % of peak DRAM throughput: HBM-bound on the traffic it causes"}
    del delta, counts, off, lon, lat, k3, k2, eps, res
    torch.cuda.empty_cache()
    # FP64 transforms of the lensing stage at nside 2048 (BASELINE.json configs[2])
    nside, lmax = 2048, 4095
    ntri = (lmax + 1) * (lmax + 2) // 2 * 2 * nside
    kap = 0.01 * torch.randn(12 * nside * nside, dtype=torch.float64, device=dev, generator=g)
    t = ev(lambda: hp.map2alm(kap, lmax=lmax, pol=False, niter=0), n=3, warm=1)
    out["map2alm niter=0 (K10)"] = {"ms": t, "TFLOP/s": 8 * ntri / t / 1e9, "frac_fp64": 8 * ntri / t / 1e9 / fp64_peak, "algorithmic_flop": 8 * ntri}
    alm = hp.map2alm(kap, lmax=lmax, pol=False, niter=0)
    t = ev(lambda: hp.alm2map_spin([alm, None], nside, 2, lmax), n=3, warm=1)
    out["alm2map_spin s=2 E-only (K11)"] = {"ms": t, "TFLOP/s": 16 * ntri / t / 1e9, "frac_fp64": 16 * ntri / t / 1e9 / fp64_peak, "algorithmic_flop": 16 * ntri}
    # the same for FOUR convergence planes on shared Wigner-d recurrences (what shear_from_convergence does with a stack)
    alm4 = torch.stack([alm, 0.5 * alm, -alm, 2.0 * alm])
    t = ev(lambda: hp.alm2map_spin_batch(alm4, nside, 2, lmax), n=3, warm=1)
    out["alm2map_spin s=2, 4 planes per recurrence (K11 batched)"] = {"ms_per_plane": t / 4, "TFLOP/s": 64 * ntri / t / 1e9, "frac_fp64": 64 * ntri / t / 1e9 / fp64_peak,
                                                                      "algorithmic_flop": 64 * ntri, "note": "includes the ring FFTs of the eight maps"}
    del alm4
    hp.clear_plans()
    torch.cuda.empty_cache()
    return out


def bench_msplit(dev, rank: int, world: int, nside: int, lmax: int, single_ms: float | None) -> dict:
    """SURVEY.md 8(e) axis 2, measured in the same run (N > 1): ONE batch of 4 maps at the bench
    size with the Legendre stage sharded by m over the N GPUs and the ring FFT by ring band, in
    both forms of the m -> ring transpose -- an NCCL all-to-all of the phase array, and the fused
    form in which the Legendre kernel stores F_m(ring) straight into the owner's buffer over
    NVLink peer mappings.  Times are CUDA events around whole transforms (max over ranks);
    ``alltoall_ms`` is the collective alone (events around the NCCL calls).  bit_identical: every
    rank's ring bands equal to the single-GPU transform of the same alm, checked on each rank."""
    import torch
    import torch.distributed as dist

    from glass_b200.dist import MSplitTransform
    from glass_b200.healpix import alm2map_batch

    nb = 4
    nalm = (lmax + 1) * (lmax + 2) // 2
    g = torch.Generator(device=dev)
    g.manual_seed(2024)  # the same alm on every rank
    alm = torch.view_as_complex(torch.randn((nb, nalm, 2), dtype=torch.float64, device=dev, generator=g))
    ref = alm2map_batch(alm, nside, lmax)
    out: dict = {"maps": nb, "nside": nside, "lmax": lmax, "n_gpus": world}

    def timed(fn, n=3, warm=2):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    if single_ms is None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(2):
            alm2map_batch(alm, nside, lmax, out=ref)
        b.record()
        torch.cuda.synchronize()
        single_ms = a.elapsed_time(b) / 2
    out["single_gpu_ms"] = single_ms
    ok = True
    for name, p2p in (("nccl_alltoall", False), ("p2p_fused", True)):
        ms = MSplitTransform(nside, lmax, max_batch=nb, device=dev, p2p=p2p)
        res = ms.alm2map(alm)
        torch.cuda.synchronize()
        same = all(bool(torch.equal(res[:, x:y], ref[:, x:y])) for x, y in ms.pixel_ranges)
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(int(flag[0]))
        t = timed(lambda: ms.alm2map(alm, out=res))
        # phases this rank hands to the others per transform: its m columns of every foreign ring
        sent = nb * sum(r for k, r in enumerate(ms.rows) if k != ms.rank) * ms.W * 16
        ent = {"ms": t, "speedup_vs_1gpu": single_ms / t, "nvlink_bytes_sent_per_rank": sent}
        if not p2p:
            send = torch.empty((nb, ms.nring, ms.W), dtype=torch.complex128, device=dev)
            recv = torch.empty((nb, world, ms.rows[ms.rank], ms.W), dtype=torch.complex128, device=dev)
            ins, outs = [r * ms.W for r in ms.rows], [ms.rows[ms.rank] * ms.W] * world

            def a2a():
                for b_ in range(nb):
                    dist.all_to_all_single(torch.view_as_real(recv[b_]).reshape(-1, 2), torch.view_as_real(send[b_]).reshape(-1, 2),
                                           output_split_sizes=outs, input_split_sizes=ins)

            ta = timed(a2a)
            ent.update({"alltoall_ms": ta, "alltoall_GB/s_per_rank": sent / ta / 1e6, "alltoall_share": ta / t})
            del send, recv
        out[name] = ent
        del ms, res
        torch.cuda.empty_cache()
    out["bit_identical"] = ok
    return out


def bench_chain(dev, rank: int, world: int) -> dict:
    """The whole user loop the north star names (examples/2-advanced/stage_4_galaxies.ipynb cell 13,
    SURVEY.md 3.5) on BASELINE.json configs[3]+[4]: 60 correlated lognormal shells at nside 4096,
    lmax 8191 -> multi-plane convergence -> shear_from_convergence (niter 3) -> 1.0e9 galaxy
    positions, redshifts, ellipticities, reduced shear; through the public API, strong scaling
    over the N GPUs (contiguous shell blocks, multi-plane recurrence pipelined over the ranks).
    ``wall_s``: maps and catalogues stay in HBM; ``e2e``: the same with every galaxy column copied
    to host memory inside the timed region (the maps never leave the device)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from run_config import run_chain

    out = {}
    for key, host in (("device", False), ("e2e", True)):
        r = run_chain(4, dev=dev, rank=rank, world=world, lensing=True, niter=3, host_catalog=host)
        out[key] = {k: r[k] for k in ("wall_s", "galaxies", "galaxies_per_s", "shells_per_s", "catalog_d2h_bytes", "stage_ms_total")}
        if key == "device":
            out["workload"] = (f"{r['shells']} shells nside={r['nside']} lmax={r['lmax']} ncorr={r['ncorr']} + MultiPlaneConvergence + "
                               f"shear_from_convergence(niter={r['niter']}) + {r['galaxies']:.3g} galaxies (positions, redshifts, ellipticities, "
                               f"reduced shear) on {world} GPU(s), strong scaling")
    return out


def run_reference(args) -> None:
    """The reference arm: the CPU port of the SAME workload (nside, lmax, ncorr of the B200 arm)
    on all host cores of the box, one shell per step, every step measured at full size -- nothing
    extrapolated.  Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nside, lmax = args.nside, args.lmax
    cores = host_threads()
    cpu = CpuShells(nside, lmax, cores)
    budget_s = 200.0  # of timed CPU work; the run must end within a few minutes
    t_first = sum(cpu.step())  # first warm-up step (thread pool start-up, first touch)
    for _ in range(max(0, min(args.warmup - 1, int(20.0 // t_first)))):
        cpu.step()
    steps = max(1, min(args.steps, int(budget_s // t_first)))
    t = [cpu.step() for _ in range(steps)]
    t_np, t_sht = (sum(x[i] for x in t) / steps for i in (0, 1))
    dt = t_np + t_sht
    value = 1.0 / dt
    sample = (
        f"each step = 1 lognormal shell at nside={nside} lmax={lmax} measured at full size "
        f"({t_np:.2f} s NumPy on 1 thread + {t_sht:.2f} s alm2map on {cores} threads per shell); {steps} timed steps"
        + ("" if steps == args.steps else f" (of the {args.steps} asked for: {budget_s:.0f} s budget)")
    )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(nside, lmax, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    import glass_b200
    from glass_b200 import _lib
    from glass_b200.healpix import get_plan
    from glass_b200.sharding import bind_to_local_cpus, shard_shells

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host threads: the per-shell host work (iternorm on (lmax+1, ncorr+1) arrays) gains nothing
    # from BLAS/OpenMP threads, and world ranks x all cores oversubscribes the box
    local_cpus = bind_to_local_cpus(local) if world > 1 else None  # NUMA-local pinned staging buffers
    per_rank = max(1, (os.cpu_count() or 1) // max(1, world))
    torch.set_num_threads(per_rank)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=per_rank)
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    nside, lmax = args.nside, args.lmax
    K, W, S = args.steps, args.warmup, SHELLS_PER_STEP
    nsteps = K + W
    nshell_total = world * S * nsteps
    gls = synthetic_gls(nshell_total, lmax, NCORR)
    fields = [glass_b200.grf.Lognormal(1.0)] * nshell_total
    mine = shard_shells(nshell_total, rank, world)  # contiguous block of S * nsteps shells per rank
    lib = _lib.load()
    plan = get_plan(nside, lmax, max_batch=4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(gls_in, consume, stats=None):
        """W warm-up steps, then exactly K timed steps; returns max-over-ranks seconds."""
        gen = glass_b200.generate(fields, gls_in, nside, ncorr=NCORR, rng=42, shells=mine, stats=stats)
        for _ in range(W * S):
            consume(next(gen))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K * S):
            consume(next(gen))
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3
        barrier()
        t = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gen.close()
        return float(t[0]), float(t[1])

    # ---- FP64 peak (roofline denominator), measured live ----
    import ctypes as C

    peak_tf, peak_ms = C.c_double(), C.c_double()
    _lib.check(lib.glb_measure_fp64_peak(local, C.byref(peak_tf), C.byref(peak_ms), None), "glb_measure_fp64_peak")

    # second FP64 figure: cuBLAS DGEMM 8192^3 through torch.matmul, rank 0 only
    dgemm_tf = None
    if rank == 0:
        try:
            n_g = 8192
            ga = torch.randn((n_g, n_g), dtype=torch.float64, device=dev)
            gb = torch.randn((n_g, n_g), dtype=torch.float64, device=dev)
            torch.matmul(ga, gb)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                torch.matmul(ga, gb)
            g1.record()
            torch.cuda.synchronize()
            dgemm_tf = 3 * 2.0 * n_g**3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
            del ga, gb
            torch.cuda.empty_cache()
        except Exception:  # a library figure for context only: never fail the bench on it
            dgemm_tf = None

    # ---- device-resident arm: gls on the device -> maps stay in HBM ----
    gls_dev = [torch.as_tensor(g).to(dev) for g in gls]
    checksum = torch.zeros((), dtype=torch.float64, device=dev)

    def consume_dev(m):
        checksum.add_(m[::4097].sum())  # touch the result; negligible work

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _lib.check(lib.glb_plan_timing_enable(plan.handle, 0), "timing")
    launches0 = None

    # warm-up happens inside timed_run; enable stage timing only for the timed region by
    # wrapping consume: switch it on at the first timed shell
    state = {"n": 0}

    def consume_dev_timed(m):
        nonlocal launches0
        state["n"] += 1
        if state["n"] == W * S:  # last warm-up shell consumed: the timed region starts next
            torch.cuda.synchronize()
            _lib.check(lib.glb_plan_timing_enable(plan.handle, 1), "timing")
            launches0 = int(lib.glb_kernel_launch_count())
        consume_dev(m)

    if W == 0:
        _lib.check(lib.glb_plan_timing_enable(plan.handle, 1), "timing")
        launches0 = int(lib.glb_kernel_launch_count())
    dev_s, _wall = timed_run(gls_dev, consume_dev_timed)
    launches = int(lib.glb_kernel_launch_count()) - launches0
    ms3 = (C.c_double * 3)()
    l3 = (C.c_int64 * 3)()
    nm = C.c_int64()
    _lib.check(lib.glb_plan_timing_read(plan.handle, ms3, l3, C.byref(nm)), "timing")
    _lib.check(lib.glb_plan_timing_enable(plan.handle, 0), "timing")
    clk = clocks.stop() if rank == 0 else None
    value = world * K * S / dev_s

    # ---- end-to-end arm: NumPy gls in, NumPy maps out (D2H inside the timed region) ----
    stats: dict = {}
    sink = {"x": 0.0}

    def consume_host(m):
        assert isinstance(m, np.ndarray)
        sink["x"] += float(m[0])

    _e2e_dev_s, e2e_wall = timed_run(gls, consume_host, stats)
    e2e_value = world * K * S / e2e_wall
    npix = 12 * nside * nside
    h2d = (lmax + 1) * (NCORR + 1) * 8 * S  # iternorm weights of S shells (the only per-step input)
    d2h = npix * 8 * S

    # ---- what the box can copy: all ranks device -> pinned host at once (the ceiling of e2e) ----
    d2h_ceiling = None
    if world > 1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from probe_d2h import d2h_ceiling as _ceiling

            d2h_ceiling = _ceiling(dev, world, reps=2)
        except Exception as e:
            d2h_ceiling = {"failed": f"{type(e).__name__}: {e}"}

    # ---- the m <-> ring split of one transform (N > 1) and the whole north-star chain ----
    from glass_b200.healpix import clear_plans

    msplit = chain = None
    del gls_dev, plan
    clear_plans()
    torch.cuda.empty_cache()
    if world > 1 and not args.no_msplit:
        try:
            msplit = bench_msplit(dev, rank, world, nside, lmax, dev_s * 1e3 / K if S == 4 else None)
        except Exception as e:  # never take the headline down
            msplit = {"failed": f"{type(e).__name__}: {e}"}
        clear_plans()
        torch.cuda.empty_cache()
    if not args.no_chain and nside == NSIDE:
        try:
            chain = bench_chain(dev, rank, world)
        except Exception as e:
            free, total = torch.cuda.mem_get_info(dev)
            chain = {"failed": f"{type(e).__name__}: {e}", "mem_free_GB": free / 1e9, "torch_allocated_GB": torch.cuda.memory_allocated(dev) / 1e9,
                     "torch_reserved_GB": torch.cuda.memory_reserved(dev) / 1e9}
        clear_plans()
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (Legendre synthesis) ----
    nalm = (lmax + 1) * (lmax + 2) // 2
    ntri = nalm * 2 * nside
    leg_launches = max(int(l3[1]), 1)
    leg_ms = ms3[1] / leg_launches
    maps_per_launch = nm.value / leg_launches
    # F = 4 N_tri (1 + 1/B) per map for B maps on one recurrence
    alg_flop = 4.0 * ntri * (1.0 + 1.0 / maps_per_launch) * maps_per_launch
    achieved = alg_flop / (leg_ms * 1e-3) / 1e12
    fft_ms = ms3[2] / max(int(l3[2]), 1)
    fft_bytes = (4 * nside - 1) * (lmax + 1) * 16 * maps_per_launch + npix * 8 * maps_per_launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    total_ms = dev_s * 1e3 / K
    # DRAM traffic per launch from the committed ncu --set full capture of this same workload
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic.update(json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))))
    except OSError:
        pass

    def traffic_of(kernel):
        t = traffic.get(kernel)
        if t and t.get("nside") == nside and t.get("lmax") == lmax and t.get("maps_per_launch") == int(round(maps_per_launch)):
            return t["bytes_per_launch"]
        return None

    int8 = int(round(maps_per_launch)) == 8  # groups of eight run the contraction on the INT8 tensor cores
    roofline = {
        "kernel": "sht_legendre_ozaki_kernel<32>" if int8 else "sht_legendre_synth_kernel",
        "bound": "fp64",
        "achieved": achieved,
        "peak": peak_tf.value,
        "unit": "TFLOP/s",
        "frac": achieved / peak_tf.value,
        "traffic": traffic_of("sht_legendre_ozaki_kernel" if int8 else "sht_legendre_synth_kernel"),
        "bound_note": (
            "SURVEY 8(d) counts this stage in algorithmic FP64 flops against the FP64 vector peak (tcgen05 has no FP64 mode; DMMA "
            "measured at the same 37 TFLOP/s on the same units, tools/microbench/dmma_mix.cu).  With eight maps per launch only the "
            "recurrence (2 DFMA per l-pair and ring pair, run twice) stays on that pipe: the contraction over l is an exact Ozaki-type "
            "integer product on tcgen05.mma kind::i8 with TMEM accumulators (six base-256 digits per operand, 21 digit products), so "
            "the fraction exceeds 1; what limits the kernel is instruction issue of the digit cutting (ncu: profiles/r02_ncu_int8_legendre.txt), "
            "neither 'hbm' nor 'tensor'"
            if int8
            else "FP64 vector pipe (tcgen05 has no FP64 mode; DMMA measured at the same 37 TFLOP/s on the same units, "
            "tools/microbench/dmma_mix.cu), so neither 'hbm' nor 'tensor' applies"
        ),
        "peak_source": "measured live in this run: register-resident DFMA chains on all SMs "
        "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2)",
        "peak_cublas_dgemm_8192": dgemm_tf,
        "algorithmic_flop_per_launch": alg_flop,
        "maps_per_launch": maps_per_launch,
        "ms_per_launch": leg_ms,
        "share_of_step": ms3[1] / (dev_s * 1e3),
    }
    if int8:
        # executed tensor work: 21 digit products of 128 x (4 B) x 32 MMAs per half tile; N_tri / 2 (ring pair, l-pair) units
        int8_ops = 2.0 * 21 * 4 * maps_per_launch * (ntri / 2.0)
        roofline["tensor_int8"] = {
            "ops_per_launch": int8_ops,
            "achieved_TOPS": int8_ops / (leg_ms * 1e-3) / 1e12,
            "dense_int8_TOPS_estimate": 2.0 * peaks.get("bf16_tflops", 2250.0),
            "estimate_source": "twice the measured dense bf16 rate of MEASURED_PEAKS.json (no INT8 entry there)",
            "note": "upper bound on the executed INT8 work (tiles of silent rings issue no MMA); the tensor pipe is ~18 % busy (ncu)",
        }
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": total_ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "dtype_note": ("FP64 in and out; for groups of eight shells the contraction over l of the Legendre stage is an exact integer product of 48-bit "
                       "fixed-point digits on the INT8 tensor cores (phases within 1e-11 of the FP64 kernel, maps within the 1e-10 bar: tests/test_gpu_int8.py)"),
        "data": "synthetic",
        "config": workload_config(nside, lmax, world),
        "host_binding": (f"rank 0 bound to {len(local_cpus)} GPU-local cores (NVML affinity)" if local_cpus else "none"),
        "clocks": clk,
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h,
            "note": "glass_b200.generate with NumPy gls -> NumPy maps; wall clock incl. device->host copies, max over ranks",
            "d2h_GB/s": e2e_value * npix * 8 / 1e9,
            "d2h_ceiling": d2h_ceiling,
        },
        "gpu_launches": launches,
        "roofline": roofline,
        "stages_ms_per_step": {
            "prep": ms3[0] / K,
            "legendre": ms3[1] / K,
            "ringfft_lognormal": ms3[2] / K,
            "other (draw, combine, host)": total_ms - (ms3[0] + ms3[1] + ms3[2]) / K,
        },
        "roofline_ringfft": {
            "kernel": "sht_ringfft_synth_kernel (3 size classes per launch group)",
            "bound": "hbm",
            "achieved": fft_bytes / (fft_ms * 1e-3) / 1e9,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": fft_bytes / (fft_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": traffic_of("sht_ringfft_synth_kernel"),
        },
    }
    if msplit is not None:
        line["msplit"] = msplit
    if chain is not None:
        line["chain"] = chain
    if world == 1 and not args.no_extra:
        try:
            line["other_stages"] = extra_stage_rooflines(dev, hbm_peak, peak_tf.value)
        except Exception as e:  # secondary numbers must not take the headline down
            line["other_stages"] = {"failed": str(e)}
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(nside, lmax)
        except Exception as e:  # the CPU arm must not take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nside", type=int, default=NSIDE, help="development override (the metric is quoted at 4096)")
    ap.add_argument("--lmax", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary per-stage measurements")
    ap.add_argument("--no-msplit", action="store_true", help="skip the m-split transform measurement (N > 1)")
    ap.add_argument("--no-chain", action="store_true", help="skip the full-chain (config 4 + lensing + 1e9 galaxies) measurement")
    args = ap.parse_args()
    if args.lmax is None:
        args.lmax = 2 * args.nside - 1
    if args.warmup < 3 and args.impl == "b200" and args.nside == NSIDE:
        print("note: W >= 3 warm-up steps are required for a valid number", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
