"""Lensing-stage timing probe (CUDA events, third call of each):

    python tools/probe_lensing.py [nside] [niter] [nplanes]

shear_from_convergence on one plane and on a stack of ``nplanes`` (the batched refinement syntheses
and the batched spin-2 synthesis), and the spin-2 synthesis alone for 1, 2 and 4 map pairs
(GLB_SPIN4_R=1|2 selects the ring-pairs-per-thread variant of the four-map kernel)."""
import sys

import torch

sys.path.insert(0, ".")
import glass_b200  # noqa: E402
from glass_b200 import healpix as hp  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nplanes = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lmax = 2 * nside - 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
kap = 0.01 * torch.randn((nplanes, 12 * nside * nside), dtype=torch.float64, device=dev, generator=g)


def timed(fn, n=3):
    t = None
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
    return t


t1 = timed(lambda: glass_b200.shear_from_convergence(kap[0], lmax, discretized=False, niter=niter))
print(f"nside {nside} lmax {lmax} niter {niter}: shear_from_convergence, one plane {t1:.1f} ms", flush=True)
if nplanes > 1:
    tn = timed(lambda: glass_b200.shear_from_convergence(kap, lmax, discretized=False, niter=niter))
    print(f"nside {nside} lmax {lmax} niter {niter}: shear_from_convergence, stack of {nplanes}: {tn / nplanes:.1f} ms per plane", flush=True)
nalm = (lmax + 1) * (lmax + 2) // 2
alm = torch.view_as_complex(torch.randn((4, nalm, 2), dtype=torch.float64, device=dev, generator=g))
ntri = nalm * 2 * nside
for nb in (1, 2, 4):
    t = timed(lambda: hp.alm2map_spin_batch(alm[:nb], nside, 2, lmax))
    print(f"spin-2 synthesis (Legendre + ring FFT), {nb} map pair(s): {t / nb:.1f} ms per pair, {16 * ntri * nb / t / 1e9:.1f} TFLOP/s algorithmic", flush=True)
for nb in (1, 4):
    t = timed(lambda: hp.map2alm(list(kap[:nb]) if nb > 1 else kap[0], lmax=lmax, pol=False, niter=0))
    print(f"map2alm niter=0, {nb} map(s): {t / nb:.1f} ms per map, {8 * ntri * nb / t / 1e9:.1f} TFLOP/s algorithmic", flush=True)
