"""Per-stage timing of the sampling / lensing kernels (development probe): achieved GB/s vs HBM peak."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import glass_b200
from glass_b200.points import _Population, linear_bias, ARCMIN2_SPHERE
from glass_b200 import healpix as hp

def ev(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

dev = torch.device("cuda", 0)
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
npix = 12 * nside * nside
peak = 6459.3
g = torch.Generator(device=dev); g.manual_seed(1)
delta = torch.expm1(0.5 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g) - 0.125)
vis = torch.rand(npix, dtype=torch.float64, device=dev, generator=g)
ngal_per_pix = 0.083
ngal = ngal_per_pix / (ARCMIN2_SPHERE / npix)
pop = None
def counts():
    global pop
    pop = _Population(delta, vis, ngal * 2, 1.2, linear_bias, False, 42, 0, None, dev)
t = ev(counts)
print(f"K6+K7 counts+scan (nside {nside}, vis, incl. .item() sync): {t:.3f} ms -> {npix*40/t/1e6:.0f} GB/s algorithmic (40 B/pix) = {npix*40/t/1e6/peak*100:.0f}% of HBM peak; total gal {pop.total}")
tot = pop.total
lon = lat = None
def fill():
    global lon, lat
    lon, lat, _ = pop.fill(0, npix, tot, None)
t = ev(fill)
print(f"K8 fill positions: {t:.3f} ms -> {(npix*16+tot*16)/t/1e6:.0f} GB/s ({(npix*16+tot*16)/t/1e6/peak*100:.0f}% of peak) for {tot} galaxies")
# multiplane update
k3 = torch.zeros(npix, dtype=torch.float64, device=dev); k2 = torch.rand(npix, dtype=torch.float64, device=dev)
from glass_b200 import _lib
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
t = ev(lambda: _lib.check(lib.glb_multiplane_update(k3.data_ptr(), k2.data_ptr(), delta.data_ptr(), 0.0, npix, 0.3, 0.01, st)))
print(f"K9 multiplane update: {t:.3f} ms -> {npix*32/t/1e6:.0f} GB/s ({npix*32/t/1e6/peak*100:.0f}% of peak)")
eps = glass_b200.ellipticity_intnorm(tot, 0.27, rng=1, xp=torch)
t = ev(lambda: glass_b200.ellipticity_intnorm(tot, 0.27, rng=1, xp=torch))
print(f"ellipticity_intnorm: {t:.3f} ms -> {tot*16/t/1e6:.0f} GB/s")
t = ev(lambda: glass_b200.galaxy_shear(lon, lat, eps, k2, k2, k2))
print(f"K12 galaxy_shear: {t:.3f} ms -> {tot*72/t/1e6:.0f} GB/s algorithmic (72 B/gal) ({tot*72/t/1e6/peak*100:.0f}% of peak)")
if nside <= 2048 or len(sys.argv) > 2:
    lmax = 2 * nside - 1
    kap = 0.01 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g)
    for niter in (0, 3):
        t = ev(lambda: hp.map2alm(kap, lmax=lmax, pol=False, niter=niter), n=2)
        ntri = (lmax + 1) * (lmax + 2) // 2 * 2 * nside
        print(f"map2alm lmax={lmax} niter={niter}: {t:.2f} ms -> {8*ntri*(1+2*niter)/t/1e9:.1f} TF/s algorithmic")
    alm = hp.map2alm(kap, lmax=lmax, pol=False, niter=0)
    t = ev(lambda: hp.alm2map_spin([alm, None], nside, 2, lmax), n=2)
    print(f"alm2map_spin(2) E-only: {t:.2f} ms -> {16*ntri/t/1e9:.1f} TF/s algorithmic (16 N_tri)")
    t = ev(lambda: glass_b200.shear_from_convergence(kap, lmax, discretized=False), n=2)
    print(f"shear_from_convergence (niter=3): {t:.2f} ms")
