"""Development probe: K6+K7 (counts + offsets), K8 (fill), K12 (galaxy_shear) at nside 4096
through the C-ABI, CUDA-event timed, with a correctness check of the chained scan."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import glass_b200
from glass_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
peak = 6459.3


def ev(fn, n=7, warm=3):
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


nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
npix = 12 * nside * nside
g = torch.Generator(device=dev)
g.manual_seed(7)
delta = torch.expm1(0.5 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g) - 0.125)
vis = (torch.rand(npix, dtype=torch.float64, device=dev, generator=g) > 0.5).double()
counts = torch.empty(npix, dtype=torch.int64, device=dev)
off = torch.empty(npix + 1, dtype=torch.int64, device=dev)
ws = torch.empty(int(lib.glb_points_workspace_bytes(npix)), dtype=torch.uint8, device=dev)

for scale, v, label in ((0.083, None, "0.083 gal/pix"), (0.083, vis, "0.083 gal/pix, half-sky vis"), (3.0, None, "3 gal/pix"), (30.0, None, "30 gal/pix (PTRS)")):
    def k67():
        _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), v.data_ptr() if v is not None else None, 1, 1.2, scale, 0, None,
                                         C.c_uint64(42), C.c_uint32(0), None, counts.data_ptr(), off.data_ptr(), None, 0, None, ws.data_ptr(), st))
    t = ev(k67)
    by = npix * (32 + (8 if v is not None else 0))
    ref = torch.cumsum(counts, 0)
    ok = bool((off[1:] == ref).all().item()) and int(off[0].item()) == 0
    lam = torch.clamp((1.2 * delta + 1) * scale * (v if v is not None else 1.0), min=0)
    print(f"K6+K7 [{label}]: {t:.3f} ms -> {by/t/1e6:.0f} GB/s algorithmic = {by/t/1e6/peak*100:.0f}% of HBM peak; scan ok={ok}; "
          f"mean count {counts.double().mean().item():.5f} vs lambda {lam.mean().item():.5f}; var {counts.double().var().item():.5f}")

_lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, 1.2, 0.083, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                 counts.data_ptr(), off.data_ptr(), None, 0, None, ws.data_ptr(), st))
tot = int(off[-1].item())
lon = torch.empty(tot, dtype=torch.float64, device=dev)
lat = torch.empty(tot, dtype=torch.float64, device=dev)


def k8():
    _lib.check(lib.glb_points_fill(nside, counts.data_ptr(), off.data_ptr(), 0, npix, None, None, C.c_uint64(42), C.c_uint32(0),
                                   lon.data_ptr(), lat.data_ptr(), None, st))


t = ev(k8)
by = npix * 8 + tot * 16
print(f"K8 fill: {t:.3f} ms -> {by/t/1e6:.0f} GB/s ({by/t/1e6/peak*100:.0f}% of peak), {tot} galaxies")
k2 = torch.rand(npix, dtype=torch.float64, device=dev, generator=g)
eps = glass_b200.ellipticity_intnorm(tot, 0.27, rng=1, xp=torch)
res = torch.empty(tot, dtype=torch.complex128, device=dev)
t = ev(lambda: _lib.check(lib.glb_galaxy_shear(nside, lon.data_ptr(), lat.data_ptr(), None, eps.data_ptr(), tot, k2.data_ptr(),
                                               k2.data_ptr(), k2.data_ptr(), 1, res.data_ptr(), st)))
print(f"K12 galaxy_shear: {t:.3f} ms -> {tot*72/t/1e6:.0f} GB/s ({tot*72/t/1e6/peak*100:.0f}% of peak)")
