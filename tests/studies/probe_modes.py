"""Per-mode accuracy of the nside-4096 synthesis (development probe): Legendre stage against the
80-bit exact-geometry lambda_lm, then the full map."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import healpix_ref as H
from glass_b200 import _lib
from glass_b200.healpix import get_plan, alm2map_batch
import test_gpu_fullsize as tf

NSIDE, LMAX = 4096, 8191
dev = torch.device("cuda", 0)
ri = H.ring_info(NSIDE)
zx, sx = tf.exact_ring_geometry(NSIDE)
nring = 4 * NSIDE - 1
nphi = torch.as_tensor(ri["nphi"], device=dev)
ring = torch.repeat_interleave(torch.arange(nphi.numel(), device=dev), nphi)
j = torch.arange(12 * NSIDE * NSIDE, device=dev) - torch.as_tensor(ri["start"], device=dev)[ring]
nphi_p = nphi[ring]
shifted = torch.as_tensor(ri["shifted"].astype(np.int64), device=dev)[ring]
pl = get_plan(NSIDE, LMAX, 1, dev)
modes = [(0, 0, 1.0 + 0j), (1, 0, -0.7 + 0j), (8000, 0, 0.9 + 0j), (8191, 0, 0.4 + 0j), (8191, 8191, 1.1 + 0.5j),
         (5000, 3000, -0.6 + 0.8j), (8191, 4000, 0.5 + 0.1j), (7000, 6999, 0.2 - 0.9j), (6001, 17, 0.3 + 0.3j), (8190, 0, 1.0 + 0j), (4001, 1, 1.0 + 0j)]
for l, m, a in modes:
    alm = np.zeros((1, H.alm_size(LMAX)), dtype=np.complex128)
    alm[0, H.alm_index(LMAX, l, m)] = a
    d_alm = torch.as_tensor(alm).to(dev)
    phase = torch.zeros((1, nring, LMAX + 1), dtype=torch.complex128, device=dev)
    _lib.check(pl.lib.glb_debug_alm2phase(pl.handle, d_alm.data_ptr(), 1, phase.data_ptr(), pl.stream_ptr()), "alm2phase")
    torch.cuda.synchronize()
    lam = lam_np = tf.lam_single(l, m, zx, sx)
    F = phase[0, :, m].cpu().numpy()
    aa = a if m > 0 else complex(a.real, 0.0)
    want = aa * lam_np
    e = np.abs(F - want)
    # rings beyond mlim are not written (zero): ignore where the reference is negligible
    scale = np.abs(want).max()
    rr = int(e.argmax())
    print(f"mode ({l},{m}): legendre max err {e.max()/scale:.3e} at ring {rr} (|want| there {abs(want[rr])/scale:.2e})", flush=True)
    got = alm2map_batch(d_alm, NSIDE, LMAX)[0]
    lam_t = torch.as_tensor(lam_np, device=dev)[ring]
    if m == 0:
        wantm = a.real * lam_t
    else:
        num = (2 * m * j + m * shifted) % (2 * nphi_p)
        ang = math.pi * num.to(torch.float64) / nphi_p.to(torch.float64)
        wantm = 2.0 * lam_t * (a.real * torch.cos(ang) - a.imag * torch.sin(ang))
    d = (got - wantm).abs()
    ip = int(d.argmax())
    print(f"            map max err {float(d.max())/float(wantm.abs().max()):.3e} at pixel {ip} ring {int(ring[ip])} j {int(j[ip])} nphi {int(nphi_p[ip])}", flush=True)
