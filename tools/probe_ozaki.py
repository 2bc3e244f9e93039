"""Legendre stage on the INT8 tensor cores (csrc/sht_ozaki.cu) against the FP64 kernel: accuracy of
the phases F_m(ring) and time per launch (development probe, not the bench).

    python tools/probe_ozaki.py nside:lmax[,nside:lmax...] [maps=4,8] [reps]
"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from glass_b200 import _lib
from glass_b200.healpix import get_plan


def ev_time(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


dev = torch.device("cuda", 0)
cfgs = [(int(a), int(b)) for a, b in (s.split(":") for s in sys.argv[1].split(","))]
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for nside, lmax in cfgs:
    nalm = (lmax + 1) * (lmax + 2) // 2
    nring = 4 * nside - 1
    ntri = nalm * 2 * nside
    pl = get_plan(nside, lmax, 4, dev)
    for B in batches:
        g = torch.Generator(device=dev); g.manual_seed(B)
        l = torch.cat([torch.arange(m, lmax + 1, device=dev) for m in range(lmax + 1)]) if lmax <= 2048 else None
        alm = torch.view_as_complex(torch.randn((B, nalm, 2), dtype=torch.float64, device=dev, generator=g)).contiguous()
        if l is not None:
            alm = alm * (1e-2 * (l + 1.0) ** -1.5).sqrt()  # the bench's power-law spectrum
        ref = torch.zeros((B, nring, lmax + 1), dtype=torch.complex128, device=dev)
        got = torch.full((B, nring, lmax + 1), float("nan"), dtype=torch.complex128, device=dev)

        def fp64():
            for b0 in range(0, B, 4):
                _lib.check(pl.lib.glb_debug_alm2phase(pl.handle, alm[b0:].data_ptr(), 4, ref[b0:].data_ptr(), pl.stream_ptr()), "fp64")

        def int8():
            _lib.check(pl.lib.glb_debug_alm2phase_int8(pl.handle, alm.data_ptr(), B, got.data_ptr(), pl.stream_ptr()), "int8")

        fp64(); int8(); torch.cuda.synchronize()
        scale = ref.abs().amax(dim=(1, 2), keepdim=True)
        err = ((got - ref).abs() / scale).amax(dim=(1, 2))
        per_m = ((got - ref).abs().amax(dim=1) / ref.abs().amax(dim=1).clamp_min(1e-300)).amax(dim=0)
        worst_m = int(per_m.argmax())
        line = f"nside={nside} lmax={lmax} B={B}: max |F_int8 - F_fp64| / max |F| per map = {[f'{e:.1e}' for e in err.tolist()]}, worst m {worst_m} ({float(per_m.max()):.1e} of that m's max)"
        if reps > 0:
            t64, t8 = ev_time(fp64, reps), ev_time(int8, reps)
            line += f"; fp64 {t64:.2f} ms, int8 {t8:.2f} ms ({t64 / t8:.2f}x), algorithmic {4.0 * ntri * (B + B / 4) / t8 / 1e9:.1f} TFLOP/s at the B=4 count"
        print(line, flush=True)
        del alm, ref, got
        torch.cuda.empty_cache()
