"""Quick per-stage timing of the scalar synthesis (development probe, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from glass_b200 import _lib
from glass_b200.healpix import get_plan, alm2map_batch

def ev_time(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts)/len(ts)

dev = torch.device("cuda", 0)
cfgs = [(int(a), int(b)) for a, b in (s.split(":") for s in sys.argv[1].split(","))] if len(sys.argv) > 1 else [(1024, 2047)]
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
for nside, lmax in cfgs:
    nalm = (lmax+1)*(lmax+2)//2
    nring = 4*nside-1
    ntri = nalm * 2*nside
    for B in batches:
        g = torch.Generator(device=dev); g.manual_seed(1)
        alm = torch.randn((B, nalm, 2), dtype=torch.float64, device=dev, generator=g)
        alm = torch.view_as_complex(alm).contiguous()
        pl = get_plan(nside, lmax, B, dev)
        phase = torch.zeros((B, nring, lmax+1), dtype=torch.complex128, device=dev)
        mp = torch.empty((B, 12*nside*nside), dtype=torch.float64, device=dev)
        t_leg = ev_time(lambda: _lib.check(pl.lib.glb_debug_alm2phase(pl.handle, alm.data_ptr(), B, phase.data_ptr(), pl.stream_ptr())))
        t_fft = ev_time(lambda: _lib.check(pl.lib.glb_debug_phase2map(pl.handle, phase.data_ptr(), B, mp.data_ptr(), pl.stream_ptr())))
        t_all = ev_time(lambda: alm2map_batch(alm, nside, lmax, out=mp))
        flop = 8.0*ntri*B
        print(f"nside={nside} lmax={lmax} B={B}: alm2phase {t_leg[0]:.2f} ms, phase2map {t_fft[0]:.2f} ms, alm2map {t_all[0]:.2f} ms "
              f"({t_all[0]/B:.2f} ms/map) -> {flop/t_leg[0]/1e9:.1f} TF/s algorithmic(8*Ntri) on legendre", flush=True)
        # sanity: power check
        v = mp[0].var().item(); 
        print("   map var", v, "expected ~", (2*np.arange(lmax+1)+1).sum()/(4*np.pi)*2 if False else float(((2*torch.arange(lmax+1)+1).sum()/(4*np.pi)).item())*1.0)
        del alm, phase, mp
        torch.cuda.empty_cache()
