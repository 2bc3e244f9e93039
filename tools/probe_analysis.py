import sys, torch
sys.path.insert(0, ".")
from glass_b200 import healpix as hp
nside = int(sys.argv[1]); lmax = 2*nside-1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
kap = 0.01 * torch.randn(12*nside*nside, dtype=torch.float64, device=dev, generator=g)
def ev(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best=1e9
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
t = ev(lambda: hp.map2alm(kap, lmax=lmax, pol=False, niter=0))
ntri=(lmax+1)*(lmax+2)//2*2*nside
print(f"nside {nside} map2alm niter=0: {t:.2f} ms -> {8*ntri/t/1e9:.1f} TF/s algorithmic")
