"""shear_from_convergence timing probe: python tools/probe_lensing.py [nside] [niter] (second call is timed)."""
import sys, torch
sys.path.insert(0, ".")
import glass_b200
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lmax = 2 * nside - 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
kap = 0.01 * torch.randn(12 * nside * nside, dtype=torch.float64, device=dev, generator=g)
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g1, g2 = glass_b200.shear_from_convergence(kap, lmax, discretized=False, niter=niter)
    b.record()
    torch.cuda.synchronize()
    print(f"nside {nside} lmax {lmax} niter {niter} call {it}: shear_from_convergence {a.elapsed_time(b):.1f} ms", flush=True)
