import sys, torch
sys.path.insert(0, ".")
import glass_b200
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lmax = 2 * nside - 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
kap = 0.01 * torch.randn(12 * nside * nside, dtype=torch.float64, device=dev, generator=g)
for _ in range(2):
    g1, g2 = glass_b200.shear_from_convergence(kap, lmax, discretized=False, niter=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
torch.cuda.synchronize()
