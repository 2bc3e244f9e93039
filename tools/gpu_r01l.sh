set -x
python -m pytest tests/test_gpu_lensing.py tests/test_gpu_sht.py -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import glass_b200
nside, lmax = 4096, 8191
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
kap = 0.01 * torch.randn((4, 12 * nside * nside), dtype=torch.float64, device=dev, generator=g)
for it in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g1, g2 = glass_b200.shear_from_convergence(kap, lmax, discretized=False, niter=3); b.record()
    torch.cuda.synchronize()
    print(f"4 planes batched, nside {nside}: shear_from_convergence {a.elapsed_time(b)/4:.1f} ms per plane (call {it})", flush=True)
PY
