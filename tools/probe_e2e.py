import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import glass_b200
from bench import synthetic_gls
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lmax = 2 * nside - 1
n = 32
gls = synthetic_gls(n, lmax, 3)
fields = [glass_b200.grf.Lognormal()] * n
t0 = time.perf_counter()
gen = glass_b200.generate(fields, gls, nside, ncorr=3, rng=42)
last = t0
for i, m in enumerate(gen):
    now = time.perf_counter()
    st = torch._C._host_emptyCache if False else None
    print(f"shell {i:2d} at {now-t0:7.3f} s (+{(now-last)*1e3:7.1f} ms) {type(m).__name__}", flush=True)
    last = now
    del m
