import time, torch
dev = torch.device("cuda", 0)
n = 12 * 4096 * 4096
src = torch.randn(n, dtype=torch.float64, device=dev)
for trial in range(3):
    t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
    print(f"pinned alloc {n*8/1e9:.2f} GB: {t1-t0:.3f} s")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"   D2H {n*8/dt/1e9:.1f} GB/s")
    del h
# pageable
hp = torch.empty(n, dtype=torch.float64)
torch.cuda.synchronize(); t0 = time.perf_counter(); hp.copy_(src); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pageable D2H {n*8/dt/1e9:.1f} GB/s")
# two concurrent copies on two streams
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): h1.copy_(src, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(src, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"2 concurrent D2H total {2*n*8/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); a = h1.numpy(); x = float(a[0]); print("numpy view", time.perf_counter() - t0)
