"""K6+K7 (points_count_scan_kernel) tile-size / occupancy variants: python tools/probe_k6_variants.py lib.so [...]
Each library is a full libglassb200 built with -DGLB_PT_ITEMS / -DGLB_PT_MINBLOCKS (tools/microbench/variants)."""
import ctypes as C
import sys

import torch

dev = torch.device("cuda", 0)
nside = 4096
npix = 12 * nside * nside
g = torch.Generator(device=dev)
g.manual_seed(7)
delta = torch.expm1(0.5 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g) - 0.125)
counts = torch.empty(npix, dtype=torch.int64, device=dev)
off = torch.empty(npix + 1, dtype=torch.int64, device=dev)
gpix = torch.empty(int(0.2 * npix), dtype=torch.int64, device=dev)
tot = torch.zeros(1, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.glb_points_workspace_bytes.restype = C.c_size_t
    lib.glb_points_workspace_bytes.argtypes = [i64]
    lib.glb_points_counts.argtypes = [i64, vp, vp, C.c_int, dbl, dbl, C.c_int, vp, C.c_uint64, C.c_uint32, vp, vp, vp, vp, i64, vp, vp, vp]
    ws = torch.empty(int(lib.glb_points_workspace_bytes(npix)), dtype=torch.uint8, device=dev)
    for label, cnt, of, gp, cap in (("scan", counts.data_ptr(), off.data_ptr(), None, 0), ("list", None, None, gpix.data_ptr(), gpix.numel())):
        for scale in (0.083, 3.0):
            if label == "list" and scale > 0.1:
                continue
            best = 1e9
            for _ in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = lib.glb_points_counts(npix, delta.data_ptr(), None, 1, 1.2, scale, 0, None, 42, 0, None, cnt, of, gp, cap, tot.data_ptr(), ws.data_ptr(), st)
                b.record()
                torch.cuda.synchronize()
                assert rc == 0
                best = min(best, a.elapsed_time(b))
            print(f"{path.split('/')[-1]} {label} {scale} gal/pix: {best:.3f} ms, total {int(tot)}", flush=True)
