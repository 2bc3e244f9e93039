"""2-GPU (or more) check of the m-split transform; launched by torchrun:
    torchrun --nproc-per-node 2 tests/dist/msplit_check.py [nside ...]
Every rank compares its ring bands with the single-GPU transform (bit-identical) and rank 0
prints timings."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from glass_b200 import _lib  # noqa: E402
from glass_b200.dist import MSplitTransform  # noqa: E402
from glass_b200.healpix import alm2map_batch  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    p2p = "--p2p" in sys.argv  # fused Legendre + transpose over peer memory instead of the NCCL all-to-all
    sizes = [int(a) for a in sys.argv[1:] if a != "--p2p"] or [8, 64, 256]
    ok = True
    for nside in sizes:
        lmax = 2 * nside - 1 if nside > 8 else 3 * nside - 1
        nalm = (lmax + 1) * (lmax + 2) // 2
        for nb in (1, 4) if nside <= 1024 else ((4,) if nside <= 4096 else (1,)):  # nside 8192: one map (memory of the single-GPU reference)
            g = torch.Generator(device=dev)
            g.manual_seed(1234 + nside)  # same alm on every rank
            alm = torch.view_as_complex(torch.randn((nb, nalm, 2), dtype=torch.float64, device=dev, generator=g))
            tr = [(_lib.T_LOGNORMAL, 0.1, 1.0)] * nb if nside <= 1024 else None
            ms = MSplitTransform(nside, lmax, max_batch=nb, device=dev, p2p=p2p)
            out = ms.alm2map(alm, transforms=tr)
            torch.cuda.synchronize()
            ref = alm2map_batch(alm, nside, lmax, transforms=tr)
            covered = 0
            for a, b in ms.pixel_ranges:
                same = torch.equal(out[:, a:b], ref[:, a:b])
                ok &= bool(same)
                covered += b - a
            tot = torch.tensor([covered], device=dev)
            dist.all_reduce(tot)
            ok &= int(tot.item()) == 12 * nside * nside
            full = ms.gather(out)
            ok &= bool(torch.equal(full, ref))
            if p2p:  # both receive buffers, and a changed input through a buffer that was used before
                for scale in (1.0, -0.5, 2.0):
                    out2 = ms.alm2map(alm * scale, transforms=None)
                    torch.cuda.synchronize()
                    ref2 = alm2map_batch(alm * scale, nside, lmax, transforms=None)
                    ok &= all(bool(torch.equal(out2[:, a:b], ref2[:, a:b])) for a, b in ms.pixel_ranges)
            # timing
            for _ in range(2):
                ms.alm2map(alm, transforms=tr, out=out)
            dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3):
                ms.alm2map(alm, transforms=tr, out=out)
            torch.cuda.synchronize(); dist.barrier(); t = (time.perf_counter() - t0) / 3
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3):
                alm2map_batch(alm, nside, lmax, transforms=tr, out=ref)
            torch.cuda.synchronize(); t1 = (time.perf_counter() - t0) / 3
            if rank == 0:
                print(f"nside={nside} lmax={lmax} nb={nb} world={world} {'peer stores' if p2p else 'all-to-all'}: m-split {t*1e3:.2f} ms vs single GPU {t1*1e3:.2f} ms "
                      f"(speed-up {t1/t:.2f}x), bands identical: {ok}", flush=True)
            del ms
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MSPLIT_OK" if int(flag.item()) == 1 else "MSPLIT_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
