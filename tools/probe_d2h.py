"""Device -> host copy ceiling of the box, the denominator of the end-to-end (NumPy out) numbers:

    python tools/probe_d2h.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/probe_d2h.py                      # N GPUs copying at the same time

Every rank copies nside-4096 maps (1.6 GB) from its GPU into page-locked host buffers, all ranks
at once; prints one JSON line with the aggregate GB/s (bytes of all ranks / slowest rank's time) for
whole-map copies on one stream, whole-map copies on two streams, and 256 MB chunks on two streams,
with and without binding the rank to its GPU's NUMA-local cores before the buffers are touched."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def d2h_ceiling(dev, world: int, reps: int = 4, npix: int = 12 * 4096 * 4096) -> dict:
    """Aggregate pinned D2H rate over ``world`` ranks copying concurrently (GB/s)."""
    import torch.distributed as dist

    src = torch.empty(npix, dtype=torch.float64, device=dev).normal_()
    bufs = [torch.empty(npix, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    for b in bufs:
        b.zero_()  # first touch on this rank's cores
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    out = {}

    def run(name, fn):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = world * reps * 2 * npix * 8 / float(t[0]) / 1e9

    def one_stream():
        for b in bufs:
            b.copy_(src, non_blocking=True)

    def two_streams():
        for b, s in zip(bufs, streams):
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                b.copy_(src, non_blocking=True)

    def chunks():
        c = (256 << 20) // 8
        for i, a in enumerate(range(0, npix, c)):
            s = streams[i & 1]
            with torch.cuda.stream(s):
                for b in bufs:
                    b[a : a + c].copy_(src[a : a + c], non_blocking=True)

    run("one_stream_GB/s", one_stream)
    run("two_streams_GB/s", two_streams)
    run("chunks_256MB_two_streams_GB/s", chunks)
    return out


def main():
    import torch.distributed as dist

    from glass_b200.sharding import bind_to_local_cpus

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = {"n_gpus": world, "bytes_per_copy": 12 * 4096 * 4096 * 8, "unbound": d2h_ceiling(dev, world)}
    cpus = bind_to_local_cpus(local) if world > 1 else None
    res["host_binding"] = f"{len(cpus)} NUMA-local cores per rank" if cpus else "none"
    if cpus:
        res["bound"] = d2h_ceiling(dev, world)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
