"""2-GPU (or more) check of the shell-sharded multi-plane recurrence; launched by torchrun:
    torchrun --nproc-per-node 2 tests/dist/multiplane_check.py [nside] [nshell]
Every rank runs its block through glass_b200.dist.multi_plane_block and compares every kappa_i
with the single-GPU recurrence on the same planes (bit-identical)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import glass_b200  # noqa: E402
from glass_b200.dist import multi_plane_block  # noqa: E402
from glass_b200.sharding import shard_shells  # noqa: E402


class MockCosmology:  # reference tests/fixtures/domain.py:36-97
    Omega_m0 = 0.3
    hubble_distance = 4.4e3

    def H_over_H0(self, z):  # noqa: N802
        return (self.Omega_m0 * (1 + z) ** 3 + 1 - self.Omega_m0) ** 0.5

    def transverse_comoving_distance(self, z, z2=None):
        if z2 is None:
            return self.hubble_distance * np.asarray(z) * 1_000
        return self.hubble_distance * (np.asarray(z2) - np.asarray(z)) * 1_000


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    nside = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nshell = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    npix = 12 * nside * nside
    g = torch.Generator(device=dev)
    g.manual_seed(99)  # the same planes on every rank
    deltas = [torch.randn(npix, dtype=torch.float64, device=dev, generator=g) for _ in range(nshell)]
    dz = 1.0 / (nshell + 1)
    wins = [glass_b200.RadialWindow(np.array([i, i + 1.0, i + 2.0]) * dz, np.array([0.0, 1.0, 0.0]), (i + 1.0) * dz) for i in range(nshell)]
    mine = list(shard_shells(nshell, rank, world))
    conv = glass_b200.MultiPlaneConvergence(MockCosmology())
    conv._like = torch.empty(npix, dtype=torch.float64, device=dev)
    kappas = multi_plane_block(conv, [deltas[i] for i in mine], [wins[i] for i in mine])
    serial = glass_b200.MultiPlaneConvergence(MockCosmology())
    ok = len(kappas) == len(mine)
    for i in range(nshell):
        serial.add_window(deltas[i], wins[i])
        if i in mine:
            ok &= bool(torch.equal(kappas[mine.index(i)], serial.kappa))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIPLANE_OK" if int(flag[0]) == 1 else "MULTIPLANE_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
