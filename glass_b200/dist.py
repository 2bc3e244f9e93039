"""
glass_b200.dist -- ONE spherical-harmonic synthesis spread over the GPUs of a box
(axis 2): the Legendre stage is sharded by m, the ring FFT and everything in
pixel space by ring band, with one NCCL all-to-all of the phase array over NVLink between
them.  Use it when a single shell has to finish quickly (latency) or for the largest nside;
for throughput over many shells the communication-free shell sharding of
``glass_b200.generate(..., shells=...)`` is the better split.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import healpix as hp
from .sharding import msplit_layout, owned_pixel_ranges


class MSplitTransform:
    """alm (replicated on every rank) -> the rank's ring bands of the map."""

    def __init__(self, nside: int, lmax: int, group=None, max_batch: int = 4, device=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nside, self.lmax = int(nside), int(lmax)
        self.plan = hp.Plan(nside, lmax, max_batch=max_batch, device=device)  # private plan: dist state is per plan
        self.layout = msplit_layout(nside, lmax, self.world)
        lay = self.layout
        mine = np.asarray(lay["rings"][self.rank], dtype=np.int32)
        rowmap = np.ascontiguousarray(lay["rowmap"], dtype=np.int32)
        _lib.check(
            self.plan.lib.glb_dist_setup(self.plan.handle, self.world, self.rank, rowmap.ctypes.data, mine.ctypes.data, int(mine.size)),
            "glb_dist_setup",
        )
        self.W = lay["W"]
        self.rows = lay["rows"]
        self.nring = 4 * self.nside - 1
        self.pixel_ranges = owned_pixel_ranges(self.nside, lay, self.rank)

    def alm2map(self, alms: torch.Tensor, transforms=None, out: torch.Tensor | None = None) -> torch.Tensor:
        """alms [nb, nalm] complex128 CUDA (same on every rank), nb in {1, 2, 4}.  Returns
        [nb, npix]; only this rank's ``pixel_ranges`` are written."""
        pl, dev = self.plan, alms.device
        nb = alms.shape[0]
        alms = alms.contiguous()
        send = torch.empty((nb, self.nring, self.W), dtype=torch.complex128, device=dev)
        rows_me = self.rows[self.rank]
        recv = torch.empty((nb, self.world, rows_me, self.W), dtype=torch.complex128, device=dev)
        if out is None:
            out = torch.zeros((nb, pl.npix), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st = pl.stream_ptr()
            _lib.check(pl.lib.glb_dist_alm2phase(pl.handle, alms.data_ptr(), nb, send.data_ptr(), st), "glb_dist_alm2phase")
            in_splits = [r * self.W for r in self.rows]
            out_splits = [rows_me * self.W] * self.world
            for b in range(nb):  # the m -> ring transpose: one all-to-all per map over NVLink
                dist.all_to_all_single(
                    torch.view_as_real(recv[b]).reshape(-1, 2),
                    torch.view_as_real(send[b]).reshape(-1, 2),
                    output_split_sizes=out_splits,
                    input_split_sizes=in_splits,
                    group=self.group,
                )
            kinds, params, _keep = hp._transform_args(transforms)
            _lib.check(
                pl.lib.glb_dist_phase2map(pl.handle, recv.data_ptr(), nb, out.data_ptr(), kinds, params, st),
                "glb_dist_phase2map",
            )
        return out

    def gather(self, maps: torch.Tensor) -> torch.Tensor:
        """Assemble the full maps on every rank (sum of the disjoint bands)."""
        full = torch.zeros_like(maps)
        for a, b in self.pixel_ranges:
            full[:, a:b] = maps[:, a:b]
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        return full
