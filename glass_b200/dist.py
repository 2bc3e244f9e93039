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


# ------------------------------------------------------------------------------------------
# Multi-plane convergence over shell-sharded ranks (axis 1 + lensing).
#
# With contiguous blocks of shells per rank (sharding.shard_shells) the matter planes are
# produced without communication, but the convergence recurrence (glass/lensing.py:580-586)
#     kappa_{i+1} = (1 - t) kappa_{i-1} + t kappa_i + f delta_{i-1}
# runs through the shells in order.  Its state is small -- five scalars and three maps
# (delta3, kappa2, kappa3) -- and the update is one HBM pass (K9, ~1 ms per plane at nside 4096),
# so the recurrence is run as a PIPELINE over the ranks: rank r receives the state from rank
# r-1, adds the planes of its block (keeping a copy of every kappa_i), passes the state on to
# rank r+1 over NVLink (3 x 1.6 GB at nside 4096) and only then starts the expensive per-shell
# work (kappa -> shear transforms, galaxies), which is again communication-free.  The chain
# costs (block recurrence + one hand-off) per rank in sequence -- ~15 ms per hop at nside 4096
# -- against ~0.7 s per shell of transforms that then run in parallel on all ranks.  The
# result is bit-identical to the single-process recurrence: the same kernel sees the same bits.
# ------------------------------------------------------------------------------------------

_MP_SCALARS = ("z2", "z3", "x3", "w3", "r23")
_MP_MAPS = ("delta3", "kappa2", "kappa3")


def send_multi_plane_state(conv, dst: int, *, like: torch.Tensor, group=None) -> None:
    """Send the recurrence state of ``conv`` (a ``MultiPlaneConvergence``) to rank ``dst``.
    ``like`` fixes device/dtype/shape of the maps for a state that has no planes yet."""
    has = conv.kappa2 is not None
    head = torch.tensor([float(getattr(conv, k)) for k in _MP_SCALARS] + [1.0 if has else 0.0], dtype=torch.float64, device=like.device)
    dist.send(head, dst, group=group)
    if has:
        for k in _MP_MAPS:
            dist.send(torch.as_tensor(getattr(conv, k)).to(like.device).contiguous(), dst, group=group)


def recv_multi_plane_state(conv, src: int, *, like: torch.Tensor, group=None) -> None:
    """Receive the state sent by :func:`send_multi_plane_state` into ``conv``."""
    head = torch.empty(len(_MP_SCALARS) + 1, dtype=torch.float64, device=like.device)
    dist.recv(head, src, group=group)
    vals = head.cpu().tolist()
    for k, v in zip(_MP_SCALARS, vals):
        setattr(conv, k, v)
    if vals[-1] != 0.0:
        for k in _MP_MAPS:
            buf = torch.empty_like(like)
            dist.recv(buf, src, group=group)
            conv._set_state_map(k, buf)


def multi_plane_block(conv, deltas, windows, *, group=None, keep=True):
    """Add this rank's block of matter planes to the box-wide multi-plane recurrence.

    ``deltas[i]`` / ``windows[i]`` are the planes of the rank's CONTIGUOUS block of shells
    (``sharding.shard_shells(..., mode="block")``), ranks ordered by redshift.  Returns the list
    of convergence planes ``kappa_i`` after each shell of the block (copies when ``keep``, since
    the recurrence recycles its two buffers, glass/lensing.py:580).  An empty block just
    forwards the state."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(deltas) != len(windows):
        raise ValueError("mismatch between number of planes and windows")
    like = deltas[0] if len(deltas) else conv._like
    if like is None:
        raise ValueError("an empty block needs conv._like (a map-shaped tensor) to size the hand-off")
    if rank > 0:
        recv_multi_plane_state(conv, rank - 1, like=like, group=group)
    kappas = []
    for d, w in zip(deltas, windows):
        conv.add_window(d, w)
        k = conv.kappa
        kappas.append(k.clone() if (keep and isinstance(k, torch.Tensor)) else (k.copy() if keep else k))
    if rank + 1 < world:
        send_multi_plane_state(conv, rank + 1, like=like, group=group)
    return kappas
