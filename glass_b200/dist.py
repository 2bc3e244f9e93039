"""
glass_b200.dist -- ONE spherical-harmonic synthesis spread over the GPUs of a box
(axis 2): the Legendre stage is sharded by m, the ring FFT and everything in
pixel space by ring band.  The m -> ring transpose between them is either one NCCL all-to-all
of the phase array over NVLink (default), or -- ``p2p=True`` -- fused into the Legendre kernel,
which then stores every F_m(ring) straight into the owning rank's receive buffer through
NVLink peer mappings, so the transfer overlaps the FP64 work and only a barrier remains.  Use it when a single shell has to finish quickly (latency) or for the largest nside;
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

    def __init__(self, nside: int, lmax: int, group=None, max_batch: int = 4, device=None, p2p: bool = False):
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
        self.p2p = bool(p2p)
        if self.p2p:
            self._p2p_setup(min(int(max_batch), 4))

    def _p2p_setup(self, nb_max: int) -> None:
        """Receive buffers in IPC-exportable memory, handles exchanged over the process group,
        peers opened (C ABI: glb_dist_p2p_alloc / glb_dist_p2p_open)."""
        pl = self.plan
        mine = (C.c_ubyte * 64)()
        _lib.check(pl.lib.glb_dist_p2p_alloc(pl.handle, nb_max, mine), "glb_dist_p2p_alloc")
        every = gather_bytes(bytes(mine), self.group, pl.torch_device)
        rows = np.ascontiguousarray(self.rows, dtype=np.int32)
        _lib.check(pl.lib.glb_dist_p2p_open(pl.handle, b"".join(every), rows.ctypes.data), "glb_dist_p2p_open")
        self._flag = torch.zeros(1, dtype=torch.int32, device=pl.torch_device)
        self._buffer = 0
        dist.barrier(group=self.group)  # nobody stores before every receive buffer is zeroed and mapped

    def _alm2map_p2p(self, alms: torch.Tensor, transforms, out: torch.Tensor) -> torch.Tensor:
        pl, nb = self.plan, alms.shape[0]
        buf, self._buffer = self._buffer, self._buffer ^ 1
        with torch.cuda.device(alms.device):
            st = pl.stream_ptr()
            _lib.check(pl.lib.glb_dist_alm2phase_p2p(pl.handle, alms.data_ptr(), nb, buf, st), "glb_dist_alm2phase_p2p")
            # barrier ordered on the stream: when it completes here, the Legendre kernel of every
            # rank has ended, i.e. all stores into this rank's receive buffer are done.  With the two
            # buffers alternating it also keeps a fast rank from overwriting the buffer that a slow
            # rank's ring FFT of the previous transform is still reading.
            dist.all_reduce(self._flag, group=self.group)
            recv = C.c_void_p()
            _lib.check(pl.lib.glb_dist_p2p_recv(pl.handle, buf, C.byref(recv)), "glb_dist_p2p_recv")
            kinds, params, _keep = hp._transform_args(transforms)
            _lib.check(pl.lib.glb_dist_phase2map(pl.handle, recv, nb, out.data_ptr(), kinds, params, st), "glb_dist_phase2map")
        return out

    def alm2map(self, alms: torch.Tensor, transforms=None, out: torch.Tensor | None = None) -> torch.Tensor:
        """alms [nb, nalm] complex128 CUDA (same on every rank), nb in {1, 2, 4}.  Returns
        [nb, npix]; only this rank's ``pixel_ranges`` are written."""
        pl, dev = self.plan, alms.device
        nb = alms.shape[0]
        alms = alms.contiguous()
        if out is None:
            out = torch.zeros((nb, pl.npix), dtype=torch.float64, device=dev)
        if self.p2p:
            return self._alm2map_p2p(alms, transforms, out)
        send = torch.empty((nb, self.nring, self.W), dtype=torch.complex128, device=dev)
        rows_me = self.rows[self.rank]
        recv = torch.empty((nb, self.world, rows_me, self.W), dtype=torch.complex128, device=dev)
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


def gather_bytes(mine: bytes, group=None, device=None) -> list[bytes]:
    """All-gather of one fixed-size byte string per rank (the IPC handles of the receive buffers);
    ``device``: where the staging tensors live (a CUDA device under NCCL, None/cpu under gloo)."""
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, t, group=group)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


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
