"""Array plumbing shared by the host-side mirrors: NumPy <-> CUDA tensor conversion and the
leading-axes broadcasting rule of ``glass/arraytools.py:47-110``."""

from __future__ import annotations

import numpy as np
import torch

from . import healpix as hp


def is_cuda(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def pick_device(*arrays) -> tuple[torch.device, bool]:
    """(device, on_device): on_device is True when any input is a CUDA tensor."""
    for a in arrays:
        if is_cuda(a):
            return a.device, True
    return torch.device("cuda", hp._device_index()), False


def to_dev(x, device, dtype=torch.float64) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    a = np.ascontiguousarray(x)
    if not a.flags.writeable:  # broadcast views: torch warns about wrapping read-only memory
        a = a.copy()
    return torch.as_tensor(a).to(device=device, dtype=dtype).contiguous()


def to_np(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def shape_of(x) -> tuple[int, ...]:
    return tuple(x.shape) if hasattr(x, "shape") else ()


def broadcast_leading_axes(*args):
    """glass/arraytools.py:47-110: broadcast all but the last ``n`` axes of each input.
    Returns (dims, *trailing shapes); the arrays themselves are indexed lazily with
    :func:`take_leading` so nothing full-size is materialised."""
    shapes, trails = [], []
    for a, n in args:
        s = shape_of(a)
        i = len(s) - n
        shapes.append(s[:i])
        trails.append(s[i:])
    dims = tuple(np.broadcast_shapes(*shapes))
    return dims, shapes, trails


def take_leading(a, lead_shape, dims, k):
    """Element k (index tuple over ``dims``) of ``a`` whose leading axes ``lead_shape``
    broadcast to ``dims``."""
    if not lead_shape:
        return a
    off = len(dims) - len(lead_shape)
    idx = tuple(0 if lead_shape[i] == 1 else k[off + i] for i in range(len(lead_shape)))
    return a[idx]


def host_slices(batches):
    """NumPy views for ``(lon, lat, n)`` device batches with ONE device->host copy per underlying
    allocation: consecutive batches that are slices of the same device arrays (the position
    sampler writes a whole population at once) are copied together into fresh host arrays and
    handed out as views of them, instead of one blocking pageable copy per array and batch."""
    base = None  # (lon storage ptr, lat storage ptr) -> host copies
    for lon, lat, n in batches:
        key = (lon.untyped_storage().data_ptr(), lat.untyped_storage().data_ptr())
        if base is None or base[0] != key:
            def whole(t):
                full = torch.empty(0, dtype=t.dtype, device=t.device).set_(t.untyped_storage())
                return full.cpu().numpy()

            base = (key, whole(lon), whole(lat))
        o1, o2 = lon.storage_offset(), lat.storage_offset()
        yield base[1][o1 : o1 + n], base[2][o2 : o2 + n], n


def nvtx(name: str):
    """Decorator: an NVTX range around a public entry point, so that a timeline (nsys, ncu --nvtx) shows the stages of the
    user loop by their GLASS names.  A few hundred nanoseconds per call; nothing when no tool listens."""
    import functools

    def wrap(fn):
        @functools.wraps(fn)
        def inner(*a, **k):
            if not torch.cuda.is_available():  # (the host-flow tests of the CPU suite)
                return fn(*a, **k)
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()

        return inner

    return wrap
