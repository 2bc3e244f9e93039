"""
glass_b200.user -- mirror of ``glass/user.py``: ``save_cls`` / ``load_cls`` and the catalogue
writer ``write_catalog`` (glass/user.py:89-205), the sink of the per-galaxy stream
(SURVEY.md 8f, rank 1).

The reference appends every batch to a FITS binary table through ``fitsio`` (not installed
here).  This writer produces the same kind of file -- an empty primary HDU followed by one
BINTABLE extension, one row per galaxy -- with NumPy only, and takes the columns as they come
out of the kernels: CUDA tensors are copied to pinned host memory on a side stream (double
buffered), so the device->host transfer and the file write of batch i overlap with the kernels
of batch i+1 instead of sitting in the sampling loop.  NAXIS2 is patched when the file is closed.
"""

from __future__ import annotations

import os
from contextlib import contextmanager

import numpy as np
import torch

_BLOCK = 2880

# NumPy kind/itemsize -> (FITS TFORM code, big-endian dtype)
_TFORM = {
    ("f", 8): ("D", ">f8"),
    ("f", 4): ("E", ">f4"),
    ("i", 8): ("K", ">i8"),
    ("i", 4): ("J", ">i4"),
    ("i", 2): ("I", ">i2"),
    ("u", 1): ("B", "u1"),
    ("c", 16): ("M", ">c16"),
    ("c", 8): ("C", ">c8"),
}


def save_cls(filename, cls) -> None:
    """Save a list of Cls to file (glass/user.py:41-62): ``values`` and ``split`` in an .npz."""
    cls = [np.asarray(cl.detach().cpu() if isinstance(cl, torch.Tensor) else cl) for cl in cls]
    split = np.cumsum([cl.shape[0] for cl in cls[:-1]])
    values = np.concatenate(cls)
    np.savez(filename, values=values, split=split)


def load_cls(filename):
    """Load a list of Cls from file (glass/user.py:65-86)."""
    with np.load(filename) as npz:
        values = npz["values"]
        split = npz["split"]
    return np.split(values, split)


def _card(key: str, value, comment: str = "") -> bytes:
    if isinstance(value, bool):
        v = ("T" if value else "F").rjust(20)
    elif isinstance(value, (int, np.integer)):
        v = str(int(value)).rjust(20)
    else:
        v = ("'" + str(value).replace("'", "''").ljust(8) + "'").ljust(20)
    s = f"{key:<8}= {v}"
    if comment:
        s += " / " + comment
    return s[:80].ljust(80).encode("ascii")


def _header(cards) -> bytes:
    h = b"".join(cards) + b"END".ljust(80)
    return h + b" " * (-len(h) % _BLOCK)


class _FitsWriter:
    """Appends rows to one BINTABLE extension (glass/user.py:89-166).  ``write(**columns)`` takes
    equally long 1-D arrays (NumPy or torch, host or CUDA); the column set and dtypes are fixed
    by the first call."""

    def __init__(self, fh, ext: str | None = None) -> None:
        self.fh = fh
        self.ext = ext
        self.names: list[str] | None = None
        self.rowtype: np.dtype | None = None
        self.nrows = 0
        self.header_pos = None
        self.header_len = 0
        self._stream = None
        self._slots = [None, None]  # double-buffered pinned staging: (dict name -> tensor, event, n)
        self._turn = 0

    # ---- staging -----------------------------------------------------------------------
    def _flush_slot(self, k: int) -> None:
        slot = self._slots[k]
        if slot is None or slot[2] is None:
            return
        bufs, ev, n = slot
        ev.synchronize()
        self._append_host({name: buf[:n].numpy() for name, buf in bufs.items()})
        self._slots[k] = (bufs, ev, None)

    def _stage_cuda(self, columns: dict) -> None:
        dev = next(t.device for t in columns.values() if isinstance(t, torch.Tensor) and t.is_cuda)
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        k = self._turn
        self._turn ^= 1
        self._flush_slot(k)  # the copy issued two calls ago: long finished, write it out
        n = next(iter(columns.values())).shape[0]
        slot = self._slots[k]
        if slot is None or any(slot[0][name].shape[0] < n for name in columns):
            bufs = {name: torch.empty(max(n, 1), dtype=t.dtype, pin_memory=True) for name, t in columns.items()}
        else:
            bufs = slot[0]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        done = torch.cuda.Event()
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            for name, t in columns.items():
                bufs[name][:n].copy_(t, non_blocking=True)
                t.record_stream(self._stream)
            done.record(self._stream)
        self._slots[k] = (bufs, done, n)
        # the OTHER slot's copy was issued one call ago: write it while this one is in flight
        self._flush_slot(k ^ 1)

    # ---- file --------------------------------------------------------------------------
    def _start_table(self, host_cols: dict) -> None:
        self.names = list(host_cols)
        fields, cards = [], []
        for i, name in enumerate(self.names, 1):
            a = host_cols[name]
            key = (a.dtype.kind, a.dtype.itemsize)
            if key not in _TFORM:
                raise TypeError(f"column {name}: unsupported dtype {a.dtype}")
            code, be = _TFORM[key]
            fields.append((name, be))
            cards += [_card(f"TTYPE{i}", name), _card(f"TFORM{i}", code)]
        self.rowtype = np.dtype(fields)
        head = [
            _card("XTENSION", "BINTABLE", "binary table extension"),
            _card("BITPIX", 8),
            _card("NAXIS", 2),
            _card("NAXIS1", self.rowtype.itemsize, "bytes per row"),
            _card("NAXIS2", 0, "number of rows"),
            _card("PCOUNT", 0),
            _card("GCOUNT", 1),
            _card("TFIELDS", len(self.names)),
        ]
        if self.ext is not None:
            cards.append(_card("EXTNAME", self.ext))
        self.header_pos = self.fh.tell()
        h = _header(head + cards)
        self.header_len = len(h)
        self.fh.write(h)

    def _append_host(self, host_cols: dict) -> None:
        if self.names is None:
            self._start_table(host_cols)
        if list(host_cols) != self.names:
            raise ValueError("columns differ from the first write")
        n = len(next(iter(host_cols.values())))
        rows = np.empty(n, dtype=self.rowtype)
        for name in self.names:
            a = host_cols[name]
            if a.shape != (n,):
                raise ValueError("columns must be one-dimensional and equally long")
            rows[name] = a
        rows.tofile(self.fh)
        self.nrows += n

    def write(self, data=None, /, **columns) -> None:
        """Append rows.  ``data``: a structured NumPy array (written as it is); ``columns``:
        name=array pairs (glass/user.py:131-166)."""
        if data is not None:
            data = np.asarray(data)
            self._drain()
            self._append_host({name: data[name] for name in data.dtype.names})
        if columns:
            cols = {k: (v if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in columns.items()}
            if any(isinstance(v, torch.Tensor) and v.is_cuda for v in cols.values()):
                dev = next(v.device for v in cols.values() if isinstance(v, torch.Tensor) and v.is_cuda)
                cols = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)).to(dev).reshape(-1) for k, v in cols.items()}
                self._stage_cuda(cols)
            else:
                self._drain()
                self._append_host({k: (v.numpy() if isinstance(v, torch.Tensor) else v).reshape(-1) for k, v in cols.items()})

    def _drain(self) -> None:
        # oldest first: the slot that will be reused next holds the older copy
        self._flush_slot(self._turn)
        self._flush_slot(self._turn ^ 1)

    def close(self) -> None:
        self._drain()
        if self.names is None:
            return
        end = self.fh.tell()
        pad = -(end - self.header_pos - self.header_len) % _BLOCK
        self.fh.write(b"\0" * pad)
        self.fh.seek(self.header_pos + 4 * 80)  # the NAXIS2 card
        self.fh.write(_card("NAXIS2", self.nrows, "number of rows"))
        self.fh.seek(0, os.SEEK_END)


@contextmanager
def write_catalog(filename, *, ext: str | None = None):
    """
    Write a catalogue into a FITS file (glass/user.py:169-205)::

        with write_catalog("catalog.fits") as out:
            ...
            out.write(RA=lon, DEC=lat, E1=eps1, E2=eps2, WHT=w)
    """
    with open(filename, "wb") as fh:
        fh.write(_header([_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)]))
        w = _FitsWriter(fh, ext)
        try:
            yield w
        finally:
            w.close()


def read_catalog(filename) -> dict:
    """Read back the first BINTABLE of a file written by :func:`write_catalog` (columns as
    native-endian NumPy arrays).  Minimal reader for tests and quick looks, not a FITS library."""
    with open(filename, "rb") as fh:
        raw = fh.read()

    def parse(pos):
        cards = {}
        while True:
            block = raw[pos : pos + _BLOCK]
            pos += _BLOCK
            for i in range(0, _BLOCK, 80):
                c = block[i : i + 80].decode("ascii")
                if c.startswith("END"):
                    return cards, pos
                if c[8:10] == "= ":
                    v = c[10:].split(" / ")[0].strip()
                    cards[c[:8].strip()] = v[1:-1].rstrip() if v.startswith("'") else v
    _, pos = parse(0)
    cards, pos = parse(pos)
    if cards.get("XTENSION") != "BINTABLE":
        raise ValueError("no binary table extension")
    inv = {code: be for code, be in _TFORM.values()}
    fields = [(cards[f"TTYPE{i}"], inv[cards[f"TFORM{i}"]]) for i in range(1, int(cards["TFIELDS"]) + 1)]
    rows = np.frombuffer(raw, dtype=np.dtype(fields), count=int(cards["NAXIS2"]), offset=pos)
    out = {name: np.ascontiguousarray(rows[name]).astype(rows[name].dtype.newbyteorder("=")) for name, _ in fields}
    out["__extname__"] = cards.get("EXTNAME")
    return out
