#!/usr/bin/env python
"""Timing of the Gaussian-spectra pre-step on one GPU: table construction,
one C_l <-> C(theta) DGEMM pair, and solve_gaussian_spectra for S lognormal shells at lmax.

    python tools/probe_solver.py [nshell=60] [lmax=8191] [ncorr=3]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glass_b200 as glass  # noqa: E402
from glass_b200 import grf, transformcl as tcl  # noqa: E402


def main():
    nshell = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 8191
    ncorr = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda", 0)
    n = lmax + 1
    l = torch.arange(n, dtype=torch.float64, device=dev)
    base = 1e-2 * (l + 1.0) ** -1.5  # target (lognormal) spectra, monopole non-zero
    spectra = []
    for i in range(nshell):
        for j in range(i, -1, -1):
            spectra.append(0.5 ** (i - j) * base if i - j <= ncorr else base[:0])
    fields = [grf.Lognormal(1.0 - 0.005 * i) for i in range(nshell)]
    ncols = sum(1 for s in spectra if s.shape[0])
    out = {"nshell": nshell, "lmax": lmax, "ncorr": ncorr, "spectra": ncols, "n_padded": 3 * n}

    def sync_time(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t0

    _, out["tables_s"] = sync_time(lambda: (tcl._tables(n, dev), tcl._tables(3 * n, dev)))
    cols = torch.stack([s for s in spectra if s.shape[0]], dim=1)
    big = torch.nn.functional.pad(cols, (0, 0, 0, 2 * n))
    tcl.cltocorr_dev(big)
    (c, t_f) = sync_time(lambda: tcl.cltocorr_dev(big))
    (b, t_b) = sync_time(lambda: tcl.corrtocl_dev(c))
    flop = 2.0 * (3 * n) ** 2 * ncols
    out["cltocorr_ms"], out["corrtocl_ms"] = t_f * 1e3, t_b * 1e3
    out["dgemm_tflops"] = [flop / t_f / 1e12, flop / t_b / 1e12]
    out["roundtrip_err"] = float((b - big).abs().max() / big.abs().max())
    gls, out["solve_s"] = sync_time(lambda: glass.solve_gaussian_spectra(fields, spectra))
    # the solution reproduces the targets: realised spectrum of shell 0
    g = torch.as_tensor(gls[0], device=dev)
    rl = tcl.corrtocl_dev(grf.corr(fields[0], fields[0], tcl.cltocorr_dev(torch.nn.functional.pad(g, (0, 2 * n)))))[:n]
    out["max_rel_err_cl_shell0"] = float(((rl - spectra[0]) / spectra[0]).abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
