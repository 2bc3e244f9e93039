"""CPU tests (-m "not gpu"): host-side logic of glass_b200 (no kernel launches) and the C ABI."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import synthetic_gls

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/glass_b200.h is the contract: every glb_* function it declares must be
    exported by libglassb200.so and bound in glass_b200._lib.SIGNATURES."""
    from glass_b200 import _lib, build

    build.build()
    hdr = open(os.path.join(ROOT, "include", "glass_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(glb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.fspath(_lib.lib_path()))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib2 = _lib.load()
    assert lib2.glb_version().startswith(b"glass_b200")
    assert lib2.glb_status_string(-10) == b"covariance matrix is not positive definite"


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch

    import glass_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception, match="CUDA device"):
        next(glass_b200.generate([glass_b200.grf.Normal()], [np.ones(4)], 4))
    with pytest.raises(Exception, match="CUDA device"):
        glass_b200.healpix.alm2map(np.zeros(6, dtype=complex), 2, pol=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "glass_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_iternorm_cls2cov_host_mirror():
    """cls2cov (host) bit-exact against the reference's rows; the product's iternorm recursion -- K1's
    per-multipole body, csrc/iternorm_core.cuh, compiled for the host -- against the weights the
    reference's source produced (tolerance: another summation order than NumPy's matmul)."""
    import glass_b200
    from helpers import native_iternorm_rows

    def close(got, want):
        got, want = np.stack(got), np.asarray(want)
        return got.shape == want.shape and np.abs(got - want).max() <= 1e-14 * max(np.abs(want).max(), 1.0)

    cov = np.array([[1.0, 0.2, 0.1], [0.2, 0.5, 0.2], [0.1, 0.2, 0.3]])
    for k in (0, 1, 2):
        rows = [np.pad(cov[i, i::-1][: min(i, k) + 1], (0, k + 1 - min(i + 1, k + 1))) for i in range(3)]
        assert close(native_iternorm_rows(rows), GOLD[f"iternorm_k{k}"])
    for name, (nshell, lmax, ncorr, ragged) in {"a": (4, 12, 2, False), "b": (5, 9, None, False), "c": (4, 10, 1, True)}.items():
        nc = nshell - 1 if ncorr is None else ncorr
        gls = synthetic_gls(nshell, lmax, nc, ragged)
        got = np.stack([c.copy() for c in glass_b200.cls2cov(gls, lmax + 1, nshell, nc)])
        assert np.array_equal(got, GOLD[f"cls2cov_{name}"])
        assert close(native_iternorm_rows(c.copy() for c in glass_b200.cls2cov(gls, lmax + 1, nshell, nc)), GOLD[f"iternorm_{name}"])
    # the same buffer is re-yielded (tests/core/test_fields.py:239-242)
    gen = glass_b200.cls2cov([np.array(a) for a in ([1.0, 0.5, 0.3], [0.8, 0.4, 0.2], [0.7, 0.6, 0.1], [0.9, 0.5, 0.3], [0.6, 0.3, 0.2], [0.8, 0.7, 0.4])], 3, 3, 2)
    c1 = next(gen)
    c1_copy = c1.copy()
    c2 = next(gen)
    assert c1 is c2 and np.array_equal(c1_copy[:, 0], [0.5, 0.25, 0.15]) and np.array_equal(c1[:, 0], [0.4, 0.2, 0.1])
    with pytest.raises(ValueError, match="negative values in cl"):
        next(glass_b200.cls2cov([np.array([-1.0, 0.5, 0.3])], 3, 1, 0))
    with pytest.raises(ValueError, match="empty covariance"):
        list(glass_b200.iternorm([np.ones(0)]))
    with pytest.raises(ValueError, match="not positive definite"):
        native_iternorm_rows([np.array([1.0, 0.0]), np.array([0.1, 1.0])])
    # degenerate shells (s = 0: identically zero monopole, perfectly correlated shells) follow the reference
    from oracle import glass_ref as G

    rows = [np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]), np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 0.0]]),
            np.array([[0.0, 0.0, 0.0], [2.0, 1.0, 1.0]]), np.array([[0.0, 0.0, 0.0], [3.0, 1.0, 1.0]])]
    w = np.stack(native_iternorm_rows(rows))
    assert np.array_equal(w, np.stack(G.iternorm(rows))) and not np.isnan(w).any()


def test_getcl_multalm_misc():
    """tests/core/test_fields.py:441-465, 649-655; tests/core/test_harmonics.py:15-52."""
    import glass_b200
    from glass_b200.fields import _glass_to_healpix_alm, _inv_triangle_number

    cls = [np.array([i, j], dtype=float) for i in range(10) for j in range(i, -1, -1)]
    for i in range(10):
        for j in range(10):
            assert np.array_equal(np.sort(glass_b200.getcl(cls, i, j)), [min(i, j), max(i, j)])
            assert np.array_equal(glass_b200.getcl(cls, i, j, lmax=0), [max(i, j)])
            r = glass_b200.getcl(cls, i, j, lmax=50)
            assert r.shape[0] == 51 and np.all(r[2:] == 0)
    assert np.array_equal(glass_b200.multalm(np.arange(1.0, 7.0), np.array([2.0, 0.5, 1.0])), [2.0, 1.0, 1.5, 4.0, 5.0, 6.0])
    assert glass_b200.multalm(np.array([]), np.array([])).size == 0
    inp = np.array([0, 10, 11, 20, 21, 22, 30, 31, 32, 33], dtype=complex)
    assert np.array_equal(_glass_to_healpix_alm(inp), np.array([0, 10, 20, 30, 11, 21, 31, 22, 32, 33], dtype=complex))
    for n in range(2000):
        assert _inv_triangle_number(n * (n + 1) // 2) == n
    for t in [2, 4, 5, 7, 8, 9, 11, 12, 13, 14, 16, 17, 18, 19, 20]:
        with pytest.raises(ValueError, match="not a triangle number"):
            _inv_triangle_number(t)
    with pytest.raises(ValueError, match="invalid number of spectra: 4"):
        glass_b200.nfields_from_nspectra(4)
    x = GOLD["lognormal_x"]
    assert np.array_equal(glass_b200.grf.Lognormal(0.7)(x.copy(), 0.35), GOLD["lognormal_y"])
    assert np.array_equal(glass_b200.grf.SquaredNormal(0.3, 1.5)(x.copy(), 0.35), GOLD["sqnormal_y"])
    assert glass_b200.points.ARCMIN2_SPHERE == float(GOLD["ARCMIN2_SPHERE"])
    shells = [glass_b200.RadialWindow(np.zeros(2), np.zeros(2), z) for z in (0.5, 1.0)]
    lf = glass_b200.lognormal_fields(shells, lambda z: 2 * z)
    assert [f.lamda for f in lf] == [1.0, 2.0] and len(glass_b200.gaussian_fields(shells)) == 2


def _gloo_worker(rank, world, port, nshell, ncorr, q):
    import torch.distributed as dist

    from glass_b200.sharding import max_over_ranks, neighbours_needed, shard_shells

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = list(shard_shells(nshell, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    tmax = max_over_ranks(1.0 + rank)
    need = sorted(neighbours_needed(mine, ncorr))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, gathered, tmax, need))


def test_shell_sharding_world2_gloo():
    """N>1 host logic on CPU: ranks partition the shells, timing is the max over ranks."""
    import torch.multiprocessing as mp

    world, nshell, ncorr = 2, 11, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, nshell, ncorr, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, tmax, need in res:
        allsh = sorted(s for g in gathered for s in g)
        assert allsh == list(range(nshell))  # a partition: every shell exactly once
        assert set(gathered[0]).isdisjoint(gathered[1])
        assert tmax == 2.0  # max over ranks
        mine = gathered[rank]
        assert set(mine) <= set(need) and all(0 <= s < nshell for s in need)
        assert all(any(0 <= j - s <= ncorr for j in mine) for s in need)
        # contiguous blocks: only the ncorr shells before the block are regenerated
        assert mine == list(range(mine[0], mine[-1] + 1)) and len(need) <= len(mine) + ncorr


def _msplit_worker(rank, world, port, nside, lmax, q):
    import torch
    import torch.distributed as dist

    from glass_b200.sharding import msplit_layout

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lay = msplit_layout(nside, lmax, world)
    nring, W = 4 * nside - 1, lay["W"]
    # what the Legendre stage of this rank writes: value(ring, m) at [rowmap[ring], m // world]
    send = torch.full((nring, W), -1.0, dtype=torch.float64)
    for ring in range(nring):
        for m in range(rank, lmax + 1, world):
            send[int(lay["rowmap"][ring]), m // world] = ring * 10000.0 + m
    rows_me = lay["rows"][rank]
    recv = torch.empty((world, rows_me, W), dtype=torch.float64)
    dist.all_to_all_single(recv.view(-1), send.view(-1), output_split_sizes=[rows_me * W] * world,
                           input_split_sizes=[r * W for r in lay["rows"]])
    # what the Fourier stage of this rank reads: F(ring, m) = recv[m % world, local row, m // world]
    ok = True
    for row, ring in enumerate(lay["rings"][rank]):
        for m in range(lmax + 1):
            ok &= recv[m % world, row, m // world].item() == ring * 10000.0 + m
    # the handle exchange of the peer-store form: one fixed-size byte string per rank, in rank order
    from glass_b200.dist import gather_bytes

    every = gather_bytes(bytes([rank + 1]) * 128, None, None)
    ok &= every == [bytes([r + 1]) * 128 for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def _p2p_host_flow_worker(rank, world, port, q):
    """MSplitTransform(p2p=True) end to end on the HOST side, under gloo, against a fake library that
    records the C-ABI calls: handle exchange, open, alternating buffers, barrier, phase2map."""
    import contextlib
    import ctypes as C

    import torch
    import torch.distributed as dist

    import glass_b200.dist as gd
    import glass_b200.healpix as hp

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    calls = []

    class FakeLib:
        def glb_dist_setup(self, handle, w, r, rowmap, mine, n):
            calls.append(("setup", w, r, n))
            return 0

        def glb_dist_p2p_alloc(self, handle, nb, out):
            assert len(out) == 64
            for i in range(64):
                out[i] = 16 * (rank + 1)
            calls.append(("alloc", nb))
            return 0

        def glb_dist_p2p_open(self, handle, blob, rows_ptr):
            assert blob == b"".join(bytes([16 * (r + 1)]) * 64 for r in range(world))
            rows = (C.c_int32 * world).from_address(rows_ptr)
            calls.append(("open", list(rows)))
            return 0

        def glb_dist_alm2phase_p2p(self, handle, alm, nb, buf, st):
            calls.append(("legendre", nb, buf))
            return 0

        def glb_dist_p2p_recv(self, handle, buf, out):
            out._obj.value = 1000 + buf
            calls.append(("recv", buf))
            return 0

        def glb_dist_phase2map(self, handle, recv, nb, out, kinds, params, st):
            calls.append(("fft", recv.value, nb))
            return 0

    class FakePlan:
        def __init__(self, nside, lmax, max_batch=1, device=None):
            self.lib, self.handle, self.npix = FakeLib(), C.c_void_p(1), 12 * nside * nside
            self.torch_device = torch.device("cpu")

        def stream_ptr(self):
            return 0

    hp.Plan = FakePlan
    torch.cuda.device = lambda d: contextlib.nullcontext()
    ms = gd.MSplitTransform(8, 15, max_batch=4, p2p=True)
    alm = torch.zeros((2, 136), dtype=torch.complex128)
    for _ in range(3):
        out = ms.alm2map(alm)
    ok = out.shape == (2, 768)
    rows = ms.layout["rows"]
    ok &= calls[:3] == [("setup", world, rank, rows[rank]), ("alloc", 4), ("open", rows)]
    want = []
    for i in range(3):
        want += [("legendre", 2, i % 2), ("recv", i % 2), ("fft", 1000 + i % 2, 2)]
    ok &= calls[3:] == want
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_msplit_peer_store_host_flow_world2_gloo():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_p2p_host_flow_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _r, ok in res)


def test_msplit_peer_store_addressing():
    """Address algebra of the fused Legendre + transpose (sht_legendre_synth_kernel<.., P2P>):
    writer g puts F(ring, m) at ((b*world + g)*rows_d + (row - rowstart[d]))*W + m // world of rank
    d's receive buffer, d = owner of the ring's row.  Replayed on the host for every (g, ring, m,
    b): each rank's buffer must end up exactly as the all-to-all form delivers it,
    [b][src rank][local row][slot], every needed entry written once."""
    from glass_b200.sharding import msplit_layout

    for nside, lmax, world, nb in [(4, 9, 2, 2), (8, 20, 3, 1), (16, 31, 8, 4), (2, 5, 2, 1)]:
        lay = msplit_layout(nside, lmax, world)
        nring, W, rows = 4 * nside - 1, lay["W"], lay["rows"]
        rowstart = np.concatenate([[0], np.cumsum(rows)])
        assert rowstart[-1] == nring
        bufs = [np.full(nb * world * max(rows[d], 1) * W, np.nan) for d in range(world)]
        for g in range(world):  # the writer's kernel
            for ring in range(nring):
                row = int(lay["rowmap"][ring])
                d = 0
                while row >= rowstart[d + 1]:
                    d += 1
                for m in range(g, lmax + 1, world):
                    for b in range(nb):
                        off = ((b * world + g) * rows[d] + (row - rowstart[d])) * W + m // world
                        assert np.isnan(bufs[d][off])  # nobody else writes here
                        bufs[d][off] = (b * 1000 + ring) * 10000.0 + m
        for d in range(world):  # the reader: glb_dist_phase2map's view of its receive buffer
            recv = bufs[d].reshape(nb, world, max(rows[d], 1), W)
            for lr, ring in enumerate(lay["rings"][d]):
                for m in range(lmax + 1):
                    for b in range(nb):
                        assert recv[b, m % world, lr, m // world] == (b * 1000 + ring) * 10000.0 + m


def test_msplit_alltoall_layout_world2_gloo():
    """m -> ring transpose: after the all-to-all every rank holds all m of its own rings."""
    import torch.multiprocessing as mp

    world, nside, lmax = 2, 4, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_msplit_worker, args=(r, world, port, nside, lmax, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _r, ok in res)


def test_msplit_layout_partitions():
    from glass_b200.sharding import msplit_layout, owned_pixel_ranges

    for nside, world in [(4, 2), (8, 3), (64, 8), (48, 5), (2, 2), (4096, 8)]:
        lay = msplit_layout(nside, 2 * nside - 1, world)
        nring = 4 * nside - 1
        assert sorted(r for rs in lay["rings"] for r in rs) == list(range(nring))
        assert sorted(lay["rowmap"]) == list(range(nring))
        segs = sorted(s for d in range(world) for s in owned_pixel_ranges(nside, lay, d))
        assert segs[0][0] == 0 and segs[-1][1] == 12 * nside**2
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        # north/south mirrors stay on the same rank
        for rs in lay["rings"]:
            s = set(rs)
            assert all((nring - 1 - r) in s for r in rs)


def test_discretized_and_effective_cls_golden():
    """glass/fields.py:239-300 and 607-694 against vectors made by executing the reference
    (its healpy pixel window replaced by a supplied array)."""
    import os

    import glass_b200
    from helpers import synthetic_gls

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    gls = synthetic_gls(4, 12, 3)
    for tag, kw in {"lmax": {"lmax": 8}, "ncorr": {"ncorr": 1}, "all": {"lmax": 9, "ncorr": 2, "nside": 4, "pixwin": gold["dcl_pw"]}}.items():
        res = glass_b200.discretized_cls(gls, **kw)
        assert np.array_equal(np.array([r.shape[0] for r in res]), gold[f"dcl_{tag}_len"])
        assert np.array_equal(np.concatenate(res), gold[f"dcl_{tag}"])
    assert glass_b200.discretized_cls([]) == []
    # without a table the window comes from hp.pixwin(nside, lmax=lmax) (glass/fields.py:288-289)
    import glass_b200.fields as F

    asked = []
    saved = F.hp.pixwin
    F.hp.pixwin = lambda ns, lmax=None, pol=False: asked.append((ns, lmax)) or gold["dcl_pw"]
    try:
        res = glass_b200.discretized_cls(gls, lmax=9, ncorr=2, nside=4)
    finally:
        F.hp.pixwin = saved
    assert asked == [(4, 9)] and np.array_equal(np.concatenate(res), gold["dcl_all"])
    assert np.array_equal(glass_b200.effective_cls(gls, gold["ecl_w1"]), gold["ecl_auto"])
    assert np.array_equal(glass_b200.effective_cls(gls, gold["ecl_w1"], gold["ecl_w2"], lmax=7), gold["ecl_cross"])
    with pytest.raises(ValueError, match="shape mismatch between fields and weights1"):
        glass_b200.effective_cls(gls, np.ones((3, 2)))


def test_spectra_helpers_and_position_weights_golden():
    """Spectra-order helpers (glass/fields.py:563-604, 897-1052) and position_weights
    (glass/points.py:610-651) against vectors from executing the reference's source
    (tests/golden/make_golden.py --spectra): bit-identical."""
    import glass_b200 as glass

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_spectra.npz"))
    for n in (1, 2, 3, 5):
        assert np.array_equal(glass.spectra_indices(n), g[f"indices_{n}"])
    labels = list(np.arange(15))
    assert np.array_equal(np.array([(i, j, int(c)) for i, j, c in glass.enumerate_spectra(labels)]), g["enum_15"])
    assert np.array_equal(np.array(glass.glass_to_healpix_spectra(labels)), g["g2h_15"])
    assert np.array_equal(np.array(glass.healpix_to_glass_spectra(labels)), g["h2g_15"])
    assert glass.healpix_to_glass_spectra(glass.glass_to_healpix_spectra(labels)) == labels
    with pytest.raises(ValueError, match="invalid number of spectra: 4"):
        glass.glass_to_healpix_spectra(labels[:4])
    assert np.array_equal(np.array([glass.lognormal_shift_hilbert2011(float(z)) for z in g["hilbert_z"]]), g["hilbert_shift"])
    gls = np.split(g["cov_gls"], np.cumsum(g["cov_gls_len"])[:-1])
    assert np.array_equal(glass.cov_from_spectra(gls), g["cov_full"])
    assert np.array_equal(glass.cov_from_spectra(gls, lmax=5), g["cov_lmax5"])
    assert np.array_equal(glass.cov_from_spectra(gls, lmax=20), g["cov_lmax20"])
    assert glass.check_posdef_spectra(gls) == bool(g["posdef_good"]) is True
    bad = [np.array([1.0, 1.0]), np.array([1.0, 1.0]), np.array([0.5, 1.5])]
    assert glass.check_posdef_spectra(bad) == bool(g["posdef_bad"]) is False
    assert np.array_equal(glass.position_weights(g["pw_d1"]), g["pw_1"])
    assert np.array_equal(glass.position_weights(g["pw_d1"], g["pw_b1"]), g["pw_1b"])
    assert np.array_equal(glass.position_weights(g["pw_d1"], 1.7), g["pw_1f"])
    assert np.array_equal(glass.position_weights(g["pw_d2"], g["pw_b2"]), g["pw_2b"])
    assert np.array_equal(glass.position_weights(g["pw_d2"], g["pw_b1"]), g["pw_2b1"])
    for tag in ("", "_wide"):
        w = glass.RadialWindow(g[f"eb{tag}_za"], g[f"eb{tag}_wa"], 1.0)
        assert np.array_equal(np.asarray(glass.effective_bias(g["eb_z"], g["eb_bz"], w)), g[f"eb{tag}"])
    import torch

    t = glass.position_weights(torch.as_tensor(g["pw_d2"]), torch.as_tensor(g["pw_b2"]))
    assert np.allclose(t.numpy(), g["pw_2b"], rtol=1e-14, atol=0)  # torch sums in another order


@pytest.fixture()
def transforms_on_cpu(monkeypatch):
    """The spectra solver is device-agnostic torch + GEMM; the PRODUCT insists on a CUDA device
    (transformcl._compute_device raises without one).  For the CPU suite the device choice is
    patched so that the same code runs on CPU tensors."""
    import torch

    import glass_b200.transformcl as tcl

    monkeypatch.setattr(tcl, "_compute_device", lambda *a: (torch.device("cpu"), False))
    tcl.clear_tables()
    yield tcl
    tcl.clear_tables()


def test_transformcl_needs_cuda():
    import torch

    import glass_b200
    from glass_b200._lib import GlassB200Error

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(GlassB200Error, match="no CPU fallback"):
        glass_b200.transformcl.cltocorr(np.ones(8))
    with pytest.raises(GlassB200Error, match="no CPU fallback"):
        glass_b200.grf.solve(np.ones(8), glass_b200.grf.Lognormal())


def test_transformcl_pair(transforms_on_cpu):
    """cltocorr / corrtocl (the transformcl interface, glass/grf/_solver.py:100-130): definition
    by direct Legendre sums, exact inverse, Gauss-Legendre projection of a non-polynomial function,
    columns transformed together."""
    from scipy.special import eval_legendre

    tcl = transforms_on_cpu
    for n in (1, 2, 3, 8, 101, 400):
        rng = np.random.default_rng(n)
        cl = rng.standard_normal(n) / (1 + np.arange(n)) ** 2
        c = tcl.cltocorr(cl)
        x = np.cos(tcl.theta(n))
        direct = sum((2 * l + 1) / (4 * np.pi) * cl[l] * eval_legendre(l, x) for l in range(n))
        assert np.abs(c - direct).max() <= 1e-13 * max(np.abs(direct).max(), 1e-300)
        assert np.abs(tcl.corrtocl(c) - cl).max() <= 1e-14 * np.abs(cl).max()
    both = tcl.cltocorr(np.stack([cl, -2 * cl], 1))
    assert np.allclose(both[:, 0], c, rtol=1e-13, atol=1e-18) and np.allclose(both[:, 1], -2 * c, rtol=1e-13, atol=1e-18)
    n = 64
    f = np.exp(np.cos(tcl.theta(n)))
    xg, wg = np.polynomial.legendre.leggauss(200)
    ref = np.array([2 * np.pi * np.sum(wg * np.exp(xg) * eval_legendre(l, xg)) for l in range(20)])
    assert np.abs(tcl.corrtocl(f)[:20] - ref).max() < 1e-12
    assert abs(tcl.cltovar(cl) - np.sum((2 * np.arange(cl.size) + 1) / (4 * np.pi) * cl)) < 1e-18
    assert tcl.cltocorr(np.zeros(0)).shape == (0,)
    with pytest.raises(NotImplementedError):
        tcl.cltocorr(cl, closed=True)


def test_grf_solve_reference_properties(transforms_on_cpu):
    """The reference's own solver tests (tests/core/grf/test_solver.py:24-131) and the transformation
    pairs of tests/core/grf/test_transformations.py, on the batched GPU formulation."""
    from glass_b200 import grf

    lmax = 100
    ell = np.arange(lmax + 1)
    cl = 1e-2 / (2 * ell + 1) ** 2
    rng = np.random.default_rng(42)
    t = grf.Lognormal(rng.random())
    # test_one_transformation, test_pad, test_initial, test_no_iterations
    assert np.array_equal(grf.solve(cl, t)[0], grf.solve(cl, t, t)[0])
    assert grf.solve(cl, t, pad=2 * cl.shape[0])[1].shape[0] == 3 * cl.shape[0]
    with pytest.raises(ValueError, match="pad must be a positive integer"):
        grf.solve(cl, t, pad=-1)
    assert np.array_equal(grf.solve(cl, t)[0], grf.solve(cl, t, initial=grf.compute(cl, t))[0])
    assert np.array_equal(grf.compute(cl, grf.Lognormal()), grf.solve(cl, grf.Lognormal(), maxiter=0)[0])
    # test_lognormal
    gl0, cltol = rng.random(), 1e-7
    gl, cl_, info = grf.solve(cl, grf.Lognormal(), grf.Lognormal(), monopole=gl0, cltol=cltol)
    assert info > 0 and gl[0] == gl0
    assert np.allclose(cl_[1 : cl.shape[0]], cl[1:], atol=0.0, rtol=cltol)
    assert np.allclose(grf.compute(cl_, grf.Lognormal(), grf.Lognormal())[1 : gl.shape[0]], gl[1:])
    # test_monopole
    c2 = cl.copy()
    c2[0] = rng.random()
    gl, cl_out, _ = grf.solve(c2, grf.Lognormal(), monopole=None, gltol=1e-8)
    assert gl[0] != 0.0 and np.allclose(cl_out[0], c2[0])
    gl, cl_out, _ = grf.solve(c2, grf.Lognormal(), monopole=gl0, gltol=1e-8)
    assert gl[0] == gl0 and not np.allclose(cl_out[0], c2[0])
    # corr / icorr / dcorr pairs: inverse and derivative relations, NotImplemented dispatch
    x = np.linspace(-0.2, 0.5, 29)
    for t1, t2 in [(grf.Normal(), grf.Normal()), (grf.Lognormal(0.6), grf.Lognormal(1.4)), (grf.Lognormal(0.6), grf.Normal()),
                   (grf.Normal(), grf.Lognormal(0.6)), (grf.SquaredNormal(0.8, 1.2), grf.SquaredNormal(0.6, 0.7))]:
        y = grf.corr(t1, t2, x)
        assert np.allclose(grf.icorr(t1, t2, y), x, rtol=1e-12, atol=1e-14)
        h = 1e-6
        assert np.allclose((grf.corr(t1, t2, x + h) - grf.corr(t1, t2, x - h)) / (2 * h), grf.dcorr(t1, t2, x), rtol=1e-7)
    with pytest.raises(NotImplementedError, match="SquaredNormal x Normal"):
        grf.corr(grf.SquaredNormal(0.8), grf.Normal(), x)


def test_multi_plane_host_scalars_golden(monkeypatch):
    """The REAL glass_b200.MultiPlaneConvergence on CPU tensors with only the K9 kernel replaced
    (by the reference's three NumPy passes acting on the same buffers): window weight, extrapolation
    law t, lensing weight f, buffer cycling -- against the kappa planes the reference's own source
    produced with MockCosmology (golden mpc_*), bit for bit; plus the error of a non-increasing
    source redshift."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.lensing as L

    def arr(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    class FakeLib:
        def glb_multiplane_update(self, k3, k2, d2, _zero, n, t, f, st):
            kappa3, kappa2 = arr(k3, n), arr(k2, n)
            kappa3 *= 1 - np.asarray(t)  # glass/lensing.py:584-586
            kappa3 += np.asarray(t) * kappa2
            if d2 is not None:
                kappa3 += np.asarray(f) * arr(d2, n)
            return 0

    monkeypatch.setattr(L._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(L.A, "pick_device", lambda *a: (torch.device("cpu"), True))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    deltas = gold["mpc_deltas"]
    conv = L.MultiPlaneConvergence(_MockCosmo())
    assert conv.kappa is None and conv.delta is None
    for i in range(deltas.shape[0]):
        w = types.SimpleNamespace(za=np.array([i, i + 1.0, i + 2.0]), wa=np.array([0.0, 1.0, 0.0]), zeff=i + 1.0)
        conv.add_window(torch.as_tensor(deltas[i].copy()), w)
        assert np.array_equal(conv.kappa.numpy(), gold["mpc_kappas"][i]), i
        assert conv.zsrc == i + 1.0 and conv.wlens == 1.0 and np.array_equal(conv.delta.numpy(), deltas[i])
    with pytest.raises(ValueError, match="source redshift must be increasing"):
        conv.add_plane(torch.as_tensor(deltas[0].copy()), 5.0)


def test_generate_host_flow_golden(monkeypatch):
    """The REAL glass_b200.generate / _generate_grf on CPU tensors with the kernels replaced by their
    definitions (l-major -> m-major, sum_i z_i w[l, i] + the m = 0 fix) and the synthesis by a
    recorder: iternorm weights, the z history and its trimming, the `mis` offset into the weights,
    shell batching and the fused transformation descriptors -- the alm handed to the synthesis
    against the alm the reference's own source handed to healpy.alm2map (golden grf_alm_*);
    shell selection (sharding) and the deferred error on a bad spectrum."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.fields as F
    from glass_b200 import _lib as L
    from glass_b200 import grf
    from glass_b200.rng import Deviates
    from helpers import native_iternorm, synthetic_gls

    k1 = native_iternorm()

    def c128(ptr, n):
        return np.ctypeslib.as_array((C.c_double * (2 * n)).from_address(ptr)).view(np.complex128)

    def f64(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    def order(lmax):  # m-major position -> l-major index, and l of every m-major entry
        ls = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        ms = np.concatenate([np.full(lmax + 1 - m, m) for m in range(lmax + 1)])
        return ls * (ls + 1) // 2 + ms, ls

    class FakeLib:
        def glb_iternorm_step(self, n, k, first, row, m, a, s_, tmp, w, flag, st):
            return k1.iternorm_step_host(n, k, first, row, m, a, s_, tmp, w, flag)  # K1's body, host build

        def glb_alm_glass_to_healpix(self, lmax, src, dst, st):
            n = (lmax + 1) * (lmax + 2) // 2
            c128(dst, n)[:] = c128(src, n)[order(lmax)[0]]
            return 0

        def glb_alm2map_prepare(self, plan, alm, nmaps, slot, st):  # the split form is the INT8 path's: not on this fake
            return L.GLB_ERR_UNSUPPORTED

        def glb_alm_combine(self, lmax, nterms, zptrs, w, stride, out, st):
            n = (lmax + 1) * (lmax + 2) // 2
            ls = order(lmax)[1]
            wmat = f64(w, (lmax + 1) * stride).reshape(lmax + 1, stride)
            alm = sum(c128(zptrs[i], n) * wmat[ls, i] for i in range(nterms))
            alm[: lmax + 1] = alm[: lmax + 1].real + alm[: lmax + 1].imag + 0j
            c128(out, n)[:] = alm
            return 0

    fed = []

    def alm2map_batch(alms, nside, lmax=None, transforms=None, out=None):
        fed.append((alms.numpy().copy(), list(transforms)))
        return torch.zeros((alms.shape[0], 12 * nside * nside), dtype=torch.float64)

    monkeypatch.setattr(F._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(F, "_pick_device", lambda gls: (torch.device("cpu"), True))
    monkeypatch.setattr(F.hp, "alm2map_batch", alm2map_batch)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    fake_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0, wait_stream=lambda s: None, wait_event=lambda e: None)  # noqa: E731
    monkeypatch.setattr(torch.cuda, "current_stream", fake_stream)
    monkeypatch.setattr(torch.cuda, "Stream", fake_stream)
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: types.SimpleNamespace(record=lambda s=None: None, synchronize=lambda: None))
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    for name, (nshell, lmax, ncorr, ragged) in {"a": (4, 12, 2, False), "b": (5, 9, None, False), "c": (4, 10, 1, True)}.items():
        nc = nshell - 1 if ncorr is None else ncorr
        gls = synthetic_gls(nshell, lmax, nc, ragged)
        rng = np.random.default_rng(42)
        n = (lmax + 1) * (lmax + 2) // 2
        zs = [rng.standard_normal((n, 2)) @ np.array([1, 1j]) for _ in range(nshell)]
        fed.clear()
        maps = list(F._generate_grf(gls, 4, ncorr=ncorr, rng=Deviates(normal_alm=zs)))
        assert len(maps) == nshell and all(m.shape == (192,) for m in maps)
        # the weights come from K1's arithmetic (another summation order than NumPy's matmul): 1e-13
        close = lambda a, b: np.abs(a - b).max() <= 1e-13 * np.abs(b).max()  # noqa: E731
        assert close(np.concatenate([a for a, _t in fed]), gold[f"grf_alm_{name}"])
        assert [a.shape[0] for a, _t in fed] == [nshell]  # one batch: nshell <= SHT_BATCH = 8 shells
        assert all(t == (L.T_NORMAL, 0.0, 1.0) for _a, tr in fed for t in tr)
        # a rank's shard: only its shells are synthesised, from the same deviates
        fed.clear()
        mine = [1, 3]
        list(F.generate([grf.Normal()] * nshell, gls, 4, ncorr=ncorr, rng=Deviates(normal_alm=zs), shells=mine))
        assert close(np.concatenate([a for a, _t in fed]), gold[f"grf_alm_{name}"][mine])
    # fused descriptors: Lognormal(lamda) -> (kind, var/2, lamda) with var = sum (2l+1)/(4 pi) g_l of the auto-spectrum
    gls = synthetic_gls(3, 8, 2)
    fields = [grf.Lognormal(0.7), grf.Normal(), grf.SquaredNormal(0.3, 1.5)]
    fed.clear()
    zs = [np.zeros(45, dtype=complex)] * 3
    list(F.generate(fields, gls, 4, rng=Deviates(normal_alm=zs)))
    var0 = F.cltovar(gls[0])
    assert fed[0][1] == [(L.T_LOGNORMAL, var0 / 2, 0.7), (L.T_NORMAL, 0.0, 1.0), (L.T_SQUARED_NORMAL, 0.3, 1.5)]
    with pytest.raises(ValueError, match="mismatch between number of fields and gls"):
        next(F.generate(fields[:2], gls, 4))
    # a negative auto-spectrum in the third shell: the first two shells are still produced, then the error
    bad = [g.copy() for g in gls]
    bad[3][2] = -1.0
    got = []
    with pytest.raises(ValueError, match="negative values in cl"):
        for m in F.generate(fields, bad, 4, rng=Deviates(normal_alm=zs)):
            got.append(m)
    assert len(got) == 2


def test_ellipticity_host_flow_golden(monkeypatch):
    """ellipticity_intnorm (glass/shapes.py:288-362) with the kernel replaced by its definition
    (e = sigma_eta n, e *= tanh(r/2)/r): the host side -- admissibility check, the sigma -> sigma_eta
    fit, population order and offsets -- against the reference's own source on the same normals."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.shapes as S
    from glass_b200.rng import Deviates

    def c128(ptr, n):
        return np.ctypeslib.as_array((C.c_double * (2 * n)).from_address(ptr)).view(np.complex128)

    calls = []

    class FakeLib:
        def glb_ellipticity(self, mode, sigma_eta, normals, n, seed, call, pos, out, st):
            calls.append((mode, sigma_eta, n, pos.value))
            e = c128(normals, n) * sigma_eta
            r = np.hypot(e.real, e.imag)
            e = e * np.where(r > 0, np.divide(np.tanh(r / 2), np.where(r > 0, r, 1.0)), 1.0)
            c128(out, n)[:] = e
            return 0

    monkeypatch.setattr(S._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(S.hp, "_device_index", lambda device=None: 0)
    real_device = torch.device
    monkeypatch.setattr(S.torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    got = S.ellipticity_intnorm(500, 0.256, rng=Deviates(normal=gold["eps_normals"]))
    assert isinstance(got, np.ndarray) and np.array_equal(got, gold["eps_intnorm"])
    s = 0.256
    assert calls == [(0, s * ((8 + 5 * s**2) / (2 - 4 * s**2)) ** 0.5, 500, 0)]
    # two populations: consecutive runs of the output, each with its own sigma_eta
    calls.clear()
    got2 = S.ellipticity_intnorm(np.array([200, 300]), np.array([0.256, 0.1]), rng=Deviates(normal=gold["eps_normals"]))
    assert got2.shape == (500,) and np.array_equal(got2[:200], gold["eps_intnorm"][:200])
    assert [(c[2], c[3]) for c in calls] == [(200, 0), (300, 200)]
    for bad in (-0.1, 0.5**0.5, 1.0):
        with pytest.raises(ValueError, match="sigma must be between 0 and sqrt"):
            S.ellipticity_intnorm(10, bad)


def test_gaussian_phz_host_flow_golden(monkeypatch):
    """gaussian_phz (glass/galaxies.py:350-455) with the kernel replaced by its definition (normal
    draw z + (1+z) sigma_0 n; later rounds replace only the out-of-bounds entries; count of those
    left): bounds handling, validation and the rejection rounds of the product's host side against
    the reference's own source on the same normals (golden phz_*)."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.galaxies as gal
    from glass_b200.rng import Deviates

    def f64(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    rounds_run = []

    class FakeLib:
        def glb_gaussian_phz(self, z, s_arr, s_val, lo_arr, lo_val, hi_arr, hi_val, normals, redraw_only, n, seed, call, out, nbad, st):
            zz = f64(z, n)
            sig = (1 + zz) * (f64(s_arr, n) if s_arr else s_val)
            lo = f64(lo_arr, n) if lo_arr else lo_val
            hi = f64(hi_arr, n) if hi_arr else hi_val
            o = f64(out, n)
            draw = zz + sig * f64(normals, n)  # Generator.normal(loc, scale) = loc + scale * standard_normal
            if redraw_only:
                bad = (o < lo) | (o > hi)
                o[bad] = draw[bad]
            else:
                o[:] = draw
            np.ctypeslib.as_array((C.c_int64 * 1).from_address(nbad))[0] = np.count_nonzero((o < lo) | (o > hi))
            rounds_run.append(int(redraw_only))
            return 0

    monkeypatch.setattr(gal._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(gal.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    got = gal.gaussian_phz(gold["phz_z"], 0.2, lower=0.1, upper=1.2, rng=Deviates(normal=list(gold["phz_normals"])))
    assert got.shape == gold["phz_z"].shape and np.array_equal(got, gold["phz_out"])
    assert rounds_run[0] == 0 and len(rounds_run) > 2 and all(r == 1 for r in rounds_run[1:])
    with pytest.raises(ValueError, match="requires lower < upper"):
        gal.gaussian_phz(gold["phz_z"], 0.2, lower=1.0, upper=0.5)
    with pytest.raises(ValueError, match="lower and upper must best scalars or have the same shape as z"):
        gal.gaussian_phz(gold["phz_z"], 0.2, lower=np.zeros(3), upper=np.ones(3))


def test_uniform_positions_host_flow_golden(monkeypatch):
    """uniform_positions (glass/points.py:543-607) with the kernel replaced by its definition
    (lon = -180 + 360 u1, lat = degrees(asin(-1 + 2 u2))): Poisson totals, population order, the
    count arrays, against the reference's own source replayed from the same stream (golden up_*)."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.points as P
    from glass_b200.rng import Deviates

    def f64(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    class FakeLib:
        def glb_uniform_positions(self, n, u1, u2, seed, stream, lon, lat, st):
            f64(lon, n)[:] = -180.0 + 360.0 * f64(u1, n)
            f64(lat, n)[:] = np.degrees(np.arcsin(-1.0 + 2.0 * f64(u2, n)))
            return 0

    monkeypatch.setattr(P._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(P.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    pos = {"i": 0}

    def uniforms(n):
        i = pos["i"]
        pos["i"] += n
        return g["up_u_lon"][i : i + n], g["up_u_lat"][i : i + n]

    res = list(P.uniform_positions(g["up_ngal"], rng=Deviates(poisson=[g["up_totals"]], uniform=uniforms)))
    assert len(res) == 2
    assert np.array_equal(np.concatenate([r[0] for r in res]), g["up_lon"])
    assert np.array_equal(np.concatenate([r[1] for r in res]), g["up_lat"])
    assert np.array_equal(np.stack([r[2] for r in res]), g["up_count"])
    # scalar density: the count is a plain int
    pos["i"] = 0
    (lon, lat, cnt), = list(P.uniform_positions(float(g["up_ngal"][0]), rng=Deviates(poisson=[g["up_totals"][:1]], uniform=uniforms)))
    assert isinstance(cnt, int) and cnt == int(g["up_totals"][0]) and lon.shape == (cnt,)


def test_galaxy_shear_host_flow_golden(monkeypatch):
    """galaxy_shear (glass/galaxies.py:271-347) with the kernel replaced by its definition (pixel lookup,
    three gathers, reduced-shear formula from the oracle): argument plumbing, broadcasting of scalar
    maps / ellipticities, the ipix shortcut -- against the reference's own source (golden gs_*)."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.galaxies as gal
    from oracle import glass_ref as G
    from oracle import healpix_ref as H

    def f64(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    def c128(ptr, n):
        return f64(ptr, 2 * n).view(np.complex128)

    used_ipix = []

    class FakeLib:
        def glb_galaxy_shear(self, nside, lon, lat, ipix, eps, n, kappa, g1, g2, reduced, out, st):
            npix = 12 * nside * nside
            used_ipix.append(ipix is not None)
            if ipix is None:
                lo, la = f64(lon, n), f64(lat, n)
            else:  # pixel centres stand for the positions: the same pixels are looked up
                p = np.ctypeslib.as_array((C.c_int64 * n).from_address(ipix))
                lo, la = H.ring2ang_uv(nside, p, np.full(n, 0.5), np.full(n, 0.5), lonlat=True)
            c128(out, n)[:] = G.galaxy_shear(lo, la, c128(eps, n), f64(kappa, npix), f64(g1, npix), f64(g2, npix), reduced_shear=bool(reduced))
            return 0

    monkeypatch.setattr(gal._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(gal.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    args = (g["gs_lon"], g["gs_lat"], g["gs_eps"], g["gs_kappa"], g["gs_g1"], g["gs_g2"])
    assert np.array_equal(gal.galaxy_shear(*args, reduced_shear=True), g["gs_reduced"])
    assert np.array_equal(gal.galaxy_shear(*args, reduced_shear=False), g["gs_plain"])
    ipix = H.ang2pix(4, g["gs_lon"], g["gs_lat"], lonlat=True)
    assert np.array_equal(gal.galaxy_shear(*args, ipix=ipix), g["gs_reduced"]) and used_ipix == [False, False, True]
    # one intrinsic ellipticity for all galaxies broadcasts like in the reference
    one = gal.galaxy_shear(g["gs_lon"], g["gs_lat"], np.complex128(0.1 + 0.2j), g["gs_kappa"], g["gs_g1"], g["gs_g2"])
    assert one.shape == g["gs_lon"].shape
    assert np.array_equal(one, G.galaxy_shear(g["gs_lon"], g["gs_lat"], np.full(g["gs_lon"].shape, 0.1 + 0.2j), g["gs_kappa"], g["gs_g1"], g["gs_g2"]))


def test_displace_host_flow_golden(monkeypatch):
    """displace / displacement (glass/points.py:654-772) with the kernels replaced by the reference's
    formulas: the complex and the (2, n) form of alpha, broadcasting of a scalar displacement, array
    kinds -- against the reference's own source (golden displace file), bit for bit."""
    import contextlib
    import ctypes as C
    import math
    import types

    import torch

    import glass_b200.points as P

    def f64(ptr, n, stride=1):
        return np.ctypeslib.as_array((C.c_double * (n * stride)).from_address(ptr))[::stride]

    class FakeLib:
        def glb_displace(self, lon, lat, a1, a2, stride, deflect, n, out_lon, out_lat, st):
            assert not deflect
            alpha1, alpha2 = f64(a1, n, stride), f64(a2, n, stride)
            t = f64(lat, n) / 180 * math.pi
            ct, st_ = np.sin(t), np.cos(t)
            a, g = np.hypot(alpha1, alpha2), np.arctan2(alpha2, alpha1)
            ca, sa, cg, sg = np.cos(a), np.sin(a), np.cos(g), np.sin(g)
            tp = np.arctan2(ct * ca + st_ * sa * cg, np.hypot(ct * sa - st_ * ca * cg, st_ * sg))
            d = np.arctan2(sa * sg, st_ * ca - ct * sa * cg)
            f64(out_lon, n)[:] = f64(lon, n) + d / math.pi * 180
            f64(out_lat, n)[:] = tp / math.pi * 180
            return 0

        def glb_displacement(self, from_lon, from_lat, to_lon, to_lat, n, out, st):
            a, b = np.radians(f64(from_lat, n)), np.radians(f64(to_lat, n))
            g = np.radians(f64(to_lon, n) - f64(from_lon, n))
            sa, ca, sb, cb, sg, cg = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(g), np.cos(g)
            r = np.arctan2(np.hypot(cb * sg, ca * sb - sa * cb * cg), sa * sb + ca * cb * cg)
            x = np.arctan2(cb * sg, ca * sb - sa * cb * cg)
            f64(out, 2 * n).view(np.complex128)[:] = r * np.exp(1j * x)
            return 0

    monkeypatch.setattr(P._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(P.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_displace.npz"))
    for alpha in (g["alpha"], np.stack([g["alpha"].real, g["alpha"].imag])):
        lon, lat = P.displace(g["lon"], g["lat"], alpha)
        assert isinstance(lon, np.ndarray)
        assert np.array_equal(lon, g["displace_lon"]) and np.array_equal(lat, g["displace_lat"])
    lon1, lat1 = P.displace(g["lon"][:5], g["lat"][:5], np.complex128(0.01 + 0.02j))  # scalar alpha broadcasts
    lon2, lat2 = P.displace(g["lon"][:5], g["lat"][:5], np.full(5, 0.01 + 0.02j))
    assert np.array_equal(lon1, lon2) and np.array_equal(lat1, lat2)
    with pytest.raises(ValueError, match="leading axis of size 2"):
        P.displace(g["lon"][:3], g["lat"][:3], np.zeros((3, 3)))
    assert np.array_equal(P.displacement(g["lon"], g["lat"], g["to_lon"], g["to_lat"]), g["displacement"])


def test_positions_from_delta_host_flow_golden(monkeypatch):
    """The REAL glass_b200.positions_from_delta on CPU tensors with the C-ABI calls replaced by
    their definitions (counts supplied, exclusive scan or galaxy list, the kernel's own cut-rule
    header compiled for the host, np.repeat + pixel -> angle): broadcasting of
    the population axes, iteration order, the one-hot batch counts, batch cuts and the concatenated
    positions against the reference's own source run on the same count maps (golden positions file)."""
    import contextlib
    import ctypes as C
    import itertools
    import types

    import torch

    import glass_b200.points as P
    from glass_b200.rng import Deviates
    from helpers import native_points_cuts
    from oracle import healpix_ref as H

    def f64(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    def i64(ptr, n):
        return np.ctypeslib.as_array((C.c_int64 * n).from_address(ptr))

    seen = []
    walker = native_points_cuts()  # csrc/points_cuts.cuh on the host: scan and list walkers

    class FakeLib:
        glb_points_cuts = staticmethod(walker.glb_points_cuts)
        glb_points_cuts_list = staticmethod(walker.glb_points_cuts_list)

        def glb_points_workspace_bytes(self, npix):
            return 64

        def glb_points_counts(self, npix, d, v, code, bias, scale, rm, cin, seed, stream, nbar, counts, off, gpix, cap, total, ws, st):
            c = i64(cin, npix)
            if not (gpix and cap < c.sum()):  # a list that was too short is filled again by a second call
                seen.append((code, bias, bool(rm), v is not None, stream.value))
            if counts:
                i64(counts, npix)[:] = c
            if off:
                o = i64(off, npix + 1)
                o[0] = 0
                o[1:] = np.cumsum(c)
            if gpix:
                full = np.repeat(np.arange(npix), c)
                k = min(cap, full.size)
                i64(gpix, max(cap, 1))[:k] = full[:k]
            i64(total, 1)[0] = c.sum()
            return 0

        def glb_points_fill_list(self, nside, gpix, g0, g1, u, v, seed, stream, lon, lat, st):
            n = g1 - g0
            pix = i64(gpix, g1)[g0:g1]
            lo, la = H.ring2ang_uv(nside, pix, f64(u, n), f64(v, n), lonlat=True)
            f64(lon, n)[:], f64(lat, n)[:] = lo, la
            return 0

        def glb_points_fill(self, nside, counts, off, start, stop, u, v, seed, stream, lon, lat, ipix, st):
            npix = 12 * nside * nside
            c, o = i64(counts, npix), i64(off, npix + 1)
            n = int(o[stop] - o[start])
            pix = np.repeat(np.arange(start, stop), c[start:stop])
            lo, la = H.ring2ang_uv(nside, pix, f64(u, n), f64(v, n), lonlat=True)
            f64(lon, n)[:], f64(lat, n)[:] = lo, la
            return 0

    monkeypatch.setattr(P._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(P.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_positions.npz"))
    centre = lambda n: (np.full(n, 0.5), np.full(n, 0.5))  # noqa: E731
    # both layouts of the counts: the galaxy list (sparse maps) and the per-pixel scan
    for (tag, kw), dens in itertools.product({"lin": {}, "loglin_rm": {"bias_model": P.loglinear_bias, "remove_monopole": True}}.items(), (1.0, 0.0)):
        monkeypatch.setattr(P, "LIST_MODE_MAX_DENSITY", dens)
        seen.clear()
        res = list(P.positions_from_delta(g["ngal"], g["delta"], g["bias"], g["vis"], batch=40,
                                          rng=Deviates(poisson=list(g[f"{tag}_counts"]), uv=centre), **kw))
        assert np.array_equal(np.stack([c for _lo, _la, c in res]), g[f"{tag}_batch_count"])
        assert np.array_equal(np.concatenate([lo for lo, _la, _c in res]), g[f"{tag}_lon"])
        assert np.array_equal(np.concatenate([la for _lo, la, _c in res]), g[f"{tag}_lat"])
        # six populations in C order of dims (3, 2): bias varies along the first axis, every one sees vis
        assert [s[1] for s in seen] == [0.5, 0.5, 1.0, 1.0, 1.7, 1.7] and [s[4] for s in seen] == list(range(6))
        assert all(s[0] == (2 if tag == "loglin_rm" else 1) and s[2] == (tag == "loglin_rm") and s[3] for s in seen)
    res = list(P.positions_from_delta(2e-3, g["delta"], None, None, batch=1000, rng=Deviates(poisson=[g["scalar_counts"]], uv=centre)))
    assert all(isinstance(c, int) for _lo, _la, c in res) and [c for _lo, _la, c in res] == list(g["scalar_batch_count"])
    assert np.array_equal(np.concatenate([lo for lo, _la, _c in res]), g["scalar_lon"])
    with pytest.raises(TypeError, match="bias_model must be callable"):
        next(P.positions_from_delta(1e-3, g["delta"], bias_model=0))


def test_redshifts_host_cdf_golden(monkeypatch):
    """redshifts_from_nz (glass/galaxies.py:188-268, 77-89): the host side of the product -- broadcast
    of count / z / nz, cumulative-trapezoid CDF, its normalisation -- with the inverse-CDF kernel
    replaced by its definition interp(u, cdf, z) on the same buffers, against the reference's own
    source fed the same uniform deviates (golden z_*), bit for bit."""
    import contextlib
    import ctypes as C
    import types

    import torch

    import glass_b200.galaxies as gal
    from glass_b200.rng import Deviates

    def arr(ptr, n):
        return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))

    class FakeLib:
        def glb_redshifts_from_cdf(self, cdf, z, ncdf, u, n, seed, call, pos, out, st):
            arr(out, n)[:] = np.interp(arr(u, n), arr(cdf, ncdf), arr(z, ncdf))
            return 0

    monkeypatch.setattr(gal._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(gal.A, "pick_device", lambda *a: (torch.device("cpu"), False))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    got = gal.redshifts_from_nz(400, gold["z_grid"], gold["z_nz"], rng=Deviates(uniform=gold["z_uniform"]), warn=False)
    assert isinstance(got, np.ndarray) and np.array_equal(got, gold["z_samples"])
    with pytest.warns(UserWarning, match="redshifts_from_nz"):
        gal.redshifts_from_nz(3, gold["z_grid"], gold["z_nz"], rng=Deviates(uniform=gold["z_uniform"]))
    # two populations with their own n(z): runs concatenated in population order
    nz2 = np.stack([gold["z_nz"], gold["z_nz"][::-1]])
    got2 = gal.redshifts_from_nz(np.array([150, 250]), gold["z_grid"], nz2, rng=Deviates(uniform=gold["z_uniform"]), warn=False)
    assert got2.shape == (400,) and np.array_equal(got2[:150], gold["z_samples"][:150])
    w = types.SimpleNamespace(za=gold["z_grid"], wa=gold["z_nz"], zeff=1.0)
    assert np.array_equal(gal.redshifts(400, w, rng=Deviates(uniform=gold["z_uniform"])), gold["z_samples"])


def test_lensing_factor_chain_golden(monkeypatch):
    """from_convergence / shear_from_convergence between the transforms (glass/lensing.py:296-371,
    403-428): the l-dependent factors and the alm handed to the spin transforms, against what the
    reference's own source produced with recording shims in place of healpy
    (tests/golden/make_golden.py --lensing-factors).  The transforms are faked here too, so the
    host-side control flow of the product runs on the CPU; oracle factors are checked alongside."""
    import torch

    import glass_b200.lensing as L
    from oracle import glass_ref as G

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_lensing_factors.npz"))
    lmax, nside = int(g["lmax"]), 4
    lof = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
    rec = {"fl": [], "scalar": [], "spin": []}

    def almxfl(alm, fl, *, inplace=False):
        rec["fl"].append(np.array(fl, copy=True))
        alm *= torch.as_tensor(np.asarray(fl)[lof])
        return alm

    def alm2map_batch(alms, nside, lmax=None, transforms=None, out=None):
        rec["scalar"].append(alms[0].numpy().copy())
        return torch.zeros((alms.shape[0], 12 * nside * nside), dtype=torch.float64)

    def alm2map_spin(alms, nside, spin, lmax):
        assert alms[1] is None  # E-only, like the reference's zero blm
        rec["spin"].append((spin, alms[0].numpy().copy()))
        return torch.zeros(12 * nside * nside, dtype=torch.float64), torch.zeros(12 * nside * nside, dtype=torch.float64)

    monkeypatch.setattr(L.hp, "almxfl", almxfl)
    monkeypatch.setattr(L.hp, "alm2map_batch", alm2map_batch)
    monkeypatch.setattr(L.hp, "alm2map_spin", alm2map_spin)
    monkeypatch.setattr(L, "_kappa_alm", lambda kappa, lmax_, niter, rw: (torch.as_tensor(g["alm0"].copy()), nside, lmax, torch.device("cpu"), True))
    kappa = np.zeros(12 * nside * nside)
    for tag, disc in (("plain", False), ("disc", True)):
        pixwin = (g["pw0"], g["pw2"]) if disc else None
        for k in rec:
            rec[k].clear()
        res = L.from_convergence(kappa, lmax, potential=True, deflection=True, shear=True, discretized=disc, pixwin=pixwin)
        assert len(res) == 3 and res[1].is_complex() and res[2].is_complex()
        assert np.array_equal(np.stack(rec["fl"]), g[f"fc_{tag}_fl"])
        assert np.array_equal(rec["scalar"][0], g[f"fc_{tag}_psi_alm"])
        assert rec["spin"][0][0] == 1 and np.array_equal(rec["spin"][0][1], g[f"fc_{tag}_alpha_alm"])
        assert rec["spin"][1][0] == 2 and np.array_equal(rec["spin"][1][1], g[f"fc_{tag}_gamma_alm"])
        for k in rec:
            rec[k].clear()
        g1g2 = L.shear_from_convergence(kappa, lmax, discretized=disc, pixwin=pixwin)
        assert len(g1g2) == 2
        assert np.array_equal(rec["fl"][0], g[f"sfc_{tag}_fl"]) and np.array_equal(rec["spin"][0][1], g[f"sfc_{tag}_alm"])
        # the oracle's restatement of the same factors
        pw = {"pw0": g["pw0"], "pw2": g["pw2"]} if disc else {}
        assert all(np.array_equal(a, b) for a, b in zip(G.from_convergence_factors(lmax, discretized=disc, **pw), g[f"fc_{tag}_fl"]))
    assert np.array_equal(G.kappa_to_shear_fl(lmax), g["sfc_plain_fl"])
    # potential only: the pixel window is not needed (and not asked for)
    for k in rec:
        rec[k].clear()
    assert len(L.from_convergence(kappa, lmax, potential=True)) == 1 and len(rec["fl"]) == 1
    # the reference's default (discretized=True, no tables given) asks hp.pixwin for both windows
    asked = []
    monkeypatch.setattr(L.hp, "pixwin", lambda ns, lmax=None, pol=False: asked.append((ns, lmax, pol)) or (g["pw0"], g["pw2"]))
    for k in rec:
        rec[k].clear()
    L.from_convergence(kappa, lmax, shear=True)
    assert asked == [(nside, lmax, True)] and np.array_equal(np.stack(rec["fl"]), g["fc_disc_fl"])


def test_solver_against_reference_source_golden(transforms_on_cpu):
    """The product's transform pair against the oracle's independent restatement
    (oracle/transformcl_ref.py), and the batched solver against vectors produced by executing the
    reference's OWN solver source on top of that oracle (tests/golden/make_golden.py --solver):
    same info flags (same stopping point and halvings), gl and the realised cl to 1e-9."""
    import glass_b200 as glass
    from glass_b200 import grf
    from oracle import transformcl_ref as tref

    tcl = transforms_on_cpu
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 41, 123):
        x = rng.standard_normal(n) / (1 + np.arange(n))
        assert np.allclose(tcl.cltocorr(x), tref.cltocorr(x), rtol=1e-12, atol=1e-14)
        c = tref.cltocorr(x)
        assert np.allclose(tcl.corrtocl(c), tref.corrtocl(c), rtol=1e-10, atol=1e-13)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_solver.npz"))
    cl, lmax = g["cl"], g["cl"].shape[0] - 1
    cases = {
        "ln": (grf.Lognormal(0.8), None, {}),
        "ln_pad": (grf.Lognormal(0.8), grf.Lognormal(1.3), {"pad": 2 * (lmax + 1)}),
        "ln_mono": (grf.Lognormal(), None, {"pad": 2 * (lmax + 1), "monopole": 0.0, "cltol": 1e-9, "gltol": 1e-9}),
        "ln_normal": (grf.Lognormal(0.6), grf.Normal(), {"pad": lmax + 1}),
        "sq": (grf.SquaredNormal(0.9, 1.1), grf.SquaredNormal(0.8, 0.7), {"pad": lmax + 1, "cltol": 1e-8}),
        "iter2": (grf.Lognormal(0.5), None, {"pad": lmax + 1, "maxiter": 2, "cltol": 1e-14, "gltol": 1e-14}),
    }
    for tag, (t1, t2, kw) in cases.items():
        gl, rl, info = grf.solve(cl.copy(), t1, t2, **kw)
        assert info == int(g[f"solve_{tag}_info"]), tag
        assert np.allclose(gl, g[f"solve_{tag}_gl"], rtol=1e-9, atol=1e-9 * np.abs(gl).max()), tag
        assert np.allclose(rl, g[f"solve_{tag}_rl"], rtol=1e-9, atol=1e-9 * np.abs(rl).max()), tag
    assert np.allclose(grf.compute(cl.copy(), grf.Lognormal(0.8)), g["compute_ln"], rtol=1e-10, atol=1e-15)
    fields = [grf.Lognormal(1.0), grf.Lognormal(0.7), grf.Normal()]
    spectra = np.split(g["sgs_spectra"], np.cumsum(g["sgs_spectra_len"])[:-1])
    gls = glass.solve_gaussian_spectra(fields, spectra)
    ref = np.split(g["sgs_gls"], np.cumsum(g["sgs_len"])[:-1])
    assert [x.shape[0] for x in gls] == [x.shape[0] for x in ref]
    for a, b in zip(gls, ref):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * max(np.abs(b).max(), 1e-300) if b.size else 0)


def test_solve_gaussian_spectra_batched_equals_single(transforms_on_cpu):
    """solve_gaussian_spectra (glass/fields.py:778-836): the batched run over all spectra follows,
    column by column, the solve the reference would run on each spectrum alone (padding 2n, zero
    monopole pinned, empty spectra passed through, mixed transformation pairs)."""
    import glass_b200 as glass
    from glass_b200 import grf

    lmax = 60
    base = 1e-2 / (2 * np.arange(lmax + 1) + 1) ** 2
    fields = [grf.Lognormal(1.0), grf.Lognormal(0.7), grf.Normal(), grf.Lognormal(1.3)]
    spectra = []
    for i in range(4):
        for j in range(i, -1, -1):
            c = 0.5 ** (i - j) * base * (1 + 0.1 * i)
            if i == 1:
                c[0] = 0.0
            spectra.append(c if i - j <= 2 else np.zeros(0))
    gls = glass.solve_gaussian_spectra(fields, spectra)
    assert len(gls) == len(spectra)
    for k, (i, j, cl) in enumerate(glass.enumerate_spectra(spectra)):
        if cl.shape[0] == 0:
            assert gls[k].shape == (0,)
            continue
        g, _, info = grf.solve(cl, fields[i], fields[j], pad=2 * cl.shape[0], monopole=0.0 if cl[0] == 0 else None)
        assert info > 0
        assert np.abs(g - gls[k]).max() <= 1e-12 * np.abs(g).max()
        if cl[0] == 0:
            assert gls[k][0] == 0.0
    with pytest.raises(ValueError, match="mismatch between number of fields and spectra"):
        glass.solve_gaussian_spectra(fields, spectra[:-1])
    assert glass.solve_gaussian_spectra([], []) == []
    cg = glass.compute_gaussian_spectra(fields, spectra)
    assert np.array_equal(cg[0], grf.compute(spectra[0], fields[0], fields[0])) and cg[-1].shape == (0,)
    with pytest.warns(DeprecationWarning):
        assert len(glass.lognormal_gls(spectra[:3])) == 3


def test_solve_columns_stop_at_different_iterations(transforms_on_cpu):
    """Round-1 advisor finding: columns of one batch that stop at different Gauss-Newton iterations
    (amplitudes differing by orders of magnitude; near-zero cross-spectra converge at once) or take
    different step halvings must each follow the solve the reference runs on that spectrum alone
    (glass/grf/_solver.py:114-146), incl. the per-column transformation parameters and the ``info``
    flags (bit 0 is tested at the top of every iteration, so a column can end with info == 3)."""
    import torch

    import glass_b200 as glass
    from glass_b200 import grf

    lmax = 48
    ell = np.arange(lmax + 1)
    base = 1.0 / (2 * ell + 1) ** 2
    fields = [grf.Lognormal(1.0)] * 3
    amps = {(0, 0): 3e-1, (1, 1): 1e-3, (1, 0): 1e-9, (2, 2): 5e-2, (2, 1): 2e-6, (2, 0): 1e-2}
    spectra = [amps[(i, j)] * base for i in range(3) for j in range(i, -1, -1)]
    gls = glass.solve_gaussian_spectra(fields, spectra)
    infos = []
    for k, (i, j, cl) in enumerate(glass.enumerate_spectra(spectra)):
        g, _, info = grf.solve(cl, fields[i], fields[j], pad=2 * cl.shape[0])
        infos.append(info)
        assert np.abs(g - gls[k]).max() <= 1e-12 * np.abs(g).max()
    # mixed lamdas (per-column parameters) and SquaredNormal columns with halvings
    for mk in (lambda r: grf.Lognormal(0.3 + r), lambda r: grf.SquaredNormal(0.2 + r, 0.5 + r)):
        rng = np.random.default_rng(7)
        fs = [mk(rng.random()) for _ in range(3)]
        sp = [10.0 ** rng.uniform(-6, -1) * base * 0.3 ** (i - j) for i in range(3) for j in range(i, -1, -1)]
        got = glass.solve_gaussian_spectra(fs, sp)
        for k, (i, j, cl) in enumerate(glass.enumerate_spectra(sp)):
            g, _, info = grf.solve(cl, fs[i], fs[j], pad=2 * cl.shape[0])
            assert np.abs(g - got[k]).max() <= 1e-11 * np.abs(g).max(), (k, info)
    # the flags of the batched run are the flags of the single runs, column by column
    cols = torch.as_tensor(np.stack(spectra, 1))
    t = grf.Lognormal(torch.ones(cols.shape[1], dtype=torch.float64))
    _, _, info_b = grf.solve_columns(cols, t, t, pad=2 * (lmax + 1))
    assert info_b.tolist() == infos
    assert len(set(infos)) > 1 or len({tuple(np.round(g, 3)) for g in gls}) > 1
    # info == 3: the step converges (bit 1) in the same iteration that brings clerr under cltol
    _gl, _rl, info3 = grf.solve_columns(cols[:, :1], grf.Lognormal(1.0), grf.Lognormal(1.0), pad=2 * (lmax + 1), cltol=1e-3, gltol=0.1)
    assert int(info3[0]) == 3
    _gl, _rl, info2 = grf.solve_columns(cols[:, :1], grf.Lognormal(1.0), grf.Lognormal(1.0), pad=2 * (lmax + 1), cltol=1e-8, gltol=0.1)
    assert int(info2[0]) == 2


def test_regularized_spectra_golden(transforms_on_cpu):
    """cov_clip / nearcorr / cov_nearest (glass/algorithm.py:111-277) and regularized_spectra
    (glass/fields.py:1055-1112) against vectors from executing the reference's source; the batched
    eigendecomposition is another library than the reference's LAPACK call, hence a tolerance."""
    import glass_b200 as glass
    from glass_b200 import algorithm

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_spectra.npz"))
    close = lambda a, b: np.allclose(a, b, rtol=1e-10, atol=1e-12 * np.abs(b).max())  # noqa: E731
    covs = g["reg_cov"]
    assert close(algorithm.cov_clip(covs), g["reg_clip"])
    assert close(algorithm.cov_clip(covs, rtol=0.1), g["reg_clip_rtol"])
    assert close(algorithm.cov_nearest(covs), g["reg_nearest"])
    corr = covs / np.sqrt(np.einsum("...ii,...jj->...ij", covs, covs))
    near = algorithm.nearcorr(corr)
    assert close(near, g["reg_nearcorr"])
    assert np.allclose(np.einsum("...ii->...i", near), 1.0) and np.linalg.eigvalsh(near).min() > -1e-12
    with pytest.raises(ValueError, match="negative values on the diagonal"):
        algorithm.cov_nearest(-covs)
    with pytest.raises(ValueError, match="non-square matrix"):
        algorithm.nearcorr(np.ones((3, 2)))
    bad = np.split(g["reg_gls"], np.cumsum(g["reg_gls_len"])[:-1])
    assert not glass.check_posdef_spectra(bad)
    for m in ("nearest", "clip"):
        reg = glass.regularized_spectra(bad, method=m)
        assert close(np.stack(reg), g[f"reg_spectra_{m}"])
        assert glass.check_posdef_spectra([np.asarray(r) * (1 + 1e-12 * (i == 0 or i == 1 or i == 3)) for i, r in enumerate(reg)])
    assert close(np.stack(glass.regularized_spectra(bad, lmax=5, method="clip")), g["reg_spectra_lmax5"])
    with pytest.raises(ValueError, match="unknown method 'foo'"):
        glass.regularized_spectra(bad, method="foo")


def test_redshifts_from_bins_order_golden(monkeypatch):
    """redshifts_from_bins (glass/galaxies.py:122-185): the tally / run / scatter logic against the
    reference executed with a seeded NumPy stream.  The per-bin draw (the redshift kernel, covered by
    the GPU tests) is replaced here by its definition interp(u, cdf, z) fed from the same stream."""
    import glass_b200.galaxies as gal

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_spectra.npz"))
    stream = iter(g["rb_uniform"])

    def draw(count, z, nz, *, rng=None, warn=True):
        cdf = gal._cumulative_trapezoid(nz, z)
        cdf /= cdf[-1]
        return np.interp(np.array([next(stream) for _ in range(int(count))]), cdf, z)

    monkeypatch.setattr(gal, "redshifts_from_nz", draw)
    nzd = {5: g["rb_nz"][1], 2: g["rb_nz"][0], 9: g["rb_nz"][2]}  # dictionary order is not label order
    out = gal.redshifts_from_bins(g["rb_bins"], g["rb_z"], nzd)
    assert out.shape == g["rb_bins"].shape
    # every bin receives exactly the reference's run of draws; WITHIN a bin the reference hands them out
    # in the tie order of NumPy's unstable argsort (which depends on the CPU's sort kernels), here in
    # galaxy order -- the draws of a run are i.i.d., so only the per-bin sets are comparable
    for label in (2, 5, 9):
        sel = g["rb_bins"] == label
        assert np.array_equal(np.sort(out[sel]), np.sort(g["rb_out"][sel]))
    first = {label: np.flatnonzero(g["rb_bins"] == label) for label in (2, 5, 9)}
    runs = np.split(g["rb_uniform"], np.cumsum([first[2].size, first[5].size])[:2])
    for label, u in zip((2, 5, 9), runs):  # galaxy order within the bin = draw order
        cdf = gal._cumulative_trapezoid(nzd[label], g["rb_z"])
        assert np.array_equal(out[first[label]], np.interp(u, cdf / cdf[-1], g["rb_z"]))


def test_batch_cut_rule_closed_form_on_cpu(monkeypatch):
    """points._Population.cuts -- the closed form of the reference's 1000-pixel stepping loop
    (glass/points.py:409-437), walked by glb_points_cuts; here the kernel's header
    csrc/points_cuts.cuh compiled for the host stands in for the launch -- against the batch sizes recorded from the
    reference's source (golden) and against the oracle's restatement of the loop on corner cases
    (exact fit at a group boundary, empty batches before an oversize pixel, trailing zeros,
    random sparse and dense maps, batch = 1)."""
    import torch

    from glass_b200.points import _Population
    from helpers import native_points_cuts
    from oracle import glass_ref as G

    fake = native_points_cuts()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: __import__("types").SimpleNamespace(cuda_stream=0))

    def cuts_of(counts, batch):
        pop = object.__new__(_Population)
        pop.npix, pop.lib, pop.device = counts.size, fake, torch.device("cpu")
        pop.off = torch.as_tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
        pop.total = int(counts.sum())
        got = list(pop.cuts(batch))
        assert got == list(pop.cuts(batch, chunk=3))  # the walk resumes correctly between chunks
        return [n for _a, _b, n in got], [(a, b) for a, b, _n in got]

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    for batch in (1_000_000, 500, 37, 1):
        sizes, _ = cuts_of(gold["pt_counts"], batch)
        assert sizes == list(gold[f"pt_batches_{batch}"]), batch
    npix = 12 * 16**2
    cases = []
    c = np.zeros(npix, dtype=np.int64)
    c[0], c[1500], c[1501], c[2999] = 5, 9, 1, 2
    cases.append((c, 5))
    c = np.zeros(npix, dtype=np.int64)
    c[999], c[1000], c[2500] = 3, 2, 4
    cases.append((c, 3))
    cases.append((np.random.default_rng(0).poisson(0.01, npix), 2))
    cases.append((np.random.default_rng(1).poisson(2.0, npix), 1000))
    cases.append((np.random.default_rng(2).poisson(0.5, npix), 1))
    cases.append((np.random.default_rng(3).poisson(0.3, 5 * npix), 977))
    rng = np.random.default_rng(0)
    for it in range(120):  # random maps incl. exact fits and batch = 1 (ranges of the LAST batch end with its group too)
        n = int(rng.integers(1, 4000))
        c = rng.poisson(10 ** rng.uniform(-3, 0.7), n)
        if c.sum() == 0:
            continue
        batch = 1 if it % 5 == 0 else 1 + int(rng.random() ** 2 * 2 * c.sum())
        if it % 7 == 0:
            batch = max(1, int(c[: rng.integers(1, n + 1)].sum()))
        if c.sum() / batch <= 300:
            cases.append((c, batch))
    for counts, batch in cases:
        ref = G.batch_cuts(counts, batch)
        sizes, ranges = cuts_of(counts, batch)
        assert sizes == [r[2] for r in ref], batch
        assert ranges == [(r[0], r[1]) for r in ref], batch
        assert sum(sizes) == counts.sum()


def test_points_cuts_core_host_build_and_run(tmp_path):
    """csrc/points_cuts.cuh (the batch-cut rule the device walks for glb_points_cuts) on the host:
    the native fuzz against the reference's 1000-pixel stepping loop restated in C++
    (tests/native/points_cuts_host.cpp, ~3000 maps incl. exact fits, oversize pixels, batch = 1)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "points_cuts_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(root, "tests", "native", "points_cuts_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "points_cuts ok" in r.stdout, r.stdout + r.stderr


def test_fft_core_host_build_and_run(tmp_path):
    """The shared-memory FFT passes of the ring-FFT kernels (csrc/fft_core.cuh) are plain
    per-thread functions: compile them for the host and check every pass, thread by thread,
    against a long-double reference (DIF, reordering DIF tail, DIT, fused Bluestein middle)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "fft_core_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", str(exe), os.path.join(root, "tests", "native", "fft_core_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "fft_core ok" in r.stdout, r.stdout + r.stderr


class _CpuMultiPlane:
    """Test double with the state layout of glass_b200.MultiPlaneConvergence and the reference's
    arithmetic (glass/lensing.py:535-586) on CPU tensors: exercises the rank-to-rank hand-off of
    glass_b200.dist.multi_plane_block without a GPU."""

    def __init__(self, cosmo):
        self.cosmo = cosmo
        self.z2 = self.z3 = self.x3 = self.w3 = 0.0
        self.r23 = 1.0
        self.delta3 = self.kappa2 = self.kappa3 = None
        self._like = None

    def _set_state_map(self, name, value):
        setattr(self, name, value)

    def add_window(self, delta, w):
        za, wa, zeff = w
        self.add_plane(delta, zeff, float(np.trapezoid(wa, za) / np.interp(zeff, za, wa)))

    def add_plane(self, delta, zsrc, wlens=1.0):
        import torch

        delta2, self.delta3 = self.delta3, delta
        z1, self.z2, self.z3 = self.z2, self.z3, zsrc
        w2, self.w3 = self.w3, wlens
        x2, self.x3 = self.x3, self.cosmo.transverse_comoving_distance(self.z3) / self.cosmo.hubble_distance
        r12 = self.r23
        r13, self.r23 = np.asarray(self.cosmo.transverse_comoving_distance([z1, self.z2], self.z3)) / (self.cosmo.hubble_distance * self.x3)
        t = float(r13 / r12)
        f = float(3 * self.cosmo.Omega_m0 / 2 * x2 * self.r23 * (1 + self.z2) / self.cosmo.H_over_H0(self.z2) * w2)
        if self.kappa2 is None:
            self.kappa2, self.kappa3 = torch.zeros_like(delta), torch.zeros_like(delta)
        self.kappa2, self.kappa3 = self.kappa3, self.kappa2
        self.kappa3 *= 1 - t
        self.kappa3 += t * self.kappa2
        if delta2 is not None:
            self.kappa3 += f * delta2

    @property
    def kappa(self):
        return self.kappa3


class _MockCosmo:  # reference tests/fixtures/domain.py:36-97
    Omega_m0 = 0.3
    hubble_distance = 4.4e3

    def H_over_H0(self, z):  # noqa: N802
        return (self.Omega_m0 * (1 + z) ** 3 + 1 - self.Omega_m0) ** 0.5

    def transverse_comoving_distance(self, z, z2=None):
        if z2 is None:
            return self.hubble_distance * np.asarray(z) * 1_000
        return self.hubble_distance * (np.asarray(z2) - np.asarray(z)) * 1_000


def _mp_inputs(nshell, npix):
    import torch

    g = torch.Generator().manual_seed(7)
    deltas = [torch.randn(npix, dtype=torch.float64, generator=g) for _ in range(nshell)]
    dz = 1.0 / (nshell + 1)
    wins = [(np.array([i, i + 1.0, i + 2.0]) * dz, np.array([0.0, 1.0, 0.0]), (i + 1.0) * dz) for i in range(nshell)]
    return deltas, wins


def _multiplane_worker(rank, world, port, nshell, npix, q):
    import torch
    import torch.distributed as dist

    from glass_b200.dist import multi_plane_block
    from glass_b200.sharding import shard_shells

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    deltas, wins = _mp_inputs(nshell, npix)
    mine = shard_shells(nshell, rank, world)
    conv = _CpuMultiPlane(_MockCosmo())
    conv._like = torch.empty(npix, dtype=torch.float64)
    kappas = multi_plane_block(conv, [deltas[i] for i in mine], [wins[i] for i in mine])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, list(mine), [k.numpy() for k in kappas]))


@pytest.mark.parametrize("world,nshell", [(2, 7), (3, 2)])
def test_multi_plane_pipeline_gloo(world, nshell):
    """Shell-sharded multi-plane recurrence: the state handed from rank to rank (five scalars,
    three maps) reproduces the serial recurrence bit for bit, including an empty last block."""
    import torch.multiprocessing as mp

    npix = 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_multiplane_worker, args=(r, world, port, nshell, npix, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    deltas, wins = _mp_inputs(nshell, npix)
    conv = _CpuMultiPlane(_MockCosmo())
    serial = []
    for d, w in zip(deltas, wins):
        conv.add_window(d, w)
        serial.append(conv.kappa.clone().numpy())
    seen = set()
    for rank, mine, kappas in res:
        assert len(mine) == len(kappas)
        for i, k in zip(mine, kappas):
            assert np.array_equal(k, serial[i])
            seen.add(i)
    assert seen == set(range(nshell))


def test_expm1_fast_host_build_and_run(tmp_path):
    """csrc/expm1_fast.cuh (the lognormal transform fused into the ring-FFT store) against
    80-bit expm1l on the host: <= 2 ulp over [-40, 40], tiny arguments, reduction boundaries,
    and the library's results for out-of-range / non-finite arguments."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "expm1_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(root, "tests", "native", "expm1_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "expm1_fast ok" in r.stdout, r.stdout + r.stderr


def test_legendre_tables_host_build_and_run(tmp_path):
    """csrc/sht_tables.cuh + dd.cuh (double-double coefficient tables of the Legendre stages):
    every entry within 0.51 ulp of the 80-bit value and no systematic offset in the product
    a_k a_{k-1} (the k^2-amplified error source found at lmax 8191)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "tables_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "native", "tables_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "tables ok" in r.stdout, r.stdout + r.stderr


def test_write_catalog_fits_roundtrip_host(tmp_path):
    """glass.write_catalog mirror (glass/user.py:169-205) with host arrays: a conforming FITS
    file (2880-byte blocks, primary HDU + one BINTABLE), rows in write order, NAXIS2 patched."""
    import glass_b200
    from glass_b200.user import read_catalog

    rng = np.random.default_rng(4)
    path = tmp_path / "cat.fits"
    parts = []
    with glass_b200.write_catalog(path, ext="CATALOG") as out:
        for n in (5, 0, 1234, 77):
            cols = {"RA": rng.uniform(0, 360, n), "DEC": rng.uniform(-90, 90, n), "Z": rng.random(n).astype(np.float32),
                    "ID": rng.integers(0, 2**40, n), "G": rng.standard_normal(n) + 1j * rng.standard_normal(n)}
            out.write(**cols)
            parts.append(cols)
    raw = path.read_bytes()
    assert len(raw) % 2880 == 0 and raw[:30].startswith(b"SIMPLE  =                    T")
    assert raw[2880:2880 + 20] == b"XTENSION= 'BINTABLE'"
    cat = read_catalog(path)
    assert cat["__extname__"] == "CATALOG"
    for name in ("RA", "DEC", "Z", "ID", "G"):
        want = np.concatenate([p[name] for p in parts])
        assert cat[name].dtype == want.dtype and np.array_equal(cat[name], want)
    # save_cls / load_cls (glass/user.py:41-86)
    cls = [np.arange(5.0), np.zeros(0), np.arange(3.0) + 10]
    glass_b200.save_cls(tmp_path / "cls.npz", cls)
    back = glass_b200.load_cls(tmp_path / "cls.npz")
    assert len(back) == 3 and all(np.array_equal(a, b) for a, b in zip(cls, back))
    with pytest.raises(ValueError, match="columns differ"):
        with glass_b200.write_catalog(tmp_path / "bad.fits") as out:
            out.write(A=np.zeros(2))
            out.write(B=np.zeros(2))


def test_legendre_recurrence_host_replay(tmp_path):
    """Host replay of the synthesis kernel's recurrence arithmetic at nside 4096, lmax 8191
    (double-double tables, per-ring recurrence variable, FMA chain) against 80-bit arithmetic on
    exact ring geometry: 9e-12 of the function's maximum -- the figure measured on the GPU by
    tests/test_gpu_fullsize.py -- and >= 10x worse for the two discarded formulations."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "legendre_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "native", "legendre_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "legendre recurrence ok" in r.stdout, r.stdout + r.stderr


def test_spin_recurrence_host_replay(tmp_path):
    """Host replay of the spin-2 synthesis kernel's recurrence arithmetic at nside 4096, lmax 8191
    (per-ring variable x or t = 1 - x) against 80-bit arithmetic on exact ring geometry."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "spin_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "native", "spin_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "spin recurrence ok" in r.stdout, r.stdout + r.stderr


def test_pixwin_generator_against_definition(monkeypatch):
    """glass_b200.pixwin (hp.pixwin, glass/healpix.py:313-356, generated instead of read from healpy's
    data files): the pair-moment series against the DEFINITION evaluated by brute force -- per pixel
    the quadrature average of every Y_lm (scipy) and spin-2 harmonic (oracle, Goldberg form), summed
    over m, averaged over pixels -- at nside 1 and 2; the closed forms of l = 1, 2; convergence of
    the triangle-split quadrature; nside > EXACT_NSIDE_MAX scaled in (l + 1/2)/nside; defaults and
    errors.  The device kernel behind the geometry is replaced by the oracle's pixel -> angle."""
    import torch
    from scipy.special import sph_harm_y

    import glass_b200.pixwin as PW
    from oracle import healpix_ref as H

    def positions(nside, ipix, u, v):
        th, ph = H.ring2ang_uv(nside, ipix.numpy().ravel(), u.numpy().ravel(), v.numpy().ravel())
        return torch.as_tensor(th).reshape(ipix.shape), torch.as_tensor(ph).reshape(ipix.shape)

    monkeypatch.setattr(PW, "_compute_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(PW, "_positions", positions)
    monkeypatch.setattr(PW, "_Q", 7)
    PW._pair_moments.cache_clear()

    def brute(nside, lmax, q=7):
        uu, vv, ww = PW._pixel_rule(q)  # the same nodes: the comparison isolates the pair-moment algebra
        npix = 12 * nside * nside
        wt, wp = np.zeros(lmax + 1), np.zeros(lmax + 1)
        for p in range(npix):
            th, ph = H.ring2ang_uv(nside, np.full(ww.size, p), uu, vv)
            for l in range(lmax + 1):
                for m in range(-l, l + 1):
                    wt[l] += 4 * np.pi / (2 * l + 1) * abs((ww * sph_harm_y(l, m, th, ph)).sum()) ** 2
                    if l >= 2:
                        wp[l] += 4 * np.pi / (2 * l + 1) * abs((ww * H.sYlm_goldberg(2, l, m, th, ph)).sum()) ** 2
        return np.sqrt(wt / npix), np.sqrt(wp / npix)

    for nside, lmax in ((1, 4), (2, 6)):
        wt, wp = PW.pixwin(nside, lmax=lmax, pol=True)
        bt, bp = brute(nside, lmax)
        assert np.abs(wt - bt).max() < 1e-13 and np.abs(wp - bp).max() < 1e-13
        assert wt[0] == 1.0 and wp[0] == wp[1] == 0.0
    # the quadrature itself: spectral convergence of the triangle-split rule (12 against 16 nodes per
    # axis), and the converged values at nside 2 that the GPU test checks on the device
    PW._pair_moments.cache_clear()
    monkeypatch.setattr(PW, "_Q", 16)
    w16q = PW.pixwin(2, lmax=8, pol=True)
    PW._pair_moments.cache_clear()
    monkeypatch.setattr(PW, "_Q", None)
    wt, wp = PW.pixwin(2, lmax=8, pol=True)
    assert np.abs(wt - w16q[0]).max() < 1e-14 and np.abs(wp - w16q[1]).max() < 1e-14
    assert np.abs(wt - [1.0, 0.977303, 0.93310702, 0.86971852, 0.79038278, 0.69905215, 0.60011811, 0.49813949, 0.39760902]).max() < 1e-8
    assert PW.pixwin(4).shape == (12,)
    # lowest multipoles in closed form: P_1 = 1 - 2x, P_2 = 1 - 6x + 6x^2 in x = sin^2(gamma/2)
    x0, mt, _mp = PW._pair_moments(8)
    w8 = PW.pixwin(8, lmax=2)
    assert abs(w8[1] ** 2 - (1 - 2 * x0 * mt[1])) < 1e-15 and abs(w8[2] ** 2 - (1 - 6 * x0 * mt[1] + 6 * x0**2 * mt[2])) < 1e-15
    # scaling above the exact range: nside 16 from the nside-8 moments against nside 16 itself
    w16 = PW.pixwin(16, lmax=40, pol=True)
    monkeypatch.setattr(PW, "EXACT_NSIDE_MAX", 8)
    s16 = PW.pixwin(16, lmax=40, pol=True)
    assert np.abs(s16[0] - w16[0]).max() < 2e-3 and np.abs(s16[1] - w16[1])[2:].max() < 5e-3 and s16[1][0] == s16[1][1] == 0.0
    with pytest.raises(ValueError, match="tabulated up to"):
        PW.pixwin(4, lmax=17)
    with pytest.raises(ValueError, match="power of two"):
        PW.pixwin(12)
    PW._pair_moments.cache_clear()


def test_vmap_and_rotator_host_side():
    """glass/observations.py:88-94 raises before any array work; the coordinate matrices are healpy's constants
    (rotations to 1e-9, the poles where the almanac puts them); no device is needed for either."""
    import glass_b200
    from glass_b200.healpix import _coordconv_matrix

    with pytest.raises(TypeError, match="galactic stripe must be a pair of numbers"):
        glass_b200.vmap_galactic_ecliptic(4, galactic=(1,))
    with pytest.raises(TypeError, match="ecliptic stripe must be a pair of numbers"):
        glass_b200.vmap_galactic_ecliptic(4, ecliptic=(1, 2, 3))
    for c in ("GC", "CE", "EG", ("G", "C"), "gq"):
        M = _coordconv_matrix(c)
        assert np.abs(M @ M.T - np.eye(3)).max() < 2e-9
    assert np.array_equal(_coordconv_matrix("C"), np.eye(3)) and np.array_equal(_coordconv_matrix(None), np.eye(3))
    assert np.allclose(_coordconv_matrix("GQ"), _coordconv_matrix("GC"))
    x, y, z = _coordconv_matrix("GC") @ np.array([0.0, 0.0, 1.0])
    assert abs(np.degrees(np.arctan2(y, x)) % 360 - 192.86) < 0.01 and abs(np.degrees(np.arcsin(z)) - 27.13) < 0.01
    with pytest.raises(TypeError):
        _coordconv_matrix("GX")


def test_int8_digit_arithmetic_host_build_and_run(tmp_path):
    """csrc/oz_digits.cuh (the digit arithmetic of the INT8 tensor-core Legendre kernel: magic-constant fixed point, the byte
    transposes that lay the digit planes out for the tensor core, the range check, the int32 -> double trick) built for the
    host and checked bit for bit -- the device code is the same source with the intrinsics in place of their host shims."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "oz_digits_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "native", "oz_digits_host.cpp")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "oz_digits ok" in r.stdout, r.stdout + r.stderr
