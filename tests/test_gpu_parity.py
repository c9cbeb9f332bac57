"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs.  Tolerances: complex128 1e-10 relative on localdensity / DensityCurrents (the
north-star bar; most checks are far tighter), complex64 1e-5.  Reference scenarios cited per
test (paths relative to /root/reference)."""
import ctypes as C
import warnings
from importlib import import_module

import numpy as np
import pytest
import scipy.sparse as sp

import lm_b200 as lm
from oracle import evolution as EV
from oracle import fields as F
from oracle import lattice as L
from oracle import observables as OB
from oracle import operators as OP
from oracle import spectrum as SP

pytestmark = pytest.mark.gpu
_lib = import_module("lm_b200._lib")


@pytest.fixture(scope="module")
def ctx():
    return lm.default_context("c128")


@pytest.fixture(scope="module")
def ctx64():
    return lm.default_context("c64")


def _rand_block(n, m, seed=1234, orth=True):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    if orth and m <= n:
        a, _ = np.linalg.qr(a)
    return np.ascontiguousarray(a)


def _relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


# ------------------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("M", [1, 2, 3, 17, 32, 33, 64, 200, 515])
def test_spmm_matches_oracle(ctx, M):
    H = OP.qwz(L.square_lattice(7, 6), field=F.LandauGauge(0.13))
    dev = lm.DeviceHam.from_csc(ctx, H, 2)
    X = np.asfortranarray(_rand_block(H.shape[0], M, orth=False))
    Y = np.zeros_like(X, order="F")
    _lib.check(_lib.load().lm_spmm(dev.handle, _lib.ptr(X), _lib.ptr(Y), H.shape[0], M))
    assert _relerr(Y, H @ X) < 1e-14


@pytest.mark.parametrize("case", ["square_t123", "haldane", "honeycomb_pbc"])
def test_spmm_stencils(ctx, case):
    if case == "square_t123":
        H = OP.tightbinding_hamiltonian(L.square_lattice(9, 8), t1=1, t2=0.4, t3=0.2)
    elif case == "haldane":
        H = OP.haldane(L.honeycomb_lattice(7, 6), 1.0, 0.2, 0.1)
    else:
        H = OP.haldane(L.honeycomb_lattice(4, 4, periodic=(1, 2)), 1.0, 0.3, 0.0)
    dev = lm.DeviceHam.from_csc(ctx, H, 1)
    X = np.asfortranarray(_rand_block(H.shape[0], 40, orth=False))
    Y = np.zeros_like(X, order="F")
    _lib.check(_lib.load().lm_spmm(dev.handle, _lib.ptr(X), _lib.ptr(Y), H.shape[0], 40))
    assert _relerr(Y, H @ X) < 1e-14
    assert abs(dev.to_csc() - H).max() == 0     # CSC round trip through the ELL layout


def test_csc_julia_one_based(ctx):
    """index_base = 1 (Julia SparseMatrixCSC) is accepted as is."""
    H = OP.tightbinding_hamiltonian(L.square_lattice(5, 5), field=F.LandauGauge(0.2))
    colptr = np.ascontiguousarray(H.indptr + 1, np.int64)
    rowval = np.ascontiguousarray(H.indices + 1, np.int64)
    nz = np.ascontiguousarray(H.data, np.complex128)
    h = C.c_void_p()
    lib = _lib.load()
    _lib.check(lib.lm_ham_create_csc(ctx.handle, 25, 1, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz), 1, C.byref(h)))
    X = np.asfortranarray(_rand_block(25, 6, orth=False))
    Y = np.zeros_like(X, order="F")
    _lib.check(lib.lm_spmm(h, _lib.ptr(X), _lib.ptr(Y), 25, 6))
    assert _relerr(Y, H @ X) < 1e-14
    cp2, rv2 = np.zeros(26, np.int64), np.zeros(H.nnz, np.int64)
    _lib.check(lib.lm_ham_get_csc(h, _lib.ptr(cp2), _lib.ptr(rv2), None))
    assert np.array_equal(cp2, colptr) and np.array_equal(rv2, rowval)
    lib.lm_ham_destroy(h)


@pytest.mark.parametrize("path", [0, 1, 2, 3, 4, 5])   # 3 = site-blocked (n_int = 2), 4 = TMA quad mapping, 5 = register-tiled stencil
@pytest.mark.parametrize("model", ["square", "qwz_pbc", "haldane"])
def test_all_spmm_kernels_on_plan(ctx, path, model):
    """Every SpMM kernel generation (consecutive-row gather, TMA-staged tiles, tile-order gather)
    against the oracle on device-built Hamiltonians with site coordinates, ragged widths."""
    lib = _lib.load()
    if model == "square":
        Hd = lm.tightbinding_hamiltonian(lm.SquareLattice(23, 17), t1=1, t2=0.3, field=lm.LandauGauge(0.07))
        Ho = OP.tightbinding_hamiltonian(L.square_lattice(23, 17), t1=1, t2=0.3, field=F.LandauGauge(0.07))
    elif model == "qwz_pbc":
        Hd = lm.qwz(lm.SquareLattice(14, 15, boundaries=[("axis1", True)]), field=lm.LandauGauge(0.5))
        Ho = OP.qwz(L.square_lattice(14, 15, periodic=(1,)), field=F.LandauGauge(0.5))
    else:
        Hd = lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1, field=lm.SymmetricGauge(0.03))
        Ho = OP.haldane(L.honeycomb_lattice(13, 11), 1.0, 0.2, 0.1, field=F.SymmetricGauge(0.03))
    dev = Hd.device(ctx)
    N = Ho.shape[0]
    try:
        lib.lm_dbg_set_apply_path(path)
        for M in (16, 31, 32, 40, 64, 100, 131, 300):
            X = _rand_block(N, M, seed=M, orth=False)
            x = lm.DeviceState.from_psi(X, ctx=ctx)
            y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
            _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
            assert _relerr(y.download(), Ho @ X) < 1e-14, (path, model, M)
        # one propagator step per kernel path (product form, Horner, Chebyshev epilogues)
        X = _rand_block(N, 40, seed=5)
        want = EV.exact_propagator(Ho, 0.3) @ X
        for method in ("taylor", "taylor_horner", "chebyshev", "chebyshev_clenshaw"):
            st = lm.DeviceState.from_psi(X, ctx=ctx)
            sol = lm.B200Exp(tol=1e-14, method=method, ctx=ctx)
            sol.update_solver(Hd, 0.3)
            sol.step(st)
            assert _relerr(st.download(), want) < 2e-13, (path, model, method)
    finally:
        lib.lm_dbg_set_apply_path(-1)


def test_csc_hamiltonian_with_site_coords(ctx):
    """A raw CSC Hamiltonian (host-assembled closure path) gets the tile plan through
    lm_ham_set_site_coords and must give the same results as without it."""
    lo, l = L.honeycomb_lattice(9, 8), lm.HoneycombLattice(9, 8)
    Ho = OP.haldane(lo, 1.0, 0.2, 0.1, field=F.LandauGauge(0.05))
    X = _rand_block(Ho.shape[0], 48, seed=3, orth=False)
    outs = []
    for coords in (None, l.coords):
        dev = lm.DeviceHam.from_csc(ctx, Ho, 1, coords=coords)
        Y = np.zeros_like(X, order="F")
        Xf = np.asfortranarray(X)
        _lib.check(_lib.load().lm_spmm(dev.handle, _lib.ptr(Xf), _lib.ptr(Y), Ho.shape[0], 48))
        outs.append(Y)
        st = lm.DeviceState.from_psi(X, np.linspace(0.1, 1, 48), ctx=ctx)
        V = lm.DensityCurrents(dev, st).pair_values()[2]
        outs.append(V)
    assert _relerr(outs[0], Ho @ X) < 1e-14 and _relerr(outs[2], Ho @ X) < 1e-14
    assert np.abs(outs[1] - outs[3]).max() < 1e-13
    with pytest.raises(lm.ArgumentError):
        lm.DeviceHam.from_csc(ctx, Ho, 1, coords=l.coords[:5])


# ------------------------------------------------------------------------------ register-tiled stencil kernel
def _stencil_info(dev):
    lib = _lib.load()
    i, rc, sw, m = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
    lib.lm_dbg_stencil_info.argtypes = [C.c_void_p] * 5
    _lib.check(lib.lm_dbg_stencil_info(dev.handle, C.byref(i), C.byref(rc), C.byref(sw), C.byref(m)))
    return i.value, rc.value, sw.value, m.value


STENCIL_CASES = {
    # name: (device Hamiltonian, oracle Hamiltonian, expected compiled pattern id)
    "square_nn": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(23, 17), field=lm.LandauGauge(0.07)),
                  lambda: OP.tightbinding_hamiltonian(L.square_lattice(23, 17), field=F.LandauGauge(0.07)), 0),
    "square_nnn": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(9, 21), t1=1, t2=0.3, field=lm.SymmetricGauge(0.05)),
                   lambda: OP.tightbinding_hamiltonian(L.square_lattice(9, 21), t1=1, t2=0.3, field=F.SymmetricGauge(0.05)), 1),
    "square_torus": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(8, 8, boundaries=[("axis1", True), ("axis2", True)]), field=lm.LandauGauge(0.125)),
                     lambda: OP.tightbinding_hamiltonian(L.square_lattice(8, 8, periodic=(1, 2)), field=F.LandauGauge(0.125)), 0),
    "square_3x3_torus": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(3, 3, boundaries=[("axis1", True), ("axis2", True)])),
                         lambda: OP.tightbinding_hamiltonian(L.square_lattice(3, 3, periodic=(1, 2))), 0),
    "honeycomb_nn": (lambda: lm.tightbinding_hamiltonian(lm.HoneycombLattice(7, 12), field=lm.LandauGauge(0.03)),
                     lambda: OP.tightbinding_hamiltonian(L.honeycomb_lattice(7, 12), field=F.LandauGauge(0.03)), 2),
    "qwz_pbc": (lambda: lm.qwz(lm.SquareLattice(14, 15, boundaries=[("axis1", True)]), field=lm.LandauGauge(0.5)),
                lambda: OP.qwz(L.square_lattice(14, 15, periodic=(1,)), field=F.LandauGauge(0.5)), 9),        # diagonal on-site term: exact mask
    # QWZ plus an on-site orbital-mixing term sigma_x: full 2 x 2 blocks everywhere (pattern 3)
    "qwz_sx": (lambda: lm.construct_hamiltonian(lm.SquareLattice(11, 12), 2, (np.array([[1, 0], [0, -1]], complex), 1.0), (np.array([[0, 1], [1, 0]], complex), 0.3),
                                                (np.array([[1, -1j], [-1j, -1]], complex) / 2, lm.BravaisTranslation(axis=1)),
                                                (np.array([[1, -1], [1, -1]], complex) / 2, lm.BravaisTranslation(axis=2)), field=lm.LandauGauge(0.1)),
               lambda: OP.construct_hamiltonian(L.square_lattice(11, 12), 2, [(np.array([[1, 0], [0, -1]], complex), 1.0), (np.array([[0, 1], [1, 0]], complex), 0.3),
                                                                              (np.array([[1, -1j], [-1j, -1]], complex) / 2, L.translation(axis=1, nu=2)),
                                                                              (np.array([[1, -1], [1, -1]], complex) / 2, L.translation(axis=2, nu=2))], F.LandauGauge(0.1)), 3),
    "haldane": (lambda: lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1, field=lm.SymmetricGauge(0.03)),
                lambda: OP.haldane(L.honeycomb_lattice(13, 11), 1.0, 0.2, 0.1, field=F.SymmetricGauge(0.03)), 4),
    "haldane_torus": (lambda: lm.haldane(lm.HoneycombLattice(9, 16, boundaries=[("axis1", True), ("axis2", True)]), 1.0, 0.2, 0.1),
                      lambda: OP.haldane(L.honeycomb_lattice(9, 16, periodic=(1, 2)), 1.0, 0.2, 0.1), 4),
}


@pytest.mark.parametrize("case", sorted(STENCIL_CASES))
def test_stencil_kernel_matches_oracle(ctx, case):
    """The register-tiled stencil kernel (default for unfiltered Bravais lattices, M >= 32) against
    the oracle matrix: open / periodic boundaries, ragged patch edges, ragged column chunks, every
    epilogue (plain SpMM, product-form factor, Horner, Clenshaw) through the propagators."""
    mk_dev, mk_or, want_id = STENCIL_CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    dev = Hd.device(ctx)
    sid, rc, sw, mask = _stencil_info(dev)
    assert sid == want_id, (case, sid, hex(mask))
    from oracle import stencil as ST
    # device detection == oracle restatement (ELL padding points at the own row, so the device may
    # add the diagonal bits of rows shorter than the widest one)
    rc_o, mask_o = ST.stencil_mask(Ho, *Hd.lattice.sizes)
    diag = sum(1 << (4 * rc * rc + a * rc + a) for a in range(rc))
    assert rc == rc_o and (mask | diag) == (mask_o | diag)
    lib = _lib.load()
    N = Ho.shape[0]
    for M in (32, 33, 40, 64, 100, 131):
        X = _rand_block(N, M, seed=M, orth=False)
        x = lm.DeviceState.from_psi(X, ctx=ctx)
        y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
        n0 = ctx.launch_count()
        _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
        assert ctx.launch_count() - n0 >= 1
        assert _relerr(y.download(), Ho @ X) < 1e-14, (case, M)
    X = _rand_block(N, 40, seed=5) if N >= 40 else _rand_block(N, 40, seed=5, orth=False)
    want = EV.exact_propagator(Ho, 0.3) @ X
    for method in ("taylor", "taylor_horner", "chebyshev", "chebyshev_clenshaw"):
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        sol = lm.B200Exp(tol=1e-14, method=method, ctx=ctx)
        sol.update_solver(Hd, 0.3)
        sol.step(st)
        assert _relerr(st.download(), want) < 2e-13, (case, method)


def test_stencil_not_used_for_filtered_or_long_range_lattices(ctx):
    """Patterns the stencil view cannot express keep the ELL kernels (and stay correct)."""
    # third-neighbour hops couple cells two apart
    Hd = lm.tightbinding_hamiltonian(lm.SquareLattice(12, 12), t1=1, t2=0.2, t3=0.1)
    Ho = OP.tightbinding_hamiltonian(L.square_lattice(12, 12), t1=1, t2=0.2, t3=0.1)
    dev = Hd.device(ctx)
    assert _stencil_info(dev)[0] == -1
    X = _rand_block(144, 48, orth=False)
    x = lm.DeviceState.from_psi(X, ctx=ctx)
    y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
    _lib.check(_lib.load().lm_spmm_state(dev.handle, x.handle, y.handle))
    assert _relerr(y.download(), Ho @ X) < 1e-14
    # a raw CSC Hamiltonian declared on a lattice it does not live on
    with pytest.raises(lm.ArgumentError):
        lm.DeviceHam.from_csc(ctx, Ho, 1, lattice_dims=(7, 5))
    # a raw CSC Hamiltonian with the right dims gets the stencil view
    Hn = OP.tightbinding_hamiltonian(L.square_lattice(12, 12), field=F.LandauGauge(0.1))
    dev2 = lm.DeviceHam.from_csc(ctx, Hn, 1, lattice_dims=(12, 12))
    assert _stencil_info(dev2)[0] == 0
    Y = np.zeros((144, 48), complex, order="F")
    Xf = np.asfortranarray(X)
    _lib.check(_lib.load().lm_spmm(dev2.handle, _lib.ptr(Xf), _lib.ptr(Y), 144, 48))
    assert _relerr(Y, Hn @ X) < 1e-14


def test_stencil_values_follow_field_updates_under_graph_replay(ctx):
    """The slot-ordered value copy is refreshed before a step graph is replayed: a time-dependent
    field must act on every step even when the first step graph was captured with fresh values."""
    lat, lo = lm.SquareLattice(12, 10), L.square_lattice(12, 10)
    X = _rand_block(240, 36, seed=9)
    st = lm.DeviceState.from_psi(X, ctx=ctx, n_int=2)
    y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx, n_int=2)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    want = X.copy()
    for k in range(4):
        B = 0.05 + 0.1 * k
        Hd = lm.qwz(lat, field=lm.LandauGauge(B))
        sol.update_solver(Hd, 0.2)
        if k == 0:      # eager SpMM first: the value copy is fresh when the step graph is captured
            _lib.check(_lib.load().lm_spmm_state(sol.dev.handle, st.handle, y.handle))
        sol.step(st)
        want = EV.exact_propagator(OP.qwz(lo, field=F.LandauGauge(B)), 0.2) @ want
    assert _relerr(st.download(), want) < 1e-12


@pytest.mark.parametrize("case", ["square_torus", "square_nnn", "qwz_pbc", "honeycomb_nn", "haldane", "haldane_torus"])
def test_stencil_observables_match_oracle(ctx, case):
    """Fused localdensity + DensityCurrents on the stencil view (streamed column groups, forward
    halo only, conjugate hand-off across periodic boundaries) against the oracle: every pair."""
    mk_dev, mk_or, _ = STENCIL_CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    n_int = Hd.n_int
    N = Ho.shape[0]
    lib = _lib.load()
    for M in (32, 70, 200):
        Psi = _rand_block(N, M, seed=M, orth=False) / np.sqrt(N)
        w = np.random.default_rng(M).random(M)
        st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=n_int)
        I, J, V = lm.DensityCurrents(Hd, st).pair_values()
        rho = lm.localdensity(st).values
        P = (Psi * w) @ Psi.conj().T
        assert _relerr(rho, np.real(np.diag(P)).reshape(-1, n_int).sum(1)) < 1e-13
        Hdn = Ho.toarray()
        want = np.zeros(len(I))
        for q, (i, j) in enumerate(zip(I.tolist(), J.tolist())):       # 1-based site pairs, i < j
            bi, bj = slice((i - 1) * n_int, i * n_int), slice((j - 1) * n_int, j * n_int)
            want[q] = 2 * np.imag(np.sum(Hdn[bi, bj] * P[bj, bi].T))
        assert np.abs(V - want).max() < 1e-13 * max(1.0, np.abs(want).max()), (case, M)
        # ... and the ELL-plan kernel gives the same numbers
        try:
            lib.lm_dbg_set_apply_path(2)
            V2 = lm.DensityCurrents(Hd, st).pair_values()[2]
        finally:
            lib.lm_dbg_set_apply_path(-1)
        assert np.abs(V - V2).max() < 1e-13 * max(1.0, np.abs(want).max())


def test_stencil_complex64(ctx64):
    for mk_dev, mk_or in ((lambda: lm.haldane(lm.HoneycombLattice(9, 10), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.04)),
                           lambda: OP.haldane(L.honeycomb_lattice(9, 10), 1.0, 0.2, 0.1, field=F.LandauGauge(0.04))),
                          (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(11, 7), field=lm.LandauGauge(0.04)),
                           lambda: OP.tightbinding_hamiltonian(L.square_lattice(11, 7), field=F.LandauGauge(0.04)))):
        Hd, Ho = mk_dev(), mk_or()
        dev = Hd.device(ctx64)
        assert _stencil_info(dev)[0] >= 0
        N = Ho.shape[0]
        for M in (32, 50, 64, 130):
            X = _rand_block(N, M, seed=M, orth=False)
            x = lm.DeviceState.from_psi(X, ctx=ctx64)
            y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx64)
            _lib.check(_lib.load().lm_spmm_state(dev.handle, x.handle, y.handle))
            assert _relerr(y.download(), Ho @ X) < 2e-6
        Psi = _rand_block(N, 96, seed=2)
        st = lm.DeviceState.from_psi(Psi, ctx=ctx64, n_int=Hd.n_int)
        I, J, V = lm.DensityCurrents(Hd, st).pair_values()
        P = Psi @ Psi.conj().T
        Hdn = Ho.toarray()
        want = np.array([2 * np.imag(Hdn[i - 1, j - 1] * P[j - 1, i - 1]) for i, j in zip(I.tolist(), J.tolist())])
        assert np.abs(V - want).max() < 1e-5 * max(1.0, np.abs(want).max())
        assert _relerr(lm.localdensity(st).values, np.real(np.diag(P))) < 1e-5


# ------------------------------------------------------------------------------ N4: async frame sink, region sums
def test_async_frame_sink_matches_synchronous_frames(ctx):
    """lm_observables_async / lm_frame_wait (double-buffered D2H on a second stream) deliver the
    same frames as the synchronous lm_observables while the evolution keeps stepping."""
    lat = lm.HoneycombLattice(9, 8)

    def ham(t):
        return lm.haldane(lat, 1.0, 0.2, 0.1, field=lm.LandauGauge(0.02 * t))
    Psi = _rand_block(144, 40, seed=11)
    w = np.random.default_rng(5).random(40)
    ts = [0.1 * k for k in range(1, 8)]
    frames = []
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    for t in ts:
        sol.update_solver(ham(t), 0.1)
        sol.step(st)
        frames.append((lm.localdensity(st).values.copy(), lm.DensityCurrents(ham(t), st).pair_values()[2].copy()))
    st2 = lm.DeviceState.from_psi(Psi, w, ctx=ctx)
    sink = lm.AsyncFrameSink()
    for t in ts:
        sol.update_solver(ham(t), 0.1)
        sol.step(st2)
        sink.push(sol.dev, st2, t)
    rho_seq, cur_seq = sink.finish()
    assert len(rho_seq) == len(ts) and len(cur_seq) == len(ts)
    for t, (rho, J) in zip(ts, frames):
        assert np.abs(rho_seq[t] - rho).max() < 1e-14
        assert np.abs(cur_seq[t] - J).max() < 1e-14
    # slot discipline
    lib = _lib.load()
    _lib.check(lib.lm_observables_async(sol.dev.handle, st2.handle, 0, 1))
    with pytest.raises(lm.ArgumentError):
        _lib.check(lib.lm_observables_async(sol.dev.handle, st2.handle, 0, 1))     # slot 0 not read yet
    with pytest.raises(lm.ArgumentError):
        _lib.check(lib.lm_frame_wait(ctx.handle, 1, None, None))                   # nothing enqueued in slot 1
    with pytest.raises(lm.ArgumentError):
        _lib.check(lib.lm_observables_async(sol.dev.handle, st2.handle, 2, 1))
    _lib.check(lib.lm_frame_wait(ctx.handle, 0, None, None))


def test_region_sums_on_device(ctx):
    """currentsfromto / currentsfrom (src/currents.jl:85-109) summed on the device against the
    host sums over the materialised Currents matrix; mask and index-list regions; QWZ (n_int=2)."""
    l = lm.SquareLattice(7, 6)
    Hd = lm.qwz(l, field=lm.LandauGauge(0.1))
    Psi = _rand_block(84, 36, seed=3)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx, n_int=2)
    dc = lm.DensityCurrents(Hd, st)
    full = lm.Currents(dc).currents.toarray()
    rng = np.random.default_rng(1)
    src = np.sort(rng.choice(42, 11, replace=False)) + 1
    dst = np.setdiff1d(np.arange(1, 43), src)[::2]
    want = full[np.ix_(src - 1, dst - 1)].sum()
    assert lm.currentsfromto(dc, src, dst) == pytest.approx(want, abs=1e-13)
    rest = np.setdiff1d(np.arange(1, 43), src)
    assert lm.currentsfromto(dc, src) == pytest.approx(full[np.ix_(src - 1, rest - 1)].sum(), abs=1e-13)
    mask = np.zeros(42, bool)
    mask[src - 1] = True
    assert lm.currentsfromto(dc, mask) == pytest.approx(lm.currentsfromto(dc, src), abs=1e-14)
    got = lm.currentsfrom(dc, src).values
    wantv = full[src - 1, :].sum(0)
    wantv[src - 1] = 0
    assert np.abs(got - wantv).max() < 1e-13
    # overlapping regions: i in src and j in dst counted per ordered pair, like the reference's double sum
    both = np.arange(1, 20)
    assert lm.currentsfromto(dc, both, both) == pytest.approx(full[np.ix_(both - 1, both - 1)].sum(), abs=1e-13)
    # reuse of the most recent frame: no state needed
    lib = _lib.load()
    dev = Hd.device(ctx)
    J = np.zeros(len(dev.pairs()[0]))
    _lib.check(lib.lm_observables(dev.handle, st.handle, None, _lib.ptr(J)))
    out = C.c_double()
    m = np.zeros(42, np.uint8)
    m[src - 1] = 1
    _lib.check(lib.lm_currents_fromto(dev.handle, None, _lib.ptr(m), None, 1, C.byref(out)))
    assert out.value == pytest.approx(full[np.ix_(src - 1, rest - 1)].sum(), abs=1e-13)


# ------------------------------------------------------------------------------ device Peierls phases
FIELD_CASES = {
    "nofield": (lambda: lm.NoField(), lambda: F.NoField()),
    "landau": (lambda: lm.LandauGauge(0.1), lambda: F.LandauGauge(0.1)),
    "symmetric": (lambda: lm.SymmetricGauge(0.07), lambda: F.SymmetricGauge(0.07)),
    "axial": (lambda: lm.PointFlux(0.13, (5.5, 5.5)), lambda: F.PointFlux(0.13, (5.5, 5.5))),
    "axial_collinear": (lambda: lm.PointFlux(0.2, (5.0, 5.5)), lambda: F.PointFlux(0.2, (5.0, 5.5))),
    "axial_onsite": (lambda: lm.PointFlux(0.2, (3.0, 4.0)), lambda: F.PointFlux(0.2, (3.0, 4.0))),
    "singular": (lambda: lm.PointFlux(0.3, (4.5, 4.5), gauge="singular"), lambda: F.PointFlux(0.3, (4.5, 4.5), "singular")),
    "fluxes": (lambda: lm.PointFluxes([0.1, -0.2], [(2.5, 2.5), (6.5, 7.5)]), lambda: F.PointFluxes([0.1, -0.2], [(2.5, 2.5), (6.5, 7.5)])),
    "sum": (lambda: lm.LandauGauge(0.05) + lm.PointFlux(0.3, (3.5, 3.5)), lambda: F.LandauGauge(0.05) + F.PointFlux(0.3, (3.5, 3.5))),
}


@pytest.mark.parametrize("name", sorted(FIELD_CASES))
@pytest.mark.parametrize("model", ["tb", "qwz"])
def test_device_phases_match_oracle(ctx, name, model):
    """Device-regenerated H == host-assembled reference H for every closed-form field
    (src/zoo/magneticfields.jl:15,31,72-104 incl. the 1e-11 guards and the acos fudge)."""
    fl, fo = FIELD_CASES[name]
    if model == "tb":
        Hd = lm.tightbinding_hamiltonian(lm.SquareLattice(9, 10), t1=1, t2=0.3, field=fl())
        Ho = OP.tightbinding_hamiltonian(L.square_lattice(9, 10), t1=1, t2=0.3, field=fo())
    else:
        Hd = lm.qwz(lm.SquareLattice(9, 10), field=fl())
        Ho = OP.qwz(L.square_lattice(9, 10), field=fo())
    got = Hd.device(ctx).to_csc()
    assert abs(got - Ho).max() < 5e-15


def test_device_phases_pbc_and_twist(ctx):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Hd = lm.tightbinding_hamiltonian(
            lm.SquareLattice(5, 6, boundaries=[("axis1", True), ("axis2", 0.7)]), t1=1, t2=0.3, t3=0.1,
            field=lm.PointFlux(0.2, (2.5, 2.5), gauge="singular") + lm.LandauGauge(0.2))
    Ho = OP.tightbinding_hamiltonian(L.square_lattice(5, 6, periodic=(1,), twists={2: 0.7}), t1=1, t2=0.3, t3=0.1,
                                     field=F.PointFlux(0.2, (2.5, 2.5), "singular") + F.LandauGauge(0.2))
    assert abs(Hd.device(ctx).to_csc() - Ho).max() < 5e-15
    Hh = lm.haldane(lm.HoneycombLattice(4, 5, boundaries=[("axis1", True)]), 1.0, 0.2, 0.3, field=lm.LandauGauge(0.11))
    Hho = OP.haldane(L.honeycomb_lattice(4, 5, periodic=(1,)), 1.0, 0.2, 0.3, field=F.LandauGauge(0.11))
    assert abs(Hh.device(ctx).to_csc() - Hho).max() < 5e-15


def test_field_param_update_path(ctx):
    """lm_ham_set_field_params (a few doubles) == full re-assembly, repeatedly."""
    l = lm.SquareLattice(8, 8)
    lo = L.square_lattice(8, 8)
    for B in (0.0, 0.05, 0.2, 0.05):
        got = lm.tightbinding_hamiltonian(l, field=lm.PointFlux(B, (4.5, 4.5))).device(ctx).to_csc()
        want = OP.tightbinding_hamiltonian(lo, field=F.PointFlux(B, (4.5, 4.5)))
        assert abs(got - want).max() < 5e-15
    lo_b, hi_b = lm.tightbinding_hamiltonian(l).device(ctx).spectral_bounds()
    assert lo_b <= -3.9 and hi_b >= 3.9        # Gershgorin encloses [-4, 4]


# ------------------------------------------------------------------------------ propagator
@pytest.mark.parametrize("method", ["taylor", "taylor_horner", "chebyshev", "chebyshev_clenshaw", "lanczos", "auto"])
@pytest.mark.parametrize("dt", [0.1, 0.7, 3.0, -0.4, 12.0])
def test_step_matches_exact_exponential(ctx, method, dt):
    Ho = OP.qwz(L.square_lattice(6, 5), field=F.LandauGauge(0.1))
    Psi = _rand_block(60, 13)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol = lm.B200Exp(tol=1e-14, method=method, ctx=ctx, n_int=2)
    sol.update_solver(Ho, dt)
    sol.step(st)
    want = EV.exact_propagator(Ho, dt) @ Psi
    assert _relerr(st.download(), want) < 2e-13
    assert sol.n_matvec > 0


def test_evolution_known_answer_gpu(ctx):
    """test/test_timedeps.jl:42-68: qwz(SquareLattice(10,10)) ground state over 0:0.1:10,
    component psi[2] every frame vs repeated dense exp(-i 0.1 H) products, atol 1e-10."""
    H = lm.qwz(lm.SquareLattice(10, 10))
    Hd = OP.qwz(L.square_lattice(10, 10)).toarray()
    psi = SP.groundstate(Hd)
    ts = np.arange(0, 101) * 0.1
    U = EV.exact_propagator(Hd, 0.1)
    correct, v = [], psi.copy()
    for _ in ts:
        correct.append(v[1])
        v = U @ v
    for method in ("taylor", "chebyshev", "lanczos"):      # lanczos = KrylovKitExp semantics
        vals = [m.state.data[1] for m in lm.Evolution(lm.B200Exp(method=method, ctx=ctx), H, psi)(ts)]
        assert np.abs(np.array(vals) - np.array(correct)).max() < 1e-10
    # CachedExp-style constant sparse matrix passed as a raw CSC (update_solver! identity skip)
    vals = [m[0].data[1] for m in lm.Evolution(lm.B200Exp(ctx=ctx, n_int=2), sp.csc_matrix(Hd), psi)(ts)]
    assert np.abs(np.array(vals) - np.array(correct)).max() < 1e-10


def test_stepping_semantics_and_errors(ctx):
    """src/evolution.jl:238-250,266-275: H(t_old) pairing, dt = 0 first frame, negative dt."""
    l = lm.SquareLattice(3, 3)
    lo = L.square_lattice(3, 3)
    calls = []

    def h(t):
        calls.append(t)
        return lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(t))
    psi = np.zeros(9, complex)
    psi[4] = 1
    ev = lm.Evolution(lm.B200Exp(tol=1e-14, ctx=ctx), h, psi)
    calls.clear()
    frames = [(m.state.data.copy(), m.H, m.t) for m in ev([0.0, 0.5, 1.0])]
    assert calls == [0.0, 0.0, 0.5]
    assert np.array_equal(frames[0][0], psi)
    U0 = EV.exact_propagator(OP.tightbinding_hamiltonian(lo, field=F.LandauGauge(0.0)), 0.5)
    U1 = EV.exact_propagator(OP.tightbinding_hamiltonian(lo, field=F.LandauGauge(0.5)), 0.5)
    assert np.abs(frames[2][0] - U1 @ (U0 @ psi)).max() < 1e-13
    assert frames[2][1].field.B == 0.5 and frames[2][2] == pytest.approx(1.0)
    with pytest.raises(lm.ArgumentError, match="negative time step"):
        ev.step(-0.1)
    nxt = list(ev([1.5]))                      # stateful continuation
    assert nxt[0].t == pytest.approx(1.5)
    with pytest.raises(lm.ArgumentError):
        lm.Evolution(lm.B200Exp(ctx=ctx), h)   # no states
    with pytest.raises(lm.ArgumentError):
        lm.Evolution(lm.B200Exp(ctx=ctx), h, psi, P=psi)   # named + unnamed
    # dimension mismatch surfaces as ArgumentError from the C ABI
    with pytest.raises(lm.ArgumentError, match="dimension mismatch"):
        sol = lm.B200Exp(ctx=ctx)
        sol.update_solver(lm.tightbinding_hamiltonian(lm.SquareLattice(4, 4)), 0.1)
        sol.step(lm.DeviceState.from_psi(psi, ctx=ctx))


def test_host_assembled_closure_path(ctx):
    """Arbitrary closure t -> sparse matrix: same pattern, nzval re-uploaded each step
    (lm_ham_update_values), results equal the device-phase path."""
    lo = L.square_lattice(6, 6)
    l = lm.SquareLattice(6, 6)
    Psi = _rand_block(36, 9)
    ts = np.arange(0, 6) * 0.1
    ev_a = lm.Evolution(lm.B200Exp(tol=1e-14, ctx=ctx), lambda t: OP.tightbinding_hamiltonian(lo, field=F.PointFlux(0.3 * t, (3.5, 3.5))), lm.PsiProjector(Psi))
    ev_b = lm.Evolution(lm.B200Exp(tol=1e-14, ctx=ctx), lambda t: lm.tightbinding_hamiltonian(l, field=lm.PointFlux(0.3 * t, (3.5, 3.5))), lm.PsiProjector(Psi))
    ref = EV.Evolution(lambda t: OP.tightbinding_hamiltonian(lo, field=F.PointFlux(0.3 * t, (3.5, 3.5))), [Psi], solver="exact", block=True)
    for ma, mb, (st, H, t) in zip(ev_a(ts), ev_b(ts), ref(ts)):
        assert _relerr(ma.state.download(), st[0]) < 1e-12
        assert _relerr(mb.state.download(), st[0]) < 1e-12


def test_lanczos_block_restart_and_dense_rejection(ctx):
    """Large dt forces the krylovdim = 30 restart path; dense states are rejected like
    KrylovKitExp rejects matrices (src/evolution.jl:150, docs/src/manual/evolution.md:217)."""
    Ho = OP.haldane(L.honeycomb_lattice(8, 8), 1.0, 0.3, 0.2, field=F.LandauGauge(0.05))
    Psi = _rand_block(128, 37, seed=11)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol = lm.B200Exp(tol=1e-12, method="lanczos", ctx=ctx)
    sol.update_solver(Ho, 25.0)
    sol.step(st)
    assert _relerr(st.download(), EV.exact_propagator(Ho, 25.0) @ Psi) < 1e-10
    assert sol.n_matvec > 30
    P = np.eye(128, dtype=complex) / 128
    with pytest.raises(lm.ArgumentError, match="Lanczos"):
        sol.step(lm.DeviceState.from_dense(P, ctx=ctx))


# ------------------------------------------------------------------------------ observables
@pytest.mark.parametrize("M", [1, 5, 16, 40, 600])
def test_observables_match_oracle(ctx, M):
    lo = L.square_lattice(20, 20) if M == 600 else L.square_lattice(6, 5)
    l = lm.SquareLattice(20, 20) if M == 600 else lm.SquareLattice(6, 5)
    Ho = OP.qwz(lo, field=F.LandauGauge(0.1))
    Hd = lm.qwz(l, field=lm.LandauGauge(0.1))
    N = Ho.shape[0]
    Psi = _rand_block(N, M, seed=7)
    w = np.random.default_rng(3).random(M)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=2)
    ost = OB.State(Psi, w, block=True)
    rho = lm.localdensity(st)
    assert _relerr(rho.values, OB.localdensity(ost, 2)) < 1e-13
    dc = lm.DensityCurrents(Hd, st)
    I, J, V = dc.pair_values()
    pairs = OB.site_adjacency(Ho, 2)
    assert list(zip(I.tolist(), J.tolist())) == pairs          # findnz ordering (J major, I minor)
    sub = pairs[:: max(1, len(pairs) // 60)]
    idx = [pairs.index(p) for p in sub]
    want = np.array([OB.density_current(Ho, ost, i, j, 2) for i, j in sub])
    assert np.abs(V[idx] - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    i0, j0 = sub[0]
    assert dc[i0, j0] == pytest.approx(want[0], abs=1e-13)
    assert dc[j0, i0] == pytest.approx(-want[0], abs=1e-13)      # antisymmetry (test_currents.jl:14)
    assert dc[i0, i0] == 0.0                                     # zero self current (:15)


def test_currents_reference_scenarios(ctx):
    """test/test_currents.jl:1-63 on the device: Heisenberg sum, Ket vs block representation,
    Currents(dc) == Currents(dc, adjacency), findnz filter."""
    lo, l = L.square_lattice(4, 4), lm.SquareLattice(4, 4)
    H0o, H1o = OP.qwz(lo), OP.qwz(lo, field=F.LandauGauge(0.1))
    H1 = lm.qwz(l, field=lm.LandauGauge(0.1))
    P, Psi, w = SP.densitymatrix(H0o, mu=0.0)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=2)
    dc = lm.DensityCurrents(H1, st)
    s1 = 6
    nbrs = [s1 + 4, s1 + 1, s1 - 4, s1 - 1]
    n_op = np.zeros((32, 32), complex)
    for a in range(2):
        n_op[(s1 - 1) * 2 + a, (s1 - 1) * 2 + a] = 1
    Hd = H1o.toarray()
    dens_dt = np.trace(1j * (Hd @ n_op - n_op @ Hd) @ P).real
    assert sum(dc[s1, t] for t in nbrs) == pytest.approx(dens_dt, abs=1e-13)
    assert lm.currentsfromto(dc, s1) == pytest.approx(dens_dt, abs=1e-12)
    full = lm.Currents(dc)
    want = OB.currents_matrix(H1o, P, 2)
    assert abs(full.currents - want).max() < 1e-13
    Is, Js, Vs = lm.findnz(dc)
    Io, Jo, Vo = OB.currents_findnz(H1o, P, 2)
    assert np.array_equal(Is, Io) and np.array_equal(Js, Jo) and np.abs(Vs - Vo).max() < 1e-13
    adj = [(i, j) for i, j in OB.site_adjacency(H1o, 2)]
    assert abs(lm.Currents(dc, adj).currents - full.currents).max() == 0
    gs = SP.groundstate(H0o)
    c1 = lm.Currents(lm.DensityCurrents(H1, lm.DeviceState.from_psi(gs, ctx=ctx, n_int=2)))
    c2 = OB.currents_matrix(H1o, np.outer(gs, gs.conj()), 2)
    assert abs(c1.currents - c2).max() < 1e-13
    assert abs((full + full).currents - (2 * full).currents).max() == 0


# ------------------------------------------------------------------------------ dense-P path (DMMA)
def test_dense_density_matrix_step(ctx):
    """P <- U P U' on the FP64 tensor cores == CachedExp semantics (src/evolution.jl:73-78)."""
    Ho = OP.qwz(L.square_lattice(5, 4), field=F.LandauGauge(0.1))
    N = Ho.shape[0]
    P, Psi, w = SP.densitymatrix(Ho, T=0.5, mu=0.1)          # mixed state, all columns
    st = lm.DeviceState.from_dense(P, ctx=ctx, n_int=2)
    sol = lm.B200Exp(tol=1e-14, ctx=ctx, n_int=2)
    U = EV.exact_propagator(Ho, 0.1)
    want = P.copy()
    for _ in range(3):
        sol.update_solver(Ho, 0.1)
        sol.step(st)
        want = U @ want @ U.conj().T
    assert _relerr(st.download(), want) < 1e-12
    assert _relerr(lm.localdensity(st).values, OB.localdensity(want, 2)) < 1e-12
    # Psi-block escape hatch: dense(P) materialises Psi W Psi'
    stb = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=2)
    assert _relerr(stb.dense(), P) < 1e-13


def test_readme_workflow_config1(ctx):
    """README.md:25-55 = BASELINE config 1: SquareLattice(10,10), PointFlux ramp, mu = 0 density
    matrix, 0:0.1:20, localdensity + DensityCurrents per frame.  Dense-P state (as the reference
    evolves it) and the Psi-block reformulation both track the oracle to 1e-10 relative."""
    lo, l = L.square_lattice(10, 10), lm.SquareLattice(10, 10)
    tau = 10.0

    def h_ref(t):
        return OP.tightbinding_hamiltonian(lo, field=F.PointFlux(0.2 * min(t, tau) / tau, (5.5, 5.5)))

    def h_dev(t):
        return lm.tightbinding_hamiltonian(l, field=lm.PointFlux(0.2 * min(t, tau) / tau, (5.5, 5.5)))
    P0, Psi0, w0 = SP.densitymatrix(h_ref(0.0), mu=0.0)        # shared input (degenerate Fermi level)
    ts = np.arange(0, 201) * 0.1
    pairs = OB.site_adjacency(h_ref(0.0), 1)
    ev_ref = EV.Evolution(h_ref, [P0], solver="exact")
    ev_dense = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), h_dev, P=P0)
    ev_block = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), h_dev, lm.PsiProjector(Psi0, w0))
    worst_rho = worst_j = 0.0
    for k, ((st, H, t), md, mb) in enumerate(zip(ev_ref(ts), ev_dense(ts), ev_block(ts))):
        if k % 10 and k != 200:
            continue
        rho = OB.localdensity(st[0], 1)
        Jw = np.array([OB.density_current(H, st[0], i, j, 1) for i, j in pairs])
        rb = lm.localdensity(mb.state).values
        rd = lm.localdensity(md.P).values
        Pb, Hb, tb = mb
        Jb = lm.DensityCurrents(Hb, Pb).pair_values()[2]
        worst_rho = max(worst_rho, _relerr(rb, rho), _relerr(rd, rho))
        big = np.abs(Jw) >= 1e-10
        if big.any():
            worst_j = max(worst_j, np.abs(Jb - Jw)[big].max() / np.abs(Jw).max())
        assert tb == pytest.approx(t)
    assert worst_rho < 1e-10 and worst_j < 1e-10


# ------------------------------------------------------------------------------ complex64 mode
def test_complex64_mode(ctx64):
    Ho = OP.qwz(L.square_lattice(6, 6), field=F.LandauGauge(0.1))
    Hd = lm.qwz(lm.SquareLattice(6, 6), field=lm.LandauGauge(0.1))
    Psi = _rand_block(72, 20)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx64, n_int=2)
    sol = lm.B200Exp(tol=1e-7, ctx=ctx64)
    want = Psi.copy()
    U = EV.exact_propagator(Ho, 0.1)
    for _ in range(10):
        sol.update_solver(Hd, 0.1)
        sol.step(st)
        want = U @ want
    assert _relerr(st.download(), want) < 1e-5
    rho = lm.localdensity(st).values
    assert _relerr(rho, OB.localdensity(OB.State(want, None, block=True), 2)) < 1e-5
    V = lm.DensityCurrents(Hd, st).pair_values()[2]
    pairs = OB.site_adjacency(Ho, 2)
    Jw = np.array([OB.density_current(Ho, OB.State(want, None, block=True), i, j, 2) for i, j in pairs])
    assert np.abs(V - Jw).max() < 1e-5 * max(1.0, np.abs(Jw).max())


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_config2_properties(ctx):
    """BASELINE config 2 (SquareLattice(100,100), N = 1e4, M = 5e3, complex128): size-independent
    properties - unitarity (column norms / particle number), linearity, continuity
    d rho_i/dt = sum_j J_ij by finite differences, round trip exp(-iHdt) exp(+iHdt) = 1."""
    l = lm.SquareLattice(100, 100)
    H = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.02))
    N, M = 10000, 5000
    rng = np.random.default_rng(1234)
    Psi = (rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M))) / np.sqrt(2 * N)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    rho0 = lm.localdensity(st).values
    J0 = lm.Currents(lm.DensityCurrents(H, st)).currents
    n0 = rho0.sum()
    eps = 1e-4
    sol.update_solver(H, eps)
    sol.step(st)
    rho1 = lm.localdensity(st).values
    assert abs(rho1.sum() - n0) < 1e-11 * n0                               # particle number
    drho = (rho1 - rho0) / eps
    assert np.abs(drho - np.asarray(J0.sum(axis=1)).ravel()).max() < 5e-3 * np.abs(drho).max()   # continuity (O(eps))
    sol.update_solver(H, -eps)
    sol.step(st)                                                           # back to t = 0
    sol.update_solver(H, 0.1)
    sol.step(st)
    sol.update_solver(H, -0.1)
    sol.step(st)
    back = st.download()
    assert np.abs(back - Psi).max() < 1e-12 * np.abs(Psi).max() * 50      # round trip
    # a column slice against the ORACLE-assembled H (oracle/operators.py): device CSC, SpMM, and one
    # propagation step against scipy's expm_multiply
    from scipy.sparse.linalg import expm_multiply
    Ho = OP.tightbinding_hamiltonian(L.square_lattice(100, 100), field=F.LandauGauge(0.02)).tocsc()
    got = H.device(ctx).to_csc()
    assert abs(got - Ho).max() < 5e-15
    Y = np.zeros((N, 8), complex, order="F")
    X = np.asfortranarray(Psi[:, :8])
    _lib.check(_lib.load().lm_spmm(H.device(ctx).handle, _lib.ptr(X), _lib.ptr(Y), N, 8))
    assert _relerr(Y, Ho @ X) < 1e-14
    sol.update_solver(H, 0.1)
    sol.step(st)
    assert _relerr(st.download()[:, :8], expm_multiply(-0.1j * Ho, back[:, :8])) < 1e-11


@pytest.mark.parametrize("config", ["c3_qwz300", "c4_haldane500"])
def test_full_size_headline_configs_properties(ctx, config):
    """BASELINE configs 3 / 4 at their full lattice sizes (N = 1.8e5 / 5e5) with a 256-column
    block: size-independent properties - particle number, continuity d rho_i/dt = sum_j J_ij,
    exp(+iHdt) exp(-iHdt) = 1, Lanczos (KrylovKit semantics) == product-form Taylor, and SpMM
    linearity H(aX + bY) = aHX + bHY."""
    if config == "c3_qwz300":
        H = lm.qwz(lm.SquareLattice(300, 300), field=lm.LandauGauge(0.05))
    else:
        H = lm.haldane(lm.HoneycombLattice(500, 500), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.002))
    dev = H.device(ctx)
    N, M = dev.N, 256
    rng = np.random.default_rng(1234)
    blk = (rng.standard_normal((N, 32)) + 1j * rng.standard_normal((N, 32))) / np.sqrt(2 * N)
    Psi = np.asfortranarray(np.concatenate([blk * np.exp(0.37j * k) * (1 + 0.01 * k) for k in range(M // 32)], axis=1))
    w = np.linspace(0.2, 1.0, M)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=H.n_int)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    rho0 = lm.localdensity(st).values
    I, J, V = lm.DensityCurrents(H, st).pair_values()
    n_sites = N // H.n_int
    div = np.zeros(n_sites)
    np.add.at(div, I - 1, V)          # sum_j J_ij with J_ji = -J_ij
    np.add.at(div, J - 1, -V)
    eps = 1e-4
    sol.update_solver(H, eps)
    sol.step(st)
    rho1 = lm.localdensity(st).values
    assert abs(rho1.sum() - rho0.sum()) < 1e-11 * rho0.sum()
    drho = (rho1 - rho0) / eps
    assert np.abs(drho - div).max() < 5e-3 * np.abs(drho).max()
    sol.update_solver(H, -eps)
    sol.step(st)
    st2 = st.copy()
    sol.update_solver(H, 0.1)
    sol.step(st)
    lz = lm.B200Exp(tol=1e-13, method="lanczos", ctx=ctx)
    lz.update_solver(H, 0.1)
    lz.step(st2)
    a, b = st.download(), st2.download()
    assert np.abs(a - b).max() < 1e-11 * np.abs(a).max() * 10
    # against the ORACLE at full size: the device-assembled H equals the oracle's CSC entry by entry, and
    # 8 columns of the step equal scipy's expm_multiply on the oracle-assembled H
    from scipy.sparse.linalg import expm_multiply
    if config == "c3_qwz300":
        Ho = OP.qwz(L.square_lattice(300, 300), field=F.LandauGauge(0.05)).tocsc()
    else:
        Ho = OP.haldane(L.honeycomb_lattice(500, 500), 1.0, 0.2, 0.1, field=F.LandauGauge(0.002)).tocsc()
    assert abs(dev.to_csc() - Ho).max() < 5e-15
    assert _relerr(a[:, :8], expm_multiply(-0.1j * Ho, Psi[:, :8])) < 1e-11
    sol.update_solver(H, -0.1)
    sol.step(st)
    assert np.abs(st.download() - Psi).max() < 5e-11 * np.abs(Psi).max()
    # linearity of the SpMM on device-resident blocks
    x = lm.DeviceState.from_psi(Psi[:, :64], ctx=ctx)
    y = lm.DeviceState.from_psi(Psi[:, 64:128], ctx=ctx)
    zsum = lm.DeviceState.from_psi(2.0 * Psi[:, :64] - 0.5j * Psi[:, 64:128], ctx=ctx)
    outs = []
    for s_in in (x, y, zsum):
        o = lm.DeviceState.from_psi(np.zeros((N, 64), complex), ctx=ctx)
        _lib.check(_lib.load().lm_spmm_state(dev.handle, s_in.handle, o.handle))
        outs.append(o.download())
    lin = 2.0 * outs[0] - 0.5j * outs[1]
    assert np.abs(outs[2] - lin).max() < 1e-13 * max(np.abs(lin).max(), 1e-300) * 10


def test_localexpect_and_operator_currents(ctx):
    """SURVEY 8f N3 on the device: localexpect (src/operators/latticeutils.jl:13-20) and
    LocalOperatorCurrents (src/zoo/currents.jl:150-184) against the oracle + the reference's
    identities (test/test_currents.jl:76-91, test/test_operators.jl:28)."""
    lo, l = L.square_lattice(6, 5), lm.SquareLattice(6, 5)
    Ho = OP.qwz(lo, field=F.LandauGauge(0.1))
    Hd = lm.qwz(l, field=lm.LandauGauge(0.1))
    Psi = _rand_block(60, 23, seed=9)
    w = np.random.default_rng(2).random(23)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=2)
    ost = OB.State(Psi, w, block=True)
    sz = np.array([[1, 0], [0, -1]], complex)
    sy = np.array([[0, -1j], [1j, 0]], complex)
    for op in (np.eye(2), sz, sy, sz + 0.3 * sy):
        got = lm.localexpect(op, st).values
        assert np.abs(got - OB.localexpect(op, ost, 2)).max() < 1e-13
        loc = lm.LocalOperatorCurrents(Hd, st, op)
        I, J, V = loc.pair_values()
        want = np.array([OB.operator_current(Ho, ost, op, i, j, 2) for i, j in zip(I, J)])
        assert np.abs(V - want).max() < 1e-13
    assert np.abs(lm.localexpect(np.eye(2), st).values.real - lm.localdensity(st).values).max() < 1e-13
    dc = lm.Currents(lm.DensityCurrents(Hd, st))
    one = lm.Currents(lm.LocalOperatorCurrents(Hd, st, np.eye(2)))
    up = lm.Currents(lm.LocalOperatorCurrents(Hd, st, [[1, 0], [0, 0]]))
    dn = lm.Currents(lm.LocalOperatorCurrents(Hd, st, [[0, 0], [0, 1]]))
    spin = lm.Currents(lm.LocalOperatorCurrents(Hd, st, sz))
    assert abs(one.currents - dc.currents).max() < 1e-13
    assert abs((up + dn).currents - dc.currents).max() < 1e-13
    assert abs((up - dn).currents - spin.currents).max() < 1e-13
    with pytest.raises(lm.ArgumentError):
        lm.LocalOperatorCurrents(Hd, st, np.eye(3))
    with pytest.raises(lm.ArgumentError):
        lm.localexpect(np.eye(2), lm.DeviceState.from_psi(Psi, ctx=ctx, n_int=1), n_int=1)


def test_c_abi_error_behaviour(ctx):
    """Every misuse surfaces as a non-zero status + message (-> ArgumentError), never a crash."""
    lib = _lib.load()
    h = C.c_void_p()
    colptr = np.array([0, 1, 2], np.int64)
    rowval = np.array([0, 5], np.int64)                # row index out of range
    nz = np.ones(2, np.complex128)
    assert lib.lm_ham_create_csc(ctx.handle, 2, 1, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz), 0, C.byref(h)) == 1
    assert b"out of range" in lib.lm_last_error()
    assert lib.lm_ham_create_csc(ctx.handle, 3, 2, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz), 0, C.byref(h)) == 1   # N % n_int
    assert lib.lm_ham_create_csc(ctx.handle, 2, 1, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz), 7, C.byref(h)) == 1   # index_base
    assert lib.lm_ham_create_csc(None, 2, 1, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz), 0, C.byref(h)) == 1        # NULL ctx
    s = C.c_void_p()
    psi = np.ones((4, 2), np.complex128, order="F")
    assert lib.lm_state_create_psi(ctx.handle, 0, 2, _lib.ptr(psi), None, C.byref(s)) == 1
    assert lib.lm_state_create_psi(ctx.handle, 4, 2, None, None, C.byref(s)) == 1
    H = lm.tightbinding_hamiltonian(lm.SquareLattice(2, 2))
    dev = H.device(ctx)
    st = lm.DeviceState.from_psi(psi, ctx=ctx)
    nmv = C.c_int32()
    assert lib.lm_step(dev.handle, st.handle, float("nan"), 1e-12, 0, C.byref(nmv)) == 1
    assert lib.lm_step(dev.handle, st.handle, 0.1, -1.0, 0, C.byref(nmv)) == 1
    assert lib.lm_step(dev.handle, st.handle, 0.1, 1e-12, 99, C.byref(nmv)) == 1
    assert lib.lm_step(dev.handle, st.handle, 0.0, 1e-12, 0, C.byref(nmv)) == 0 and nmv.value == 0      # dt = 0 is a no-op
    assert lib.lm_ham_update_values(dev.handle, _lib.ptr(nz)) == 1                                         # bond-mode ham
    assert b"lm_ham_set_field_params" in lib.lm_last_error()
    kinds = np.array([9], np.int32)
    assert lib.lm_ham_set_fields(dev.handle, 1, _lib.ptr(kinds), _lib.ptr(np.zeros(3))) == 1               # unknown field kind
    assert lib.lm_spmm_state(dev.handle, st.handle, st.handle) == 1                                        # aliasing
    with pytest.raises(lm.ArgumentError):
        lm.Context(device=4096)
    b, e = C.c_int64(), C.c_int64()
    assert lib.lm_shard_range(10, 3, 2, C.byref(b), C.byref(e)) == 1
    assert lib.lm_shard_range(10, 1, 2, C.byref(b), C.byref(e)) == 0 and (b.value, e.value) == (5, 10)
    # the propagator refuses absurd ||H|| dt instead of looping
    with pytest.raises(lm.ArgumentError):
        sol = lm.B200Exp(method="taylor", ctx=ctx)
        sol.update_solver(H, 1e9)
        sol.step(st)


def test_refined_spectral_bounds(ctx):
    """Opt-in Lanczos tightening: the refined interval still encloses the true spectrum (with the
    5 % margin), is tighter than Gershgorin where phases cancel, needs fewer terms and keeps the
    propagator at the exact exponential."""
    lo, l = L.square_lattice(12, 12), lm.SquareLattice(12, 12)
    Ho = OP.qwz(lo, field=F.LandauGauge(0.07))
    E = np.linalg.eigvalsh(Ho.toarray())
    Hd = lm.qwz(l, field=lm.LandauGauge(0.07))
    dev = Hd.device(ctx)
    g_lo, g_hi = dev.spectral_bounds()
    assert g_lo <= E[0] and g_hi >= E[-1]
    Psi = _rand_block(Ho.shape[0], 33, seed=4)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    st = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol.update_solver(Hd, 0.5)
    sol.step(st)
    k_gersh = sol.n_matvec
    r_lo, r_hi = dev.refine_bounds(iters=80, margin=0.05)
    assert r_lo <= E[0] and r_hi >= E[-1]
    assert (r_hi - r_lo) < 0.8 * (g_hi - g_lo)
    st2 = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol.update_solver(Hd, 0.5)
    sol.step(st2)
    assert sol.n_matvec < k_gersh
    want = EV.exact_propagator(Ho, 0.5) @ Psi
    assert _relerr(st.download(), want) < 2e-13 and _relerr(st2.download(), want) < 2e-13
    # a field change re-generates values but keeps the (phase-independent) refined enclosure valid
    Hd2 = lm.qwz(l, field=lm.LandauGauge(0.11))
    st3 = lm.DeviceState.from_psi(Psi, ctx=ctx)
    sol.update_solver(Hd2, 0.5)
    sol.step(st3)
    assert _relerr(st3.download(), EV.exact_propagator(OP.qwz(lo, field=F.LandauGauge(0.11)), 0.5) @ Psi) < 2e-13


def test_timesequence_collects_device_frames(ctx):
    """TimeSequence(f, ev, times) (src/timesequence.jl:41-43) over a device evolution, then
    differentiate: d rho/dt matches the continuity equation frame by frame."""
    l = lm.SquareLattice(6, 6)
    H = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.1))
    P0 = lm.densitymatrix(lm.tightbinding_hamiltonian(l), N=9)
    ev = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), H, P0)
    ts = np.arange(0, 21) * 0.01
    frames = lm.TimeSequence(lambda m: np.concatenate([lm.localdensity(m.state).values,
                                                       np.asarray(lm.Currents(lm.DensityCurrents(m.H, m.state)).currents.sum(axis=1)).ravel()]),
                             ev, ts)
    assert len(frames) == 21 and frames[0.0][:36].sum() == pytest.approx(9.0, abs=1e-12)
    rho = lm.TimeSequence()
    for t, v in frames:
        rho[t] = v[:36]
    drho = rho.differentiate()
    mid = 0.5 * (frames[0.1][36:] + frames[0.11][36:])      # sum_j J_ij at the interval midpoint
    assert np.abs(drho[0.105] - mid).max() < 1e-4 * max(np.abs(mid).max(), 1e-12) + 1e-7


def test_golden_config1_fixture_on_device(ctx):
    """The CUDA path against the COMMITTED golden vectors of BASELINE config 1
    (tests/golden/config1_frames.npz): complex128 localdensity and DensityCurrents within 1e-10
    relative at every stored frame, for the Psi-block and the dense-P state."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_frames.npz"))
    l = lm.SquareLattice(10, 10)
    h = lambda t: lm.tightbinding_hamiltonian(l, field=lm.PointFlux(0.2 * min(t, 10.0) / 10.0, (5.5, 5.5)))
    Psi0, w0 = g["Psi0"], g["w0"]
    P0 = (Psi0 * w0[None, :]) @ Psi0.conj().T
    ts = np.arange(0, 201) * 0.1
    frames = list(g["frames"])
    ev_b = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), h, lm.PsiProjector(Psi0, w0))
    ev_d = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), h, P0)
    for k, (mb, md) in enumerate(zip(ev_b(ts), ev_d(ts))):
        if k not in frames:
            continue
        q = frames.index(k)
        rho_b = lm.localdensity(mb.state).values
        rho_d = lm.localdensity(md.state).values
        I, J, V = lm.DensityCurrents(mb.H, mb.state).pair_values()
        assert [tuple(p) for p in g["pairs"]] == list(zip(I.tolist(), J.tolist()))
        assert _relerr(rho_b, g["rho"][q]) < 1e-10 and _relerr(rho_d, g["rho"][q]) < 1e-10
        scale = max(np.abs(g["J"][q]).max(), 1e-300)
        assert np.abs(V - g["J"][q]).max() < 1e-10 * max(scale, 1e-3)
        assert mb.t == pytest.approx(g["times"][q])


def test_edge_cases_ragged_and_degenerate_inputs(ctx):
    """Edge cases: a 1-D chain (degenerate bounding box for the tile plan), a single site with an
    EMPTY Hamiltonian, a 3-orbital random-block model (generic ELL width, no site blocking), and a
    rank that would own no column."""
    # 1-D chain, wide and narrow blocks
    lo, l = L.square_lattice(40, 1), lm.SquareLattice(40, 1)
    Ho = OP.tightbinding_hamiltonian(lo, field=F.LandauGauge(0.1))
    Hd = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.1))
    assert abs(Hd.device(ctx).to_csc() - Ho).max() < 5e-15
    for M in (1, 33, 40):
        Psi = _rand_block(40, M, seed=M)
        st = lm.DeviceState.from_psi(Psi, ctx=ctx)
        sol = lm.B200Exp(tol=1e-14, ctx=ctx)
        sol.update_solver(Hd, 0.3)
        sol.step(st)
        assert _relerr(st.download(), EV.exact_propagator(Ho, 0.3) @ Psi) < 2e-13
        rho = lm.localdensity(st).values
        assert rho.sum() == pytest.approx(M, abs=1e-11)
        assert len(lm.DensityCurrents(Hd, st).pair_values()[2]) == 39
    # single site: no bonds at all -> empty H, evolution is the identity
    l1 = lm.SquareLattice(1, 1)
    H1 = lm.tightbinding_hamiltonian(l1)
    st = lm.DeviceState.from_psi(np.array([[0.6 + 0.8j]]), ctx=ctx)
    sol = lm.B200Exp(ctx=ctx)
    sol.update_solver(H1, 0.1)
    sol.step(st)
    assert abs(st.download()[0, 0] - (0.6 + 0.8j)) < 1e-15
    assert lm.localdensity(st).values[0] == pytest.approx(1.0)
    assert len(lm.DensityCurrents(H1, st).pair_values()[2]) == 0
    # 3 internal orbitals, dense random hopping blocks + random on-site blocks
    rng = np.random.default_rng(8)
    A = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    B = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    D = rng.standard_normal((3, 3)); D = D + D.T
    l3, lo3 = lm.SquareLattice(5, 4), L.square_lattice(5, 4)
    Hd3 = lm.construct_hamiltonian(l3, 3, (D, 1), (A, lm.BravaisTranslation(axis=1)), (B, lm.BravaisTranslation(axis=2)), field=lm.SymmetricGauge(0.1))
    Ho3 = OP.construct_hamiltonian(lo3, 3, [(D, 1), (A, L.translation(axis=1)), (B, L.translation(axis=2))], field=F.SymmetricGauge(0.1))
    assert abs(Hd3.device(ctx).to_csc() - Ho3).max() < 5e-15
    Psi = _rand_block(60, 35, seed=12)
    w = rng.random(35)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=3)
    sol = lm.B200Exp(tol=1e-14, ctx=ctx)
    sol.update_solver(Hd3, 0.2)
    sol.step(st)
    want = EV.exact_propagator(Ho3, 0.2) @ Psi
    assert _relerr(st.download(), want) < 2e-13
    ost = OB.State(want, w, block=True)
    assert _relerr(lm.localdensity(st).values, OB.localdensity(ost, 3)) < 1e-12
    I, J, V = lm.DensityCurrents(Hd3, st).pair_values()
    Jw = np.array([OB.density_current(Ho3, ost, i, j, 3) for i, j in zip(I, J)])
    assert np.abs(V - Jw).max() < 1e-12 * max(1.0, np.abs(Jw).max())
    op = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    assert np.abs(lm.localexpect(op, st).values - OB.localexpect(op, ost, 3)).max() < 1e-12
    # a rank with no columns is an error, not a silent empty shard
    class FakeCtx:
        nranks, rank, precision, handle = 8, 0, ctx.precision, ctx.handle      # rank 0 of 8 gets [0, 0) of 3 columns
        def shard_range(self, M):
            return lm.shard_range(M, self.rank, self.nranks)
    with pytest.raises(lm.ArgumentError, match="owns no column"):
        lm.DeviceState.from_psi(_rand_block(40, 3), ctx=FakeCtx())
