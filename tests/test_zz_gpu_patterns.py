"""GPU tests of the staging / value-sharing variants of the register-tiled stencil kernel
(csrc/stencil.cuh): shared value loads for Hermitian operators (st_tile_herm) and tensor-map boxes
for interior patches must agree with the general row-copy kernel and with the exact exponential;
non-Hermitian operators are refused by lm_step; asynchronous value updates; the catch-all RC = 2
pattern; golden fixtures of the reduced configs 3 / 4."""
import ctypes as C
from importlib import import_module

import numpy as np
import pytest

import lm_b200 as lm
from oracle import evolution as EV
from oracle import fields as F
from oracle import lattice as L
from oracle import operators as OP

pytestmark = pytest.mark.gpu
_lib = import_module("lm_b200._lib")


def _set_flags(herm, tmap):
    lib = _lib.load()
    lib.lm_dbg_set_stencil_flags.argtypes = [C.c_int32, C.c_int32]
    lib.lm_dbg_set_stencil_flags.restype = C.c_int32
    _lib.check(lib.lm_dbg_set_stencil_flags(herm, tmap))


def _rand_block(n, m, seed):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m)))


def _relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


CASES = {
    "square": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(23, 17), field=lm.LandauGauge(0.07)),
               lambda: OP.tightbinding_hamiltonian(L.square_lattice(23, 17), field=F.LandauGauge(0.07))),
    "qwz_pbc": (lambda: lm.qwz(lm.SquareLattice(14, 15, boundaries=[("axis1", True)]), field=lm.LandauGauge(0.5)),
                lambda: OP.qwz(L.square_lattice(14, 15, periodic=(1,)), field=F.LandauGauge(0.5))),
    "haldane": (lambda: lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1, field=lm.SymmetricGauge(0.03)),
                lambda: OP.haldane(L.honeycomb_lattice(13, 11), 1.0, 0.2, 0.1, field=F.SymmetricGauge(0.03))),
}


@pytest.mark.parametrize("precision", ["c128", "c64"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_stencil_kernel_variants_agree(case, precision):
    """(herm, tmap) in {0,1}^2: general values + row copies (the round-1 kernel), shared value loads,
    tensor-map boxes, both (the default).  Different summation orders: agreement to rounding, and
    every variant against the exact exponential."""
    ctx = lm.default_context(precision)
    mk_dev, mk_or = CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    N = Ho.shape[0]
    tol = 1e-13 if precision == "c128" else 1e-6
    try:
        for M, method, dt in ((200, "chebyshev", 0.3), (333, "taylor", 0.2), (131, "chebyshev", -0.7)):
            X = _rand_block(N, M, seed=M)
            outs = []
            for herm, tmap in ((0, 0), (1, 0), (0, 1), (1, 1)):
                _set_flags(herm, tmap)
                st = lm.DeviceState.from_psi(X, ctx=ctx)
                sol = lm.B200Exp(tol=tol, method=method, ctx=ctx)
                sol.update_solver(Hd, dt)
                for _ in range(3):                            # odd factor counts swap buffers: several steps
                    sol.step(st)
                outs.append(st.download())
            U = EV.exact_propagator(Ho, dt)
            want = U @ (U @ (U @ X))
            for k, o in enumerate(outs):
                assert _relerr(o, want) < (5e-13 if precision == "c128" else 2e-4), (case, M, method, k)
                assert _relerr(o, outs[0]) < (1e-13 if precision == "c128" else 5e-5), (case, M, method, k)
    finally:
        _set_flags(-1, -1)


def test_switching_variants_keeps_graph_cache_consistent():
    """Switching the kernel flags between steps of the SAME state must not replay a stale step graph."""
    ctx = lm.default_context("c128")
    Hd = lm.tightbinding_hamiltonian(lm.SquareLattice(20, 20), field=lm.LandauGauge(0.1))
    Ho = OP.tightbinding_hamiltonian(L.square_lattice(20, 20), field=F.LandauGauge(0.1))
    X = _rand_block(400, 256, seed=3)
    st = lm.DeviceState.from_psi(X, ctx=ctx)
    sol = lm.B200Exp(tol=1e-13, ctx=ctx)
    sol.update_solver(Hd, 0.1)
    want = X
    U = EV.exact_propagator(Ho, 0.1)
    try:
        for k in range(6):
            _set_flags(k % 2, (k // 2) % 2)
            sol.step(st)
            want = U @ want
        assert _relerr(st.download(), want) < 1e-12
    finally:
        _set_flags(-1, -1)


def test_non_hermitian_operator_is_refused_by_step_and_currents():
    """The propagators, the pair currents and the shared value loads assume H = H' (ADVICE r1): a
    non-Hermitian matrix still multiplies (lm_spmm) but lm_step / DensityCurrents raise."""
    import scipy.sparse as sp
    ctx = lm.default_context("c128")
    H = lm.tightbinding_hamiltonian(lm.SquareLattice(8, 8), field=lm.LandauGauge(0.1))
    A = sp.csc_matrix(H.data).astype(np.complex128)
    B = A.copy()
    B.data = B.data.copy()
    B.data[3] *= 1.0 + 1e-6                                   # one entry off its mirror
    dev = lm.DeviceHam.from_csc(ctx, B, 1)
    X = _rand_block(64, 40, seed=5)
    st = lm.DeviceState.from_psi(X, ctx=ctx)
    lib = _lib.load()
    Y = np.empty_like(np.asfortranarray(X))
    _lib.check(lib.lm_spmm(dev.handle, _lib.ptr(np.asfortranarray(X)), _lib.ptr(Y), 64, 40))
    assert _relerr(Y, B @ X) < 1e-13
    nmv = C.c_int32()
    with pytest.raises(_lib.ArgumentError, match="not Hermitian"):
        _lib.check(lib.lm_step(dev.handle, st.handle, 0.1, 1e-12, 0, C.byref(nmv)))
    rho = np.empty(64)
    J = np.empty(max(len(dev.pairs()[0]), 1))
    with pytest.raises(_lib.ArgumentError, match="not Hermitian"):
        _lib.check(lib.lm_observables(dev.handle, st.handle, _lib.ptr(rho), _lib.ptr(J)))
    # back to Hermitian values: accepted again
    _lib.check(lib.lm_ham_update_values(dev.handle, _lib.ptr(np.ascontiguousarray(A.data))))
    _lib.check(lib.lm_step(dev.handle, st.handle, 0.1, 1e-12, 0, C.byref(nmv)))


def test_async_value_updates_match_synchronous_ones_and_trip_on_a_wider_spectrum():
    """lm_ham_update_values_async: same results as the synchronous update while the values stay inside
    the planned enclosure (a gauge-field ramp), a sticky error at the next synchronising call when they do not."""
    import scipy.sparse as sp
    ctx = lm.default_context("c128")
    lat = lm.SquareLattice(16, 12)
    Hs = [sp.csc_matrix(lm.tightbinding_hamiltonian(lat, field=lm.LandauGauge(0.02 * k)).data).astype(np.complex128) for k in range(5)]
    X = _rand_block(192, 64, seed=9)
    lib = _lib.load()
    nmv = C.c_int32()
    outs = []
    for use_async in (False, True):
        dev = lm.DeviceHam.from_csc(ctx, Hs[0], 1, coords=lat.coords, lattice_dims=lat.sizes)
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        keep = []
        for k in range(1, 5):
            nz = np.ascontiguousarray(Hs[k].data)
            keep.append(nz)                                   # the async call borrows the buffer until the next sync
            _lib.check((lib.lm_ham_update_values_async if use_async else lib.lm_ham_update_values)(dev.handle, _lib.ptr(nz)))
            _lib.check(lib.lm_step(dev.handle, st.handle, 0.1, 1e-12, 0, C.byref(nmv)))
        _lib.check(lib.lm_ctx_synchronize(ctx.handle))
        outs.append(st.download())
    assert np.array_equal(outs[0], outs[1])
    want = X
    for k in range(1, 5):
        want = EV.exact_propagator(Hs[k].toarray(), 0.1) @ want
    assert _relerr(outs[1], want) < 1e-12
    # values that leave the enclosure: reported by the next synchronising call, then cleared
    nz = np.ascontiguousarray(3.0 * Hs[1].data)
    _lib.check(lib.lm_ham_update_values_async(dev.handle, _lib.ptr(nz)))
    with pytest.raises(_lib.ArgumentError, match="enclosure"):
        _lib.check(lib.lm_ctx_synchronize(ctx.handle))
    _lib.check(lib.lm_ctx_synchronize(ctx.handle))
    _lib.check(lib.lm_ham_update_values(dev.handle, _lib.ptr(np.ascontiguousarray(Hs[1].data))))


def test_synthetic_block_and_column_norms():
    """lm_state_create_psi_synth against its numpy restatement (tests/synth.py) incl. a column shard,
    and lm_state_column_norms2 against numpy."""
    from synth import synth_block
    ctx = lm.default_context("c128")
    lib = _lib.load()
    N, M = 300, 70
    want = synth_block(N, M, 0, seed=1234)
    for col0, m in ((0, M), (17, 33)):
        h = C.c_void_p()
        _lib.check(lib.lm_state_create_psi_synth(ctx.handle, N, m, col0, 1234, C.byref(h)))
        out = np.empty((N, m), dtype=np.complex128, order="F")
        _lib.check(lib.lm_state_download_psi(h, _lib.ptr(out)))
        assert np.array_equal(out, want[:, col0:col0 + m])
        n2 = np.empty(m)
        _lib.check(lib.lm_state_column_norms2(h, _lib.ptr(n2)))
        assert np.allclose(n2, (np.abs(out) ** 2).sum(axis=0), rtol=1e-13)
        assert np.all(np.abs(n2 - 1.0) < 0.2)
        _lib.check(lib.lm_state_destroy(h))


# ------------------------------------------------------------------------------ catch-all RC = 2 stencil (pattern 5)
def _stencil_id(dev):
    lib = _lib.load()
    i, rc, sw, m = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
    lib.lm_dbg_stencil_info.argtypes = [C.c_void_p] * 5
    _lib.check(lib.lm_dbg_stencil_info(dev.handle, C.byref(i), C.byref(rc), C.byref(sw), C.byref(m)))
    return i.value


@pytest.mark.parametrize("case", ["honeycomb_t123", "honeycomb_t123_torus"])
def test_catch_all_two_row_stencil_matches_oracle(case):
    """A two-rows-per-cell pattern outside the model-specific masks (honeycomb with third-neighbour
    hops) runs on the catch-all RC = 2 stencil kernel (pattern 5, 18 slots per row, one cell per
    thread in the observables kernel): SpMM, propagator, localdensity and DensityCurrents vs the oracle."""
    ctx = lm.default_context("c128")
    if case == "honeycomb_t123":
        Hd = lm.tightbinding_hamiltonian(lm.HoneycombLattice(9, 11), t1=1, t2=0.2, t3=0.1, field=lm.LandauGauge(0.04))
        Ho = OP.tightbinding_hamiltonian(L.honeycomb_lattice(9, 11), t1=1, t2=0.2, t3=0.1, field=F.LandauGauge(0.04))
    else:
        Hd = lm.tightbinding_hamiltonian(lm.HoneycombLattice(8, 7, boundaries=[("axis1", True), ("axis2", True)]), t1=1, t2=0.2, t3=0.1)
        Ho = OP.tightbinding_hamiltonian(L.honeycomb_lattice(8, 7, periodic=(1, 2)), t1=1, t2=0.2, t3=0.1)
    dev = Hd.device(ctx)
    assert _stencil_id(dev) == 5
    lib = _lib.load()
    N = Ho.shape[0]
    for M in (32, 45, 100):
        X = _rand_block(N, M, seed=M)
        x = lm.DeviceState.from_psi(X, ctx=ctx)
        y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
        _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
        assert _relerr(y.download(), Ho @ X) < 1e-14, (case, M)
    X = _rand_block(N, 40, seed=5)
    want = EV.exact_propagator(Ho, 0.3) @ X
    for method in ("taylor", "chebyshev", "chebyshev_clenshaw"):
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        sol = lm.B200Exp(tol=1e-14, method=method, ctx=ctx)
        sol.update_solver(Hd, 0.3)
        sol.step(st)
        assert _relerr(st.download(), want) < 2e-13, (case, method)
    Hdn = Ho.toarray()
    for M in (32, 70):
        Psi = _rand_block(N, M, seed=M) / np.sqrt(N)
        w = np.random.default_rng(M).random(M)
        st = lm.DeviceState.from_psi(Psi, w, ctx=ctx)
        I, J, V = lm.DensityCurrents(Hd, st).pair_values()
        P = (Psi * w) @ Psi.conj().T
        assert _relerr(lm.localdensity(st).values, np.real(np.diag(P))) < 1e-13
        want_j = np.array([2 * np.imag(Hdn[i - 1, j - 1] * P[j - 1, i - 1]) for i, j in zip(I.tolist(), J.tolist())])
        assert np.abs(V - want_j).max() < 1e-13 * max(1.0, np.abs(want_j).max()), (case, M)


# ------------------------------------------------------------------------------ online choice of the schedule
@pytest.mark.parametrize("method", ["auto", "lanczos"])
@pytest.mark.parametrize("name", ["config3s", "config4s"])
def test_golden_reduced_config_fixtures_on_device(name, method):
    """The CUDA path against the COMMITTED golden vectors of the reduced BASELINE configs 3 (QWZ,
    Landau field ramped and regenerated on the device every step) and 4 (Haldane on a honeycomb
    cylinder): complex128 localdensity and DensityCurrents within 1e-10 relative at every stored frame."""
    import os
    ctx = lm.default_context("c128")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + "_frames.npz"))
    if name == "config3s":
        l, n_int = lm.SquareLattice(12, 10), 2
        h = lambda t: lm.qwz(l, field=lm.LandauGauge(0.1 * min(t, 1.0)))
    else:
        l, n_int = lm.HoneycombLattice(9, 8, boundaries=[("axis1", True)]), 1
        h = lambda t: lm.haldane(l, 1.0, 0.2, 0.1, field=lm.LandauGauge(0.03))
    frames = list(g["frames"])
    ev = lm.Evolution(lm.B200Exp(tol=1e-13, method=method, ctx=ctx), h, lm.PsiProjector(g["Psi0"], g["w0"], lattice=l, n_int=n_int))
    seen = 0
    for k, m in enumerate(ev(np.arange(0, 21) * 0.1)):
        if k not in frames:
            continue
        q = frames.index(k)
        I, J, V = lm.DensityCurrents(m.H, m.state).pair_values()
        assert [tuple(p) for p in g["pairs"]] == list(zip(I.tolist(), J.tolist()))
        assert _relerr(lm.localdensity(m.state).values, g["rho"][q]) < 1e-10
        assert np.abs(V - g["J"][q]).max() < 1e-10 * max(np.abs(g["J"][q]).max(), 1e-3)
        assert m.t == pytest.approx(g["times"][q])
        seen += 1
    assert seen == len(frames)


@pytest.mark.parametrize("name", ["config3s", "config4s"])
def test_golden_reduced_config_fixtures_complex64_mode(name):
    """The optional complex64 mode against the same golden vectors at the stated 1e-5 relative."""
    import os
    ctx = lm.default_context("c64")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + "_frames.npz"))
    if name == "config3s":
        l, n_int = lm.SquareLattice(12, 10), 2
        h = lambda t: lm.qwz(l, field=lm.LandauGauge(0.1 * min(t, 1.0)))
    else:
        l, n_int = lm.HoneycombLattice(9, 8, boundaries=[("axis1", True)]), 1
        h = lambda t: lm.haldane(l, 1.0, 0.2, 0.1, field=lm.LandauGauge(0.03))
    frames = list(g["frames"])
    ev = lm.Evolution(lm.B200Exp(tol=1e-7, ctx=ctx), h, lm.PsiProjector(g["Psi0"], g["w0"], lattice=l, n_int=n_int))     # the states live on the solver's context
    for k, m in enumerate(ev(np.arange(0, 21) * 0.1)):
        if k not in frames:
            continue
        q = frames.index(k)
        V = lm.DensityCurrents(m.H, m.state).pair_values()[2]
        assert _relerr(lm.localdensity(m.state).values, g["rho"][q]) < 1e-5
        assert np.abs(V - g["J"][q]).max() < 1e-5 * max(np.abs(g["J"][q]).max(), 1e-3)


def test_currents_and_local_operators_on_a_dense_density_matrix():
    """the reference's README loop and test/test_workflows.jl:29-62 run
    Currents(DensityCurrents(H, P)) on a DENSE P.  Density currents, curr[i, j], Currents(...),
    region sums, localexpect and LocalOperatorCurrents of a dense device state equal those of the
    Psi-block form of the same P."""
    from lm_b200.observables import Currents, LocalOperatorCurrents, currentsfrom, currentsfromto, localexpect
    ctx = lm.default_context("c128")
    lat = lm.SquareLattice(6, 5)
    H = lm.qwz(lat, field=lm.LandauGauge(0.1))
    N = 60
    rng = np.random.default_rng(11)
    X = _rand_block(N, 9, seed=11) / np.sqrt(N)
    w = rng.random(9)
    P = (X * w) @ X.conj().T
    sd = lm.DeviceState.from_dense(P, ctx=ctx, lattice=lat, n_int=2)
    sp_ = lm.DeviceState.from_psi(X, w, ctx=ctx, lattice=lat, n_int=2)
    Id, Jd, Vd = lm.DensityCurrents(H, sd).pair_values()
    Ip, Jp, Vp = lm.DensityCurrents(H, sp_).pair_values()
    assert np.array_equal(Id, Ip) and np.array_equal(Jd, Jp) and np.abs(Vd - Vp).max() < 1e-14
    assert np.abs(Vd).max() > 1e-4
    Hd = H.data.toarray()
    i, j = int(Id[3]), int(Jd[3])
    want = 2 * np.imag(np.sum(Hd[2 * (i - 1):2 * i, 2 * (j - 1):2 * j] * P[2 * (j - 1):2 * j, 2 * (i - 1):2 * i].T))
    assert abs(lm.DensityCurrents(H, sd)[i, j] - want) < 1e-14 and abs(lm.DensityCurrents(H, sd)[j, i] + want) < 1e-14
    assert Currents(lm.DensityCurrents(H, sd)) == Currents(lm.DensityCurrents(H, sp_)) or \
        abs(Currents(lm.DensityCurrents(H, sd)).currents - Currents(lm.DensityCurrents(H, sp_)).currents).max() < 1e-14
    assert abs(currentsfromto(lm.DensityCurrents(H, sd), [1, 2, 3]) - currentsfromto(lm.DensityCurrents(H, sp_), [1, 2, 3])) < 1e-14
    assert np.abs(currentsfrom(lm.DensityCurrents(H, sd), [1, 2, 3]).values - currentsfrom(lm.DensityCurrents(H, sp_), [1, 2, 3]).values).max() < 1e-14
    sz = np.array([[1, 0], [0, -1]], complex)
    assert np.abs(localexpect(sz, sd).values - localexpect(sz, sp_).values).max() < 1e-14
    assert np.abs(LocalOperatorCurrents(H, sd, sz).pair_values()[2] - LocalOperatorCurrents(H, sp_, sz).pair_values()[2]).max() < 1e-14
    assert np.abs(lm.localdensity(sd).values - lm.localdensity(sp_).values).max() < 1e-14


@pytest.mark.parametrize("name", ["config1", "config3s", "config4s"])
def test_reference_generated_golden_series_on_device(name):
    """The CUDA path against golden series written by the REFERENCE itself (julia/make_golden.jl).
    Skipped until tests/golden/ref/ exists (no Julia in the build container); the oracle-generated
    fixtures above cover the same configs meanwhile."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_loader
    if not ref_loader.available(name):
        pytest.skip("tests/golden/ref/%s_* not generated (run julia/make_golden.jl)" % name)
    ctx = lm.default_context("c128")
    g = ref_loader.load(name)
    _, h_dev, n_int = ref_loader.hamiltonians(name)
    lat = h_dev(0.0).lattice
    ev = lm.Evolution(lm.B200Exp(tol=1e-13, ctx=ctx), h_dev, lm.PsiProjector(g["Psi0"], g["w0"], lattice=lat, n_int=n_int))
    want_pairs = [tuple(int(x) for x in p) for p in g["pairs"]]
    for k, m in enumerate(ev(g["times"])):
        I, J, V = lm.DensityCurrents(m.H, m.state).pair_values()
        assert want_pairs == list(zip(I.tolist(), J.tolist()))
        assert _relerr(lm.localdensity(m.state).values, g["rho"][k]) < 1e-10
        assert np.abs(V - g["J"][k]).max() < 1e-10 * max(np.abs(g["J"][k]).max(), 1e-3)


def test_dense_path_large_tiles_3m_kernel():
    """N > 256: U P U' runs on the 64 x 64 double-buffered 3M DMMA kernel (k_zgemm_dmma_3m); ragged
    against the tile size on purpose.  Also the Psi W Psi' escape hatch (conjugated operand)."""
    ctx = lm.default_context("c128")
    Ho = OP.tightbinding_hamiltonian(L.square_lattice(12, 22), field=F.LandauGauge(0.07))
    N = Ho.shape[0]
    rng = np.random.default_rng(3)
    X = _rand_block(N, 40, seed=3) / np.sqrt(N)
    w = rng.random(40)
    P = (X * w) @ X.conj().T
    st = lm.DeviceState.from_dense(P, ctx=ctx, n_int=1)
    sol = lm.B200Exp(tol=1e-14, ctx=ctx, n_int=1)
    U = EV.exact_propagator(Ho, 0.1)
    want = P.copy()
    for _ in range(2):
        sol.update_solver(Ho, 0.1)
        sol.step(st)
        want = U @ want @ U.conj().T
    assert _relerr(st.download(), want) < 1e-12
    stb = lm.DeviceState.from_psi(X, w, ctx=ctx, n_int=1)
    assert _relerr(stb.dense(), P) < 1e-13


@pytest.mark.parametrize("case", ["qwz", "haldane_pbc", "square_field"])
def test_device_eigensolver_lowest_states(case):
    """N1: lm_eigs_lowest - Chebyshev-filtered subspace iteration on the device -
    against the host eigen-decomposition: eigenvalues, residuals, orthonormality, and the density of the
    resulting Fermi sphere (invariant under rotations inside degenerate levels)."""
    ctx = lm.default_context("c128")
    if case == "qwz":
        Hd, Ho, n_int = lm.qwz(lm.SquareLattice(12, 11)), OP.qwz(L.square_lattice(12, 11)), 2
    elif case == "haldane_pbc":
        Hd = lm.haldane(lm.HoneycombLattice(9, 8, boundaries=[("axis1", True)]), 1.0, 0.2, 0.1)
        Ho, n_int = OP.haldane(L.honeycomb_lattice(9, 8, periodic=(1,)), 1.0, 0.2, 0.1), 1
    else:
        Hd = lm.tightbinding_hamiltonian(lm.SquareLattice(17, 13), field=lm.LandauGauge(0.11))
        Ho, n_int = OP.tightbinding_hamiltonian(L.square_lattice(17, 13), field=F.LandauGauge(0.11)), 1
    E, V = np.linalg.eigh(Ho.toarray())
    for nev in (1, 10):
        vals, st = lm.eigs_lowest(Hd, nev, tol=1e-10, ctx=ctx)
        assert np.abs(vals - E[:nev]).max() < 1e-9, (case, nev, vals - E[:nev])
        X = st.download()
        assert X.shape == (Ho.shape[0], nev)
        assert np.abs(X.conj().T @ X - np.eye(nev)).max() < 1e-12
        R = Ho @ X - X * vals[None, :]
        assert np.linalg.norm(R, axis=0).max() < 1e-9 * max(abs(E[0]), abs(E[-1])) * 2
        assert np.all(lm.eigs_lowest.info["residuals"] < 1e-9 * 10)
        if nev == 10 and E[10] - E[9] > 1e-6:                  # closed shell: the projector is unique
            rho = lm.localdensity(st).values
            want = (np.abs(V[:, :10]) ** 2).sum(1).reshape(-1, n_int).sum(1)
            assert _relerr(rho, want) < 1e-7
    e0, g = lm.groundstate_device(Hd, ctx=ctx)
    assert abs(e0 - E[0]) < 1e-9 and g.M == 1


def test_host_hermitian_jacobi_of_the_eigensolver():
    """small_la.h heig_jacobi (the only host arithmetic of lm_eigs_lowest: nb x nb matrices) against numpy."""
    lib = _lib.load()
    lib.lm_dbg_heig.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lm_dbg_heig.restype = C.c_int32
    rng = np.random.default_rng(8)
    for n in (1, 5, 40):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = np.ascontiguousarray(A + A.conj().T)
        w, V = np.zeros(n), np.zeros((n, n), complex)
        _lib.check(lib.lm_dbg_heig(n, _lib.ptr(A), _lib.ptr(w), _lib.ptr(V)))
        assert np.abs(w - np.linalg.eigvalsh(A)).max() < 1e-12 * max(1.0, np.abs(A).max()) * n
        assert np.abs(A @ V - V * w[None, :]).max() < 1e-12 * n * max(1.0, np.abs(A).max())
        assert np.abs(V.conj().T @ V - np.eye(n)).max() < 1e-13 * n


def test_device_eigensolver_full_size_against_closed_form():
    """lm_eigs_lowest at N = 2e5 (beyond any host eigen-decomposition): the open 500 x 400 square lattice has the
    closed-form spectrum E = 2 cos(pi k / 501) + 2 cos(pi l / 401); the 12 lowest levels, their residuals against the
    host-assembled H, and orthonormality."""
    ctx = lm.default_context("c128")
    n1, n2 = 500, 400
    H = lm.tightbinding_hamiltonian(lm.SquareLattice(n1, n2))
    k, l = np.arange(1, n1 + 1), np.arange(1, n2 + 1)
    exact = np.sort((2 * np.cos(np.pi * k / (n1 + 1))[:, None] + 2 * np.cos(np.pi * l / (n2 + 1))[None, :]).ravel())[:12]
    vals, st = lm.eigs_lowest(H, 12, tol=1e-9, ctx=ctx)
    assert np.abs(vals - exact).max() < 1e-8, vals - exact
    X = st.download()
    assert np.abs(X.conj().T @ X - np.eye(12)).max() < 1e-11
    R = H.data @ X - X * vals[None, :]
    assert np.linalg.norm(R, axis=0).max() < 4e-8


# ------------------------------------------------------------------------------ real / imaginary value class
def _set_ri(v):
    lib = _lib.load()
    lib.lm_dbg_set_stencil_ri.argtypes = [C.c_int32]
    lib.lm_dbg_set_stencil_ri.restype = C.c_int32
    _lib.check(lib.lm_dbg_set_stencil_ri(v))


def _ri_state(dev):
    lib = _lib.load()
    lib.lm_dbg_stencil_ri_state.argtypes = [C.c_void_p, C.c_void_p]
    lib.lm_dbg_stencil_ri_state.restype = C.c_int32
    f = C.c_int32(-1)
    _lib.check(lib.lm_dbg_stencil_ri_state(dev.handle, C.byref(f)))
    return f.value


RI_CASES = {
    # (device model, oracle model, in the class of its compiled pattern?)
    "square": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(21, 18)),
               lambda: OP.tightbinding_hamiltonian(L.square_lattice(21, 18)), 1),
    "square_nnn_pbc": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(12, 13, boundaries=[("axis1", True), ("axis2", True)]), t1=1, t2=0.3),
                       lambda: OP.tightbinding_hamiltonian(L.square_lattice(12, 13, periodic=(1, 2)), t1=1, t2=0.3), 1),
    "honeycomb": (lambda: lm.tightbinding_hamiltonian(lm.HoneycombLattice(11, 9)),
                  lambda: OP.tightbinding_hamiltonian(L.honeycomb_lattice(11, 9)), 1),
    "qwz": (lambda: lm.qwz(lm.SquareLattice(14, 15), 1.3), lambda: OP.qwz(L.square_lattice(14, 15), 1.3), 1),
    "qwz_pbc": (lambda: lm.qwz(lm.SquareLattice(14, 15, boundaries=[("axis1", True)])),
                lambda: OP.qwz(L.square_lattice(14, 15, periodic=(1,))), 1),
    "haldane": (lambda: lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1),
                lambda: OP.haldane(L.honeycomb_lattice(13, 11), 1.0, 0.2, 0.1), 1),
    "haldane_torus": (lambda: lm.haldane(lm.HoneycombLattice(9, 10, boundaries=[("axis1", True), ("axis2", True)]), 1.0, 0.2, 0.1),
                      lambda: OP.haldane(L.honeycomb_lattice(9, 10, periodic=(1, 2)), 1.0, 0.2, 0.1), 1),
    "haldane_field": (lambda: lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.05)),
                      lambda: OP.haldane(L.honeycomb_lattice(13, 11), 1.0, 0.2, 0.1, field=F.LandauGauge(0.05)), 0),
    "square_twist": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(12, 13, boundaries=[("axis1", 0.7)])),
                     lambda: OP.tightbinding_hamiltonian(L.square_lattice(12, 13, periodic=(1,), twists={1: 0.7})), 0),
}


@pytest.mark.parametrize("precision", ["c128", "c64"])
@pytest.mark.parametrize("case", sorted(RI_CASES))
def test_real_imaginary_value_class(case, precision):
    """Models without a magnetic field have purely real / purely imaginary entries (real hoppings, `im t2` of
    `haldane`, the `-im/2` orbital flips of `qwz`): the stencil kernel then reads scalar values.  The class is
    detected on the device (lm_dbg_stencil_ri_state); scalar and complex paths agree to rounding and both match
    the exact exponential; Peierls phases and twists stay on the complex path."""
    ctx = lm.default_context(precision)
    mk_dev, mk_or, in_class = RI_CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    N = Ho.shape[0]
    tol = 1e-13 if precision == "c128" else 1e-6
    try:
        outs = []
        for ri in (0, 1):
            _set_ri(ri)
            X = _rand_block(N, 150, seed=11)
            st = lm.DeviceState.from_psi(X, ctx=ctx)
            sol = lm.B200Exp(tol=tol, method="chebyshev", ctx=ctx)
            sol.update_solver(Hd, 0.25)
            assert _stencil_id(sol.dev) >= 0
            assert _ri_state(sol.dev) == in_class, case
            for _ in range(3):
                sol.step(st)
            outs.append(st.download())
        U = EV.exact_propagator(Ho, 0.25)
        want = U @ (U @ (U @ X))
        for o in outs:
            assert _relerr(o, want) < (5e-13 if precision == "c128" else 2e-4), case
        assert _relerr(outs[1], outs[0]) < (1e-13 if precision == "c128" else 5e-5), case
    finally:
        _set_ri(-1)


def test_value_class_follows_value_updates():
    """A field ramp from B = 0: the first operator is real (scalar path), later ones carry Peierls phases
    (complex path), then back - through synchronous AND asynchronous value updates and step-graph replays.
    The class flag lives on the device, so an asynchronous update needs no host round trip."""
    import scipy.sparse as sp
    ctx = lm.default_context("c128")
    lat = lm.HoneycombLattice(12, 10)
    Bs = [0.0, 0.0, 0.03, 0.06, 0.0, 0.0, 0.02]
    Hs = [sp.csc_matrix(lm.haldane(lat, 1.0, 0.2, 0.1, field=lm.LandauGauge(B)).data).astype(np.complex128) for B in Bs]
    X = _rand_block(240, 96, seed=21)
    lib = _lib.load()
    nmv = C.c_int32()
    want = X
    for H in Hs:
        want = EV.exact_propagator(H.toarray(), 0.1) @ want
    for use_async in (False, True):
        dev = lm.DeviceHam.from_csc(ctx, Hs[0], 1, coords=lat.coords, lattice_dims=lat.sizes)
        assert _stencil_id(dev) >= 0 and _ri_state(dev) == 1
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        keep = []
        for k, H in enumerate(Hs):
            nz = np.ascontiguousarray(H.data)
            keep.append(nz)
            _lib.check((lib.lm_ham_update_values_async if use_async else lib.lm_ham_update_values)(dev.handle, _lib.ptr(nz)))
            _lib.check(lib.lm_step(dev.handle, st.handle, 0.1, 1e-13, 0, C.byref(nmv)))
            if not use_async:
                assert _ri_state(dev) == (1 if Bs[k] == 0.0 else 0)
        _lib.check(lib.lm_ctx_synchronize(ctx.handle))
        assert _relerr(st.download(), want) < 1e-12


# ------------------------------------------------------------------------------ three / four rows per unit cell
WIDE_CASES = {
    # name: (device model, oracle model, compiled pattern id, in the real / imaginary class?)
    "kagome": (lambda: lm.tightbinding_hamiltonian(lm.KagomeLattice(9, 11), field=lm.LandauGauge(0.04)),
               lambda: OP.tightbinding_hamiltonian(L.kagome_lattice(9, 11), field=F.LandauGauge(0.04)), 6, 0),
    "kagome_torus": (lambda: lm.tightbinding_hamiltonian(lm.KagomeLattice(8, 7, boundaries=[("axis1", True), ("axis2", True)])),
                     lambda: OP.tightbinding_hamiltonian(L.kagome_lattice(8, 7, periodic=(1, 2))), 6, 1),
    "kagome_t2": (lambda: lm.tightbinding_hamiltonian(lm.KagomeLattice(10, 9), t1=1, t2=0.3, field=lm.SymmetricGauge(0.02)),
                  lambda: OP.tightbinding_hamiltonian(L.kagome_lattice(10, 9), t1=1, t2=0.3, field=F.SymmetricGauge(0.02)), 7, 0),
    "kanemele": (lambda: lm.kanemele(lm.HoneycombLattice(9, 8), 1.0, 0.2),
                 lambda: OP.kanemele(L.honeycomb_lattice(9, 8), 1.0, 0.2), 8, 1),
    "kanemele_cyl_field": (lambda: lm.kanemele(lm.HoneycombLattice(9, 8, boundaries=[("axis1", True)]), 1.0, 0.2, field=lm.LandauGauge(0.03)),
                           lambda: OP.kanemele(L.honeycomb_lattice(9, 8, periodic=(1,)), 1.0, 0.2, field=F.LandauGauge(0.03)), 8, 0),
}


@pytest.mark.parametrize("precision", ["c128", "c64"])
@pytest.mark.parametrize("case", sorted(WIDE_CASES))
def test_three_and_four_row_stencils_match_oracle(case, precision):
    """`KagomeLattice` (three sites per cell, src/zoo/lattices.jl:209) and `kanemele` (spin-1/2 honeycomb, four rows per
    cell, src/zoo/models.jl:188-194) run on the register-tiled stencil kernels (patterns 6 - 8, multi-word masks,
    2 x 2-cell tiles): SpMM, every propagator, localdensity and DensityCurrents against the oracle."""
    ctx = lm.default_context(precision)
    mk_dev, mk_or, pid, in_class = WIDE_CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    dev = Hd.device(ctx)
    assert _stencil_id(dev) == pid
    assert _ri_state(dev) == in_class
    lib = _lib.load()
    N = Ho.shape[0]
    c128 = precision == "c128"
    for M in (32, 45, 100):
        X = _rand_block(N, M, seed=M)
        x = lm.DeviceState.from_psi(X, ctx=ctx)
        y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
        _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
        assert _relerr(y.download(), Ho @ X) < (1e-14 if c128 else 1e-6), (case, M)
    X = _rand_block(N, 40, seed=5)
    want = EV.exact_propagator(Ho, 0.3) @ X
    for method in ("taylor", "chebyshev", "chebyshev_clenshaw", "taylor_horner"):
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        sol = lm.B200Exp(tol=1e-14 if c128 else 1e-6, method=method, ctx=ctx)
        sol.update_solver(Hd, 0.3)
        sol.step(st)
        assert _relerr(st.download(), want) < (2e-13 if c128 else 2e-5), (case, method)
    Hdn = Ho.toarray()
    for M in (32, 70):
        Psi = _rand_block(N, M, seed=M) / np.sqrt(N)
        w = np.random.default_rng(M).random(M)
        st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, lattice=Hd.lattice, n_int=Hd.n_int)
        I, J, V = lm.DensityCurrents(Hd, st).pair_values()
        P = (Psi * w) @ Psi.conj().T
        n = Hd.n_int
        rho = np.real(np.diag(P)).reshape(-1, n).sum(axis=1)
        assert _relerr(lm.localdensity(st).values, rho) < (1e-13 if c128 else 1e-5)
        want_j = np.array([sum(2 * np.imag(Hdn[(i - 1) * n + a, (j - 1) * n + b] * P[(j - 1) * n + b, (i - 1) * n + a])
                               for a in range(n) for b in range(n)) for i, j in zip(I.tolist(), J.tolist())])
        assert np.abs(V - want_j).max() < (1e-13 if c128 else 1e-5) * max(1.0, np.abs(want_j).max()), (case, M)


# ------------------------------------------------------------------------------ run-time specialised patterns (NVRTC)
def _set_rtc(v):
    lib = _lib.load()
    lib.lm_dbg_set_stencil_rtc.argtypes = [C.c_int32]
    lib.lm_dbg_set_stencil_rtc.restype = C.c_int32
    _lib.check(lib.lm_dbg_set_stencil_rtc(v))


def _kanemele_rashba(mod_lm):
    """Kane-Mele plus a spin-mixing nearest-neighbour term: four rows per cell, not among the compiled patterns."""
    sz = np.array([[1, 0], [0, -1]], complex)
    sx = np.array([[0, 1], [1, 0]], complex)
    if mod_lm:
        lat = lm.HoneycombLattice(9, 8, boundaries=[("axis2", True)])
        return lm.construct_hamiltonian(lat, 2, (1.0, lm.NearestNeighbor(1)), (0.2j * sz, lm.honeycomb_2nn), (0.3j * sx, lm.NearestNeighbor(1)),
                                        field=lm.LandauGauge(0.02))
    lat = L.honeycomb_lattice(9, 8, periodic=(2,))
    return OP.construct_hamiltonian(lat, 2, [(1.0, L.nearest_neighbor(lat, 1)), (0.2j * sz, L.HONEYCOMB_2NN), (0.3j * sx, L.nearest_neighbor(lat, 1))],
                                    F.LandauGauge(0.02))


RTC_CASES = {
    # name: (device model, oracle model, force an exact-mask specialisation although a compiled superset exists?)
    "kanemele_rashba": (lambda: _kanemele_rashba(True), lambda: _kanemele_rashba(False), False),
    "kagome_t123": (lambda: lm.tightbinding_hamiltonian(lm.KagomeLattice(8, 9), t1=1, t2=0.3, t3=0.1),
                    lambda: OP.tightbinding_hamiltonian(L.kagome_lattice(8, 9), t1=1, t2=0.3, t3=0.1), False),
    "triangular_exact": (lambda: lm.tightbinding_hamiltonian(lm.TriangularLattice(13, 11), field=lm.SymmetricGauge(0.03)),
                         lambda: OP.tightbinding_hamiltonian(L.triangular_lattice(13, 11), field=F.SymmetricGauge(0.03)), True),
    "honeycomb_t123_exact": (lambda: lm.tightbinding_hamiltonian(lm.HoneycombLattice(9, 11), t1=1, t2=0.2, t3=0.1),
                             lambda: OP.tightbinding_hamiltonian(L.honeycomb_lattice(9, 11), t1=1, t2=0.2, t3=0.1), True),
}


@pytest.mark.parametrize("case", sorted(RTC_CASES))
def test_run_time_specialised_stencil_patterns(case):
    """A lattice pattern that is not among the compiled ones (or, with LM_STENCIL_RTC=2, one whose compiled superset
    carries empty slots) gets stencil kernels specialised on its mask at run time (csrc/stencil_rtc.cu: NVRTC on the
    library's own headers, driver-API launch): SpMM, propagators, localdensity and DensityCurrents against the oracle."""
    ctx = lm.default_context("c128")
    mk_dev, mk_or, force = RTC_CASES[case]
    try:
        _set_rtc(2 if force else 1)
        Hd, Ho = mk_dev(), mk_or()
        assert abs(Hd.data - Ho).max() < 1e-14
        dev = Hd.device(ctx)
        assert _stencil_id(dev) >= 1000, "pattern was not specialised at run time (NVRTC unavailable?)"
        lib = _lib.load()
        N = Ho.shape[0]
        for M in (32, 45, 100):
            X = _rand_block(N, M, seed=M)
            x = lm.DeviceState.from_psi(X, ctx=ctx)
            y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
            _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
            assert _relerr(y.download(), Ho @ X) < 1e-14, (case, M)
        X = _rand_block(N, 40, seed=5)
        want = EV.exact_propagator(Ho, 0.3) @ X
        for method in ("chebyshev", "taylor", "chebyshev_clenshaw", "taylor_horner"):
            st = lm.DeviceState.from_psi(X, ctx=ctx)
            sol = lm.B200Exp(tol=1e-14, method=method, ctx=ctx)
            sol.update_solver(Hd, 0.3)
            sol.step(st)
            sol.step(st)
            assert _relerr(st.download(), EV.exact_propagator(Ho, 0.3) @ want) < 3e-13, (case, method)
        Hdn = Ho.toarray()
        n = Hd.n_int
        for M in (32, 70):
            Psi = _rand_block(N, M, seed=M) / np.sqrt(N)
            w = np.random.default_rng(M).random(M)
            st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, lattice=Hd.lattice, n_int=n)
            I, J, V = lm.DensityCurrents(Hd, st).pair_values()
            P = (Psi * w) @ Psi.conj().T
            assert _relerr(lm.localdensity(st).values, np.real(np.diag(P)).reshape(-1, n).sum(axis=1)) < 1e-13
            want_j = np.array([sum(2 * np.imag(Hdn[(i - 1) * n + a, (j - 1) * n + b] * P[(j - 1) * n + b, (i - 1) * n + a])
                                   for a in range(n) for b in range(n)) for i, j in zip(I.tolist(), J.tolist())])
            assert np.abs(V - want_j).max() < 1e-13 * max(1.0, np.abs(want_j).max()), (case, M)
    finally:
        _set_rtc(-1)


def test_spin_currents_on_kanemele():
    """`LocalOperatorCurrents(H, P, sigma_z)` and `localexpect(sigma_z, P)` on the Kane-Mele model (src/zoo/models.jl:188-194; the
    reference's spin-current use case, src/zoo/currents.jl:150-184): four rows per cell on the stencil kernels, correlators by the
    block kernel (one warp per site / bond block), against the oracle; up + down = density currents."""
    ctx = lm.default_context("c128")
    l, lo = lm.HoneycombLattice(7, 6), L.honeycomb_lattice(7, 6)
    Hd = lm.kanemele(l, 1.0, 0.2, field=lm.LandauGauge(0.04))
    Ho = OP.kanemele(lo, 1.0, 0.2, field=F.LandauGauge(0.04))
    from oracle import observables as OB
    N = Ho.shape[0]
    Psi = _rand_block(N, 41, seed=13) / np.sqrt(N)
    w = np.random.default_rng(4).random(41)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, lattice=l, n_int=2)
    ost = OB.State(Psi, w, block=True)
    sz = np.array([[1, 0], [0, -1]], complex)
    sx = np.array([[0, 1], [1, 0]], complex)
    for op in (sz, sx, sz + 0.5j * sx @ sz):
        assert np.abs(lm.localexpect(op, st).values - OB.localexpect(op, ost, 2)).max() < 1e-13
        I, J, V = lm.LocalOperatorCurrents(Hd, st, op).pair_values()
        want = np.array([OB.operator_current(Ho, ost, op, i, j, 2) for i, j in zip(I, J)])
        assert np.abs(V - want).max() < 1e-13
    dc = lm.Currents(lm.DensityCurrents(Hd, st))
    up = lm.Currents(lm.LocalOperatorCurrents(Hd, st, [[1, 0], [0, 0]]))
    dn = lm.Currents(lm.LocalOperatorCurrents(Hd, st, [[0, 0], [0, 1]]))
    assert abs((up + dn).currents - dc.currents).max() < 1e-13
