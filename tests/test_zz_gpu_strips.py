"""GPU tests of the L2-resident strip schedule of the product-form propagators (csrc/api.cu
run_factors; opt-in through LM_STEP_L2_MB / lm_dbg_set_step_l2_kb).  The schedule only re-orders
independent column strips, so its results must equal the plain factor-by-factor schedule BIT FOR
BIT, and both must match the exact exponential.  Kept in its own file, sorted after the parity
suite: written in a session without GPU time, first run is the driver's."""
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


def _set_l2_kb(kb):
    lib = _lib.load()
    lib.lm_dbg_set_step_l2_kb.argtypes = [C.c_int64]
    lib.lm_dbg_set_step_l2_kb.restype = C.c_int32
    _lib.check(lib.lm_dbg_set_step_l2_kb(kb))


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
def test_strip_schedule_is_bitwise_equal_to_plain_schedule(case, precision):
    ctx = lm.default_context(precision)
    mk_dev, mk_or = CASES[case]
    Hd, Ho = mk_dev(), mk_or()
    N = Ho.shape[0]
    tol = 1e-13 if precision == "c128" else 1e-6
    try:
        for M, method, dt in ((200, "chebyshev", 0.3), (333, "taylor", 0.2), (131, "chebyshev", -0.7)):
            X = _rand_block(N, M, seed=M)
            outs, launches = [], []
            for kb in (0, 2 * N * 16 * 64 // 1024 + 1):      # plain, then strips of 64 columns
                _set_l2_kb(kb)
                st = lm.DeviceState.from_psi(X, ctx=ctx)
                sol = lm.B200Exp(tol=tol, method=method, ctx=ctx)
                sol.update_solver(Hd, dt)
                n0 = ctx.launch_count()
                for _ in range(3):                            # odd factor counts swap buffers: several steps
                    sol.step(st)
                launches.append(ctx.launch_count() - n0)
                outs.append(st.download())
            assert launches[1] > launches[0], (case, M, launches)      # the strip schedule really ran
            assert np.array_equal(outs[0], outs[1]), (case, M, method, _relerr(outs[1], outs[0]))
            U = EV.exact_propagator(Ho, dt)
            want = U @ (U @ (U @ X))
            assert _relerr(outs[1], want) < (5e-13 if precision == "c128" else 2e-4), (case, M, method)
    finally:
        _set_l2_kb(-1)


def test_strip_schedule_keeps_graph_cache_consistent():
    """Switching the schedule between steps of the SAME state must not replay a stale step graph."""
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
            _set_l2_kb(0 if k % 2 == 0 else 2 * 400 * 16 * 64 // 1024 + 1)
            sol.step(st)
            want = U @ want
        assert _relerr(st.download(), want) < 1e-12
    finally:
        _set_l2_kb(-1)


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
def test_online_schedule_choice_samples_every_candidate_and_settles():
    """LM_STEP_L2_MB=auto (here: lm_dbg_set_step_l2_kb(-2)): the first steps of a (Hamiltonian, dt,
    tol, method) run one candidate schedule each - plain, strips of two budgets - timed with events;
    afterwards the fastest is used.  All candidates are bit-identical, so the sampled steps are
    ordinary steps: the trajectory must equal the plain one bit for bit."""
    ctx = lm.default_context("c128")
    lib = _lib.load()
    lib.lm_dbg_set_autotune_kb.argtypes = [C.c_int64, C.c_int64]
    lib.lm_dbg_step_schedule.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    Hd = lm.haldane(lm.HoneycombLattice(9, 8), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.05))
    Ho = OP.haldane(L.honeycomb_lattice(9, 8), 1.0, 0.2, 0.1, field=F.LandauGauge(0.05))
    N = Ho.shape[0]
    X = _rand_block(N, 300, seed=11)
    kb64, kb128 = 2 * N * 16 * 64 // 1024 + 1, 2 * N * 16 * 128 // 1024 + 1
    try:
        _set_l2_kb(0)
        ref = lm.DeviceState.from_psi(X, ctx=ctx)
        sol0 = lm.B200Exp(tol=1e-13, ctx=ctx)
        sol0.update_solver(Hd, 0.2)
        for _ in range(6):
            sol0.step(ref)
        _lib.check(lib.lm_dbg_set_autotune_kb(kb64, kb128))
        _set_l2_kb(-2)
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        sol = lm.B200Exp(tol=1e-13, ctx=ctx)
        sol.update_solver(Hd, 0.2)
        kb, cal = C.c_int64(), C.c_int32()
        per_step = []
        for k in range(6):
            n0 = ctx.launch_count()
            sol.step(st)
            per_step.append(ctx.launch_count() - n0)
            _lib.check(lib.lm_dbg_step_schedule(st.handle, C.byref(kb), C.byref(cal)))
            assert cal.value == (1 if k < 2 else 0), (k, cal.value)
        # step 0 plain (K launches), step 1 strips of 64 columns (5 strips), step 2 strips of 128 (3 strips)
        assert per_step[1] == 5 * per_step[0] and per_step[2] == 3 * per_step[0], per_step
        assert kb.value in (0, 64, 128)          # the strip width the choice settled on (0 = plain)
        assert len(set(per_step[3:])) == 1 and per_step[3] in per_step[:3]
        assert np.array_equal(st.download(), ref.download())
        U = EV.exact_propagator(Ho, 0.2)
        want = X
        for _ in range(6):
            want = U @ want
        assert _relerr(st.download(), want) < 1e-12
        # a new dt is a new calibration; a block too narrow for strips settles at once on the plain schedule
        sol.update_solver(Hd, 0.1)
        sol.step(st)
        _lib.check(lib.lm_dbg_step_schedule(st.handle, C.byref(kb), C.byref(cal)))
        assert cal.value == 1
        narrow = lm.DeviceState.from_psi(X[:, :40].copy(), ctx=ctx)
        sol.step(narrow)
        _lib.check(lib.lm_dbg_step_schedule(narrow.handle, C.byref(kb), C.byref(cal)))
        assert cal.value == 0 and kb.value == 0
    finally:
        _set_l2_kb(-1)
        lib.lm_dbg_set_autotune_kb(0, 0)


def test_strip_candidates_fill_the_machine_evenly():
    """The widths the online choice samples: multiples of the kernel's 32-column chunk whose two strip
    buffers stay under 60 % of the L2, ranked by wave efficiency (patches x chunks against the CTAs
    resident on all SMs).  C2 (square 100 x 100, 5000 columns): 13 x 13 patches of 8 x 8 cells, 4
    CTAs per SM -> 592 resident; 7 chunks = 1183 CTAs fill two waves (1184) almost exactly, 6 chunks =
    1014 CTAs are next (0.86).  C4-sized lattices have no candidate (the plain schedule)."""
    ctx = lm.default_context("c128")
    lib = _lib.load()
    lib.lm_dbg_strip_candidates.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    out, n = (C.c_int64 * 2)(), C.c_int32()
    dev = lm.tightbinding_hamiltonian(lm.SquareLattice(100, 100)).device(ctx)
    _lib.check(lib.lm_dbg_strip_candidates(dev.handle, 5000, out, C.byref(n)))
    assert n.value == 2 and list(out) == [224, 192], (n.value, list(out))
    for ms in out:
        assert ms % 32 == 0 and 2 * 10**4 * ms * 16 <= 0.6 * 126 * 2**20
    _lib.check(lib.lm_dbg_strip_candidates(dev.handle, 96, out, C.byref(n)))      # 31 MB of buffers: L2-resident as it is
    assert n.value == 0
    big = lm.haldane(lm.HoneycombLattice(200, 200), 1.0, 0.2, 0.1).device(ctx)     # N = 8e4: a 64-column strip pair is 164 MB
    _lib.check(lib.lm_dbg_strip_candidates(big.handle, 4096, out, C.byref(n)))
    assert n.value == 0


# ------------------------------------------------------------------------------ golden fixtures of reduced configs 3 / 4
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
