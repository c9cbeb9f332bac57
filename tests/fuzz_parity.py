#!/usr/bin/env python
"""Randomised parity sweep of the library against the oracle, on the CPU build of the library
(tests/cpu_emul/build_emul_lib.py) or on a real device.

    LM_EMUL_LIB=<liblm_b200_emul.so> python tests/fuzz_parity.py [first_seed] [n_cases]     # no GPU
    python tests/fuzz_parity.py 0 200                                                         # on a B200

Every case draws a lattice (square / honeycomb / kagome, 3..18 cells per axis, open / periodic / twisted
boundaries), a model (tight binding with t1 / t2 / t3, QWZ, Haldane, Kane-Mele), a field (Landau, symmetric,
axial and singular point fluxes, sums), a block width 1..150, a precision, a propagator method, a
step and the schedule (plain or L2-resident strips) and checks, through the C ABI: the device
assembled H, H X, a few evolution steps against the exact exponential, localdensity and
DensityCurrents against the dense formulas, and (small N) the dense-P path U P U'.  Test infrastructure (imports oracle/).

Tolerances: the axial point flux is ill-conditioned by construction when the flux point is nearly
collinear with a bond (acos(c / (1 + 1e-11)) near c = 1, src/zoo/magneticfields.jl:80-84: one ulp of c
moves the phase by ~2e-11 * flux / 2 pi), so assembled values are compared at 5e-12 for that field
and at 5e-14 otherwise."""
import os, sys, time, ctypes as C, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
if os.environ.get('LM_EMUL_LIB'):
    import conftest
    assert conftest._emulated_library()
import lm_b200 as lm
from importlib import import_module
_lib = import_module("lm_b200._lib")
from oracle import evolution as EV, fields as F, lattice as L, operators as OP
warnings.simplefilter("ignore")
lib = _lib.load()
lib.lm_dbg_set_stencil_flags.argtypes = [C.c_int32, C.c_int32]
lib.lm_dbg_set_stencil_ri.argtypes = [C.c_int32]
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctxs = {"c128": lm.default_context("c128"), "c64": lm.default_context("c64")}
def relerr(a, b): return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
nfail = 0
t00 = time.time()
for case in range(seed0, seed0 + ncases):
    rng = np.random.default_rng(case)
    n1, n2 = int(rng.integers(3, 19)), int(rng.integers(3, 19))
    per = [bool(rng.integers(0, 2)), bool(rng.integers(0, 2))]
    tw = float(rng.uniform(0.1, 2.0)) if rng.integers(0, 3) == 0 else None
    model = ["tb1", "tb12", "tb123", "qwz", "hc1", "haldane", "hc123", "kagome1", "kagome12", "kanemele"][int(rng.integers(0, 10))]
    fk = int(rng.integers(0, 5))
    B = float(rng.uniform(-0.2, 0.2))
    px, py = float(rng.uniform(1, n1)), float(rng.uniform(1, n2))
    mkf = [(lambda m: m.NoField()), (lambda m: m.LandauGauge(B)), (lambda m: m.SymmetricGauge(B)),
           (lambda m: m.PointFlux(B, (px, py))), (lambda m: m.LandauGauge(B) + m.PointFlux(0.3 * B, (px, py), "singular") if m is F else m.LandauGauge(B) + m.PointFlux(0.3 * B, (px, py), gauge="singular"))][fk]
    bl, bo_per, bo_tw = [], [], {}
    for ax in (1, 2):
        if per[ax - 1]:
            if tw is not None and ax == 2:
                bl.append(("axis2", tw)); bo_tw[2] = tw
            else:
                bl.append(("axis%d" % ax, True)); bo_per.append(ax)
    honey = model in ("hc1", "haldane", "hc123", "kanemele")
    kag = model in ("kagome1", "kagome12")
    latd = (lm.KagomeLattice if kag else lm.HoneycombLattice if honey else lm.SquareLattice)(n1, n2, boundaries=bl)
    lato = (L.kagome_lattice if kag else L.honeycomb_lattice if honey else L.square_lattice)(n1, n2, periodic=tuple(bo_per), twists=bo_tw or None)
    try:
        if model == "tb1": Hd, Ho = lm.tightbinding_hamiltonian(latd, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, field=mkf(F))
        elif model == "tb12": Hd, Ho = lm.tightbinding_hamiltonian(latd, t1=1, t2=0.3, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, t1=1, t2=0.3, field=mkf(F))
        elif model == "tb123": Hd, Ho = lm.tightbinding_hamiltonian(latd, t1=1, t2=0.3, t3=0.1, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, t1=1, t2=0.3, t3=0.1, field=mkf(F))
        elif model == "qwz": Hd, Ho = lm.qwz(latd, field=mkf(lm)), OP.qwz(lato, field=mkf(F))
        elif model == "hc1": Hd, Ho = lm.tightbinding_hamiltonian(latd, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, field=mkf(F))
        elif model == "kagome1": Hd, Ho = lm.tightbinding_hamiltonian(latd, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, field=mkf(F))
        elif model == "kagome12": Hd, Ho = lm.tightbinding_hamiltonian(latd, t1=1, t2=0.3, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, t1=1, t2=0.3, field=mkf(F))
        elif model == "kanemele": Hd, Ho = lm.kanemele(latd, 1.0, 0.2, field=mkf(lm)), OP.kanemele(lato, 1.0, 0.2, field=mkf(F))
        elif model == "haldane": Hd, Ho = lm.haldane(latd, 1.0, 0.2, 0.1, field=mkf(lm)), OP.haldane(lato, 1.0, 0.2, 0.1, field=mkf(F))
        else: Hd, Ho = lm.tightbinding_hamiltonian(latd, t1=1, t2=0.2, t3=0.1, field=mkf(lm)), OP.tightbinding_hamiltonian(lato, t1=1, t2=0.2, t3=0.1, field=mkf(F))
    except Exception as e:
        print("case %d: construction raised %r" % (case, e)); continue
    prec = "c64" if rng.integers(0, 4) == 0 else "c128"
    ctx = ctxs[prec]
    eps = 1.0 if prec == "c128" else 1e8
    N = Ho.shape[0]
    M = int(rng.choice([1, 2, 3, 7, 16, 31, 32, 33, 40, 64, 65, 100, 131, 150]))
    desc = "case %d: %s %dx%d per=%s tw=%s field=%d %s M=%d" % (case, model, n1, n2, per, tw, fk, prec, M)
    try:
        dev = Hd.device(ctx)
        got = dev.to_csc()
        e0 = abs(got - Ho).max()
        assert e0 < (5e-12 if fk == 3 else 5e-14) * (1 if prec == "c128" else 1e8), ("assembly", e0)
        if fk == 3 and prec == "c128":
            Ho = got.astype(np.complex128)          # kernels are checked against the values the device assembled (see the tolerance note)
        X = (rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M)))
        x = lm.DeviceState.from_psi(X, ctx=ctx); y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
        _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
        e1 = relerr(y.download(), Ho @ X)
        assert e1 < 3e-14 * eps, ("spmm", e1)
        method = ["auto", "taylor", "chebyshev", "taylor_horner", "chebyshev_clenshaw", "lanczos"][int(rng.integers(0, 6))]
        dt = float(rng.choice([0.1, 0.37, -0.5, 1.3]))
        kb = int(rng.integers(0, 8))                 # stencil kernel variant: bit 0 shared value loads (Hermitian), bit 1 tensor-map boxes, bit 2 no real / imaginary class scalars
        lib.lm_dbg_set_stencil_flags(kb & 1, (kb >> 1) & 1)
        lib.lm_dbg_set_stencil_ri(0 if kb & 4 else -1)
        st = lm.DeviceState.from_psi(X, ctx=ctx)
        sol = lm.B200Exp(tol=1e-13 if prec == "c128" else 1e-6, method=method, ctx=ctx)
        sol.update_solver(Hd, dt)
        nst = int(rng.integers(1, 4))
        for _ in range(nst): sol.step(st)
        U = EV.exact_propagator(Ho, dt); want = X
        for _ in range(nst): want = U @ want
        e2 = relerr(st.download(), want)
        assert e2 < 1e-12 * eps * 3, ("step", method, dt, kb, e2)
        lib.lm_dbg_set_stencil_flags(-1, -1)
        lib.lm_dbg_set_stencil_ri(-1)
        w = rng.random(M)
        Xn = X / np.sqrt(N)
        so = lm.DeviceState.from_psi(Xn, w, ctx=ctx, n_int=Hd.n_int)
        I, J, V = lm.DensityCurrents(Hd, so).pair_values()
        rho = lm.localdensity(so).values
        P = (Xn * w) @ Xn.conj().T
        n = Hd.n_int
        e3 = relerr(rho, np.real(np.diag(P)).reshape(-1, n).sum(1))
        Hdn = Ho.toarray()
        wantj = np.array([2 * np.imag(np.sum(Hdn[(i - 1) * n:i * n, (j - 1) * n:j * n] * P[(j - 1) * n:j * n, (i - 1) * n:i * n].T)) for i, j in zip(I.tolist(), J.tolist())])
        e4 = np.abs(V - wantj).max() / max(1.0, np.abs(wantj).max()) if len(V) else 0.0
        assert e3 < 1e-13 * eps and e4 < 1e-13 * eps, ("observables", e3, e4)
        if N <= 160 and prec == "c128" and case % 3 == 0:
            # dense density matrix: P <- U P U' (FP64 DMMA path, CachedExp semantics) and the Psi W Psi' escape hatch
            Pd = (Xn * w) @ Xn.conj().T
            sd = lm.DeviceState.from_dense(Pd, ctx=ctx, n_int=n)
            sol2 = lm.B200Exp(tol=1e-13, ctx=ctx, n_int=n)
            sol2.update_solver(Hd, dt)
            sol2.step(sd); sol2.step(sd)
            wantd = U @ (U @ Pd @ U.conj().T) @ U.conj().T
            e5 = relerr(sd.download(), wantd)
            e6 = relerr(so.dense(), Pd)
            e7 = relerr(lm.localdensity(sd).values, np.real(np.diag(wantd)).reshape(-1, n).sum(1))
            assert e5 < 1e-12 and e6 < 1e-13 and e7 < 1e-12, ("dense", e5, e6, e7)
    except Exception as e:
        nfail += 1
        lib.lm_dbg_set_stencil_flags(-1, -1)
        lib.lm_dbg_set_stencil_ri(-1)
        print("FAIL", desc, "->", repr(e)[:300], flush=True)
print("fuzz: %d cases, %d failures, %.0fs" % (ncases, nfail, time.time() - t00))
sys.exit(1 if nfail else 0)
