#!/usr/bin/env python
"""Small driver for ncu captures of the non-headline kernels: fused observables (Haldane),
site-blocked SpMM (QWZ), dense U P U' on the FP64 tensor cores, device Peierls phase regeneration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import lm_b200 as lm  # noqa: E402

ctx = lm.Context()
rng = np.random.default_rng(0)
which = sys.argv[1]
if which == "obs":
    H = lm.haldane(lm.HoneycombLattice(300, 300), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.001))
    N = H.structure.dim
    psi = np.asfortranarray((rng.standard_normal((N, 512)) + 1j * rng.standard_normal((N, 512))) / 30)
    st = lm.DeviceState.from_psi(psi, ctx=ctx)
    for _ in range(3):
        lm.localdensity(st)
        lm.DensityCurrents(H, st).pair_values()
elif which == "sites":
    H = lm.qwz(lm.SquareLattice(300, 300), field=lm.LandauGauge(0.01))
    N = H.structure.dim
    psi = np.asfortranarray((rng.standard_normal((N, 1024)) + 1j * rng.standard_normal((N, 1024))) / 30)
    st = lm.DeviceState.from_psi(psi, ctx=ctx)
    sol = lm.B200Exp(ctx=ctx)
    for k in range(2):
        sol.update_solver(lm.qwz(lm.SquareLattice(300, 300), field=lm.LandauGauge(0.01 + 0.001 * k)), 0.1)
        sol.step(st)
elif which == "dense":
    l = lm.SquareLattice(32, 32)
    H = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.05))
    P0 = lm.densitymatrix(lm.tightbinding_hamiltonian(l), mu=0.0).dense()
    st = lm.DeviceState.from_dense(P0, ctx=ctx)
    sol = lm.B200Exp(ctx=ctx)
    for _ in range(3):
        sol.update_solver(H, 0.1)
        sol.step(st)
    print("tr P =", lm.localdensity(st).values.sum())
elif which == "stencil":
    # one plain SpMM launch per register-tile variant of the stencil kernel (Haldane 500 x 500)
    import ctypes as C
    from importlib import import_module
    _lib = import_module("lm_b200._lib")
    lib = _lib.load()
    variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    kind = sys.argv[4] if len(sys.argv) > 4 else "haldane"
    H = lm.haldane(lm.HoneycombLattice(500, 500), 1.0, 0.2, 0.1) if kind == "haldane" else lm.qwz(lm.SquareLattice(300, 300), field=lm.LandauGauge(0.01))
    dev = H.device(ctx)
    blk = (rng.standard_normal((dev.N, 32)) + 1j * rng.standard_normal((dev.N, 32))) / 30
    x = lm.DeviceState.from_psi(np.asfortranarray(np.tile(blk, (1, M // 32))), ctx=ctx, shard=False)
    y = x.copy()
    lib.lm_dbg_set_apply_path.argtypes = [C.c_int32]
    lib.lm_dbg_set_stencil_variant.argtypes = [C.c_int32]
    for v in variants:
        lib.lm_dbg_set_apply_path(5 if v >= 0 else 2)
        lib.lm_dbg_set_stencil_variant(v)
        _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle))
    ctx.synchronize()
elif which == "stencil_obs":
    # fused observables on the stencil view (Haldane 500 x 500)
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    H = lm.haldane(lm.HoneycombLattice(500, 500), 1.0, 0.2, 0.1)
    blk = (rng.standard_normal((H.structure.dim, 32)) + 1j * rng.standard_normal((H.structure.dim, 32))) / 30
    st = lm.DeviceState.from_psi(np.asfortranarray(np.tile(blk, (1, M // 32))), ctx=ctx, shard=False)
    for _ in range(2):
        lm.DensityCurrents(H, st).pair_values()
elif which == "stencil_step":
    # two propagation steps (K product-form factors each, MODE 3) on the default kernel path
    kind = sys.argv[2] if len(sys.argv) > 2 else "haldane"
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    H = lm.haldane(lm.HoneycombLattice(500, 500), 1.0, 0.2, 0.1) if kind == "haldane" else lm.qwz(lm.SquareLattice(300, 300), field=lm.LandauGauge(0.01))
    blk = (rng.standard_normal((H.structure.dim, 32)) + 1j * rng.standard_normal((H.structure.dim, 32))) / 30
    st = lm.DeviceState.from_psi(np.asfortranarray(np.tile(blk, (1, M // 32))), ctx=ctx, shard=False)
    sol = lm.B200Exp(ctx=ctx)
    for _ in range(2):
        sol.update_solver(H, 0.1)
        sol.step(st)
ctx.synchronize()
