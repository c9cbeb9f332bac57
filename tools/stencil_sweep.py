#!/usr/bin/env python
"""Register-tiled stencil kernel (csrc/stencil.cuh): parity against the host-assembled matrix on
small ragged lattices, then timing of every compiled register-tile variant against the ELL gather
kernels at the benchmark sizes.  One JSON line per measurement.

    python tools/stencil_sweep.py [--M 1024] [--reps 10] [--variants 0,1,2,3,4]
"""
import argparse
import ctypes as C
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import lm_b200 as lm  # noqa: E402
from importlib import import_module  # noqa: E402

_lib = import_module("lm_b200._lib")


def peak():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(f))["hbm_gbs"]) if os.path.exists(f) else 6650.0


def stencil_info(lib, dev):
    i, rc, sw, m = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
    lib.lm_dbg_stencil_info.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    _lib.check(lib.lm_dbg_stencil_info(dev.handle, C.byref(i), C.byref(rc), C.byref(sw), C.byref(m)))
    return dict(id=i.value, rc=rc.value, sw=sw.value, mask=hex(m.value))


def small_cases():
    yield "haldane13x11", lm.haldane(lm.HoneycombLattice(13, 11), 1.0, 0.2, 0.1, field=lm.SymmetricGauge(0.03))
    yield "haldane_pbc9x16", lm.haldane(lm.HoneycombLattice(9, 16, boundaries=[("axis1", True), ("axis2", True)]), 1.0, 0.2, 0.1, field=lm.LandauGauge(0.02))
    yield "qwz_pbc14x15", lm.qwz(lm.SquareLattice(14, 15, boundaries=[("axis1", True)]), field=lm.LandauGauge(0.5))
    yield "square23x17", lm.tightbinding_hamiltonian(lm.SquareLattice(23, 17), field=lm.LandauGauge(0.07))
    yield "square_pbc8x8", lm.tightbinding_hamiltonian(lm.SquareLattice(8, 8, boundaries=[("axis1", True), ("axis2", True)]), field=lm.LandauGauge(0.125))


def parity(ctx, variants):
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for name, H in small_cases():
        dev = H.device(ctx)
        info = stencil_info(lib, dev)
        Hs = H.data
        N = Hs.shape[0]
        for v in variants:
            worst = 0.0
            ok = True
            for M in (32, 40, 100, 131):
                X = (rng.random((N, M)) - 0.5) + 1j * (rng.random((N, M)) - 0.5)
                x = lm.DeviceState.from_psi(X, ctx=ctx)
                y = lm.DeviceState.from_psi(np.zeros_like(X), ctx=ctx)
                lib.lm_dbg_set_apply_path(5)
                lib.lm_dbg_set_stencil_variant(v)
                st = lib.lm_spmm_state(dev.handle, x.handle, y.handle)
                if st != 0:
                    ok = False
                    break
                want = Hs @ X
                worst = max(worst, float(np.abs(y.download() - want).max() / np.abs(want).max()))
                # one product-form step (MODE 3) against the default ELL path
                a = lm.DeviceState.from_psi(X, ctx=ctx)
                b = lm.DeviceState.from_psi(X, ctx=ctx)
                nmv = C.c_int32()
                _lib.check(lib.lm_step(dev.handle, a.handle, 0.1, 1e-12, 0, C.byref(nmv)))
                lib.lm_dbg_set_apply_path(2)
                _lib.check(lib.lm_step(dev.handle, b.handle, 0.1, 1e-12, 0, C.byref(nmv)))
                worst = max(worst, float(np.abs(a.download() - b.download()).max()))
            lib.lm_dbg_set_apply_path(-1)
            lib.lm_dbg_set_stencil_variant(-1)
            print(json.dumps(dict(check="parity", case=name, variant=v, compiled=ok, relerr=worst, **info)), flush=True)


def timing(ctx, cases, variants, reps):
    lib = _lib.load()
    pk = peak()
    for kind, n, M in cases:
        if kind == "square":
            H = lm.tightbinding_hamiltonian(lm.SquareLattice(n, n))
        elif kind == "qwz":
            H = lm.qwz(lm.SquareLattice(n, n), field=lm.LandauGauge(0.01))
        elif kind == "kagome":
            H = lm.tightbinding_hamiltonian(lm.KagomeLattice(n, n), field=lm.LandauGauge(0.01))
        elif kind == "kagome2":
            H = lm.tightbinding_hamiltonian(lm.KagomeLattice(n, n), t1=1, t2=0.3)
        elif kind == "kagome3":
            H = lm.tightbinding_hamiltonian(lm.KagomeLattice(n, n), t1=1, t2=0.3, t3=0.1)
        elif kind == "kmr":        # Kane-Mele + spin-mixing NN term: run-time specialised pattern
            sz, sx = np.array([[1, 0], [0, -1]], complex), np.array([[0, 1], [1, 0]], complex)
            H = lm.construct_hamiltonian(lm.HoneycombLattice(n, n), 2, (1.0, lm.NearestNeighbor(1)), (0.2j * sz, lm.honeycomb_2nn), (0.3j * sx, lm.NearestNeighbor(1)))
        elif kind == "kanemele":
            H = lm.kanemele(lm.HoneycombLattice(n, n), 1.0, 0.2)
        elif kind == "kanemele_field":
            H = lm.kanemele(lm.HoneycombLattice(n, n), 1.0, 0.2, field=lm.LandauGauge(0.01))
        else:
            H = lm.haldane(lm.HoneycombLattice(n, n), 1.0, 0.2, 0.1)
        dev = H.device(ctx)
        N = dev.N
        rng = np.random.default_rng(1)
        blk = ((rng.random((N, 32)) - 0.5) + 1j * (rng.random((N, 32)) - 0.5))
        psi = np.asfortranarray(np.tile(blk, (1, (M + 31) // 32))[:, :M])
        x = lm.DeviceState.from_psi(psi, ctx=ctx, shard=False)
        y = x.copy()
        del psi
        bytes_spmm = 2.0 * N * M * 16 + dev.nnz * 20 + 4.0 * (N + 1)

        def timeit(fn, r):
            for _ in range(3):
                fn()
            ctx.synchronize()
            ctx.timer_start()
            for _ in range(r):
                fn()
            return ctx.timer_stop() / r

        nmv = C.c_int32()
        runs = [("ell", -1, -1)] + [("stencil", 5, v) for v in variants]
        for tag, path, v in runs:
            lib.lm_dbg_set_apply_path(path if path >= 0 else (3 if kind in ("qwz", "kanemele", "kanemele_field", "kmr") else 2))
            lib.lm_dbg_set_stencil_variant(v)
            if lib.lm_spmm_state(dev.handle, x.handle, y.handle) != 0:
                continue
            ms = timeit(lambda: _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle)), reps)
            ms_step = timeit(lambda: _lib.check(lib.lm_step(dev.handle, x.handle, 0.1, 1e-12, 0, C.byref(nmv))), max(3, reps // 3))
            K = nmv.value
            print(json.dumps(dict(check="timing", kind=kind, n=n, M=M, kernel=tag, variant=v, spmm_ms=ms,
                                  spmm_frac=bytes_spmm / ms / 1e6 / pk, step_ms=ms_step, K=K,
                                  step_frac=K * bytes_spmm / ms_step / 1e6 / pk, steps_s=1e3 / ms_step)), flush=True)
        lib.lm_dbg_set_apply_path(-1)
        lib.lm_dbg_set_stencil_variant(-1)
        # fused observables: stencil kernel (default) against the ELL-plan kernel
        rho = np.zeros(N // dev.n_int)
        J = np.zeros(max(1, len(dev.pairs()[0])))
        for tag, path in (("obs_stencil", -1), ("obs_ell_plan", 2)):
            lib.lm_dbg_set_apply_path(path)
            ms = timeit(lambda: _lib.check(lib.lm_observables(dev.handle, x.handle, _lib.ptr(rho), _lib.ptr(J))), max(3, reps // 2))
            print(json.dumps(dict(check="timing", kind=kind, n=n, M=M, kernel=tag, obs_ms=ms,
                                  obs_frac=(N * M * 16.0) / ms / 1e6 / pk, rho_sum=float(rho.sum()), j_abs=float(np.abs(J).sum()))), flush=True)
        lib.lm_dbg_set_apply_path(-1)
        del x, y, dev, timeit
        gc.collect()        # free the big device buffers now, not in the middle of the next case's timing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,8,9,10,11,12")
    ap.add_argument("--cases", default="haldane:500,qwz:300,square:100")
    ap.add_argument("--skip-parity", action="store_true")
    args = ap.parse_args()
    variants = [int(v) for v in args.variants.split(",")]
    ctx = lm.Context(precision="c128")
    lib = _lib.load()
    lib.lm_dbg_set_stencil_variant.argtypes = [C.c_int32]
    lib.lm_dbg_set_apply_path.argtypes = [C.c_int32]
    if not args.skip_parity:
        parity(ctx, variants)
    cases = []
    for c in args.cases.split(","):
        kind, n = c.split(":")
        cases.append((kind, int(n), args.M if kind != "square" else max(args.M, 2048)))
    timing(ctx, cases, variants, args.reps)


if __name__ == "__main__":
    main()
