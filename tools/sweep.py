#!/usr/bin/env python
"""SpMM / propagator bandwidth sweep (BASELINE config 5) and kernel-variant timing.

Times device-resident lm_spmm_state (plain Y = H X) and lm_step with the library's CUDA-event
timer, reports algorithmic GB/s (2 N M s + nnz (s+4) + 4 (N+1) per SpMM-equivalent) against the
measured HBM copy bandwidth.  Usage:
    python tools/sweep.py --lattice square --n 100 --M 5000 [--precision c128] [--reps 20]
    python tools/sweep.py --preset c5      # the full sweep table (CSV on stdout)
"""
import argparse
import ctypes as C
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


def make_ham(kind, n):
    if kind == "square":
        return lm.tightbinding_hamiltonian(lm.SquareLattice(n, n))
    if kind == "square_phase":
        return lm.tightbinding_hamiltonian(lm.SquareLattice(n, n), field=lm.LandauGauge(0.0123))
    if kind == "qwz":
        return lm.qwz(lm.SquareLattice(n, n), field=lm.LandauGauge(0.01))
    if kind == "haldane":
        return lm.haldane(lm.HoneycombLattice(n, n), 1.0, 0.2, 0.1)
    raise SystemExit("unknown lattice kind")


def run_case(ctx, kind, n, M, reps, tol=1e-12, what=("spmm", "step", "triad", "obs")):
    lib = _lib.load()
    H = make_ham(kind, n)
    dev = H.device(ctx)
    N = dev.N
    esz = 16 if ctx.precision == _lib.LM_C128 else 8
    # device-generated block (uniform complex in [-1, 1]^2): no multi-GB host arrays
    x = lm.DeviceState.synthetic(N, M, ctx=ctx, seed=1, shard=False)
    y = x.copy()
    z = x.copy()
    bytes_spmm = 2.0 * N * M * esz + dev.nnz * (esz + 4) + 4.0 * (N + 1)
    out = dict(kind=kind, n=n, N=N, M=M, W=dev.W, nnz=dev.nnz, precision="c128" if esz == 16 else "c64")
    pk = peak()

    def timeit(fn, r):
        for _ in range(3):
            fn()
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(r):
            fn()
        return ctx.timer_stop() / r

    if "spmm" in what:
        ms = timeit(lambda: _lib.check(lib.lm_spmm_state(dev.handle, x.handle, y.handle)), reps)
        out.update(spmm_ms=ms, spmm_gbs=bytes_spmm / ms / 1e6, spmm_frac=bytes_spmm / ms / 1e6 / pk)
    if "triad" in what:
        lib.lm_dbg_triad.argtypes = [C.c_void_p] * 3
        ms = timeit(lambda: _lib.check(lib.lm_dbg_triad(x.handle, z.handle, y.handle)), reps)
        out.update(triad_ms=ms, triad_gbs=3.0 * N * M * esz / ms / 1e6)
    if "refine" in what:
        out.update(bounds_gershgorin=dev.spectral_bounds())
        out.update(bounds_refined=dev.refine_bounds())
    if "step" in what:
        nmv = C.c_int32()
        for method, tag in ((0, "auto"), (2, "taylor"), (1, "cheb"), (5, "clenshaw")):
            ms = timeit(lambda: _lib.check(lib.lm_step(dev.handle, x.handle, 0.1, tol, method, C.byref(nmv))), max(3, reps // 4))
            K = nmv.value
            out.update({tag + "_ms": ms, tag + "_K": K, tag + "_gbs": K * bytes_spmm / ms / 1e6,
                        tag + "_frac": K * bytes_spmm / ms / 1e6 / pk, tag + "_steps_s": 1e3 / ms})
    if "obs" in what:
        rho = np.zeros(N // dev.n_int)
        J = np.zeros(max(1, len(dev.pairs()[0])))
        ms = timeit(lambda: _lib.check(lib.lm_observables(dev.handle, x.handle, _lib.ptr(rho), _lib.ptr(J))), max(3, reps // 4))
        out.update(obs_ms=ms, obs_gbs=(N * M * esz + 8.0 * (len(rho) + len(J))) / ms / 1e6)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", default="square")
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--M", type=int, default=5000)
    ap.add_argument("--precision", default="c128")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--preset", default="")
    ap.add_argument("--what", default="spmm,step,triad,obs")
    args = ap.parse_args()
    ctx = lm.Context(precision=args.precision)
    what = tuple(args.what.split(","))
    cases = []
    if args.preset == "c5":
        for n in (50, 100, 200, 300, 500, 700, 1000):
            for M in (1, 8, 64, 512, 2048, 8192):
                if 3.0 * n * n * M * (16 if args.precision == "c128" else 8) <= 40e9:
                    cases.append(("square_phase", n, M))
    elif args.preset == "configs":
        cases = [("square", 100, 5000), ("qwz", 300, 4096), ("haldane", 500, 4096), ("haldane", 500, 512)]
    else:
        cases = [(args.lattice, args.n, args.M)]
    for kind, n, M in cases:
        r = run_case(ctx, kind, n, M, args.reps, what=what)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
