#!/usr/bin/env python
"""Timing of localexpect / LocalOperatorCurrents on a QWZ lattice: the generic correlator kernel.
    LM_CORR_BLOCKS=0|1 python tools/n3_bench.py [--n 300] [--M 1024]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import lm_b200 as lm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=300)
ap.add_argument("--M", type=int, default=1024)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
ctx = lm.Context(precision="c128")
lat = lm.SquareLattice(args.n, args.n)
H = lm.qwz(lat, field=lm.LandauGauge(0.01))
N = H.structure.dim
st = lm.DeviceState.synthetic(N, args.M, ctx=ctx, seed=7, lattice=lat, n_int=2)
sz = np.array([[1, 0], [0, -1]], complex)
out = {"blocks": int(os.environ.get("LM_CORR_BLOCKS", "1")), "N": N, "M": args.M}
for name, fn in (("localexpect_ms", lambda: lm.localexpect(sz, st)), ("operator_currents_ms", lambda: lm.LocalOperatorCurrents(H, st, sz).pair_values()),
                 ("density_currents_ms", lambda: lm.DensityCurrents(H, st).pair_values())):
    fn(); fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        fn()
    ctx.synchronize()
    out[name] = 1e3 * (time.perf_counter() - t0) / args.reps
out["psi_read_ms_at_peak"] = N * args.M * 16 / 6.54e12 * 1e3
print(json.dumps(out))
