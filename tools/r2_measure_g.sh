#!/bin/bash
# Round-2 GPU call G: suite with the device eigensolver, its timing at full size.
set -u
OUT=gpurun_out/r2g
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
LM_DEBUG_PLAN=1 timeout 900 python - <<'PY' 2>&1 | tee "$OUT/eigs.txt"
import time, numpy as np, lm_b200 as lm
ctx = lm.default_context("c128")
for name, H, nev in (("haldane 500x500 (N = 5e5)", lm.haldane(lm.HoneycombLattice(500, 500), 1.0, 0.2, 0.1), 16),
                     ("qwz 300x300 (N = 1.8e5)", lm.qwz(lm.SquareLattice(300, 300)), 16),
                     ("square 1000x1000 (N = 1e6), ground state", lm.tightbinding_hamiltonian(lm.SquareLattice(1000, 1000)), 1)):
    H.device(ctx)
    ctx.synchronize()
    t0 = time.perf_counter()
    E, st = lm.eigs_lowest(H, nev, tol=1e-8, ctx=ctx)
    ctx.synchronize()
    print("EIGS %s: %d lowest levels in %.3f s, %d iterations, worst residual %.2e, E0 = %.10f" % (name, nev, time.perf_counter() - t0, lm.eigs_lowest.info["iterations"], lm.eigs_lowest.info["residuals"].max(), E[0]), flush=True)
PY
echo "== done"
