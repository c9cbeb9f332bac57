#!/bin/bash
# Round-2 GPU call C: CTA-shape experiments of the staged stencil kernel (more, smaller CTAs per SM),
# new GPU tests (dense-P observables, async updates, non-Hermitian refusal).
set -u
OUT=gpurun_out/r2c
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
echo "== 1. new GPU tests"
timeout 900 python -m pytest tests/test_zz_gpu_patterns.py -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-24s %9.2f %s  e2e %9.2f  frac %.3f  launches %s  clk %s %s parity %s" % (sys.argv[2], d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d.get("parity_check", {}).get("max_rel")))
except Exception as e:
    print("%-24s FAILED (%s)" % (sys.argv[2], e))
PY
}
echo "== 2. shapes: c4 / c3 shards (M = 512)"
for v in 2 13 14 15 18; do
    run c4_m512_v$v LM_STENCIL_VARIANT=$v -- --workload c4 --M 512 --steps 30 --warmup 5
    run c3_m512_v$v LM_STENCIL_VARIANT=$v -- --workload c3 --M 512 --steps 30 --warmup 5
done
echo "== 3. shapes: c2"
for v in 7 16 17 18; do
    run c2_v$v LM_STENCIL_VARIANT=$v -- --workload c2 --steps 40 --warmup 5
    run c2_m625_v$v LM_STENCIL_VARIANT=$v -- --workload c2 --M 625 --steps 100 --warmup 10
done
echo "== 4. c4 whole block with the two best candidates"
for v in 13 15 18; do
    run c4_m4096_v$v LM_STENCIL_VARIANT=$v -- --workload c4 --steps 12 --warmup 3
done
echo "== done"
