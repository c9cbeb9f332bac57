#!/bin/bash
# Round-2 GPU call H: deferred scaling of the factor chain (LM_STEP_DEFER_SCALE) A/B on c4 / c3 / c2, same box.
set -u
OUT=gpurun_out/r2h
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_patterns.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3 | tee "$OUT/pytest_gpu.txt"
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print('%-22s %9.2f steps/s  e2e %9.2f  frac %.3f  clk %s %s parity %s' % (sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check']['max_rel']))" "$OUT/$name.json" "$name" || tail -3 "$OUT/$name.err"
}
for rep in 1 2; do
  for d in 0 1; do
    run c4_m4096_d${d}_$rep LM_STEP_DEFER_SCALE=$d -- --workload c4 --steps 20 --warmup 3
    run c4_m512_d${d}_$rep LM_STEP_DEFER_SCALE=$d -- --workload c4 --M 512 --steps 30 --warmup 5
  done
done
for d in 0 1; do
  run c3_d$d LM_STEP_DEFER_SCALE=$d -- --workload c3 --steps 20 --warmup 3
  run c2_d$d LM_STEP_DEFER_SCALE=$d -- --workload c2 --steps 40 --warmup 5
  run c2_c64_d$d LM_STEP_DEFER_SCALE=$d -- --workload c2 --steps 40 --warmup 5 --precision c64
done
echo "== done"
