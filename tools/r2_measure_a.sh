#!/bin/bash
# Round-2 GPU call A: GPU parity suite on the round-1 code + the strip schedule x PDL grid the
# round-1 verdict asked for (trimmed from tools/r2_measure.sh to fit ~20 GPU-minutes).
set -u
OUT=gpurun_out/r2a
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee "$OUT/gpu.txt"
free -g | head -2 | tee -a "$OUT/gpu.txt"; nproc | tee -a "$OUT/gpu.txt"
echo "== 1. GPU parity suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee "$OUT/pytest_gpu.txt"
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-24s %9.2f %s  e2e %9.2f  frac %.3f  launches %s  clk %s %s" % (sys.argv[2], d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("%-24s FAILED (%s)" % (sys.argv[2], e))
PY
}
echo "== 2. C2 whole block: plain vs strips x PDL"
run c2_plain          LM_STEP_L2_MB=0  LM_STEP_PDL=0 -- --workload c2 --steps 40 --warmup 5
run c2_pdl            LM_STEP_L2_MB=0  LM_STEP_PDL=1 -- --workload c2 --steps 40 --warmup 5
for mb in 40 56 72; do
    run c2_l2_${mb}       LM_STEP_L2_MB=$mb LM_STEP_PDL=0 -- --workload c2 --steps 40 --warmup 5
    run c2_l2_${mb}_pdl   LM_STEP_L2_MB=$mb LM_STEP_PDL=1 -- --workload c2 --steps 40 --warmup 5
done
run c2_auto_pdl       LM_STEP_L2_MB=auto LM_STEP_PDL=1 LM_DEBUG_PLAN=1 -- --workload c2 --steps 40 --warmup 5
echo "== 3. C2 narrow shard (1 of 8 GPUs)"
run c2_m625_plain     LM_STEP_L2_MB=0  LM_STEP_PDL=0 -- --workload c2 --steps 100 --warmup 10 --M 625
run c2_m625_pdl       LM_STEP_L2_MB=0  LM_STEP_PDL=1 -- --workload c2 --steps 100 --warmup 10 --M 625
run c2_m625_l2_56_pdl LM_STEP_L2_MB=56 LM_STEP_PDL=1 -- --workload c2 --steps 100 --warmup 10 --M 625
run c2_m625_auto_pdl  LM_STEP_L2_MB=auto LM_STEP_PDL=1 LM_DEBUG_PLAN=1 -- --workload c2 --steps 100 --warmup 10 --M 625
echo "== 3b. single ket (ELL kernel)"
run c2_m1_plain       LM_STEP_PDL=0 -- --workload c2 --steps 400 --warmup 20 --M 1
run c2_m1_pdl         LM_STEP_PDL=1 -- --workload c2 --steps 400 --warmup 20 --M 1
echo "== 4. C3 / C4 shard (M = 512): PDL"
run c3_m512_plain     LM_STEP_PDL=0 -- --workload c3 --steps 20 --warmup 3 --M 512
run c3_m512_pdl       LM_STEP_PDL=1 -- --workload c3 --steps 20 --warmup 3 --M 512
run c4_m512_plain     LM_STEP_PDL=0 -- --workload c4 --steps 20 --warmup 3 --M 512
run c4_m512_pdl       LM_STEP_PDL=1 -- --workload c4 --steps 20 --warmup 3 --M 512
grep -h "step schedule" "$OUT"/*.err | tee "$OUT/auto_choices.txt"
echo "== done"
