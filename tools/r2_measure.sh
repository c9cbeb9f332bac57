#!/bin/bash
# One-call GPU measurement plan for the opt-in schedules added without GPU time at the end of
# round 1 (L2-resident strip schedule, programmatic dependent launch, catch-all RC = 2 stencil).
# Run on the GPU box from the repository root:
#     gpurun --timeout 1500 -- 'bash tools/r2_measure.sh'
# Everything lands in gpurun_out/r2/ (JSON lines per run, ncu CSVs).  Nothing here changes defaults:
# read the numbers, then flip the defaults in csrc/api.cu (strip_columns / pdl_env) if they win.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1

echo "== 1. GPU parity suite (new tests sort last: tests/test_zz_gpu_strips.py)"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"

echo "== 1b. randomised parity sweep against the oracle ON THE DEVICE (300 cases; the same sweep runs on the CPU build in the CPU suite)"
python tests/fuzz_parity.py 0 300 2>&1 | tail -4 | tee "$OUT/fuzz_device.txt"

run() {  # name, env..., -- bench args
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python - "$OUT/$name.json" "$name" <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-28s %9.2f %s  e2e %9.2f  frac %.3f  launches %s" % (sys.argv[2], d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches")))
except Exception as e:
    print("%-28s FAILED (%s)" % (sys.argv[2], e))
EOF
}

echo "== 2. C2 (bench default): plain vs L2-resident strips x PDL"
run c2_plain            LM_STEP_L2_MB=0  LM_STEP_PDL=0 -- --steps 40 --warmup 5
run c2_pdl              LM_STEP_L2_MB=0  LM_STEP_PDL=1 -- --steps 40 --warmup 5
for mb in 24 40 56 72 96; do
    run c2_l2_${mb}         LM_STEP_L2_MB=$mb LM_STEP_PDL=0 -- --steps 40 --warmup 5
    run c2_l2_${mb}_pdl     LM_STEP_L2_MB=$mb LM_STEP_PDL=1 -- --steps 40 --warmup 5
done
run c2_auto             LM_STEP_L2_MB=auto LM_STEP_PDL=0 LM_DEBUG_PLAN=1 -- --steps 40 --warmup 5
run c2_auto_pdl         LM_STEP_L2_MB=auto LM_STEP_PDL=1 LM_DEBUG_PLAN=1 -- --steps 40 --warmup 5
run c2_c64_plain        LM_STEP_L2_MB=0  LM_STEP_PDL=0 -- --steps 40 --warmup 5 --precision c64
run c2_c64_l2_56_pdl    LM_STEP_L2_MB=56 LM_STEP_PDL=1 -- --steps 40 --warmup 5 --precision c64

echo "== 3. narrow shards (one of 8 GPUs' share): launch-bound regime"
run c2_m625_plain       LM_STEP_L2_MB=0  LM_STEP_PDL=0 -- --steps 100 --warmup 10 --M 625
run c2_m625_pdl         LM_STEP_L2_MB=0  LM_STEP_PDL=1 -- --steps 100 --warmup 10 --M 625
run c2_m625_l2_56_pdl   LM_STEP_L2_MB=56 LM_STEP_PDL=1 -- --steps 100 --warmup 10 --M 625
run c2_m625_auto_pdl    LM_STEP_L2_MB=auto LM_STEP_PDL=1 LM_DEBUG_PLAN=1 -- --steps 100 --warmup 10 --M 625

echo "== 3b. single ket / narrow blocks (ELL kernel, launch-bound): PDL"
run c2_m1_plain         LM_STEP_PDL=0 -- --steps 400 --warmup 20 --M 1
run c2_m1_pdl           LM_STEP_PDL=1 -- --steps 400 --warmup 20 --M 1
run c2_m8_plain         LM_STEP_PDL=0 -- --steps 400 --warmup 20 --M 8
run c2_m8_pdl           LM_STEP_PDL=1 -- --steps 400 --warmup 20 --M 8

echo "== 4. C3 / C4: PDL only (N too large for strips)"
run c3_plain            LM_STEP_PDL=0 -- --workload c3 --steps 10 --warmup 3
run c3_pdl              LM_STEP_PDL=1 -- --workload c3 --steps 10 --warmup 3
run c4_plain            LM_STEP_PDL=0 -- --workload c4 --steps 6 --warmup 3
run c4_pdl              LM_STEP_PDL=1 -- --workload c4 --steps 6 --warmup 3

echo "== 5. ncu: launch list of the strip schedule, full capture of one strip factor"
LM_STEP_L2_MB=56 LM_STEP_PDL=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file "$OUT/launches_c2_l2_56_pdl.csv" python bench.py --no-cpu-baseline --steps 2 --warmup 1 > "$OUT/ncu_launches.log" 2>&1
LM_STEP_L2_MB=56 ncu --set full --clock-control none --import-source on -k regex:k_apply_stencil_tma -s 300 -c 2 \
    -o "$OUT/c2_strip_factor" python bench.py --no-cpu-baseline --steps 2 --warmup 1 > "$OUT/ncu_full.log" 2>&1
ncu -i "$OUT/c2_strip_factor.ncu-rep" --page raw --csv > "$OUT/c2_strip_factor_raw.csv" 2>/dev/null
python tools/ncu_summary.py "$OUT/c2_strip_factor_raw.csv" > "$OUT/c2_strip_factor_summary.txt" 2>&1 || true
echo "== done; results in $OUT"
