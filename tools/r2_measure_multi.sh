#!/bin/bash
# Multi-GPU half of the round-2 measurement plan (lead 4: the 4 / 8 GPU numbers
# were last taken with the ELL kernel generation).  Run on N GPUs of one box:
#     gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_measure_multi.sh 8'
# Results: gpurun_out/r2/multi_n<N>_*.json (one bench line each) and the sharded-vs-unsharded check.
set -u
N=${1:-8}
OUT=gpurun_out/r2
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== sharded vs unsharded evolution + observables (both exchange paths)"
$TR tests/multi_gpu_check.py 2>&1 | tail -6 | tee "$OUT/multi_n${N}_check.txt"
LM_OBS_P2P=0 $TR tests/multi_gpu_check.py 2>&1 | tail -3 | tee -a "$OUT/multi_n${N}_check.txt"
run() {  # name, env..., -- bench args
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" $TR bench.py --gpus "$N" --no-cpu-baseline "$@" 2> "$OUT/multi_n${N}_$name.err" | tail -1 > "$OUT/multi_n${N}_$name.json"
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print('%-20s %10.2f %s  e2e %10.2f  frac %.3f' % (sys.argv[2], d['value'], d['unit'], d['e2e']['value'], d['roofline']['frac']))" "$OUT/multi_n${N}_$name.json" "$name" || echo "$name FAILED"
}
run c2_plain     LM_STEP_PDL=0 -- --steps 100 --warmup 10
run c2_pdl       LM_STEP_PDL=1 -- --steps 100 --warmup 10
run c2_auto_pdl  LM_STEP_L2_MB=auto LM_STEP_PDL=1 LM_DEBUG_PLAN=1 -- --steps 100 --warmup 10
run c4_plain     LM_STEP_PDL=0 -- --workload c4 --steps 10 --warmup 3
run c4_pdl       LM_STEP_PDL=1 -- --workload c4 --steps 10 --warmup 3
run c3_plain     LM_STEP_PDL=0 -- --workload c3 --steps 10 --warmup 3
echo "== done"
