#!/bin/bash
# Round-2 multi-GPU call: sharded vs unsharded parity (both exchange paths, incl. the replicated-ket case)
# and the headline bench on N GPUs.   gpurun --gpus N -- 'bash tools/r2_measure_multi.sh N'
set -u
N=${1:-2}
OUT=gpurun_out/r2m
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== sharded vs unsharded evolution + observables (peer-memory exchange, then NCCL)"
timeout 600 $TR tests/multi_gpu_check.py 2>&1 | tail -4 | tee "$OUT/multi_n${N}_check.txt"
LM_OBS_P2P=0 timeout 600 $TR tests/multi_gpu_check.py 2>&1 | tail -3 | tee -a "$OUT/multi_n${N}_check.txt"
run() {
    local name=$1; shift
    timeout 1200 $TR bench.py --gpus "$N" "$@" 2> "$OUT/multi_n${N}_$name.err" | tail -1 > "$OUT/multi_n${N}_$name.json"
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print('%-14s %10.2f %s  e2e %10.2f  frac %.3f  clk %s %s parity %s' % (sys.argv[2], d['value'], d['unit'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check']))" "$OUT/multi_n${N}_$name.json" "$name" || { echo "$name FAILED"; tail -5 "$OUT/multi_n${N}_$name.err"; }
}
run c4 --steps 20 --warmup 3
run c3 --workload c3 --steps 20 --warmup 3
run c2 --workload c2 --steps 100 --warmup 10
echo "== done"
