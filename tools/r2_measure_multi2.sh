#!/bin/bash
# Multi-GPU validation of the final library: sharded vs unsharded parity (both exchange paths, replicated-ket case) and the
# DEFAULT bench command as the driver launches it (c4 with the c3 / c2 sub-objects).   gpurun --gpus N -- 'bash tools/r2_measure_multi2.sh N'
set -u
N=${1:-2}
OUT=gpurun_out/r2t
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_gpu_check.py 2>&1 | tail -4 | tee "$OUT/multi_n${N}_check.txt"
LM_OBS_P2P=0 timeout 600 $TR tests/multi_gpu_check.py 2>&1 | tail -3 | tee -a "$OUT/multi_n${N}_check.txt"
timeout 1200 $TR bench.py --gpus "$N" --steps 20 --warmup 3 2> "$OUT/multi_n${N}_default.err" | tail -1 > "$OUT/multi_n${N}_default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('c4 %10.2f %s  e2e %10.2f  frac %.3f  clk %s %s parity %s' % (d['value'], d['unit'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check'])); print({k: (round(v.get('value', 0), 2), round(v.get('e2e', 0), 2), round(v.get('roofline_frac', 0), 3), v.get('parity_max_rel')) for k, v in d.get('secondary', {}).items()})" "$OUT/multi_n${N}_default.json" || tail -5 "$OUT/multi_n${N}_default.err"
echo "== done"
