#!/bin/bash
# Round-2 GPU call W (8 GPUs): end-to-end leg with the NVLink broadcast of the H values (lm_ham_update_values_bcast, default at N > 1)
# against one PCIe upload per rank (LM_BENCH_BCAST=0), plus the multi-GPU check with the broadcast case.
set -u
N=${1:-8}
OUT=gpurun_out/r2w
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_gpu_check.py 2>&1 | tail -2 | tee "$OUT/multi_n${N}_check.txt"
for b in 1 0; do
  LM_BENCH_BCAST=$b timeout 900 $TR bench.py --gpus "$N" --steps 20 --warmup 3 --no-secondary 2> "$OUT/n${N}_bcast$b.err" | tail -1 > "$OUT/n${N}_bcast$b.json"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bcast=$b c4 %10.2f %s  e2e %10.2f  frac %.3f  clk %s %s parity %.2e' % (d['value'], d['unit'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check']['max_rel']))" "$OUT/n${N}_bcast$b.json" || tail -5 "$OUT/n${N}_bcast$b.err"
done
echo "== done"
