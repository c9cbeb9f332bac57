#!/bin/bash
# Round-2 GPU call E: suite, dense-path bench lines (c1, c1x) + ncu of the 3M DMMA GEMM, e2e with the upload stream.
set -u
OUT=gpurun_out/r2e
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
for w in c1 c1x; do
  timeout 900 python bench.py --workload $w --steps 50 --warmup 5 2> "$OUT/$w.err" | tail -1 > "$OUT/$w.json"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'steps/s %.2f e2e %.2f' % (d['value'], d['e2e']['value']), 'roofline', d['roofline']['achieved'], '/', d['roofline']['peak'], d['roofline']['unit'], 'frac %.3f' % d['roofline']['frac'], 'parity', d['parity_check']['max_rel'], 'cpu', d.get('cpu_baseline',{}).get('value'))" "$OUT/$w.json" $w || tail -5 "$OUT/$w.err"
done
LM_DENSE_3M_MIN=100000 timeout 900 python bench.py --workload c1x --steps 50 --warmup 5 --no-cpu-baseline 2> "$OUT/c1x_old.err" | tail -1 > "$OUT/c1x_old.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('c1x with the round-1 32x32 kernel: steps/s %.2f' % d['value'], 'TFLOP/s %.2f' % d['roofline']['achieved'])" "$OUT/c1x_old.json"
ncu --set full --clock-control none --import-source on -k regex:k_zgemm_dmma_3m -s 4 -c 2 -o "$OUT/c1x_zgemm" \
    python bench.py --no-cpu-baseline --workload c1x --steps 3 --warmup 3 > "$OUT/ncu_zgemm.log" 2>&1
for M in 512 4096; do
  timeout 900 python bench.py --no-cpu-baseline --workload c4 --M $M --steps 20 --warmup 3 2> "$OUT/c4_m$M.err" | tail -1 > "$OUT/c4_m$M.json"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('c4 M', sys.argv[2], 'steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'])" "$OUT/c4_m$M.json" $M
done
echo "== done"
