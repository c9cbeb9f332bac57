#!/bin/bash
# Round-2 GPU call F: suite with the full-size oracle comparisons, the published disc workload, default bench line.
set -u
OUT=gpurun_out/r2f
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
for w in pub4 pub5; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 2> "$OUT/$w.err" | tail -1 > "$OUT/$w.json"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'total s %.4f (best %.4f) published %.2f  steps/s %.0f' % (d['value'], d['config']['best_seconds'], d['config']['published_seconds'], d['config']['steps_per_second']), 'launches', d['gpu_launches'], 'parity', d['parity_check'])" "$OUT/$w.json" $w || tail -5 "$OUT/$w.err"
done
timeout 900 python bench.py 2> "$OUT/default.err" | tail -1 > "$OUT/default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('default', d['config']['workload'][:40], 'steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'], d['parity_check']['max_rel'], d['cpu_baseline'])" "$OUT/default.json"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2> "$OUT/reference.err" | tail -1 > "$OUT/reference.json"; cut -c1-300 "$OUT/reference.json"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== done"
