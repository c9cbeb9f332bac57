#!/bin/bash
# Round-2 GPU call K: BASELINE config 5 sweep (square lattices 50^2 .. 1000^2 with Landau phases, block widths 1 .. 8192,
# complex128 and complex64) and the config table with the round-2 kernels; final suite + smoke + default bench line.
set -u
OUT=gpurun_out/r2k
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/smoke.txt"
timeout 1200 python tools/sweep.py --preset c5 --reps 12 --what spmm,step,obs > "$OUT/sweep_c5_c128_r2.jsonl" 2> "$OUT/sweep_c5_c128.err"
timeout 1200 python tools/sweep.py --preset c5 --reps 12 --precision c64 --what spmm,step,obs > "$OUT/sweep_c5_c64_r2.jsonl" 2> "$OUT/sweep_c5_c64.err"
timeout 900 python tools/sweep.py --preset configs --reps 12 --what spmm,step,obs > "$OUT/sweep_configs_c128_r2.jsonl" 2> "$OUT/sweep_configs.err"
timeout 900 python tools/sweep.py --preset configs --reps 12 --precision c64 --what spmm,step,obs > "$OUT/sweep_configs_c64_r2.jsonl" 2>> "$OUT/sweep_configs.err"
wc -l "$OUT"/*.jsonl
timeout 900 python bench.py 2> "$OUT/default.err" | tail -1 > "$OUT/default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('default steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'], d['parity_check']['max_rel'])" "$OUT/default.json"
echo "== done"
