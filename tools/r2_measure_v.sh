#!/bin/bash
# Round-2 GPU call V: final library (value classes, run-time patterns, automatic L2 look-ahead on two-row stencils):
# full GPU suite + smoke, default bench line, look-ahead on the three / four-row patterns, ncu launch list + full capture.
set -u
OUT=gpurun_out/r2v
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/smoke.txt"
timeout 900 python bench.py 2> "$OUT/default.err" | tail -1 > "$OUT/default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('default steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'], d['parity_check']['max_rel'], {k: (round(v.get('value', 0), 2), round(v.get('roofline_frac', 0), 3)) for k, v in d.get('secondary', {}).items()})" "$OUT/default.json"
for pf in 0 -1; do
  LM_STENCIL_PF=$pf timeout 900 python tools/stencil_sweep.py --skip-parity --M 1024 --reps 10 --variants 19 --cases kagome2:400,kanemele:300 2> /dev/null | python -c "
import json,sys
for l in sys.stdin:
    d = json.loads(l)
    if d.get('kernel') == 'stencil': print('pf=$pf', d['kind'], 'spmm frac %.3f step frac %.3f' % (d['spmm_frac'], d['step_frac']))
"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_r2_bench_default_final.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > "$OUT/ncu_launches.log" 2>&1
python tools/ncu_summary.py launches "$OUT/launches_r2_bench_default_final.csv" | head -16
ncu --set full --clock-control none --import-source on -k regex:k_apply_stencil_tma -s 12 -c 1 -o "$OUT/c4_m4096_stencil_final" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > "$OUT/ncu_full.log" 2>&1
python tools/ncu_summary.py full "$OUT/c4_m4096_stencil_final.ncu-rep" | head -26
echo "== done"
