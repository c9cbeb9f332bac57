#!/bin/bash
# Round-2 GPU call Q: three / four rows per unit cell (kagome, Kane-Mele) on the stencil kernels: parity tests, then
# timing against the ELL kernels they ran on before; full GPU suite + smoke + default bench line with the new library.
set -u
OUT=gpurun_out/r2q
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_zz_gpu_patterns.py -m gpu -x -q -k "three_and_four or value_class" 2>&1 | tail -4 | tee "$OUT/pytest_wide.txt"
timeout 900 python tools/stencil_sweep.py --skip-parity --M 1024 --reps 10 --variants 19 --cases kagome:400,kagome2:400,kanemele:300,kanemele_field:300 > "$OUT/wide_patterns.jsonl" 2> "$OUT/wide_patterns.err"
python - <<'PY'
import json
for l in open("gpurun_out/r2q/wide_patterns.jsonl"):
    d = json.loads(l)
    if "spmm_ms" in d: print(d["kind"], d["n"], d["M"], d["kernel"], d["variant"], "spmm %.3f ms frac %.3f | step %.2f ms frac %.3f" % (d["spmm_ms"], d["spmm_frac"], d["step_ms"], d["step_frac"]))
    else: print(d["kind"], d["kernel"], "obs %.3f ms frac %.3f" % (d["obs_ms"], d["obs_frac"]))
PY
tail -3 "$OUT/wide_patterns.err"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/smoke.txt"
timeout 900 python bench.py 2> "$OUT/default.err" | tail -1 > "$OUT/default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('default steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'], d['parity_check']['max_rel'], {k: (v.get('value'), v.get('roofline_frac')) for k, v in d.get('secondary', {}).items()})" "$OUT/default.json"
echo "== done"
