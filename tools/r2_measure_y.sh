#!/bin/bash
# Round-2 GPU call Y: the shipped library - full GPU suite, smoke, the default bench command and the reference arm.
set -u
OUT=gpurun_out/r2y
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/smoke.txt"
timeout 900 python bench.py 2> "$OUT/default.err" | tail -1 > "$OUT/default.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('default steps/s %.2f e2e %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks'], d['parity_check']['max_rel'], {k: (round(v.get('value', 0), 2), round(v.get('e2e', 0), 2), round(v.get('roofline_frac', 0), 3)) for k, v in d.get('secondary', {}).items()}, d['cpu_baseline'])" "$OUT/default.json"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> "$OUT/reference.err" | tail -1 > "$OUT/reference.json"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('reference arm', d['value'], d['unit'], d['cpu_baseline'])" "$OUT/reference.json"
echo "== done"
