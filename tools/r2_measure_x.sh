#!/bin/bash
# Round-2 GPU call X: exact QWZ pattern (id 9: diagonal on-site term, 9 entries per row) against the full-block pattern (id 3, hidden
# from the matcher with LM_STENCIL_SKIP=0x200 -> id 3 is used) on c3; final GPU suite + smoke with the shipped library.
set -u
OUT=gpurun_out/r2x
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
for skip in 0 0x200 0 0x200; do
  LM_STENCIL_SKIP=$skip timeout 600 python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline 2>> "$OUT/c3.err" | tail -1 > "$OUT/c3_skip$skip.json"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('skip=$skip c3 %.3f steps/s e2e %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['clocks']['sm_mhz'], 'parity %.2e' % d['parity_check']['max_rel'])" "$OUT/c3_skip$skip.json"
done
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/smoke.txt"
echo "== done"
