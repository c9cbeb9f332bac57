#!/bin/bash
# Round-2 GPU call P: the real / imaginary value class of the stencil kernel (scalar values, two FMAs per element).
# New GPU tests, then A/B on config 4 (M = 4096 block and the 512-column shard) and configs 3 / 2, then one ncu capture.
set -u
OUT=gpurun_out/r2p
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_zz_gpu_patterns.py -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_patterns.txt"
for ri in 1 0; do
  LM_STENCIL_RI=$ri timeout 600 python bench.py --steps 12 --warmup 3 --no-secondary --no-cpu-baseline 2> "$OUT/c4_ri$ri.err" | tail -1 > "$OUT/c4_ri$ri.json"
  LM_STENCIL_RI=$ri timeout 600 python bench.py --steps 20 --warmup 3 --M 512 --no-cpu-baseline 2> "$OUT/c4_m512_ri$ri.err" | tail -1 > "$OUT/c4_m512_ri$ri.json"
  LM_STENCIL_RI=$ri timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline 2> "$OUT/c2_ri$ri.err" | tail -1 > "$OUT/c2_ri$ri.json"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p/c*_ri*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "steps/s %.3f e2e %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "parity %.2e" % d["parity_check"]["max_rel"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
ncu --set full --clock-control none --import-source on -k regex:k_apply_stencil_tma -s 12 -c 1 -o "$OUT/c4_m512_stencil_ri" \
    python bench.py --steps 2 --warmup 3 --M 512 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
python tools/ncu_summary.py full "$OUT/c4_m512_stencil_ri.ncu-rep" | head -26
echo "== done"
