#!/bin/bash
# Round-2 GPU call R: run-time specialised stencil patterns (NVRTC): parity tests, timing against the ELL kernels.
set -u
OUT=gpurun_out/r2r
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export LM_DEBUG_PLAN=0
timeout 1200 python -m pytest tests/test_zz_gpu_patterns.py -m gpu -x -q -k "run_time_specialised" 2>&1 | tail -15 | tee "$OUT/pytest_rtc.txt"
ls -la ~/.cache/lm_b200 2>/dev/null | tail -5
timeout 900 python tools/stencil_sweep.py --skip-parity --M 1024 --reps 10 --variants 19 --cases kmr:300,kagome3:400 > "$OUT/rtc_patterns.jsonl" 2> "$OUT/rtc_patterns.err"
python - <<'PY'
import json
for l in open("gpurun_out/r2r/rtc_patterns.jsonl"):
    d = json.loads(l)
    if "spmm_ms" in d: print(d["kind"], d["n"], d["M"], d["kernel"], d["variant"], "spmm %.3f ms frac %.3f | step %.2f ms frac %.3f" % (d["spmm_ms"], d["spmm_frac"], d["step_ms"], d["step_frac"]))
    else: print(d["kind"], d["kernel"], "obs %.3f ms frac %.3f" % (d["obs_ms"], d["obs_frac"]))
PY
tail -3 "$OUT/rtc_patterns.err"
echo "== done"
