#!/bin/bash
# Round-2 GPU call S: with the value-class path the c4 kernel waits mostly on its staging barrier - re-measure the two
# latency knobs that were neutral before (L2 prefetch of a later CTA's box, programmatic dependent launch) and the CTA shapes.
set -u
OUT=gpurun_out/r2s
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 12 --warmup 3 --no-secondary --no-cpu-baseline 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
}
run base LM_STENCIL_PF=0
run pf1wave LM_STENCIL_PF=-1
run pdl LM_STEP_PDL=1
run pf_pdl LM_STENCIL_PF=-1 LM_STEP_PDL=1
run base2 LM_STENCIL_PF=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s/*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "steps/s %.3f e2e %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
echo "== done"
