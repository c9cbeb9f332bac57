#!/bin/bash
# Round-2 GPU call U: L2 prefetch of a later CTA's box (LM_STENCIL_PF) re-measured with the value-class path: distance sweep on c4,
# and on / off on the 512-column shard, c3 (complex values), c2.
set -u
OUT=gpurun_out/r2u
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { # name, bench args (quoted), env...
  local name=$1; local bargs=$2; shift; shift
  env "$@" timeout 600 python bench.py $bargs --warmup 3 --no-secondary --no-cpu-baseline 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
}
for pf in 0 148 222 444 666 888; do run c4_pf$pf "--steps 12" LM_STENCIL_PF=$pf; done
for pf in 0 -1; do
  run c4m512_pf$pf "--steps 24 --M 512" LM_STENCIL_PF=$pf
  run c3_pf$pf "--steps 16 --workload c3" LM_STENCIL_PF=$pf
  run c2_pf$pf "--steps 60 --workload c2" LM_STENCIL_PF=$pf
  run c2m625_pf$pf "--steps 100 --workload c2 --M 625" LM_STENCIL_PF=$pf
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2u/*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "steps/s %.3f e2e %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
echo "== done"
