#!/bin/bash
# Round-2 GPU call B: new stencil kernel (shared value loads for Hermitian H + tensor-map boxes),
# the headline bench on c4, the (herm, tmap) grid on the shard sizes, ncu of the c4 kernel.
set -u
OUT=gpurun_out/r2b
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
echo "== 1. GPU parity suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee "$OUT/pytest_gpu.txt"
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 900 python bench.py "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-24s %9.2f %s  e2e %9.2f  frac %.3f  launches %s  clk %s %s parity %s" % (sys.argv[2], d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d.get("parity_check", {}).get("max_rel")))
except Exception as e:
    print("%-24s FAILED (%s)" % (sys.argv[2], e))
PY
}
echo "== 2. headline: bench.py defaults (c4, M = 4096) + the reference arm"
run c4_default  X=1 --
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2> "$OUT/c4_reference.err" | tail -1 > "$OUT/c4_reference.json"; cat "$OUT/c4_reference.json" | cut -c1-400
echo "== 3. (herm, tmap) grid, one of 8 shards (M = 512) of c4 / c3, c2 whole"
for h in 0 1; do for t in 0 1; do
    run c4_m512_h${h}_t${t} LM_STENCIL_HERM=$h LM_STENCIL_TMAP=$t -- --no-cpu-baseline --workload c4 --M 512 --steps 30 --warmup 5
    run c3_m512_h${h}_t${t} LM_STENCIL_HERM=$h LM_STENCIL_TMAP=$t -- --no-cpu-baseline --workload c3 --M 512 --steps 30 --warmup 5
    run c2_h${h}_t${t}      LM_STENCIL_HERM=$h LM_STENCIL_TMAP=$t -- --no-cpu-baseline --workload c2 --steps 40 --warmup 5
done; done
run c4_m4096_h0_t0 LM_STENCIL_HERM=0 LM_STENCIL_TMAP=0 -- --no-cpu-baseline --workload c4 --steps 20 --warmup 3
run c3_default X=1 -- --no-cpu-baseline --workload c3 --steps 20 --warmup 3
echo "== 4. ncu: launch list + full capture of the c4 kernel (M = 512)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_c4_m512.csv" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_stencil_tma -s 20 -c 2 -o "$OUT/c4_m512_stencil" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_full.log" 2>&1
LM_STENCIL_HERM=0 LM_STENCIL_TMAP=0 ncu --set full --clock-control none -k regex:k_apply_stencil_tma -s 20 -c 1 -o "$OUT/c4_m512_stencil_h0t0" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_full0.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_observe_stencil -s 3 -c 1 -o "$OUT/c4_m512_observe" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_obs.log" 2>&1
ls -la "$OUT" | tail -5
echo "== done"
