#!/bin/bash
# Round-2 GPU call J: values through L1 (ld.global.nc) instead of shared memory: variants 20 (RC = 2) / 19 (RC = 1)
# against the defaults 2 / 7.  Library built with -DLM_STENCIL_SHAPES.
set -u
OUT=gpurun_out/r2j
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print('%-22s %9.2f steps/s  e2e %9.2f  frac %.3f  clk %s %s parity %s' % (sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check']['max_rel']))" "$OUT/$name.json" "$name" || tail -3 "$OUT/$name.err"
}
for v in 2 20; do
    run c4_m512_v$v LM_STENCIL_VARIANT=$v -- --workload c4 --M 512 --steps 30 --warmup 5
    run c4_m4096_v$v LM_STENCIL_VARIANT=$v -- --workload c4 --steps 20 --warmup 3
    run c3_v$v LM_STENCIL_VARIANT=$v -- --workload c3 --steps 20 --warmup 3
done
for v in 7 19; do
    run c2_v$v LM_STENCIL_VARIANT=$v -- --workload c2 --steps 40 --warmup 5
done
LM_STENCIL_VARIANT=20 ncu --set full --clock-control none -k regex:k_apply_stencil_tma -s 20 -c 1 -o "$OUT/c4_m512_stencil_vldg" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_full.log" 2>&1
echo "== done"
