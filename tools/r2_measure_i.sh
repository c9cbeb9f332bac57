#!/bin/bash
# Round-2 GPU call I: L2 prefetch of a later CTA's box (LM_STENCIL_PF = distance in CTAs; 0 = off, -1 = one resident wave).
set -u
OUT=gpurun_out/r2i
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_patterns.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3 | tee "$OUT/pytest_gpu.txt"
run() {
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"
    python -c "import json,sys; d=json.load(open(sys.argv[1])); print('%-22s %9.2f steps/s  e2e %9.2f  frac %.3f  clk %s %s parity %s' % (sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity_check']['max_rel']))" "$OUT/$name.json" "$name" || tail -3 "$OUT/$name.err"
}
for pf in 0 -1 222 888 1776; do
    run c4_m512_pf$pf LM_STENCIL_PF=$pf -- --workload c4 --M 512 --steps 30 --warmup 5
done
for pf in 0 -1 888; do
    run c4_m4096_pf$pf LM_STENCIL_PF=$pf -- --workload c4 --steps 20 --warmup 3
    run c3_pf$pf LM_STENCIL_PF=$pf -- --workload c3 --steps 20 --warmup 3
    run c2_pf$pf LM_STENCIL_PF=$pf -- --workload c2 --steps 40 --warmup 5
done
ncu --set full --clock-control none -k regex:k_apply_stencil_tma -s 20 -c 1 -o "$OUT/c4_m512_stencil_pf" \
    python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 3 --warmup 3 > "$OUT/ncu_full.log" 2>&1
echo "== done"
