#!/bin/bash
# Round-2 GPU call M: ncu launch list + one full capture of the dominant kernel for the DEFAULT bench command (c4, M = 4096).
set -u
OUT=gpurun_out/r2m2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_r2_bench_default.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
python tools/ncu_summary.py launches "$OUT/launches_r2_bench_default.csv" | head -16
ncu --set full --clock-control none --import-source on -k regex:k_apply_stencil_tma -s 12 -c 1 -o "$OUT/c4_m4096_stencil" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
python tools/ncu_summary.py full "$OUT/c4_m4096_stencil.ncu-rep" | head -26
ncu --set full --clock-control none -k regex:k_observe_stencil -s 3 -c 1 -o "$OUT/c4_m4096_observe" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_obs.log" 2>&1
python tools/ncu_summary.py full "$OUT/c4_m4096_observe.ncu-rep" | head -12
echo "== done"
