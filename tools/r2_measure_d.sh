#!/bin/bash
# Round-2 GPU call D: GPU suite on the current build, fused-observables kernel durations (tensor-map boxes +
# recursive-halving reduction), full capture of k_observe_stencil.
set -u
OUT=gpurun_out/r2d
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
for M in 512 1024; do
  for t in 0 1; do
    LM_STENCIL_TMAP=$t ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_observe_stencil -c 20 --csv --log-file "$OUT/obs_c4_m${M}_t${t}.csv" \
        python bench.py --no-cpu-baseline --workload c4 --M $M --steps 4 --warmup 3 > "$OUT/obs_c4_m${M}_t${t}.log" 2>&1
    python tools/ncu_summary.py launches "$OUT/obs_c4_m${M}_t${t}.csv" | tail -3
  done
done
ncu --set full --clock-control none --import-source on -k regex:k_observe_stencil -s 3 -c 1 -o "$OUT/c4_m1024_observe" \
    python bench.py --no-cpu-baseline --workload c4 --M 1024 --steps 3 --warmup 3 > "$OUT/ncu_obs.log" 2>&1
python bench.py --no-cpu-baseline --workload c4 --M 512 --steps 30 --warmup 5 2> "$OUT/c4_m512.err" | tail -1 | tee "$OUT/c4_m512.json" | cut -c1-300
echo "== done"
