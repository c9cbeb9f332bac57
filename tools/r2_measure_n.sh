#!/bin/bash
# Round-2 GPU call N: fused observables with patch-fastest grid order and short column groups (LM_OBS_CPG).
set -u
OUT=gpurun_out/r2n
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_patterns.py -m gpu -x -q -k "observ or currents or golden or workflow or region or async" 2>&1 | tail -3
for M in 512 1024 4096; do
  for cpg in 1000 32 16 8; do
    LM_OBS_CPG=$cpg ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_observe_stencil -c 6 --csv --log-file "$OUT/obs_c4_m${M}_cpg${cpg}.csv" \
        python bench.py --no-cpu-baseline --workload c4 --M $M --steps 2 --warmup 3 > "$OUT/obs.log" 2>&1
    python - "$OUT/obs_c4_m${M}_cpg${cpg}.csv" $M $cpg <<'PY'
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
t, b = [], []
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"].startswith("gpu__time"):
        t.append(v / 1e6 if r["Metric Unit"] == "ns" else v / 1e3 if r["Metric Unit"] == "us" else v)
    else:
        b.append(v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[r["Metric Unit"]])
print("c4 M=%s cpg<=%s: k_observe_stencil %.3f ms, DRAM read %.2f GB (full-width launches)" % (sys.argv[2], sys.argv[3], t[0], b[0]))
PY
  done
done
for w in c2 c3; do
  for cpg in 1000 16; do
    LM_OBS_CPG=$cpg ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_observe_stencil -c 3 --csv --log-file "$OUT/obs_${w}_cpg${cpg}.csv" \
        python bench.py --no-cpu-baseline --workload $w --steps 2 --warmup 3 > "$OUT/obs.log" 2>&1
    echo "$w cpg<=$cpg: $(grep -v '^==' "$OUT/obs_${w}_cpg${cpg}.csv" | tail -1 | awk -F'","' '{print $(NF-1), $NF}')"
  done
done
echo "== done"
