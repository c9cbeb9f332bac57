#!/bin/bash
# Round-2 GPU call L: why is lm_observables slow for (complex64, square 1000^2, M = 512)?  launch lists of the odd sweep cases.
set -u
OUT=gpurun_out/r2l
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
for args in "--lattice square_phase --n 1000 --M 512 --precision c64" "--lattice square_phase --n 700 --M 512 --precision c64" "--lattice square_phase --n 1000 --M 512 --precision c128" "--lattice haldane --n 500 --M 4096 --precision c64" "--lattice haldane --n 500 --M 4096 --precision c128"; do
  tag=$(echo $args | tr -d ' -' )
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/l_$tag.csv" python tools/sweep.py $args --reps 4 --what obs > "$OUT/l_$tag.log" 2>&1
  echo "== $args"; tail -1 "$OUT/l_$tag.log" | cut -c1-200
  python tools/ncu_summary.py launches "$OUT/l_$tag.csv" | head -8
done
echo "== done"
