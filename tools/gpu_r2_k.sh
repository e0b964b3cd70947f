#!/usr/bin/env bash
out=gpurun_out/r2k; mkdir -p "$out"
for args in "13 70 11 11 0" "13 70 11 5 0" "13 70 7 11 0" "13 70 9 5 0" "13 70 10 5 0" "13 70 12 5 0" "13 72 11 5 0" "13 64 11 5 0" "16 70 11 5 0"; do
  n=$(echo $args | tr ' ' '_')
  timeout 120 python tools/repro_ring.py $args > "$out/$n.log" 2>&1
  echo "== $args rc=$? $(grep -h 'bit-exact\|markers' "$out/$n.log" | tr '\n' ' ' | cut -c1-260)" | tee -a "$out/steps.log"
done
