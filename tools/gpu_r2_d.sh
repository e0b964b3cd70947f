#!/usr/bin/env bash
set -u
out=gpurun_out/r2d
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
export MAREX_POOL_RING=0
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
run ncu_reg 600 ncu --set full --clock-control none --import-source on -k regex:"shift_reg" -s 2 -c 1 -o "$out/prof_shift_reg" $B
MAREX_SHIFT_REG=0 run ncu_smem 600 ncu --set full --clock-control none --import-source on -k regex:"shift_daily" -s 2 -c 1 -o "$out/prof_shift_daily" $B
MAREX_SHIFT_REG=0 run ncu_rest 600 ncu --set full --clock-control none --import-source on -k regex:"hobday_band|compare_bins" -s 4 -c 2 -o "$out/prof_band_compare" $B
ls -la "$out"
