#!/usr/bin/env bash
# round 2, final validation: GPU tests, smoke, the bench line of record, launch list and full captures of every kernel of the step
set -u
out=gpurun_out/r2final
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench_full 900 python bench.py --steps 20 --warmup 3
run bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" $B
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"shift_daily|digitize_doy|hobday_band|compare_bins|transpose" -s 15 -c 5 -o "$out/prof_main" $B
run ncu_exact 900 ncu --set full --clock-control none -k regex:"hobday_exact_win|col_minmax" -s 2 -c 2 -o "$out/prof_exact" $B --workload 0.25deg_40yr_shifting_hobday_exact
run ncu_icon 900 ncu --set full --clock-control none -k regex:"hobday_hist_kernel" -s 1 -c 1 -o "$out/prof_hist" $B --workload icon_1Mi_cells_30yr_shifting_hobday_approx
run ncu_global 900 ncu --set full --clock-control none -k regex:"global_hist_fast" -s 1 -c 1 -o "$out/prof_global" $B --workload icon_1Mi_cells_30yr_detrend_global
grep -h '"metric"' "$out"/bench_*.log | cut -c1-600
tail -3 "$out/pytest_gpu.log"; tail -2 "$out/smoke.log"
