#!/usr/bin/env bash
set -u
out=gpurun_out/r2n
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench_full 900 python bench.py --steps 10 --warmup 3
run bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" $B
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"shift_daily|hobday_band_kernel<2, 64|compare_bins|transpose" -s 8 -c 4 -o "$out/prof_main" $B
grep -h '"metric"' "$out"/bench_*.log | cut -c1-1500
tail -3 "$out/pytest_gpu.log"; tail -2 "$out/smoke.log"
