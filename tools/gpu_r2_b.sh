#!/usr/bin/env bash
# round 2, call B: ring band kernel + lane-arithmetic compare: parity, bench, launch list, full captures.
set -u
out=gpurun_out/r2b
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 900 python -m pytest tests -m gpu -q
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run bench_default 300 $B
MAREX_POOL_RING=0 run bench_noring 300 $B
MAREX_POOL_TMA=0 run bench_ring_notma 300 $B
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_bench.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"shift_daily|hobday_ring|compare_bins" -s 9 -c 3 -o "$out/prof_main" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'].get('tuning_env'), round(d['ms_per_step'], 2), {k: round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'])
"
tail -5 "$out/pytest_gpu.log"
