#!/usr/bin/env bash
# Exact-percentile queue kernel: parity tests, config-3 bench, a light ncu capture (speed-of-light, occupancy, warp states).
out=gpurun_out/r2exact; mkdir -p "$out"; : > "$out/steps.log"
run() { name=$1; lim=$2; shift 2; echo "== $name" >> "$out/steps.log"; ( time timeout "$lim" "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" >> "$out/steps.log"; }
run pytest_exact 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exact"
tail -3 "$out/pytest_exact.log"
W3=0.25deg_40yr_shifting_hobday_exact
run bench_exact 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $W3
grep -h '"metric"' "$out/bench_exact.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms_per_step'], 2), {k.replace('marex_',''): round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'], (d.get('parity') or {}).get('ok'))
"
if [ "${XQ_NCU:-0}" = 1 ]; then
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity"
run ncu_queue 600 ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section InstructionStats --section SourceCounters --import-source on --clock-control none -k regex:"hobday_exact_queue" -s 1 -c 1 -o "$out/prof_queue2" $B --workload $W3
fi
cat "$out/steps.log"
