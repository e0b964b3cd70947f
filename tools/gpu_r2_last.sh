#!/usr/bin/env bash
# round 2, last validation after the exact-percentile queue kernel: GPU tests, smoke, the bench lines of record (config 2
# with e2e / cpu baseline / parity, reference arm, config 3), e2e chunk-count probes, a light ncu capture of the queue kernel
set -u
out=gpurun_out/r2last
mkdir -p "$out"; : > "$out/steps.log"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench_full 900 python bench.py --steps 20 --warmup 3
run bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
W3=0.25deg_40yr_shifting_hobday_exact
run bench_config3 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload $W3
run bench_icon 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload icon_1Mi_cells_30yr_shifting_hobday_approx
for c in 16 24; do run bench_e2e_chunks$c 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-chunks $c; done
run ncu_queue 600 ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section InstructionStats --section SourceCounters --import-source on --clock-control none -k regex:"hobday_exact_queue" -s 1 -c 1 -o "$out/prof_queue3" python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --workload $W3
grep -h '"metric"' "$out"/bench_*.log | cut -c1-300
tail -3 "$out/pytest_gpu.log"; tail -2 "$out/smoke.log"
