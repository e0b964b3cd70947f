#!/usr/bin/env bash
# round 2, closing run: the anomaly / band / exact kernels with their default windows as compile-time constants.
# Full GPU suite, smoke, the bench lines of record (config 2 with e2e / cpu baseline / parity; config 3).
set -u
out=gpurun_out/r2last2
mkdir -p "$out"; : > "$out/steps.log"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run bench_full 900 python bench.py --steps 20 --warmup 3
run bench_config3 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload 0.25deg_40yr_shifting_hobday_exact
run pytest_gpu 1500 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
grep -h '"metric"' "$out"/bench_*.log | cut -c1-200
tail -3 "$out/pytest_gpu.log"; tail -2 "$out/smoke.log"
