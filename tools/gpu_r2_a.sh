#!/usr/bin/env bash
# round 2, call A: parity of the rewritten anomaly kernel (fused digitize), the compare from bin codes, and a sweep of the
# anomaly kernel's shapes on config 2.
set -u
out=gpurun_out/r2a
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 600 python -m pytest tests -m gpu -x -q
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run bench_default 300 $B
MAREX_SHIFT_V=2 MAREX_SHIFT_R=4 MAREX_SHIFT_CPS=1 run bench_v2r4c1 300 $B
MAREX_SHIFT_V=1 MAREX_SHIFT_R=4 MAREX_SHIFT_CPS=2 run bench_v1r4c2 300 $B
MAREX_SHIFT_V=2 MAREX_SHIFT_R=2 MAREX_SHIFT_CPS=2 run bench_v2r2c2 300 $B
MAREX_SHIFT_V=4 MAREX_SHIFT_R=2 MAREX_SHIFT_CPS=2 run bench_v4r2c2 300 $B
MAREX_SHIFT_V=4 MAREX_SHIFT_R=4 MAREX_SHIFT_CPS=1 run bench_v4r4c1 300 $B
MAREX_SHIFT_V=4 MAREX_SHIFT_R=4 MAREX_SHIFT_CPS=2 run bench_v4r4c2 300 $B
MAREX_SHIFT_V=2 MAREX_SHIFT_R=4 MAREX_SHIFT_CPS=3 run bench_v2r4c3 300 $B
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'].get('tuning_env'), round(d['ms_per_step'], 2), {k: round(v['ms'], 2) for k, v in d['stages'].items()})
"
tail -5 "$out/pytest_gpu.log"
