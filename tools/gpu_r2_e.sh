#!/usr/bin/env bash
set -u
out=gpurun_out/r2e
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run repro_ring 120 python tools/repro_ring.py 16 64
run repro_ring2 120 python tools/repro_ring.py 40 72
run pytest_gpu 900 python -m pytest tests -m gpu -q
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run bench_default 400 $B
MAREX_POOL_RING=0 run bench_noring 400 $B --no-parity
run ncu_ring 600 ncu --set full --clock-control none --import-source on -k regex:"hobday_ring" -s 1 -c 1 -o "$out/prof_ring" $B --steps 1 --no-parity
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'].get('tuning_env'), round(d['ms_per_step'], 2), {k: round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'], d.get('parity'))
"
tail -5 "$out/pytest_gpu.log"; cat "$out/repro_ring.log" | tail -3
