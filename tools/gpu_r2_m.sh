#!/usr/bin/env bash
set -u
out=gpurun_out/r2m
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
run bench_default 400 $B
MAREX_POOL_RING=0 run bench_noring 400 $B --no-parity
run bench_exact 400 $B --workload 0.25deg_40yr_shifting_hobday_exact
run bench_icon_hobday 400 $B --workload icon_1Mi_cells_30yr_shifting_hobday_approx
run ncu_ring 600 ncu --set full --clock-control none --import-source on -k regex:"hobday_ring" -s 1 -c 1 -o "$out/prof_ring" $B --steps 1 --no-parity
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], d['config'].get('tuning_env'), round(d['ms_per_step'], 2), {k.replace('marex_',''): round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'], (d.get('parity') or {}).get('ok'))
"
tail -5 "$out/pytest_gpu.log"
