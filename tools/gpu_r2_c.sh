#!/usr/bin/env bash
set -u
out=gpurun_out/r2c
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
run sanitizer_ring 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/repro_ring.py 16 64
run pytest_shift 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shifting or rolling or fused or streamed or preprocess"
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
MAREX_POOL_RING=0 run bench_reg 300 $B
MAREX_POOL_RING=0 MAREX_SHIFT_REG=0 run bench_noreg 300 $B
MAREX_POOL_RING=0 MAREX_SHIFT_D=24 run bench_reg_d24 300 $B
MAREX_POOL_RING=0 MAREX_SHIFT_D=16 run bench_reg_d16 300 $B
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'].get('tuning_env'), round(d['ms_per_step'], 2), {k: round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'])
"
tail -5 "$out/pytest_shift.log"; tail -30 "$out/sanitizer_ring.log"
