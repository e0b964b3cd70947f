#!/usr/bin/env bash
# anomaly kernel with S / W / strip length as compile-time constants: parity tests, then config 2 with and without it
out=gpurun_out/r2spec; mkdir -p "$out"; : > "$out/steps.log"
run() { name=$1; lim=$2; shift 2; echo "== $name" >> "$out/steps.log"; ( time timeout "$lim" "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" >> "$out/steps.log"; }
run pytest_shift 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shifting or climatology or constants_folded or fused_bin"
tail -3 "$out/pytest_shift.log"
run bench_spec 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu
MAREX_SHIFT_GENERIC=1 run bench_generic 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity
grep -h '"metric"' "$out/bench_spec.log" "$out/bench_generic.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms_per_step'], 2), {k.replace('marex_',''): round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'], (d.get('parity') or {}).get('ok'))
"
cat "$out/steps.log" | paste - -
