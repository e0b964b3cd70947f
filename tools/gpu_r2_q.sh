#!/usr/bin/env bash
set -u
out=gpurun_out/r2q; mkdir -p "$out"
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity"
r() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; timeout 400 "$@" > "$out/$name.log" 2>&1
  grep -h '"metric"' "$out/$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms_per_step'], 2), {k.replace('marex_',''): round(v['ms'], 2) for k, v in d['stages'].items()}, d['extreme_events'])
" | tee -a "$out/steps.log"; }
r fused_contig $B
MAREX_POOL_CONTIG=0 r fused_nocontig $B
MAREX_FUSE_DIGITIZE=0 r unfused $B
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "banded or digitize or hobday_approx or streamed" > "$out/pytest.log" 2>&1; tail -2 "$out/pytest.log" | tee -a "$out/steps.log"
