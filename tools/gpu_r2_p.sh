#!/usr/bin/env bash
set -u
out=gpurun_out/r2p; mkdir -p "$out"
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity"
v() { name=$1; flags=$2; shift 2; echo "== $name [$flags] $*" | tee -a "$out/steps.log"
  MAREX_NVCC_FLAGS="$flags" python -c "from marex_b200 import _build; _build.build(force=True)" > "$out/build_$name.log" 2>&1
  timeout 300 $B "$@" > "$out/bench_$name.log" 2>&1
  grep -h '"metric"' "$out/bench_$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms_per_step'], 2), {k.replace('marex_',''): round(v['ms'], 2) for k, v in d['stages'].items()})
" | tee -a "$out/steps.log"; }
v default ""
v default_exact "" --workload 0.25deg_40yr_shifting_hobday_exact
v nd512_2_exact "-DMAREX_SD_MAXTHREADS=512 -DMAREX_SD_MINBLOCKS=2" --workload 0.25deg_40yr_shifting_hobday_exact
