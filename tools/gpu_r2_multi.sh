#!/usr/bin/env bash
# multi-GPU runs of round 2: `gpurun --gpus N -- bash tools/gpu_r2_multi.sh N`
set -u
N=${1:-2}
out=gpurun_out/r2multi_$N; mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$out/gpus.csv" 2>&1
if [ "$N" = "2" ]; then
  run pytest_two_gpu 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "non_current_device"
  run stage1_sharded 300 $TR tools/stage1_sharded_check.py
fi
run bench_config2 900 $TR bench.py --gpus $N --steps 10 --warmup 3
run bench_config2_gather_thr 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-parity --gather-thresholds
if [ "$N" = "8" ]; then
  run bench_config5 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --workload 0.1deg_30yr_shifting_hobday_approx_per_gpu
  run bench_config4 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --workload icon_1Mi_cells_30yr_detrend_global
fi
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], 'N', d['n_gpus'], round(d['ms_per_step'], 2), '%.3e' % d['value'], (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}).get('value'))
"
