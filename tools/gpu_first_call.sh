#!/usr/bin/env bash
# One gpurun call that measures everything the previous round left unmeasured:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_first_call.sh'      (about 15 GPU-minutes when nothing hangs)
# Every step is bounded by its own timeout and writes into gpurun_out/first/; nothing here runs under a profiler except
# the two ncu passes at the end, whose printed numbers are never bench values.
set -u
out=gpurun_out/first
mkdir -p "$out"
run() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; ( time timeout "$@" ) > "$out/$name.log" 2>&1; echo "rc=$?" | tee -a "$out/$name.log" "$out/steps.log"; }

# 1. parity: the whole GPU suite, then the experiments (float32 sums in the shifting-baseline kernel)
run pytest_gpu 400 python -m pytest tests -m gpu -x -q
MAREX_TEST_EXPERIMENTAL=1 run pytest_experimental 200 python -m pytest tests/test_gpu_parity.py tests/test_track_gpu.py -m gpu -q -k "float32_sums or lean_variant or disk_variant or pack_kernel"

# 2. config 2, device-resident: default (float64 sums) against MAREX_SHIFT_ACC=f32
run bench_default 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu
MAREX_SHIFT_ACC=f32 run bench_shift_f32 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu
MAREX_SHIFT_LEAN=1 run bench_shift_lean 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu
MAREX_SHIFT_LEAN=1 MAREX_SHIFT_ACC=f32 run bench_shift_lean_f32 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu

# 3. tracker stage 1: the measured disk kernel against its third variant
run bench_stage1 120 python tools/bench_stage1.py --days 2048 --out "$out/bench_stage1.json"
MAREX_MORPH_DISK=3 run bench_stage1_disk3 120 python tools/bench_stage1.py --days 2048 --out "$out/bench_stage1_disk3.json"
run bench_stage1_bool 120 python tools/bench_stage1.py --days 1024 --input bool --out "$out/bench_stage1_bool.json"
MAREX_MORPH_PACK=1 run bench_stage1_bool_pack 120 python tools/bench_stage1.py --days 1024 --input bool --out "$out/bench_stage1_bool_pack.json"
MAREX_MORPH_DISK=4 run bench_stage1_disk4 120 python tools/bench_stage1.py --days 2048 --out "$out/bench_stage1_disk4.json"

# (with `gpurun --gpus 2`: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/stage1_sharded_check.py)

# 4. launch lists (share of the step per kernel)
run ncu_bench 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu
run ncu_stage1 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_stage1.csv" python tools/bench_stage1.py --days 256 --reps 1

grep -h '"metric"' "$out"/bench_default.log "$out"/bench_shift_f32.log "$out"/bench_shift_lean.log "$out"/bench_shift_lean_f32.log | cut -c1-400
tail -3 "$out/pytest_gpu.log" "$out/pytest_experimental.log"
