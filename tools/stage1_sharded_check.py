"""Two or more GPUs: tracker stage 1 after a space-sharded detection, re-sharded by time over NCCL (marex_b200.track.
stage1_time_sharded), checked against the single-GPU result and timed with CUDA events.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/stage1_sharded_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from bench_stage1 import synthetic_bits
    from marex_b200 import track

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    T, ny, nx, R, T_fill = 1024, 720, 1440, 8, 2
    assert ny % world == 0
    rows = ny // world
    mask = np.ones((ny, nx), bool)
    mask[: ny // 12] = False
    bits = synthetic_bits(T, ny, nx, 0.05, dev)  # same seed on every rank: the full mask, of which a rank keeps its band
    wpr = nx // 32
    band = bits.reshape(T, ny, wpr)[:, rank * rows : (rank + 1) * rows].reshape(T, rows * wpr).contiguous()
    ref = track.MaskFiller(mask, R, T_fill, device=dev).run(from_bits=(bits, T), packed=True)
    for _ in range(2):
        out, (lo, hi) = track.stage1_time_sharded(band, T, rows, mask, R, T_fill, packed=True, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, (lo, hi) = track.stage1_time_sharded(band, T, rows, mask, R, T_fill, packed=True, device=dev)
    e1.record()
    torch.cuda.synchronize()
    ok = bool((out == ref[lo:hi]).all())
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"stage1_time_sharded": {"n_gpus": world, "days": T, "grid": [ny, nx], "R_fill": R, "T_fill": T_fill,
                                                  "ms_max_over_ranks": float(ms.item()), "bit_exact_vs_single_gpu": bool(flag.item()),
                                                  "gridpoint_days_per_s": ny * nx * T / (float(ms.item()) * 1e-3)}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
