#!/usr/bin/env python
"""
bench.py -- gridpoint-days/s of the full preprocess_data pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the whole hot path (validation numbers + shifting-baseline anomalies +
approximate 5x5-pooled Hobday thresholds + compare) over one synthetic field that is already
resident in HBM.  Default workload: BASELINE.json configs[1], 0.25 deg global daily SST
(1440 x 720, 1982-2021, T = 14610), shifting_baseline + hobday_extreme p95, approximate.
N > 1 (torchrun): weak scaling, every rank owns one such field as a latitude band of a global
(720*N) x 1440 grid, loads its 2-row pooling halo, and the small outputs (thresholds, mask,
count) are gathered with NCCL.  One JSON line is printed by rank 0.

`--impl reference` times the CPU restatement of the reference (oracle/marex_oracle.py -- the
reference itself needs xarray/dask/flox/xhistogram, none of which is installed) on all host
cores, on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time as _time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ny, nx, start, end, kwargs)
    "0.25deg_40yr_shifting_hobday_approx": (720, 1440, "1982-01-01", "2022-01-01", dict()),
    "0.25deg_40yr_shifting_hobday_exact": (720, 1440, "1982-01-01", "2022-01-01", dict(method_percentile="exact")),
    "1deg_40yr_shifting_hobday_approx": (180, 360, "1982-01-01", "2022-01-01", dict()),
    "1deg_40yr_shifting_hobday_exact": (180, 360, "1982-01-01", "2022-01-01", dict(method_percentile="exact")),
    "smoke_0.5deg_20yr": (90, 180, "2000-01-01", "2020-01-01", dict(window_year_baseline=5)),
    # BASELINE.json configs[3] per GPU: 8 Mi-cell unstructured mesh / 8 GPUs, 30 yr, detrend_fixed + global p95
    "icon_1Mi_cells_30yr_detrend_global": (1, 1 << 20, "1991-01-01", "2021-01-01",
                                           dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme")),
    "icon_1Mi_cells_30yr_shifting_hobday_approx": (1, 1 << 20, "1991-01-01", "2021-01-01", dict()),
    # BASELINE.json configs[4] per GPU: 0.1 deg (3600 x 1800) / 8 GPUs = 225 rows, 30 yr
    "0.1deg_30yr_shifting_hobday_approx_per_gpu": (225, 3600, "1991-01-01", "2021-01-01", dict()),
}
PUBLISHED_PER_CORE = 3.1e4  # gridpoint-days/s/core, docs/modules/detect.rst:729-732


def algorithmic_bytes_per_cell(T, T_out, hobday=True):
    """read x once, write dat_anomaly f32 + extreme_events bool + thresholds + mask."""
    return 4 * T + 4 * T_out + 1 * T_out + (4 * 366 if hobday else 8) + 1


# per-stage algorithmic bytes per gridpoint: what one C-ABI call must read + write at minimum
# (DESIGN.md section 4).  A stage may launch several kernels (the pooled-threshold call runs the
# digitize kernel and the banded histogram kernel); scratch traffic is not algorithmic.
def kernel_bytes_per_cell(T, T_out):
    return {
        "marex_shift_anomaly_daily_f32": 4 * T + 4 * T_out + 1 + 4,  # (+ 2 * T_out of bin codes when fused: scratch, not algorithmic)
        "marex_shift_anomaly_f32": 4 * T + 4 * T_out + 1 + 4,
        "marex_digitize_doy_f32": 4 * T_out + 2 * T_out,
        "marex_hobday_thresholds_pooled_bins": 2 * T_out + 4 * 366 + 4,
        "marex_hobday_thresholds_hist": 2 * T_out + 4 * 366 + 4,
        "marex_hobday_thresholds_exact_f32": 4 * T_out + 4 * 366,
        "marex_transpose_f32": 2 * 4 * 366,
        "marex_compare_hobday": 4 * T_out + T_out + 4 * 366,
        "marex_compare_hobday_bins": 2 * T_out + T_out + 4 * 366,
        "marex_detrend_coef_f64": 4 * T + 16,
        "marex_detrend_apply_f32": 8 * T + 16,
        "marex_doy_climatology_f32": 4 * T + 4 * 366,
        "marex_sub_doy_climatology_f32": 8 * T + 4 * 366,
        "marex_global_threshold_hist_f64": 4 * T + 8,
        "marex_global_threshold_hist_fast_f64": 4 * T + 8,
        "marex_global_threshold_exact_f64": 4 * T + 8,
        "marex_compare_global": 5 * T + 8,
    }


STAGE_KERNELS = {
    "marex_shift_anomaly_daily_f32": ["shift_daily_kernel<1, 4, 2, MODE, float, false, S = 21, W = 15, D = 48> (run-time windows otherwise)"],
    "marex_digitize_doy_f32": ["digitize_doy_kernel"],
    "marex_hobday_thresholds_pooled_bins": ["init_band_kernel", "hobday_band_kernel<P, 64, 16, NY> (NY = 15 / 25 folded, else run-time)", "hobday_band_kernel<P, 128, 16> (retry list)", "hobday_pool_tile_kernel (fall-back list)"],
    "marex_hobday_thresholds_exact_f32": ["hobday_exact_queue_kernel<64>", "hobday_exact_win_kernel (list fall-back)"],
    "marex_compare_hobday": ["compare_doy_kernel"],
    "marex_compare_hobday_bins": ["compare_bins_kernel"],
}


def ncu_traffic(workload, stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over ALL kernels the stage launches, from the
    ncu --set full captures of this very workload summarised in profiles/r02_ncu_traffic.json ({workload: {stage:
    bytes}}); None when no capture exists for the workload or the stage."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(workload, {}).get(stage)


def host_memory_available():
    """Bytes of host memory this process tree may still use: the cgroup limit when there is one
    (a container's limit is what the OOM killer enforces, not the machine's free memory)."""
    avail = None
    try:
        import psutil

        avail = float(psutil.virtual_memory().available)
    except Exception:
        pass
    for lim_p, cur_p in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_p).read().strip()
            if lim.isdigit():
                room = float(int(lim) - int(open(cur_p).read().strip()))
                avail = room if avail is None else min(avail, room)
        except Exception:
            pass
    return avail if avail is not None else 64e9


def gpu_cpu_affinity(index):
    """CPUs NVML reports as local to GPU `index` (its NUMA node), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        return cpus or None
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU side: oracle on a bounded sample, one tile per process
# ------------------------------------------------------------------------------------------
def _oracle_tile(args):
    x, time, kw = args
    from oracle import marex_oracle as mo

    r = mo.preprocess(x, time, **kw)
    return int(r["extreme_events"].sum())


def cpu_oracle_run(tiles, time, kw, cores):
    t0 = _time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_oracle_tile, [(t, time, kw) for t in tiles])
    return _time.perf_counter() - t0


def sample_tile_origins(ny, nx, n_tiles, th, tw, unstructured):
    """Where the CPU arms cut their tiles from the benchmark field: spread over the rows and columns, land included."""
    out = []
    for i in range(n_tiles):
        if unstructured:
            out.append((0, (i * 48 * 7) % (nx - 48)))
        else:
            out.append((((ny // 2) // th * th + (i // 32) * 3 * th) % (ny - th), (i * 5 * tw) % (nx - tw)))
    return out


def bench_config(args, ny, nx, T, T_out, halo, kw):
    cfg = {
        "workload": args.workload,
        "per_gpu_grid": [ny, nx],
        "days": T,
        "days_out": T_out,
        "halo_rows": halo,
        "l2": "inputs (60 GB per GPU at 0.25 deg) are far larger than L2; no flush needed",
    }
    env = {k: v for k, v in os.environ.items() if k.startswith("MAREX_")}
    if env:
        cfg["tuning_env"] = env
    cfg.update(kw)
    return cfg


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, uuid):
        self.uuid, self.proc, self.path = uuid, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"]
            if self.uuid:
                cmd += ["-i", self.uuid]
            self.proc = subprocess.Popen(cmd, stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of marEx.preprocess_data (the reference itself needs xarray / dask / flox /
    xhistogram, absent from this image) on all host cores, on tiles cut from the SAME synthetic field as the GPU arm
    (numpy twin of the counter-based generator: identical land blobs and random streams), same workload and config keys.
    Imports neither torch nor the CUDA library."""
    if rank != 0:
        return
    from marex_b200 import calendar as mcal
    from marex_b200 import synthetic

    ny, nx, start, end, kw = WORKLOADS[args.workload]
    time = np.arange(np.datetime64(start), np.datetime64(end))
    T = len(time)
    cal = mcal.build_calendar(time)
    shifting = kw.get("method_anomaly", "shifting_baseline") == "shifting_baseline"
    T_out = int((cal.year >= cal.year_val[0] + kw.get("window_year_baseline", 15)).sum()) if shifting else T
    unstructured = ny == 1
    hobday_pooled = kw.get("method_extreme", "hobday_extreme") == "hobday_extreme" and kw.get("method_percentile", "approximate") == "approximate"
    halo = 2 if (hobday_pooled and not unstructured) else 0
    cores = os.cpu_count() or 1
    th, tw = (1, 48) if unstructured else (6, 8)
    n_tiles = 4 * cores  # four 48-cell tiles per core and step
    ny_g = ny * max(1, args.gpus)
    origins = sample_tile_origins(ny, nx, n_tiles, th, tw, unstructured)
    cells = np.stack([(r0 + np.arange(th))[:, None] * nx + (c0 + np.arange(tw))[None, :] for r0, c0 in origins])
    field = synthetic.synth_sst_numpy(time, (ny_g, nx), cells, seed=2, land_fraction=0.0 if unstructured else 0.3)
    tiles = [np.ascontiguousarray(field[:, i, 0] if unstructured else field[:, i]) for i in range(n_tiles)]
    n_cells = n_tiles * th * tw
    land = float(np.mean([np.isnan(t[0]).mean() for t in tiles]))
    for _ in range(args.warmup):
        cpu_oracle_run(tiles[: max(1, cores // 4)], time, kw, cores)
    t0 = _time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_run(tiles, time, kw, cores)
    dt = (_time.perf_counter() - t0) / args.steps
    value = n_cells * T / dt
    sample = (f"{n_tiles} tiles of {th}x{tw} cells x {T} days per step, cut from the benchmark field "
              f"(land fraction of the sample {land:.2f}; each tile its own periodic domain)")
    print(
        json.dumps(
            {
                "impl": "reference",
                "metric": "gridpoint-days/s",
                "value": value,
                "unit": "gridpoint-days/s",
                "n_gpus": args.gpus,
                "steps": args.steps,
                "warmup": args.warmup,
                "ms_per_step": dt * 1e3,
                "higher_is_better": True,
                "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f32",
                "data": "synthetic",
                "config": bench_config(args, ny, nx, T, T_out, halo, kw),
                "cpu_baseline": {
                    "value": value,
                    "unit": "gridpoint-days/s",
                    "cores": cores,
                    "kind": "port",
                    "sample": sample,
                    "note": "numpy oracle, not the dask reference (xarray/dask/flox/xhistogram not installed)",
                    "published_per_core": PUBLISHED_PER_CORE,
                },
                "e2e": {"value": value, "unit": "gridpoint-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
        ),
        flush=True,
    )


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="0.25deg_40yr_shifting_hobday_approx", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=None, help="spatial pieces of the streamed host path (default: the library's own rule)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run tile check against the oracle")
    ap.add_argument("--gather-thresholds", action="store_true", help="N > 1: also gather the thresholds on rank 0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    # stdout carries exactly one JSON line: whatever libraries print there while we run (NCCL's version
    # banner, for one) is sent to stderr, and the real stdout is restored for the final line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import marex_b200
    from marex_b200 import _lib, sharding, synthetic

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    # NUMA: run (and first-touch the pinned host buffers of the e2e leg) on the CPUs next to this GPU's PCIe root
    all_cpus = os.sched_getaffinity(0)
    near_cpus = gpu_cpu_affinity(local)
    if near_cpus:
        try:
            os.sched_setaffinity(0, near_cpus & all_cpus or all_cpus)
        except Exception:
            near_cpus = None
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # NCCL_DEBUG output goes to stderr with everything else (dup2 above)
    ny, nx, start, end, kw = WORKLOADS[args.workload]
    time = np.arange(np.datetime64(start), np.datetime64(end))
    T = len(time)
    W = kw.get("window_year_baseline", 15)
    cal = marex_b200.detect.build_calendar(time)
    shifting = kw.get("method_anomaly", "shifting_baseline") == "shifting_baseline"
    hobday = kw.get("method_extreme", "hobday_extreme") == "hobday_extreme"
    T_out = int((cal.year >= cal.year_val[0] + W).sum()) if shifting else T
    unstructured = ny == 1

    # this rank's latitude band of the global (ny * world) x nx grid, with its pooling halo
    # (unstructured: a contiguous range of cells, no halo)
    halo = 0 if unstructured else sharding.effective_halo(
        kw.get("method_extreme", "hobday_extreme"), kw.get("method_percentile", "approximate"), None, True)
    ny_g = ny * world
    if unstructured:
        own_lo, own_hi, lo, hi = 0, 1, 0, 1
        x = synthetic.synth_sst(time, (world, nx), rows=(rank, rank + 1), seed=2, land_fraction=0.0, device=dev).reshape(T, nx)
    else:
        own_lo, own_hi, lo, hi = sharding.lat_band(ny_g, world, rank, halo)
        x = synthetic.synth_sst(time, (ny_g, nx), rows=(lo, hi), seed=2, device=dev)
    torch.cuda.synchronize()
    n_own = (own_hi - own_lo) * nx

    marks = []
    _lib.TRACE = lambda name: marks.append((name, _ev(torch)))

    pending = []  # NCCL work of the previous step (its results stay referenced until it has completed)

    def drain():
        for wk, _keep in pending:
            wk.wait()  # the compute stream waits; the host does not
        pending.clear()

    def step():
        res = marex_b200.preprocess_arrays(x, time, output="torch", **kw)
        out = res if unstructured else sharding.crop_owned(res, (own_lo, own_hi), (lo, hi))
        if world > 1:
            # every rank owns the same number of rows (weak scaling): the gathers need no size exchange and run
            # asynchronously on NCCL's stream, overlapping the first kernel of the next step; at most one step's
            # collectives are in flight, and the timed region ends only after the last one has completed
            # The per-shard outputs stay where a sharded writer needs them: only the ocean mask is gathered
            # (to rank 0) and the extreme count all-reduced.  Round 1 all-gathered the thresholds (1.5 GB per rank, 12 GB
            # written into every HBM per step at N = 8) although no rank needs its peers' thresholds: that was the 4 %
            # of lost scaling.  --gather-thresholds assembles them on rank 0.
            drain()
            m8 = out["mask"].to(torch.uint8).contiguous()
            glist = [torch.empty_like(m8) for _ in range(world)] if rank == 0 else None
            w2 = dist.gather(m8, glist, dst=0, async_op=True)
            w3 = dist.all_reduce(out["extreme_count"], async_op=True)
            pending.extend([(w2, (glist, m8)), (w3, out["extreme_count"])])
            if args.gather_thresholds:
                th = out["thresholds"].contiguous()
                tlist = [torch.empty_like(th) for _ in range(world)] if rank == 0 else None
                pending.append((dist.gather(th, tlist, dst=0, async_op=True), (tlist, th)))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r = step()
        del r
    drain()
    barrier()
    launches0 = _lib.launch_count()
    marks.clear()
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    clocks = ClockSampler(uuid)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stage_marks = []
    for _ in range(args.steps):
        marks.append(("step_begin", _ev(torch)))
        r = step()
        count_t = r["extreme_count"]
        del r
        stage_marks.append(list(marks))
        marks.clear()
    drain()
    e1.record()
    barrier()
    n_events = int(count_t)
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = n_own * world * T / (ms * 1e-3)

    # per-stage device time (CUDA events recorded right after each C-ABI call, same stream)
    stage_ms = {}
    for sm_ in stage_marks:
        for (n0, a), (n1, b) in zip(sm_[:-1], sm_[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + a.elapsed_time(b) / args.steps
    n_load = (hi - lo) * nx
    kb = kernel_bytes_per_cell(T, T_out)
    peak, peak_src = measured_peaks()
    dom = max((k for k in stage_ms if k in kb), key=lambda k: stage_ms[k])
    achieved = kb[dom] * n_load / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {
        "kernel": dom,
        "kernels_in_stage": STAGE_KERNELS.get(dom, [dom]),
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": ncu_traffic(args.workload, dom),
        "traffic_ratio": (ncu_traffic(args.workload, dom) / (kb[dom] * n_load)) if ncu_traffic(args.workload, dom) else None,
        "traffic_source": "profiles/r02_ncu_traffic.json: ncu --set full capture of this workload, dram read + write bytes "
                          "summed over every kernel the stage launches (null: no capture for this workload)",
        "algorithmic_bytes_per_launch": kb[dom] * n_load,
        "peak_source": peak_src,
        "ms_per_launch": stage_ms[dom],
    }
    B = algorithmic_bytes_per_cell(T, T_out, hobday)
    pipe_gbs = B * n_own / (ms * 1e-3) / 1e9
    stages = {k: {"ms": v, "GBps": (kb[k] * n_load / (v * 1e-3) / 1e9) if k in kb and v > 0 else None} for k, v in stage_ms.items()}

    # ---- parity: tiles of this very field and result against the oracle (oracle/tile_check.py) ----
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle.tile_check import check_tile

        res = marex_b200.preprocess_arrays(x, time, output="torch", **kw)
        torch.cuda.synchronize()
        n_rows_loaded = x.shape[1] if not unstructured else 1
        lay = res["thresholds_layout"]
        tiles_done, compared, err = 0, 0, None
        n = 12
        if unstructured:
            origins = [(0, 1000), (0, nx // 2 + 17), (0, nx - 48 - 5)]
        else:
            origins = [(n_rows_loaded // 7, nx // 5), (n_rows_loaded // 2 - 3, nx // 2 + 11), (n_rows_loaded - n - 4, nx - n - 9),
                       (n_rows_loaded // 3, nx - n // 2)]  # the last tile straddles the longitude seam
        try:
            for r0, c0 in origins:
                if unstructured:
                    cols = torch.arange(c0, c0 + 48, device=dev)
                    cut = lambda a, lead: a.index_select(lead, cols).cpu().numpy()  # noqa: E731
                    sp = 0
                else:
                    rows_ = slice(r0, r0 + n)
                    cols = torch.arange(c0, c0 + n, device=dev) % nx
                    cut = lambda a, lead: a[(slice(None),) * lead + (rows_,)].index_select(lead + 1, cols).cpu().numpy()  # noqa: E731
                    sp = 0
                thr = res["thresholds"]
                got = dict(
                    dat_anomaly=cut(res["dat_anomaly"], 1), extreme_events=cut(res["extreme_events"], 1).astype(bool),
                    mask=cut(res["mask"], 0).astype(bool), thresholds=cut(thr, 1 if lay == "doy_first" else sp),
                )  # fmt: skip
                compared += check_tile(cut(x, 1), time, got, **kw)
                tiles_done += 1
        except AssertionError as e:  # reported in the JSON line and as a non-zero exit code
            err = str(e).strip().splitlines()[:6]
        parity = {
            "tiles": tiles_done, "ok": err is None, "thresholds_compared": compared,
            "what": "12 x 12 tiles (48-cell runs for unstructured) of the benchmark field and of the result of one extra "
                    "untimed step against the numpy oracle: anomalies within 1e-5 of the field scale, thresholds and "
                    "events from the same anomalies bit for bit on the tile interior; the last tile straddles the seam",
            **({"error": err} if err else {}),
        }  # fmt: skip
        del res
        torch.cuda.empty_cache()

    # ---- end to end: host (pinned) buffers in, host buffers out, through the public array API ----
    e2e = None
    if not args.no_e2e:
        _lib.TRACE = None
        # host memory: this rank needs the pinned input plus the pinned outputs (~1.8x the input); with
        # several ranks on one box the per-rank field is cut to the latitude rows that fit (weak scaling
        # keeps the per-GPU work equal across ranks, so the e2e rate stays comparable; the cut is reported)
        e2e_rows = x.shape[1] if not unstructured else 1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        avail = host_memory_available() / max(1, local_world)
        need = x.numel() * 4 * 1.85
        budget = (0.7 if local_world == 1 else 0.5) * avail  # several ranks pin host memory at the same time
        if need > budget and not unstructured:
            e2e_rows = max(16, int(x.shape[1] * budget / need))
        x_e2e = x if (unstructured or e2e_rows == x.shape[1]) else x[:, :e2e_rows].contiguous()
        e2e_cells = x_e2e[0].numel()  # the host field is processed as a stand-alone periodic domain
        xh = torch.empty(x_e2e.shape, dtype=torch.float32, pin_memory=True)
        xh.copy_(x_e2e)
        torch.cuda.synchronize()
        del x, x_e2e
        torch.cuda.empty_cache()
        e2e_steps = min(args.steps, 2)
        d2h = 0
        for i in range(1 + e2e_steps):
            if i == 1:
                barrier()
                t0 = _time.perf_counter()
            res = marex_b200.preprocess_arrays(xh, time, output="pinned_reuse", chunks=args.e2e_chunks, **kw)
            d2h = sum(res[k].nbytes for k in ("dat_anomaly", "mask", "thresholds", "extreme_events"))
            h2d = int(res.get("h2d_bytes", xh.numel() * 4))
            n_chunks = int(res.get("chunks", 1))
            del res
        barrier()
        dt = (_time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t)
        e2e = {
            "value": e2e_cells * world * T / dt,
            "unit": "gridpoint-days/s",
            "gridpoints_per_rank": int(e2e_cells),
            "h2d_bytes_per_step": h2d,
            "chunks": n_chunks,
            "path": "marex_b200.preprocess_arrays(host pinned array) -> host arrays; latitude-band chunks streamed on 3 CUDA streams",
            "d2h_bytes_per_step": int(d2h),
            "steps": e2e_steps,
            "warmup": 1,
            "ms_per_step": dt * 1e3,
            "timer": "host perf_counter around the public call (copies + kernels + sync), max over ranks",
            "cpu_affinity": "GPU-local NUMA node" if near_cpus else "unchanged",
        }
        x_sample_src = xh
    else:
        x_sample_src = x

    # ---- CPU baseline on rank 0, N = 1 only: oracle on a bounded sample of this very field ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every core of the box
        except Exception:
            pass
        cores = len(all_cpus) or os.cpu_count() or 1
        th, tw = (1, 48) if unstructured else (6, 8)
        tiles = []
        for r0, c0 in sample_tile_origins(x_sample_src.shape[1] if not unstructured else 1, nx, cores * 8, th, tw, unstructured):
            if unstructured:  # ~15-30 s of CPU work on the box's cores
                tiles.append(np.ascontiguousarray(x_sample_src[:, c0 : c0 + tw].cpu().numpy()))
            else:
                tiles.append(np.ascontiguousarray(x_sample_src[:, r0 : r0 + th, c0 : c0 + tw].cpu().numpy()))
        dtc = cpu_oracle_run(tiles, time, kw, cores)
        cells = len(tiles) * th * tw
        cpu = {
            "value": cells * T / dtc,
            "unit": "gridpoint-days/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{len(tiles)} tiles of {th}x{tw} cells x {T} days cut from the benchmark field (each tile its own periodic domain), {dtc:.1f} s",
            "note": "numpy oracle, not the dask reference (xarray/dask/flox/xhistogram not installed)",
            "published_per_core": PUBLISHED_PER_CORE,
        }

    if rank == 0:
        line = {
            "metric": "gridpoint-days/s",
            "value": value,
            "unit": "gridpoint-days/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": bench_config(args, ny, nx, T, T_out, halo, kw),
            "roofline": roofline,
            "pipeline_roofline": {
                "bytes_per_gridpoint": B,
                "achieved": pipe_gbs,
                "peak": peak,
                "unit": "GB/s",
                "frac": pipe_gbs / peak,
                "peak_source": peak_src,
            },
            "stages": stages,
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "extreme_events": n_events,
            "clocks": clk,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        sys.exit(3)


def _ev(torch):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


if __name__ == "__main__":
    main()
