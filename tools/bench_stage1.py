"""Time tracker stage 1 (marex_b200.track.MaskFiller.run) kernel by kernel on one GPU.

    python tools/bench_stage1.py [--days 2048] [--ny 720 --nx 1440] [--R 8] [--T-fill 2] [--reps 5] [--out FILE.json]

The events are synthetic blobs (smoothed uniform noise thresholded at `density`) generated on the device and handed to
stage 1 as the flattened BIT mask that marex_compare_* writes.  Every C-ABI call is timed with CUDA events on the
launching stream (after one warm-up pass); the JSON line lists ms and algorithmic GB/s (bytes each pass must read +
write once) per kernel, and gridpoint-days/s of the whole stage.  A 40-year 0.25-degree run has 9,131 output days: the
stage is linear in the number of days, so `--days` only has to be large enough to fill the GPU."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_bits(T, ny, nx, density, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    N = ny * nx
    nw = (N + 31) // 32
    bits = torch.empty((T, nw), dtype=torch.int32, device=dev)
    weights = (1 << torch.arange(32, device=dev, dtype=torch.int64)).view(1, 1, 32)
    for t0 in range(0, T, 64):
        t1 = min(T, t0 + 64)
        raw = torch.rand((t1 - t0, 1, ny, nx), device=dev, generator=g)
        sm = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(raw, (3, 3, 3, 3), mode="circular"), 7, stride=1)
        thr = torch.quantile(sm[0].flatten()[:: max(1, N // 100000)], 1 - density)
        ev = (sm > thr).reshape(t1 - t0, N)
        if nw * 32 != N:
            ev = torch.nn.functional.pad(ev, (0, nw * 32 - N))
        words = (ev.view(t1 - t0, nw, 32).to(torch.int64) * weights).sum(-1)
        bits[t0:t1] = torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)  # two's complement
    return bits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=720)
    ap.add_argument("--nx", type=int, default=1440)
    ap.add_argument("--R", type=int, default=8)
    ap.add_argument("--T-fill", type=int, default=2)
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--input", default="bits", choices=["bits", "bool"], help="hand stage 1 the packed mask or the bool array")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = torch.device("cuda", 0)
    from marex_b200 import _lib, track

    T, ny, nx, R = a.days, a.ny, a.nx, a.R
    N = ny * nx
    mask = np.ones((ny, nx), bool)
    mask[: ny // 12] = False  # a polar cap of land
    mask[ny // 3 : ny // 2, nx // 5 : nx // 3] = False
    bits = synthetic_bits(T, ny, nx, a.density, dev)
    if a.input == "bool":  # the reference's layout: one byte per cell
        shifts = torch.arange(32, device=dev, dtype=torch.int32).view(1, 1, 32)
        events = torch.empty((T, ny, nx), dtype=torch.bool, device=dev)
        for t0 in range(0, T, 64):
            w = bits[t0 : t0 + 64]
            events[t0 : t0 + 64] = (((w.unsqueeze(-1) >> shifts) & 1) != 0).reshape(w.shape[0], -1)[:, :N].reshape(-1, ny, nx)
    run_kw = dict(from_bits=(bits, T)) if a.input == "bits" else dict(data_bin=events)

    # algorithmic bytes per call (read once + write once), from the arguments of the call itself
    def call_bytes(name, args):
        v = [x.value if hasattr(x, "value") else x for x in args]
        if name == "marex_morph_pad_bits":
            T_, ny_, nx_, pad = v[6], v[7], v[8], v[9]
            out = T_ * (ny_ + 2 * pad) * ((nx_ + 2 * pad + 31) // 32) * 4
            return T_ * ny_ * nx_ / 8 + out
        if name in ("marex_morph_disk", "marex_morph_disk_sep"):
            return 2 * v[2] * v[3] * ((v[4] + 31) // 32) * 4
        if name == "marex_morph_time":
            return (v[1] + v[4]) * v[2] * 4
        if name == "marex_morph_pack_u8":
            return v[1] * v[2] + v[1] * ((v[2] + 31) // 32) * 4
        if name == "marex_morph_extract":
            T_, ny_, nx_ = v[6], v[7], v[8]
            out = (T_ * ny_ * nx_ if v[9] else 0) + (T_ * ((ny_ * nx_ + 31) // 32) * 4 if v[11] else 0)
            return T_ * ny_ * nx_ / 8 + out
        return 0

    records = []
    real_call = track._call

    def timed_call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(name, *args)
        e1.record()
        records.append((name, args[6] if name.startswith("marex_morph_disk") else None, call_bytes(name, args), e0, e1))

    out = {}
    for label, packed, separable in (("direct_packed_out", True, False), ("separable_packed_out", True, True),
                                     ("separable_bool_out", False, True)):
        f = track.MaskFiller(mask, R, a.T_fill, device=dev, separable=separable)
        f.run(packed=packed, **run_kw)  # warm-up (allocator, first launches)
        torch.cuda.synchronize()
        track._call = timed_call
        whole = []
        per_kernel = {}
        for _ in range(a.reps):
            records.clear()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            f.run(packed=packed, **run_kw)
            s1.record()
            torch.cuda.synchronize()
            whole.append(s0.elapsed_time(s1))
            for i, (name, erode, nbytes, e0, e1) in enumerate(records):
                key = f"{i:02d}_{name}" + ("" if erode is None else ("_erode" if erode else "_dilate"))
                per_kernel.setdefault(key, {"ms": [], "bytes": nbytes})["ms"].append(e0.elapsed_time(e1))
        track._call = real_call
        ms = float(np.median(whole))
        out[label] = {
            "ms": ms,
            "gridpoint_days_per_s": N * T / (ms * 1e-3),
            "true_cells": f.last_count,
            "calls": {k: {"ms": float(np.median(v["ms"])), "GBps": v["bytes"] / (np.median(v["ms"]) * 1e-3) / 1e9}
                      for k, v in per_kernel.items()},  # fmt: skip
        }
    line = {
        "metric": "gridpoint-days/s of tracker stage 1 (fill_holes + fill_time_gaps) on the bit-packed mask",
        "config": {"days": T, "grid": [ny, nx], "R_fill": R, "T_fill": a.T_fill, "density": a.density, "reps": a.reps,
                   "input": "flattened bit mask on the device (the layout marex_compare_* writes)" if a.input == "bits" else "bool bytes on the device",
                   "env": {k: v for k, v in os.environ.items() if k.startswith("MAREX_")}},  # fmt: skip
        "launches": int(_lib.launch_count()),
        "results": out,
    }
    s = json.dumps(line)
    print(s)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as fh:
            fh.write(json.dumps(line, indent=1))


if __name__ == "__main__":
    main()
