"""Synthetic SST for benchmarks: generated on the device by a counter-based
kernel so any shard of a global grid can be regenerated independently (no 60 GB host staging)."""
from typing import Tuple

import numpy as np
import torch

from .calendar import decimal_year


def daily_time_axis(start: str, end: str) -> np.ndarray:
    return np.arange(np.datetime64(start), np.datetime64(end))


def synth_sst(
    time: np.ndarray,
    grid_global: Tuple[int, int],
    rows: Tuple[int, int] = None,
    seed: int = 2,
    land_fraction: float = 0.3,
    device=None,
) -> torch.Tensor:
    """float32 CUDA tensor (T, n_rows, nx) holding rows [rows[0], rows[1]) of a global
    (ny_global, nx) field: seasonal cycle + 0.02 K/yr trend + AR(1) noise, NaN land blobs."""
    from . import _lib
    from .detect import _device, _p, _stream

    dev = _device(device)
    ny_g, nx_g = grid_global
    r0, r1 = rows if rows is not None else (0, ny_g)
    T, N = len(time), (r1 - r0) * nx_g
    x = torch.empty((T, r1 - r0, nx_g), dtype=torch.float32, device=dev)
    dy = torch.from_numpy(decimal_year(time).astype(np.float32)).to(dev)
    _lib.call("marex_synth_sst_f32", _p(x), T, N, N, r0 * nx_g, ny_g, nx_g, _p(dy), int(seed), float(land_fraction), _stream())
    return x


# ---- numpy twin (CPU arms of bench.py: the reference arm must not touch the CUDA library) ----------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _u01(h: np.ndarray) -> np.ndarray:
    return ((h >> np.uint64(40)) + np.uint64(1)).astype(np.float32) * np.float32(1.0 / 16777217.0)


def synth_sst_numpy(time: np.ndarray, grid_global: Tuple[int, int], cells: np.ndarray, seed: int = 2,
                    land_fraction: float = 0.3) -> np.ndarray:
    """The field of ``synth_sst`` at the flattened global gridpoints ``cells`` (any shape), on the CPU: the same
    counter-based generator (identical land blobs, amplitudes, phases and random streams; values agree with the CUDA
    kernel to float32 rounding of its fast-math intrinsics).  Returns float32 (T,) + cells.shape."""
    ny_g, nx_g = grid_global
    cg = np.asarray(cells, dtype=np.uint64)
    shape = cg.shape
    cg = cg.reshape(-1)
    with np.errstate(over="ignore"):
        iy, ix = cg // np.uint64(nx_g), cg % np.uint64(nx_g)
        sd = np.uint64(seed)
        hb = _splitmix64(sd ^ ((iy >> np.uint64(3)) * np.uint64(1315423911) + (ix >> np.uint64(3))))
        land = _u01(hb) < np.float32(land_fraction)
        hc = _splitmix64(sd * np.uint64(0x5851F42D4C957F2D) + cg)
        latf = (iy.astype(np.float32) / np.float32(ny_g - 1)) if ny_g > 1 else np.full(cg.shape, 0.5, np.float32)
        mu = np.float32(-1.8) + np.float32(31.8) * (np.float32(1) - np.abs(np.float32(2) * latf - np.float32(1)))
        amp = np.float32(0.5) + np.float32(5.5) * _u01(hc)
        phase = _u01(_splitmix64(hc))
        rho = np.float32(0.9)
        sig = np.float32(0.6) * np.sqrt(np.float32(1) - rho * rho)
        state = _splitmix64(hc ^ np.uint64(0xD1B54A32D192ED03))
        dy = decimal_year(time).astype(np.float32)
        out = np.empty((len(time), cg.size), dtype=np.float32)
        ar = np.zeros(cg.size, dtype=np.float32)
        two_pi = np.float32(6.2831853)
        for t in range(len(time)):
            state = _splitmix64(state)
            u1, u2 = _u01(state), _u01(state * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))
            z = np.sqrt(np.float32(-2) * np.log(u1)) * np.cos(two_pi * u2)
            ar = rho * ar + sig * z
            out[t] = mu + amp * np.cos(two_pi * (dy[t] - np.floor(dy[t]) - phase)) + np.float32(0.02) * (dy[t] - dy[0]) + ar
    out[:, land] = np.nan
    return out.reshape((len(time),) + shape)
