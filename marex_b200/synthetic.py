"""Synthetic SST for benchmarks: generated on the device by a counter-based
kernel so any shard of a global grid can be regenerated independently (no 60 GB host staging)."""
from typing import Tuple

import numpy as np
import torch

from . import _lib
from .calendar import decimal_year
from .detect import _device, _p, _stream


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
    dev = _device(device)
    ny_g, nx_g = grid_global
    r0, r1 = rows if rows is not None else (0, ny_g)
    T, N = len(time), (r1 - r0) * nx_g
    x = torch.empty((T, r1 - r0, nx_g), dtype=torch.float32, device=dev)
    dy = torch.from_numpy(decimal_year(time).astype(np.float32)).to(dev)
    _lib.call("marex_synth_sst_f32", _p(x), T, N, N, r0 * nx_g, ny_g, nx_g, _p(dy), int(seed), float(land_fraction), _stream())
    return x
