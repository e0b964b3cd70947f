"""Host-side calendar tables for the kernels (reference: ``.dt.year`` / ``.dt.dayofyear``
detect.py:1605-1606, ``add_decimal_year`` detect.py:2031-2058, trim detect.py:615-641)."""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

NDOY = 366


@dataclass
class Calendar:
    time: np.ndarray  # datetime64[D] (T,)
    year: np.ndarray  # int32 (T,)
    doy: np.ndarray  # int16 (T,) 1..366
    year_val: np.ndarray  # int32 (n_years,) ascending distinct years
    tidx: np.ndarray  # int32 (n_years*366,) row of (year i, doy d) or -1

    @property
    def T(self) -> int:
        return int(self.time.shape[0])

    @property
    def n_years(self) -> int:
        return int(self.year_val.shape[0])

    @property
    def is_daily(self) -> bool:
        """Gap-free daily axis: every row is the day after the previous one (what the TMA-staged
        shifting-baseline kernel assumes; anything else takes the generic table-driven kernel)."""
        return bool(np.all(np.diff(self.time.astype("datetime64[D]").astype(np.int64)) == 1))


def year_doy(time) -> Tuple[np.ndarray, np.ndarray]:
    t = np.asarray(time).astype("datetime64[D]")
    y = t.astype("datetime64[Y]")
    year = (y.astype(np.int64) + 1970).astype(np.int32)
    doy = ((t - y.astype("datetime64[D]")).astype(np.int64) + 1).astype(np.int16)
    return year, doy


def decimal_year(time) -> np.ndarray:
    """year + elapsed_days / days_in_year in float64 (detect.py:2051-2057)."""
    t = np.asarray(time).astype("datetime64[D]")
    y = t.astype("datetime64[Y]")
    start = y.astype("datetime64[D]")
    nxt = (y + 1).astype("datetime64[D]")
    return (y.astype(np.int64) + 1970) + (t - start).astype(np.int64) / (nxt - start).astype(np.int64)


def build_calendar(time) -> Calendar:
    t = np.asarray(time).astype("datetime64[D]")
    if t.ndim != 1 or t.size == 0:
        raise ValueError("time must be a non-empty 1-D datetime64 array")
    year, doy = year_doy(t)
    year_val, yi = np.unique(year, return_inverse=True)
    key = yi.astype(np.int64) * NDOY + (doy.astype(np.int64) - 1)
    if np.unique(key).size != key.size:
        raise NotImplementedError(
            "marex_b200 needs at most one sample per (year, dayofyear): sub-daily or duplicated time steps "
            "are not supported by the day-of-year kernels"
        )
    tidx = np.full(year_val.size * NDOY, -1, dtype=np.int32)
    tidx[key] = np.arange(t.size, dtype=np.int32)
    return Calendar(t, year, doy, year_val.astype(np.int32), tidx)


def doy_csr(doy: np.ndarray, rows: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """CSR list (doy_ptr[367], doy_rows) of the rows of each day of year; ``rows`` restricts the
    list to a subset of row indices (the reference_period rows, detect.py:2342-2344)."""
    doy = np.asarray(doy).astype(np.int64)
    idx = np.arange(doy.size, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    d = doy[idx] - 1
    order = np.argsort(d, kind="stable")
    counts = np.bincount(d, minlength=NDOY)
    ptr = np.zeros(NDOY + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(counts)
    return ptr, idx[order].astype(np.int32)


def max_window_rows(doy_ptr: np.ndarray, w: int) -> int:
    """Largest number of rows any +-w//2 day-of-year window (wrap 366) holds."""
    counts = np.diff(doy_ptr.astype(np.int64))
    half = w // 2
    tot = np.zeros(NDOY, dtype=np.int64)
    for k in range(-half, half + 1):
        tot += np.roll(counts, -k)
    return int(tot.max())


def shifting_out_rows(cal: Calendar, W: int) -> Tuple[np.ndarray, np.ndarray]:
    """Rows kept by the shifting-baseline trim: year >= min_year + W (detect.py:638-641).
    Returns (out_row[T] int32 with -1 for dropped rows, keep[T] bool)."""
    keep = cal.year >= int(cal.year_val[0]) + W
    out_row = np.full(cal.T, -1, dtype=np.int32)
    out_row[keep] = np.arange(int(keep.sum()), dtype=np.int32)
    return out_row, keep
