"""Host-side calendar tables for the kernels (reference: ``.dt.year`` / ``.dt.dayofyear``
detect.py:1605-1606, ``add_decimal_year`` detect.py:2031-2058, trim detect.py:615-641)."""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

NDOY = 366


_MONTH_DAYS = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
MODEL_CALENDARS = {"noleap": 365, "365_day": 365, "all_leap": 366, "366_day": 366, "360_day": 360}


@dataclass
class ModelTime:
    """A time axis on a fixed-length model calendar (CF ``noleap`` / ``365_day``, ``all_leap`` / ``366_day``,
    ``360_day``) -- SURVEY.md 8f row 4.  numpy has no datetime type for these, so the axis is carried as
    ``year`` / ``doy`` (what the reference reads through ``.dt.year`` / ``.dt.dayofyear``, detect.py:1605-1606,
    on a cftime index) plus the length of the calendar year."""

    year: np.ndarray  # int32 (T,)
    doy: np.ndarray  # int16 (T,) 1..days_in_year
    days_in_year: int
    label: np.ndarray  # (T,) the caller's own time values (returned, trimmed, as ``time``)

    def __len__(self) -> int:
        return int(self.year.shape[0])

    @property
    def decimal_year(self) -> np.ndarray:
        return self.year.astype(np.float64) + (self.doy.astype(np.float64) - 1.0) / float(self.days_in_year)


def _model_doy0(month: int, day: int, calendar: str) -> int:
    """0-based day of year of (month, day) on a model calendar."""
    if MODEL_CALENDARS[calendar] == 360:
        return (month - 1) * 30 + (day - 1)
    md = _MONTH_DAYS.copy()
    if MODEL_CALENDARS[calendar] == 366:
        md[1] = 29
    return int(md[: month - 1].sum()) + (day - 1)


def model_time_from_cf(values, units: str, calendar: str) -> ModelTime:
    """CF ``"<unit> since YYYY-MM-DD[ hh:mm:ss]"`` numbers on a model calendar -> ``ModelTime`` (whole days;
    sub-daily samples of one day map to the same day and are rejected by ``build_calendar``)."""
    if calendar not in MODEL_CALENDARS:
        raise NotImplementedError(f"calendar '{calendar}' is not a supported model calendar {sorted(MODEL_CALENDARS)}")
    unit, _, epoch = units.partition(" since ")
    per_day = {"days": 1.0, "day": 1.0, "hours": 24.0, "hour": 24.0, "minutes": 1440.0, "seconds": 86400.0, "second": 86400.0}
    if unit.strip() not in per_day or not epoch:
        raise ValueError(f"cannot parse time units '{units}'")
    date = epoch.strip().replace("T", " ").split(" ")[0]
    ey, em, ed = (int(v) for v in date.split("-"))
    L = MODEL_CALENDARS[calendar]
    v = np.asarray(values, dtype=np.float64) / per_day[unit.strip()]
    ordinal = np.floor(v + 1e-9).astype(np.int64) + ey * L + _model_doy0(em, ed, calendar)
    return ModelTime((ordinal // L).astype(np.int32), (ordinal % L + 1).astype(np.int16), L, np.asarray(values))


@dataclass
class Calendar:
    time: np.ndarray  # datetime64[D] (T,)  (ModelTime: the caller's labels)
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

    model_dy: Optional[np.ndarray] = None  # decimal year of a ModelTime axis (None: Gregorian datetime64 axis)

    @property
    def is_daily(self) -> bool:
        """Gap-free daily proleptic-Gregorian axis: every row is the day after the previous one (what the
        TMA-staged shifting-baseline kernel assumes; model calendars, gaps and sub-sampled axes take the generic
        table-driven kernel)."""
        if self.model_dy is not None:
            return False
        return bool(np.all(np.diff(self.time.astype("datetime64[D]").astype(np.int64)) == 1))

    @property
    def decimal_year(self) -> np.ndarray:
        """detect.py:2031-2058 (Gregorian), year + (doy - 1) / days_in_year on a model calendar."""
        return self.model_dy if self.model_dy is not None else decimal_year(self.time)


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
    model_dy = None
    if isinstance(time, ModelTime):
        t, year, doy, model_dy = np.asarray(time.label), time.year.astype(np.int32), time.doy.astype(np.int16), time.decimal_year
        if year.ndim != 1 or year.size == 0 or doy.shape != year.shape or t.shape != year.shape:
            raise ValueError("ModelTime needs matching non-empty 1-D year / doy / label arrays")
        if doy.min() < 1 or doy.max() > NDOY:
            raise ValueError("day of year out of range 1..366")
    else:
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
    return Calendar(t, year, doy, year_val.astype(np.int32), tidx, model_dy)


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


def doy_slots(doy: np.ndarray, year: Optional[np.ndarray] = None) -> Tuple[int, np.ndarray]:
    """Day-of-year-major slot table of an output time axis: ``NY`` slots per day of year and
    ``slot_row[(doy - 1) * NY + k]`` = row of (doy, k) or -1.  With ``year`` given, k = year - first year (one slot
    per calendar year, the order the fused shifting-baseline kernel writes its bin codes in); without it,
    k = the running count of earlier rows with the same day of year."""
    d = np.asarray(doy).astype(np.int64) - 1
    if year is not None:
        y = np.asarray(year).astype(np.int64)
        k = y - y.min()
    else:
        order = np.argsort(d, kind="stable")
        first = np.zeros(NDOY + 1, dtype=np.int64)
        first[1:] = np.cumsum(np.bincount(d, minlength=NDOY))
        k = np.empty(d.size, dtype=np.int64)
        k[order] = np.arange(d.size) - first[d[order]]
    ny = int(k.max()) + 1
    slot = d * ny + k
    if np.unique(slot).size != slot.size:
        raise NotImplementedError("more than one time step per (year, dayofyear)")
    slot_row = np.full(NDOY * ny, -1, dtype=np.int32)
    slot_row[slot] = np.arange(d.size, dtype=np.int32)
    return ny, slot_row


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
