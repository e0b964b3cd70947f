"""
Host-side mirror of ``marEx/detect.py`` for the detection hot path, driving the sm_100a kernels
of ``libmarex_b200.so`` through ctypes.  PyTorch is used only for device buffers, streams and
host<->device copies.  There is no CPU fallback: without a CUDA device or without the built
library every entry point raises.

Two levels:

* array level (``preprocess_arrays``, ``compute_normalised_anomaly_arrays``,
  ``identify_extremes_arrays``, ``rolling_climatology_arrays``): time-major numpy / torch arrays
  plus a ``datetime64`` time axis.  This is the boundary the parity tests exercise.
* xarray level (``preprocess_data`` ... in ``marex_b200/xr_api.py``): the reference's public
  signatures (detect.py:287-313, 891-907, 1119-1133, 1511-1517, 1691-1698), DataArray in and
  Dataset out.

Validation rules, defaults and messages follow the reference line by line; each block cites it.
"""
from __future__ import annotations

import ctypes
import logging
import os
import warnings
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .calendar import NDOY, Calendar, build_calendar, decimal_year, doy_csr, doy_slots, max_window_rows, shifting_out_rows
from .exceptions import ConfigurationError, ProcessingError, create_data_validation_error

logger = logging.getLogger("marex_b200")

VALID_ANOMALY = ("detrend_harmonic", "shifting_baseline", "fixed_baseline", "detrend_fixed_baseline")
VALID_EXTREME = ("global_extreme", "hobday_extreme")


# --------------------------------------------------------------------------------------
# small plumbing helpers
# --------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("marex_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _on_device(pick):
    """Decorator: run the function with the CUDA device of its data made current, so that the library's launches,
    torch's current stream and every scratch allocation agree with the device the pointers live on (a caller may pass
    ``device="cuda:1"`` or tensors of a device that is not the current one).  ``pick(*args, **kwargs)`` returns the
    device (or None for the current device)."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            dev = pick(*args, **kwargs)
            dev = _device(dev) if not isinstance(dev, torch.device) else dev
            if dev.type != "cuda":
                return fn(*args, **kwargs)
            with torch.cuda.device(dev):
                return fn(*args, **kwargs)

        return wrapper

    return deco


def _tensor_device(t):
    return t.device if isinstance(t, torch.Tensor) and t.is_cuda else None


def _up(a: np.ndarray, dtype, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device)


_PINNED: Dict[Any, torch.Tensor] = {}


def _pinned_buffer(key: str, shape, dtype, reuse: bool) -> torch.Tensor:
    """Page-locked host tensor.  ``reuse``: one cached buffer per (key, shape, dtype), overwritten by the next call that
    asks for it (a repeated pipeline call then pays the PCIe transfer, not cudaHostAlloc); otherwise a fresh one."""
    if not reuse:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    k = (key, tuple(shape), dtype)
    buf = _PINNED.get(k)
    if buf is None:
        buf = torch.empty(shape, dtype=dtype, pin_memory=True)
        _PINNED[k] = buf
    return buf


def _to_host(t: torch.Tensor, key: str, output: str):
    """Device -> host for ``output`` "numpy" (pageable), "pinned" (fresh page-locked buffer) or "pinned_reuse"."""
    if output == "numpy":
        return t.cpu()
    buf = _pinned_buffer(key, tuple(t.shape), t.dtype, output == "pinned_reuse")
    buf.copy_(t, non_blocking=True)
    return buf


def release_host_buffers() -> None:
    """Drop the cached page-locked output buffers of ``output="pinned_reuse"`` (one buffer per output name, shape and
    dtype; arrays returned by earlier ``"pinned_reuse"`` calls become invalid)."""
    _PINNED.clear()


class _Hold(list):
    """Keeps uploaded table tensors alive until the call that uses them has been enqueued: a
    temporary freed right after ``data_ptr()`` could be handed out again by the caching
    allocator to the NEXT upload and be overwritten before the kernel launches."""

    def up(self, a: np.ndarray, dtype, device):
        t = _up(a, dtype, device)
        self.append(t)
        return _p(t)


def _to_device_field(x, device) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    """Any (T, ...space) array -> contiguous float32 CUDA tensor (T, N) + the space shape.
    ``da.astype(np.float32)`` of detect.py:600."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    space = tuple(int(s) for s in t.shape[1:])
    t = t.to(device=device, dtype=torch.float32, non_blocking=True).reshape(t.shape[0], -1).contiguous()
    return t, space


# --------------------------------------------------------------------------------------
# configuration validation.  The HEADLINE of every error is the reference's (its tests match them with regular
# expressions, tests/test_error_handling.py); rule order follows detect.py.  Details and hints are this package's own.
# --------------------------------------------------------------------------------------
def _bad_config(headline: str, why: str, *hints: str, **context) -> ConfigurationError:
    return ConfigurationError(headline, details=why, suggestions=list(hints), context=context or None)


def validate_reference_period_method(reference_period, method_anomaly: str) -> None:
    """detect.py:571-579 / 1069-1077."""
    if reference_period is not None and method_anomaly not in ("fixed_baseline", "detrend_fixed_baseline"):
        raise _bad_config(
            f"reference_period is not supported for method_anomaly='{method_anomaly}'",
            "only the two fixed-climatology methods average over a chosen span of years",
            "drop reference_period, or pick method_anomaly='fixed_baseline' / 'detrend_fixed_baseline'",
        )


def validate_anomaly_method(method_anomaly: str) -> None:
    """detect.py:1100-1116."""
    if method_anomaly not in VALID_ANOMALY:
        raise _bad_config(
            f"Unknown anomaly method '{method_anomaly}'",
            "method_anomaly selects the baseline that is removed from the data",
            "choose one of: " + ", ".join(VALID_ANOMALY),
            provided_method=method_anomaly, valid_methods=list(VALID_ANOMALY),
        )  # fmt: skip


def validate_detrend_orders(detrend_orders: Sequence[int]) -> None:
    """detect.py:2104-2126."""
    if not detrend_orders:
        raise _bad_config(
            "detrend_orders cannot be empty",
            "the trend model needs at least one polynomial term",
            "detrend_orders=[1] removes a linear trend (the default); [1, 2] adds a quadratic term",
        )
    invalid = [order for order in detrend_orders if order < 1]
    if invalid:
        raise _bad_config(
            f"Invalid polynomial orders: {invalid}",
            "orders are exponents of the centred decimal year and start at 1 (the constant is always fitted)",
            f"remove {invalid}",
        )


def validate_reference_period(reference_period, year: np.ndarray) -> np.ndarray:
    """detect.py:2334-2355: returns the row indices of the reference period."""
    start_year, end_year = reference_period
    if start_year > end_year:
        raise _bad_config(
            f"Invalid reference_period: start year ({start_year}) must be <= end year ({end_year})",
            "reference_period is (first year, last year), both inclusive",
            f"did you mean ({end_year}, {start_year})?",
        )
    rows = np.nonzero((year >= start_year) & (year <= end_year))[0]
    if rows.size == 0:
        lo, hi = int(year.min()), int(year.max())
        raise _bad_config(
            f"No data found in reference_period ({start_year}, {end_year})",
            f"the time axis covers {lo}-{hi}",
            f"choose years inside {lo}-{hi}, or reference_period=None for the whole series",
        )
    return rows


def resolve_extreme_config(
    method_extreme: str,
    threshold_percentile: float,
    window_days_hobday: Optional[int],
    window_spatial_hobday: Optional[int],
    method_percentile: str,
    precision: float,
    max_anomaly: float,
    gridded: bool,
    dimensions: Optional[Dict[str, str]] = None,
    available_dims: Optional[List[str]] = None,
) -> Optional[int]:
    """The configuration rules of ``identify_extremes`` (detect.py:1277-1503), in the reference's
    order.  Returns the effective ``window_spatial_hobday`` (5 by default on gridded data,
    detect.py:1450-1452)."""
    if method_percentile not in ("exact", "approximate"):  # :1281
        raise _bad_config(
            f"Unknown method_percentile '{method_percentile}'",
            "thresholds come either from a selection of the samples or from a histogram",
            "method_percentile='approximate' (histogram, the default) or 'exact'",
            provided_method=method_percentile, valid_methods=["exact", "approximate"],
        )  # fmt: skip
    if method_percentile == "exact":  # :1302, :1322 -- histogram parameters make no sense without a histogram
        for name, value, default in (("precision", precision, 0.01), ("max_anomaly", max_anomaly, 5.0)):
            if value != default:
                raise _bad_config(
                    f"Parameter '{name}' cannot be used with method_percentile='exact'",
                    f"{name}={value} shapes the histogram of the approximate method; the exact method has none",
                    f"leave {name} at its default, or switch to method_percentile='approximate'",
                    **{"method_percentile": method_percentile, f"provided_{name}": value, f"default_{name}": default},
                )
    if threshold_percentile < 60 and method_percentile == "approximate":  # :1342
        raise _bad_config(
            f"Percentile threshold {threshold_percentile}% is not supported with method_percentile='approximate'",
            "the histogram lumps everything below -precision into one bin, so quantiles in the lower half are not resolved",
            "use method_percentile='exact' below the 60th percentile",
            threshold_percentile=threshold_percentile, method_percentile=method_percentile, min_supported_percentile=60,
        )  # fmt: skip
    if window_spatial_hobday is not None:
        where = None
        if not gridded:  # :1367
            where = ("for unstructured grids", "pooling needs the (y, x) neighbourhood of a structured grid",
                     dict(grid_type="unstructured", dimensions=dimensions, available_dims=available_dims))  # fmt: skip
            headline = "window_spatial_hobday is not supported for unstructured grids"
        elif method_extreme != "hobday_extreme":  # :1390
            where = ("", "only the day-of-year histograms of hobday_extreme are pooled",
                     dict(method_extreme=method_extreme, compatible_methods=["hobday_extreme"]))  # fmt: skip
            headline = "window_spatial_hobday can only be used with method_extreme='hobday_extreme'"
        elif method_percentile == "exact":  # :1412
            where = ("", "the exact percentile selects among a gridpoint's own samples and pools nothing",
                     dict(method_percentile=method_percentile, compatible_methods=["approximate"]))  # fmt: skip
            headline = "window_spatial_hobday is not supported with method_percentile='exact'"
        if where is not None:
            raise _bad_config(headline, where[1], "pass window_spatial_hobday=None",
                              window_spatial_hobday=window_spatial_hobday, **where[2])  # fmt: skip
    if method_extreme == "hobday_extreme" and window_days_hobday is not None and window_days_hobday % 2 == 0:  # :1434
        raise _bad_config(
            "window_days_hobday must be an odd number",
            f"the window is centred on a day, so it spans 2k + 1 days; got {window_days_hobday}",
            f"use {window_days_hobday - 1} or {window_days_hobday + 1}",
            window_days_hobday=window_days_hobday, is_odd=False,
        )  # fmt: skip
    if method_extreme == "hobday_extreme" and window_spatial_hobday is None and gridded:
        window_spatial_hobday = 5  # detect.py:1451-1452
    if method_extreme == "hobday_extreme" and window_spatial_hobday is not None and window_spatial_hobday % 2 == 0:  # :1456
        raise _bad_config(
            "window_spatial_hobday must be an odd number",
            f"the pooling window is centred on a gridpoint, so it spans 2k + 1 cells; got {window_spatial_hobday}",
            f"use {window_spatial_hobday - 1} or {window_spatial_hobday + 1}",
            window_spatial_hobday=window_spatial_hobday, is_odd=False,
        )  # fmt: skip
    if method_extreme not in VALID_EXTREME:  # :1492
        raise _bad_config(
            f"Unknown extreme method '{method_extreme}'",
            "method_extreme selects constant (global_extreme) or day-of-year (hobday_extreme) thresholds",
            "choose one of: " + ", ".join(VALID_EXTREME),
            provided_method=method_extreme, valid_methods=list(VALID_EXTREME),
        )  # fmt: skip
    return window_spatial_hobday


def get_preprocessing_steps(
    method_anomaly: str,
    method_extreme: str,
    std_normalise: bool,
    detrend_orders: List[int],
    window_year_baseline: int,
    smooth_days_baseline: int,
    window_days_hobday: int,
    window_spatial_hobday: Optional[int],
    reference_period: Optional[Tuple[int, int]] = None,
) -> List[str]:
    """``_get_preprocessing_steps`` (detect.py:844-888): the attrs strings, pinned by
    tests/golden/ref_preprocessing_steps.json."""
    steps = []
    if method_anomaly == "detrend_harmonic":
        steps.append(f"Removed polynomial trend orders={detrend_orders} & seasonal cycle")
        if std_normalise:
            steps.append("Normalised by 30-day rolling STD")
    elif method_anomaly == "shifting_baseline":
        steps.append(f"Rolling climatology using {window_year_baseline} years")
        steps.append(f"Smoothed with {smooth_days_baseline}-day window")
    elif method_anomaly == "fixed_baseline":
        if reference_period is not None:
            steps.append(f"Daily climatology computed from {reference_period[0]}-{reference_period[1]}")
        else:
            steps.append("Daily climatology computed from full time series")
    elif method_anomaly == "detrend_fixed_baseline":
        steps.append(f"Removed polynomial trend orders={detrend_orders}")
        if reference_period is not None:
            steps.append(f"Daily climatology computed from detrended data ({reference_period[0]}-{reference_period[1]})")
        else:
            steps.append("Daily climatology computed from detrended data")
    if method_extreme == "global_extreme":
        steps.append("Global percentile threshold applied to all days")
    elif method_extreme == "hobday_extreme":
        if window_spatial_hobday is not None:
            steps.append(
                f"Day-of-year thresholds with {window_days_hobday} day window & {window_spatial_hobday} spatial neighbours"
            )
        else:
            steps.append(f"Day-of-year thresholds with {window_days_hobday} day window")
    return steps


def _raise_no_finite_data(total_values: int, n_locations: int):
    """detect.py:226-237."""
    raise create_data_validation_error(
        "Dataset contains no valid (finite) data",
        details="every value of the first time step is NaN or infinite, so no gridpoint counts as ocean",
        suggestions=["check how the field was read (fill values, units, a mask applied twice)"],
        data_info={"total_values": int(total_values), "total_spatial_locations": int(n_locations)},
    )


def _raise_invalid_ocean_values(total_invalid: int, affected: int, ocean: int, worst: int, T: int):
    """detect.py:258-279: gridpoints that are finite on the first time step must be finite throughout."""
    raise create_data_validation_error(
        f"Dataset contains {total_invalid} invalid values in {affected} ocean locations",
        details=f"Found invalid data across time series. Worst location has {worst} invalid time steps out of {T}.",
        suggestions=[
            "interpolate or fill the gaps before preprocessing, or mask those gridpoints on every time step",
            "gridpoints that are NaN on the first time step are treated as land and skipped",
        ],
        data_info={
            "total_invalid_values_in_ocean": total_invalid,
            "locations_affected": affected,
            "total_ocean_locations": ocean,
            "max_invalid_at_one_location": worst,
            "total_time_steps": int(T),
            "percentage_affected": f"{100.0 * affected / ocean:.2f}%",
        },
    )


def check_data_values(mask0: torch.Tensor, nonfinite: torch.Tensor, T: int, total_values: int) -> None:
    """``_validate_data_values`` (detect.py:205-279) from the per-cell numbers the first anomaly
    kernel produced in the same pass that read the data; one device->host transfer of four numbers."""
    m = mask0.bool()
    inv = torch.where(m, nonfinite, torch.zeros_like(nonfinite)).to(torch.int64)
    ocean, worst, total_invalid, affected = (int(v) for v in torch.stack([m.sum(), inv.max(), inv.sum(), (inv > 0).sum()]).tolist())
    if ocean == 0:
        _raise_no_finite_data(total_values, m.numel())
    if worst > 0:
        _raise_invalid_ocean_values(total_invalid, affected, ocean, worst, T)


def check_sufficient_years(cal: Calendar, W: int) -> None:
    """detect.py:615-636."""
    min_year, max_year = int(cal.year_val[0]), int(cal.year_val[-1])
    total_years = max_year - min_year + 1
    if total_years < W:
        raise create_data_validation_error(
            "Insufficient data for shifting_baseline method",
            details=f"Dataset spans {total_years} years but requires at least {W} years",
            suggestions=[
                f"lower window_year_baseline (now {W}) or supply a longer series",
                "the fixed-climatology and detrending methods need no spin-up years",
            ],
            data_info={"available_years": int(total_years), "required_years": int(W)},
        )


# --------------------------------------------------------------------------------------
# (a) anomalies
# --------------------------------------------------------------------------------------
def detrend_model(time, detrend_orders: Sequence[int], remove_harmonics: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Design matrix (K, T) and its pseudo-inverse (T, K), float64 (detect.py:2139-2169): constant,
    centred polynomial terms and, for ``detrend_harmonic``, the annual and semi-annual sine / cosine
    pairs (detect.py:2150-2159); ``remove_harmonics=False`` is detrend_fixed_baseline (detect.py:2450)."""
    dy = time.decimal_year if isinstance(time, Calendar) else decimal_year(time)
    comps = [np.ones(len(dy))]
    centered = dy - np.mean(dy)
    for order in detrend_orders:
        comps.append(centered**order)
    if remove_harmonics:
        comps.extend([np.sin(2 * np.pi * dy), np.cos(2 * np.pi * dy), np.sin(4 * np.pi * dy), np.cos(4 * np.pi * dy)])
    model = np.array(comps)
    for i in range(1, model.shape[0]):
        model[i] = model[i] - np.mean(model[i]) * model[0]
    return model, np.linalg.pinv(model)


def _shift_anomaly_call(h: "_Hold", xd: torch.Tensor, cal: Calendar, W: int, S: int, out_row: np.ndarray, mode: int,
                        out: torch.Tensor, mask0: torch.Tensor, nonfinite: torch.Tensor,
                        edges: Optional[np.ndarray] = None) -> Optional[torch.Tensor]:
    """Dispatch of kernel (a): the TMA-staged daily kernel + fix-up of gridpoints with mixed
    finite / non-finite series when the time axis is gap-free daily, else the generic
    table-driven kernel (sub-sampled or gappy axes, or windows too large for the staged kernel).
    With ``edges`` (the float32 Hobday edge table) the daily kernel also writes the day-of-year-major
    bin codes of the anomalies (fused ``np.digitize``, detect.py:2622-2631); they are returned
    ((366 * NY, pitch) uint16) or None when the generic kernel ran."""
    dev = xd.device
    T, N = xd.shape
    st = _stream()
    tidx, yv, orow = h.up(cal.tidx, np.int32, dev), h.up(cal.year_val, np.int32, dev), h.up(out_row, np.int32, dev)
    if cal.is_daily and N % 4 == 0 and xd.data_ptr() % 16 == 0 and W <= 31:
        bins, e_d, n_e, bp, t_out, y_first = None, None, 0, 0, 0, 0
        if edges is not None and mode == 0 and cal.n_years > W:
            ny_out = cal.n_years - W
            bp = (N + 7) // 8 * 8
            bins = torch.empty((NDOY * ny_out, bp), dtype=torch.uint16, device=dev)
            e_d, n_e = h.up(edges, np.float32, dev), len(edges)
            t_out, y_first = int((out_row >= 0).sum()), int(cal.year_val[0]) + W
        try:
            _lib.call(
                "marex_shift_anomaly_daily_f32", _p(xd), T, N, N, int(cal.year[0]), int(cal.doy[0]), W, S, mode, _p(out), N,
                _p(mask0), _p(nonfinite), e_d, n_e, _p(bins), bp, st,
            )  # fmt: skip
        except ProcessingError as err:
            if err.context.get("code") != _lib.ERR_UNSUPPORTED:
                raise
        else:
            work = torch.empty(N + 1, dtype=torch.int32, device=dev)
            h.append(work)
            _lib.call(
                "marex_shift_anomaly_fixup_f32", _p(xd), T, N, N, tidx, yv, cal.n_years, W, S, orow, mode, _p(out), N,
                _p(mask0), _p(nonfinite), _p(work), e_d, n_e, _p(bins), bp, t_out, y_first, st,
            )  # fmt: skip
            return bins
    _lib.call(
        "marex_shift_anomaly_f32", _p(xd), T, N, N, tidx, yv, cal.n_years, W, S, orow, mode, _p(out), N, _p(mask0),
        _p(nonfinite), st,
    )  # fmt: skip
    return None


@_on_device(lambda x, time, *a, device=None, **k: device if device is not None else _tensor_device(x))
def rolling_climatology_arrays(x, time, window_year_baseline: int = 15, smooth_days_baseline: int = 1, device=None):
    """``rolling_climatology`` (S = 1, detect.py:1511-1688) / ``smoothed_rolling_climatology``
    (detect.py:1691-1816) at array level: the per-time-step climatology, NaN for the first
    ``window_year_baseline`` years.  Returns a float32 CUDA tensor shaped like ``x``."""
    dev = _device(device)
    h = _Hold()
    cal = build_calendar(time)
    xd, space = _to_device_field(x, dev)
    T, N = xd.shape
    out = torch.full((T, N), float("nan"), dtype=torch.float32, device=dev)
    mask0 = torch.empty(N, dtype=torch.uint8, device=dev)
    nonfinite = torch.empty(N, dtype=torch.int32, device=dev)
    _shift_anomaly_call(h, xd, cal, int(window_year_baseline), int(smooth_days_baseline), np.arange(T), 1, out, mask0, nonfinite)
    return out.reshape((T,) + space)


@_on_device(lambda x_dev, *a, **k: _tensor_device(x_dev))
def compute_normalised_anomaly_arrays(
    x_dev: torch.Tensor,
    cal: Calendar,
    method_anomaly: str = "shifting_baseline",
    window_year_baseline: int = 15,
    smooth_days_baseline: int = 21,
    detrend_orders: Optional[Sequence[int]] = None,
    force_zero_mean: bool = True,
    reference_period: Optional[Tuple[int, int]] = None,
    validate: bool = True,
    in_place: bool = False,
    std_normalise: bool = False,
    hobday_edges: Optional[np.ndarray] = None,
) -> Dict[str, Any]:
    """Array-level ``compute_normalised_anomaly`` (detect.py:891-1116) for the three hot-path
    methods.  ``x_dev`` is a float32 CUDA tensor (T, N).  Returns ``dat_anomaly`` (T_out, N)
    (already trimmed for shifting_baseline, detect.py:638-641), ``mask`` (N,) bool, the kept
    rows and, when ``validate``, runs ``_validate_data_values`` on the numbers of the same pass.
    ``in_place`` lets the fixed/detrend methods overwrite ``x_dev``.  ``hobday_edges`` (the float32 edge table of
    the approximate Hobday thresholds): the shifting-baseline kernel also emits the day-of-year-major histogram
    bin codes of its anomalies (``bins``), which ``identify_extremes_arrays`` then takes instead of digitizing."""
    if detrend_orders is None:
        detrend_orders = [1]
    validate_reference_period_method(reference_period, method_anomaly)
    validate_anomaly_method(method_anomaly)
    dev = x_dev.device
    h = _Hold()
    T, N = x_dev.shape
    assert T == cal.T
    mask0 = torch.empty(N, dtype=torch.uint8, device=dev)
    nonfinite = torch.empty(N, dtype=torch.int32, device=dev)
    st = _stream()

    if method_anomaly == "shifting_baseline":
        W, S = int(window_year_baseline), int(smooth_days_baseline)
        check_sufficient_years(cal, W)
        out_row, keep = shifting_out_rows(cal, W)
        T_out = int(keep.sum())
        if T_out == 0:
            raise IndexError("shifting_baseline: no time steps remain after removing the first window_year_baseline years")
        anom = torch.empty((T_out, N), dtype=torch.float32, device=dev)
        bins = _shift_anomaly_call(h, x_dev, cal, W, S, out_row, 0, anom, mask0, nonfinite, edges=hobday_edges)
        if validate:
            check_data_values(mask0, nonfinite, T, T * N)
        return {"dat_anomaly": anom, "mask": mask0.bool(), "keep": keep, "mask_raw": mask0.bool(), "nonfinite": nonfinite, "bins": bins}

    keep = np.ones(T, dtype=bool)
    rows = None
    if reference_period is not None:
        rows = validate_reference_period(reference_period, cal.year)
    ptr, drows = doy_csr(cal.doy, rows)
    ptr_d, rows_d = _up(ptr, np.int32, dev), _up(drows, np.int32, dev)
    doy_d = _up(cal.doy, np.int16, dev)
    clim = torch.empty((NDOY, N), dtype=torch.float32, device=dev)

    if method_anomaly == "fixed_baseline":
        _lib.call("marex_doy_climatology_f32", _p(x_dev), T, N, N, _p(ptr_d), _p(rows_d), None, _p(clim), st)
        anom = x_dev if in_place else torch.empty_like(x_dev)
        _lib.call(
            "marex_sub_doy_climatology_f32", _p(x_dev), T, N, N, _p(doy_d), None, _p(clim), _p(anom), N, _p(mask0),
            _p(nonfinite), st,
        )  # fmt: skip
        if validate:
            check_data_values(mask0, nonfinite, T, T * N)
        return {"dat_anomaly": anom, "mask": mask0.bool(), "keep": keep, "mask_raw": mask0.bool(), "climatology": clim, "nonfinite": nonfinite}

    # detrend_harmonic (detect.py:2061-2296, std_normalise=False) / detrend_fixed_baseline (detect.py:2400-2462)
    harmonic = method_anomaly == "detrend_harmonic"
    validate_detrend_orders(detrend_orders)
    if 1 not in detrend_orders and len(detrend_orders) > 1:
        print("Warning: Higher-order detrending without linear term may be unstable")  # detect.py:2135-2136
    model, pmodel = detrend_model(cal, list(detrend_orders), remove_harmonics=harmonic)
    K = model.shape[0]
    coef = torch.empty((K, N), dtype=torch.float64, device=dev)
    _lib.call(
        "marex_detrend_coef_f64", _p(x_dev), T, N, N, h.up(pmodel, np.float64, dev), K, _p(coef), _p(mask0),
        _p(nonfinite), st,
    )  # fmt: skip
    mask_raw = mask0.bool()
    if validate:
        check_data_values(mask0, nonfinite, T, T * N)
    xd = x_dev if in_place else torch.empty_like(x_dev)
    mean = torch.empty(N, dtype=torch.float32, device=dev) if force_zero_mean else None
    _lib.call(
        "marex_detrend_apply_f32", _p(x_dev), T, N, N, h.up(model, np.float64, dev), K, _p(coef), _p(xd), N,
        _p(mean), st,
    )  # fmt: skip
    if harmonic:  # the detrended series is the anomaly; mask = isfinite of the raw first step (detect.py:2228)
        if force_zero_mean:
            _lib.call(
                "marex_sub_doy_climatology_f32", _p(xd), T, N, N, _p(doy_d), _p(mean), None, _p(xd), N, None, None, st
            )
        out_h = {"dat_anomaly": xd, "mask": mask_raw, "keep": keep, "mask_raw": mask_raw, "nonfinite": nonfinite}
        if std_normalise:  # detect.py:2257-2293
            sd = torch.empty((NDOY, N), dtype=torch.float32, device=dev)
            rms = torch.empty((NDOY, N), dtype=torch.float32, device=dev)
            stn = torch.empty_like(xd)
            _lib.call("marex_doy_std_f32", _p(xd), T, N, N, _p(ptr_d), _p(rows_d), _p(sd), st)
            _lib.call("marex_doy_rolling_rms_f32", _p(sd), N, 30, _p(rms), st)
            _lib.call("marex_div_doy_f32", _p(xd), T, N, N, _p(doy_d), _p(rms), _p(stn), N, st)
            std_cm = torch.empty((N, NDOY), dtype=torch.float32, device=dev)
            _lib.call("marex_transpose_f32", _p(rms), NDOY, N, _p(std_cm), st)
            out_h.update({"dat_stn": stn, "STD": std_cm})
        return out_h
    _lib.call("marex_doy_climatology_f32", _p(xd), T, N, N, _p(ptr_d), _p(rows_d), _p(mean), _p(clim), st)
    mask1 = torch.empty(N, dtype=torch.uint8, device=dev)
    _lib.call(
        "marex_sub_doy_climatology_f32", _p(xd), T, N, N, _p(doy_d), _p(mean), _p(clim), _p(xd), N, _p(mask1), None, st
    )
    return {"dat_anomaly": xd, "mask": mask1.bool(), "keep": keep, "mask_raw": mask_raw, "climatology": clim, "nonfinite": nonfinite}


# --------------------------------------------------------------------------------------
# (b) + (c) thresholds and compare
# --------------------------------------------------------------------------------------
def hobday_bins(precision: float = 0.01, max_anomaly: float = 5.0) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's float32 edge/centre expression, verbatim (detect.py:2601-2608): the kernels
    take these tables from the host so that counts match numpy bit for bit."""
    edges = np.concatenate(
        [[-np.inf], np.arange(-precision, max_anomaly + precision, precision, dtype=np.float32)], dtype=np.float32
    )
    centers = (edges[1:] + edges[:-1]) / 2
    centers[0] = 0.0
    return edges, centers.astype(np.float32)


def global_bins(precision: float = 0.01, max_anomaly: float = 5.0) -> Tuple[np.ndarray, np.ndarray]:
    """float64 edges/centres of the 1-D path (detect.py:2770-2784)."""
    edges = np.concatenate([[-np.inf], np.arange(-precision, max_anomaly + precision, precision)])
    centers = (edges[1:] + edges[:-1]) / 2
    centers[0] = 0.0
    return edges, centers


def _warn_threshold_range(vmin: float, vmax: float, upper: float, lower: float, max_anomaly: float) -> None:
    """The two UserWarnings of detect.py:2711-2730 / 2842-2861."""
    if np.isfinite(vmax) and vmax > upper:
        warnings.warn(
            f"Quantile values exceed expected range: max={vmax:.4f} > {upper:.4f}. "
            f"Consider increasing max_anomaly parameter (currently {max_anomaly:.2f}) or using a lower percentile threshold.",
            UserWarning,
            stacklevel=3,
        )
    if np.isfinite(vmin) and vmin < lower:
        warnings.warn(
            f"Quantile values below expected range in some locations: min={vmin:.4f} < {lower:.4f}. "
            "This is likely due to a constant anomaly in certain (e.g. due to sea ice). "
            "Double check the computed threshold values are correct.",
            UserWarning,
            stacklevel=3,
        )


@_on_device(lambda anom, *a, **k: _tensor_device(anom))
def identify_extremes_arrays(
    anom: torch.Tensor,
    doy: np.ndarray,
    grid: Optional[Tuple[int, int]] = None,
    method_extreme: str = "hobday_extreme",
    threshold_percentile: float = 95,
    window_days_hobday: int = 11,
    window_spatial_hobday: Optional[int] = None,
    method_percentile: str = "approximate",
    precision: float = 0.01,
    max_anomaly: float = 5.0,
    want_events: bool = True,
    want_bits: bool = False,
    n_years: Optional[int] = None,
    cells: Optional[Tuple[int, int]] = None,
    warn: bool = True,
    year: Optional[np.ndarray] = None,
    bins: Optional[torch.Tensor] = None,
) -> Dict[str, Any]:
    """Array-level ``identify_extremes`` (detect.py:1119-1503).  ``anom`` float32 CUDA (T, N);
    ``grid=(ny, nx)`` for gridded data (enables the default 5x5 pooling), ``None`` for
    unstructured.  Returns the doy-major thresholds (``thresholds_dm``), the thresholds in the
    reference's layout (``thresholds``), ``extreme_events`` (bool) and/or ``bits``.
    ``cells=(lo, hi)`` restricts the compare and every returned array to that range of flattened
    gridpoints (the streamed host path computes thresholds on a band with halo rows and keeps the
    rows it owns); ``warn=False`` returns the pre-clamp threshold range in ``stats`` instead of
    raising the reference's UserWarnings.  ``year`` (calendar year of every row) orders the day-of-year-major
    histogram bin codes by year; ``bins`` are those codes when the anomaly kernel already produced them."""
    gridded = grid is not None
    ws = resolve_extreme_config(
        method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday, method_percentile, precision,
        max_anomaly, gridded,
    )  # fmt: skip
    dev = anom.device
    h = _Hold()
    T, N = anom.shape
    st = _stream()
    doy = np.asarray(doy).astype(np.int16)
    out: Dict[str, Any] = {"window_spatial_hobday": ws}
    q = threshold_percentile / 100.0

    if method_extreme == "hobday_extreme":
        w = int(window_days_hobday)
        if n_years is not None:  # detect.py:1905-1915
            n_above = n_years * w * (ws if ws is not None else 1) ** 2 * (1.0 - threshold_percentile / 100.0)
            if n_above < 50:
                logger.warning(
                    f"Not enough samples for accurate extreme detection: {n_above} < 50. "
                    "Consider using a lower threshold_percentile, increasing your time-series size, "
                    "increasing the window_days_hobday, or using a larger window_spatial_hobday."
                    "If your time-series is very short, consider using method_percentile='exact'."
                )
        thr = torch.empty((NDOY, N), dtype=torch.float32, device=dev)
        bins_dm = None
        if method_percentile == "exact":
            ptr, rows = doy_csr(doy)
            mwr = max_window_rows(ptr, w)
            ptr_d, rows_d = _up(ptr, np.int32, dev), _up(rows, np.int32, dev)
            mm = torch.empty(2 * N, dtype=torch.float32, device=dev)
            h.append(mm)
            _lib.call(
                "marex_hobday_thresholds_exact_f32", _p(anom), T, N, N, _p(ptr_d), _p(rows_d), mwr,
                int(np.diff(ptr).max()), w, float(threshold_percentile), _p(thr), _p(mm), st,
            )  # fmt: skip
            out["thresholds"] = thr if not gridded else thr.reshape((NDOY,) + tuple(grid))
            out["thresholds_layout"] = "doy_first"
            out["exact_scratch"] = mm  # int32 view: [0] = groups of 32 gridpoints the queue kernel handed to the histogram kernel
        else:
            if w < 3:
                raise ConfigurationError(
                    "window_days_hobday must be at least 3 with method_percentile='approximate'",
                    details="The reference's wrap padding is undefined for a 1-day window (detect.py:2495-2496)",
                )
            edges, centers = hobday_bins(precision, max_anomaly)
            nb = len(centers)
            ny, nx = (grid if gridded else (1, N))
            ws_eff = int(ws) if (gridded and ws) else 1
            stats = torch.empty(2, dtype=torch.float32, device=dev)
            # histogram bin codes, day-of-year major: NY slots (years) per day of year (detect.py:2622-2631)
            NY, slot_row = doy_slots(doy, year)
            slot_d = _up(slot_row, np.int32, dev)
            edges_d = h.up(edges, np.float32, dev)
            if bins is not None and tuple(bins.shape) == (NDOY * NY, (N + 7) // 8 * 8) and year is not None:
                bins_dm = bins  # written by the anomaly kernel
            else:
                bins_dm = torch.empty((NDOY * NY, (N + 7) // 8 * 8), dtype=torch.uint16, device=dev)
                _lib.call(
                    "marex_digitize_doy_f32", _p(anom), N, N, _p(slot_d), NDOY * NY, edges_d, len(edges), _p(bins_dm),
                    bins_dm.shape[1], st,
                )  # fmt: skip
            bp = int(bins_dm.shape[1])
            mwr = w * NY
            ptr_d = _up(np.arange(NDOY + 1, dtype=np.int64) * NY, np.int32, dev)  # the slots of a day are consecutive rows
            rows_d = torch.arange(NDOY * NY, dtype=torch.int32, device=dev)
            h.extend([slot_d, ptr_d, rows_d])
            if ws_eff in (3, 5, 7) and nx >= 32 and nb <= 1024 and mwr <= 65535:
                wbytes = int(_lib.load().marex_hobday_pooled_workspace_bytes(ny, nx))
                work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
                h.append(work)
                _lib.call(
                    "marex_hobday_thresholds_pooled_bins", _p(bins_dm), NY, ny, nx, bp, _p(ptr_d), _p(rows_d),
                    h.up(centers, np.float32, dev), nb, w, ws_eff, float(q), _p(anom), float(edges[3]), _p(thr),
                    _p(stats), _p(work), wbytes, st,
                )  # fmt: skip
            else:
                _lib.call(
                    "marex_hobday_thresholds_hist", _p(bins_dm), NDOY * NY, ny, nx, bp, _p(ptr_d), _p(rows_d), mwr,
                    h.up(centers, np.float32, dev), nb, w, ws_eff, float(q), _p(anom),
                    float(edges[3]), _p(thr), _p(stats), st,
                )  # fmt: skip
            out["stats"], out["stats_bounds"] = stats, (float(edges[-2]), float(edges[3]))
            if warn:
                vmin, vmax = (float(v) for v in stats.cpu())
                _warn_threshold_range(vmin, vmax, float(edges[-2]), float(edges[3]), max_anomaly)
            thr_cm = torch.empty((N, NDOY), dtype=torch.float32, device=dev)
            _lib.call("marex_transpose_f32", _p(thr), NDOY, N, _p(thr_cm), st)
            out["thresholds"] = thr_cm if not gridded else thr_cm.reshape(tuple(grid) + (NDOY,))
            out["thresholds_layout"] = "doy_last"
        out["thresholds_dm"] = thr
    else:  # global_extreme (detect.py:2873-2923)
        thr = torch.empty(N, dtype=torch.float64, device=dev)
        if method_percentile == "exact":
            _lib.call("marex_global_threshold_exact_f64", _p(anom), T, N, N, float(q), _p(thr), st)
        else:
            edges, centers = global_bins(precision, max_anomaly)
            stats = torch.empty(2, dtype=torch.float64, device=dev)
            if T <= 65535 and len(centers) >= 16:
                eup = edges.astype(np.float32)  # smallest float32 >= each float64 edge: exact float32 binning
                low = eup.astype(np.float64) < edges
                eup[low] = np.nextafter(eup[low], np.float32(np.inf))
                e_dn = np.float32(edges[-1])
                if np.float64(e_dn) > edges[-1]:
                    e_dn = np.nextafter(e_dn, np.float32(-np.inf))
                work = torch.empty(N + 1, dtype=torch.int32, device=dev)
                h.append(work)
                _lib.call(
                    "marex_global_threshold_hist_fast_f64", _p(anom), T, N, N, h.up(edges, np.float64, dev),
                    h.up(eup, np.float32, dev), float(e_dn), h.up(centers, np.float64, dev), len(centers), float(q),
                    float(edges[3]), _p(thr), _p(stats), _p(work), st,
                )  # fmt: skip
            else:
                _lib.call(
                    "marex_global_threshold_hist_f64", _p(anom), T, N, N, h.up(edges, np.float64, dev),
                    h.up(centers, np.float64, dev), len(centers), float(q), float(edges[3]), _p(thr), _p(stats), st,
                )  # fmt: skip
            out["stats"], out["stats_bounds"] = stats, (float(edges[-2]), float(edges[3]))
            if warn:
                vmin, vmax = (float(v) for v in stats.cpu())
                _warn_threshold_range(vmin, vmax, float(edges[-2]), float(edges[3]), max_anomaly)
        out["thresholds"] = thr if not gridded else thr.reshape(tuple(grid))
        out["thresholds_layout"] = "space"
        out["thresholds_dm"] = thr

    c_lo, c_hi = cells if cells is not None else (0, N)
    n_c = c_hi - c_lo
    events = torch.empty((T, n_c), dtype=torch.uint8, device=dev) if want_events else None
    nw = (n_c + 31) // 32
    bits = torch.empty((T, nw), dtype=torch.int32, device=dev) if want_bits else None
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    a_ptr = ctypes.c_void_p(anom.data_ptr() + 4 * c_lo)
    thr_dm = out["thresholds_dm"]
    if method_extreme == "hobday_extreme":
        from_bins = (
            bins_dm is not None and n_c % 8 == 0 and c_lo % 8 == 0 and N % 4 == 0 and (bits is None or n_c % 32 == 0)
            and anom.data_ptr() % 16 == 0
        )  # fmt: skip
        if from_bins:  # 2-byte bin codes instead of the 4-byte anomalies; same result bit for bit
            _lib.call(
                "marex_compare_hobday_bins", ctypes.c_void_p(bins_dm.data_ptr() + 2 * c_lo), NY, bp, _p(slot_d), a_ptr,
                N, n_c, ctypes.c_void_p(thr_dm.data_ptr() + 4 * c_lo), N, edges_d, len(edges), _p(events), n_c,
                _p(bits), nw, _p(count), st,
            )  # fmt: skip
        else:
            ptr_c, rows_c = doy_csr(doy)
            _lib.call(
                "marex_compare_hobday", a_ptr, T, n_c, N, h.up(doy, np.int16, dev), h.up(ptr_c, np.int32, dev),
                h.up(rows_c, np.int32, dev), ctypes.c_void_p(thr_dm.data_ptr() + 4 * c_lo), N, _p(events), n_c,
                _p(bits), nw, _p(count), st,
            )  # fmt: skip
    else:
        _lib.call(
            "marex_compare_global", a_ptr, T, n_c, N, ctypes.c_void_p(thr_dm.data_ptr() + 8 * c_lo), _p(events), n_c,
            _p(bits), nw, _p(count), st,
        )  # fmt: skip
    if cells is not None:  # returned arrays cover the requested gridpoints only
        lay = out["thresholds_layout"]
        flat = out["thresholds"].reshape(N, NDOY) if lay == "doy_last" else (
            out["thresholds"].reshape(NDOY, N) if lay == "doy_first" else out["thresholds"].reshape(N))
        out["thresholds"] = flat[c_lo:c_hi] if lay != "doy_first" else flat[:, c_lo:c_hi]
    if events is not None:
        out["extreme_events"] = events.view(torch.bool)
    if bits is not None:
        out["bits"] = bits
    out["count"] = count
    return out


# The shifting-baseline kernel can write the histogram bin codes of its anomalies itself (no second pass over 38 GB of
# anomalies) -- but it is issue-bound, and the digitize arithmetic costs it more than the separate, memory-bound pass does:
# measured at 0.25 deg, fused 66.3 ms against 48.1 + 13.9 ms unfused (profiles/r02_fused_vs_unfused_digitize.json).  The
# faster arrangement is the default; MAREX_FUSE_DIGITIZE=1 selects the fused kernel (38 GB less DRAM traffic per step).
_FUSE_DIGITIZE = os.environ.get("MAREX_FUSE_DIGITIZE", "0") == "1"


def _fused_edges(method_anomaly, method_extreme, method_percentile, precision, max_anomaly) -> Optional[np.ndarray]:
    """Edge table for the digitize fused into the shifting-baseline kernel, or None when the configuration does not
    digitize (or is invalid: ``identify_extremes_arrays`` raises the reference's error later)."""
    if (
        _FUSE_DIGITIZE and method_anomaly == "shifting_baseline" and method_extreme == "hobday_extreme" and method_percentile == "approximate"
        and isinstance(precision, (int, float)) and isinstance(max_anomaly, (int, float)) and precision > 0
        and 2 <= (max_anomaly + precision) / precision <= 4000
    ):  # fmt: skip
        return hobday_bins(precision, max_anomaly)[0]
    return None


def _dataset_attrs(method_anomaly, method_extreme, threshold_percentile, std_normalise, detrend_orders,
                   window_year_baseline, smooth_days_baseline, window_days_hobday, window_spatial_hobday,
                   reference_period, force_zero_mean, method_percentile, precision, max_anomaly) -> Dict[str, Any]:
    """The Dataset attrs of detect.py:731-783."""
    attrs: Dict[str, Any] = {
        "method_anomaly": method_anomaly,
        "method_extreme": method_extreme,
        "threshold_percentile": threshold_percentile,
        "preprocessing_steps": get_preprocessing_steps(
            method_anomaly, method_extreme, std_normalise, list(detrend_orders), window_year_baseline,
            smooth_days_baseline, window_days_hobday, window_spatial_hobday, reference_period,
        ),
    }  # fmt: skip
    if method_anomaly == "detrend_harmonic":
        attrs.update({"detrend_orders": list(detrend_orders), "force_zero_mean": force_zero_mean, "std_normalise": std_normalise})
    elif method_anomaly == "shifting_baseline":
        attrs.update({"window_year_baseline": window_year_baseline, "smooth_days_baseline": smooth_days_baseline})
    elif method_anomaly == "fixed_baseline":
        if reference_period is not None:
            attrs["reference_period"] = list(reference_period)
    elif method_anomaly == "detrend_fixed_baseline":
        attrs.update({"detrend_orders": list(detrend_orders), "force_zero_mean": force_zero_mean})
        if reference_period is not None:
            attrs["reference_period"] = list(reference_period)
    if method_extreme == "hobday_extreme":
        attrs["window_days_hobday"] = window_days_hobday
    attrs.update({"method_percentile": method_percentile, "precision": precision, "max_anomaly": max_anomaly})
    return attrs


def _host_buffer(key: str, shape, dtype, output: str) -> torch.Tensor:
    """Host tensor for a streamed output."""
    if output == "numpy":
        return torch.empty(shape, dtype=dtype)
    return _pinned_buffer(key, shape, dtype, output == "pinned_reuse")


def _copy2d(dst_ptr: int, dpitch: int, src_ptr: int, spitch: int, width: int, height: int, to_device: bool, stream) -> None:
    _lib.call(
        "marex_memcpy2d_async", ctypes.c_void_p(dst_ptr), dpitch, ctypes.c_void_p(src_ptr), spitch, width, height,
        1 if to_device else 0, ctypes.c_void_p(stream.cuda_stream),
    )  # fmt: skip


def _chunk_bounds(n_units: int, n_chunks: int, align: int) -> List[Tuple[int, int]]:
    """Split [0, n_units) into at most n_chunks contiguous ranges whose interior boundaries are multiples of `align`."""
    n_chunks = max(1, min(n_chunks, n_units // max(align, 1) or 1))
    edges = sorted({min(n_units, int(round(i * n_units / n_chunks / align)) * align) for i in range(1, n_chunks)} | {0, n_units})
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def _preprocess_host_streamed(xh: torch.Tensor, cal: Calendar, gridded: bool, dev: torch.device, n_chunks: int,
                              output: str, want_events: bool, kw: Dict[str, Any]) -> Dict[str, Any]:
    """The whole pipeline on a HOST field, streamed through the GPU in spatial chunks: latitude
    bands (with the ws//2-row pooling halo re-loaded) or cell ranges.  Three streams
    overlap the strided host->device copy of chunk k+1, the kernels of chunk k and the
    device->host copy of chunk k-1, so a call costs about max(PCIe in, PCIe out, compute) instead
    of their sum, and the field may be larger than HBM.  Validation (detect.py:205-279) and the
    threshold-range warnings are evaluated once, over all chunks, as the reference does."""
    from .sharding import effective_halo

    T = int(xh.shape[0])
    space = tuple(int(v) for v in xh.shape[1:])
    method_anomaly, method_extreme = kw["method_anomaly"], kw["method_extreme"]
    if gridded:
        ny, nx = space
        unit, n_units = nx, ny
        halo = effective_halo(method_extreme, kw["method_percentile"], kw["window_spatial_hobday"], True)
        align = 1
    else:
        unit, n_units, halo, align = 1, space[0], 0, 32
    n_total = n_units * unit
    if method_anomaly == "shifting_baseline":
        check_sufficient_years(cal, int(kw["window_year_baseline"]))
        _, keep = shifting_out_rows(cal, int(kw["window_year_baseline"]))
    else:
        keep = np.ones(T, dtype=bool)
    T_out = int(keep.sum())
    if T_out == 0:
        raise IndexError("shifting_baseline: no time steps remain after removing the first window_year_baseline years")
    doy_out, year_out = cal.doy[keep], cal.year[keep]
    n_years_out = int(np.unique(year_out).size)
    hobday = method_extreme == "hobday_extreme"
    exact = kw["method_percentile"] == "exact"
    fused_edges = _fused_edges(method_anomaly, method_extreme, kw["method_percentile"], kw["precision"], kw["max_anomaly"])
    if hobday and not exact:
        thr_shape, thr_dtype, layout = (n_total, NDOY), torch.float32, "doy_last"
    elif hobday:
        thr_shape, thr_dtype, layout = (NDOY, n_total), torch.float32, "doy_first"
    else:
        thr_shape, thr_dtype, layout = (n_total,), torch.float64, "space"
    anom_h = _host_buffer("dat_anomaly", (T_out, n_total), torch.float32, output)
    ev_h = _host_buffer("extreme_events", (T_out, n_total), torch.uint8, output) if want_events else None
    thr_h = _host_buffer("thresholds", thr_shape, thr_dtype, output)
    mask_h = _host_buffer("mask", (n_total,), torch.uint8, output)

    bounds = _chunk_bounds(n_units, n_chunks, align)
    loads = [(max(0, lo - halo), min(n_units, hi + halo)) for lo, hi in bounds]
    max_load = max(b - a for a, b in loads) * unit
    slots = [torch.empty(T * max_load, dtype=torch.float32, device=dev) for _ in range(min(2, len(bounds)))]
    s_comp = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in bounds]
    ev_comp = [torch.cuda.Event() for _ in bounds]
    ev_out = [torch.cuda.Event() for _ in bounds]
    alive: Dict[int, Any] = {}
    # validation / warning statistics, accumulated on the device (no host sync inside the loop)
    z = lambda dt: torch.zeros((), dtype=dt, device=dev)  # noqa: E731
    n_ocean, n_affected, n_invalid, max_invalid, n_events = z(torch.int64), z(torch.int64), z(torch.int64), z(torch.int64), z(torch.int64)
    st_dtype = torch.float64 if not hobday else torch.float32
    smin = torch.full((), float("inf"), dtype=st_dtype, device=dev)
    smax = torch.full((), float("-inf"), dtype=st_dtype, device=dev)
    stats_bounds = None
    src_base, esz = xh.data_ptr(), 4

    def h2d(k: int) -> None:
        a, b = loads[k]
        if k >= 2:  # the slot is free once chunk k-2 has been computed and copied out
            s_in.wait_event(ev_comp[k - 2])
            s_in.wait_event(ev_out[k - 2])
        _copy2d(slots[k % 2].data_ptr(), (b - a) * unit * esz, src_base + a * unit * esz, n_total * esz,
                (b - a) * unit * esz, T, True, s_in)  # fmt: skip
        ev_in[k].record(s_in)

    def compute(k: int) -> None:
        nonlocal n_ocean, n_affected, n_invalid, max_invalid, n_events, smin, smax, stats_bounds
        (lo, hi), (a, b) = bounds[k], loads[k]
        n_load = (b - a) * unit
        c_lo, c_hi = (lo - a) * unit, (hi - a) * unit
        s_comp.wait_event(ev_in[k])
        xd = slots[k % 2][: T * n_load].view(T, n_load)
        res = compute_normalised_anomaly_arrays(
            xd, cal, method_anomaly, kw["window_year_baseline"], kw["smooth_days_baseline"], kw["detrend_orders"],
            kw["force_zero_mean"], kw["reference_period"], validate=False, in_place=method_anomaly != "shifting_baseline",
            hobday_edges=fused_edges,
        )  # fmt: skip
        anom = res["dat_anomaly"]
        ext = identify_extremes_arrays(
            anom, doy_out, (b - a, nx) if gridded else None, method_extreme, kw["threshold_percentile"],
            kw["window_days_hobday"], kw["window_spatial_hobday"], kw["method_percentile"], kw["precision"],
            kw["max_anomaly"], want_events=want_events, want_bits=False, n_years=n_years_out if k == 0 else None,
            cells=(c_lo, c_hi), warn=False, year=year_out, bins=res.get("bins"),
        )  # fmt: skip
        m = res["mask_raw"][c_lo:c_hi]
        inv = torch.where(m, res["nonfinite"][c_lo:c_hi], torch.zeros_like(res["nonfinite"][c_lo:c_hi])).to(torch.int64)
        n_ocean = n_ocean + m.sum()
        n_affected = n_affected + (inv > 0).sum()
        n_invalid = n_invalid + inv.sum()
        max_invalid = torch.maximum(max_invalid, inv.max())
        n_events = n_events + ext["count"][0]
        if "stats" in ext:
            smin, smax = torch.minimum(smin, ext["stats"][0]), torch.maximum(smax, ext["stats"][1])
            stats_bounds = ext["stats_bounds"]
        mask_u8 = res["mask"][c_lo:c_hi].to(torch.uint8)
        thr = ext["thresholds"]
        if layout == "doy_last":
            thr = thr.contiguous()
        ev_comp[k].record(s_comp)
        # ---- device -> host on its own stream ----
        s_out.wait_event(ev_comp[k])
        n_own, g0 = (hi - lo) * unit, lo * unit
        _copy2d(anom_h.data_ptr() + g0 * 4, n_total * 4, anom.data_ptr() + c_lo * 4, n_load * 4, n_own * 4, T_out, False, s_out)
        if want_events:
            e = ext["extreme_events"]
            _copy2d(ev_h.data_ptr() + g0, n_total, e.data_ptr(), n_own, n_own, T_out, False, s_out)
        if layout == "doy_last":
            _copy2d(thr_h.data_ptr() + g0 * NDOY * 4, n_own * NDOY * 4, thr.data_ptr(), n_own * NDOY * 4, n_own * NDOY * 4, 1, False, s_out)
        elif layout == "doy_first":
            _copy2d(thr_h.data_ptr() + g0 * 4, n_total * 4, thr.data_ptr(), n_load * 4, n_own * 4, NDOY, False, s_out)
        else:
            _copy2d(thr_h.data_ptr() + g0 * 8, n_own * 8, thr.data_ptr(), n_own * 8, n_own * 8, 1, False, s_out)
        _copy2d(mask_h.data_ptr() + g0, n_own, mask_u8.data_ptr(), n_own, n_own, 1, False, s_out)
        ev_out[k].record(s_out)
        alive[k] = (res, ext, mask_u8, thr, anom)  # keep the device buffers until the copies have run

    n = len(bounds)
    h2d(0)
    for k in range(n):
        if k + 1 < n:
            h2d(k + 1)  # queued before the kernels of chunk k: the copy engine runs ahead
        compute(k)
        if k >= 1:
            ev_out[k - 1].synchronize()
            alive.pop(k - 1, None)
    ev_out[n - 1].synchronize()
    alive.clear()

    # ---- _validate_data_values over the whole field (detect.py:205-279) ----
    vals = torch.stack([n_ocean, n_affected, n_invalid, max_invalid, n_events]).cpu().tolist()
    ocean, affected, total_invalid, worst, count = (int(v) for v in vals)
    if ocean == 0:  # the errors of check_data_values, with the whole field's numbers
        _raise_no_finite_data(T * n_total, n_total)
    if worst > 0:
        _raise_invalid_ocean_values(total_invalid, affected, ocean, worst, T)
    if stats_bounds is not None:
        _warn_threshold_range(float(smin), float(smax), stats_bounds[0], stats_bounds[1], kw["max_anomaly"])
    out: Dict[str, Any] = {
        "dat_anomaly": anom_h.view((T_out,) + space).numpy(),
        "mask": mask_h.view(space).numpy().view(np.bool_),
        "thresholds": (thr_h.view(space + (NDOY,)) if layout == "doy_last" else thr_h.view((NDOY,) + space) if layout == "doy_first" else thr_h.view(space)).numpy(),
        "thresholds_layout": layout,
        "time": cal.time[keep],
        "extreme_count": count,
        "chunks": len(bounds),
        "h2d_bytes": int(sum((b - a) for a, b in loads) * unit * esz * T),
    }
    if want_events:
        out["extreme_events"] = ev_h.view((T_out,) + space).numpy().view(np.bool_)
    return out


# --------------------------------------------------------------------------------------
# full pipeline at array level
# --------------------------------------------------------------------------------------
@_on_device(lambda x, *a, device=None, **k: device if device is not None else _tensor_device(x))
def preprocess_arrays(
    x,
    time,
    method_anomaly: str = "shifting_baseline",
    method_extreme: str = "hobday_extreme",
    threshold_percentile: float = 95,
    window_year_baseline: int = 15,
    smooth_days_baseline: int = 21,
    window_days_hobday: int = 11,
    window_spatial_hobday: Optional[int] = None,
    std_normalise: bool = False,
    detrend_orders: Optional[Sequence[int]] = None,
    force_zero_mean: bool = True,
    reference_period: Optional[Tuple[int, int]] = None,
    method_percentile: str = "approximate",
    precision: float = 0.01,
    max_anomaly: float = 5.0,
    device=None,
    output: str = "numpy",
    want_events: bool = True,
    want_bits: bool = False,
    gridded: Optional[bool] = None,
    chunks: Optional[int] = None,
) -> Dict[str, Any]:
    """``preprocess_data`` (detect.py:287-841) on arrays: ``x`` is ``(time, lat, lon)`` (gridded)
    or ``(time, ncells)`` (unstructured), numpy or torch, host or device; ``time`` a datetime64
    axis.  Returns ``dat_anomaly`` (T_out, ..space) float32, ``mask`` (..space) bool,
    ``thresholds`` in the reference's layout and dtype, ``extreme_events`` (T_out, ..space) bool,
    ``time`` (trimmed), ``attrs`` (the Dataset attrs, detect.py:731-783) -- numpy arrays
    (``output="numpy"``), numpy views of page-locked buffers (``output="pinned"``: fresh buffers, the result owns them;
    ``"pinned_reuse"``: cached buffers that the next such call overwrites -- for loops that consume each result before
    the next call) or CUDA tensors (``output="torch"``).  A HOST field larger than 1 GiB (or
    any host field when ``chunks`` > 1) is streamed through the GPU in ``chunks`` spatial pieces with
    copies and kernels overlapped (``_preprocess_host_streamed``); ``chunks=1`` forces one piece."""
    if detrend_orders is None:
        detrend_orders = [1]
    std_normalise = bool(std_normalise) and method_anomaly == "detrend_harmonic"  # ignored otherwise, as upstream
    dev = _device(device)
    validate_reference_period_method(reference_period, method_anomaly)
    validate_anomaly_method(method_anomaly)
    cal = build_calendar(time)
    on_host = not (isinstance(x, torch.Tensor) and x.is_cuda)
    if output not in ("numpy", "pinned", "pinned_reuse", "torch"):
        raise ValueError(f"output must be 'numpy', 'pinned', 'pinned_reuse' or 'torch', not {output!r}")
    if on_host and output != "torch" and not want_bits and not std_normalise:
        xh = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        nbytes = xh.numel() * 4
        n_chunks = chunks if chunks is not None else (1 if nbytes < (1 << 30) else int(min(16, max(2, round(nbytes / 6e9)))))
        if n_chunks > 1:
            if xh.dtype != torch.float32 or not xh.is_contiguous():
                xh = xh.to(torch.float32).contiguous()  # da.astype(np.float32), detect.py:600
            if xh.shape[0] != cal.T:
                raise ValueError("time axis length does not match the data")
            g = (xh.dim() == 3) if gridded is None else gridded
            resolve_extreme_config(
                method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday, method_percentile,
                precision, max_anomaly, g,
            )  # fmt: skip
            kw = dict(
                method_anomaly=method_anomaly, method_extreme=method_extreme, threshold_percentile=threshold_percentile,
                window_year_baseline=window_year_baseline, smooth_days_baseline=smooth_days_baseline,
                window_days_hobday=window_days_hobday, window_spatial_hobday=window_spatial_hobday,
                detrend_orders=detrend_orders, force_zero_mean=force_zero_mean, reference_period=reference_period,
                method_percentile=method_percentile, precision=precision, max_anomaly=max_anomaly,
            )  # fmt: skip
            out = _preprocess_host_streamed(xh, cal, g, dev, n_chunks, output, want_events, kw)
            out["attrs"] = _dataset_attrs(
                method_anomaly, method_extreme, threshold_percentile, std_normalise, detrend_orders, window_year_baseline,
                smooth_days_baseline, window_days_hobday, window_spatial_hobday, reference_period, force_zero_mean,
                method_percentile, precision, max_anomaly,
            )  # fmt: skip
            logger.info("Preprocessing completed successfully - %d extreme events identified", out["extreme_count"])
            return out
    owns_input = not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32)
    x_dev, space = _to_device_field(x, dev)
    if gridded is None:
        gridded = len(space) == 2
    grid = tuple(space) if gridded else None
    if x_dev.shape[0] != cal.T:
        raise ValueError("time axis length does not match the data")

    res = compute_normalised_anomaly_arrays(
        x_dev, cal, method_anomaly, window_year_baseline, smooth_days_baseline, detrend_orders, force_zero_mean,
        reference_period, validate=True, in_place=owns_input and method_anomaly != "shifting_baseline",
        std_normalise=std_normalise,
        hobday_edges=_fused_edges(method_anomaly, method_extreme, method_percentile, precision, max_anomaly),
    )  # fmt: skip
    anom, keep = res["dat_anomaly"], res["keep"]
    if owns_input and method_anomaly == "shifting_baseline":
        del x_dev
    doy_out = cal.doy[keep]
    ext = identify_extremes_arrays(
        anom, doy_out, grid, method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday,
        method_percentile, precision, max_anomaly, want_events=want_events, want_bits=want_bits,
        n_years=int(np.unique(cal.year[keep]).size), year=cal.year[keep], bins=res.pop("bins", None),
    )  # fmt: skip

    attrs = _dataset_attrs(
        method_anomaly, method_extreme, threshold_percentile, std_normalise, detrend_orders, window_year_baseline,
        smooth_days_baseline, window_days_hobday, window_spatial_hobday, reference_period, force_zero_mean,
        method_percentile, precision, max_anomaly,
    )  # fmt: skip

    T_out = anom.shape[0]
    out: Dict[str, Any] = {
        "dat_anomaly": anom.reshape((T_out,) + space),
        "mask": res["mask"].reshape(space),
        "thresholds": ext["thresholds"],
        "thresholds_layout": ext["thresholds_layout"],
        "time": cal.time[keep],
        "attrs": attrs,
        "extreme_count": ext["count"],
    }
    if want_events:
        out["extreme_events"] = ext["extreme_events"].reshape((T_out,) + space)
    if want_bits:
        out["bits"] = ext["bits"]
    if std_normalise:  # detect.py:686-715: the same extreme identification on the standardised anomalies
        ext_s = identify_extremes_arrays(
            res["dat_stn"], doy_out, grid, method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday,
            method_percentile, precision, max_anomaly, want_events=True, want_bits=False,
        )  # fmt: skip
        out["dat_stn"] = res["dat_stn"].reshape((T_out,) + space)
        out["STD"] = res["STD"].reshape(space + (NDOY,))
        out["extreme_events_stn"] = ext_s["extreme_events"].reshape((T_out,) + space)
        out["thresholds_stn"] = ext_s["thresholds"]
    logger.info("Preprocessing completed successfully - %d extreme events identified", int(ext["count"]))
    if output != "torch":
        keys = ("dat_anomaly", "mask", "thresholds", "extreme_events", "bits", "dat_stn", "STD", "extreme_events_stn", "thresholds_stn")
        host = {k: _to_host(out[k], k, output) for k in keys if k in out}
        out["extreme_count"] = int(out["extreme_count"])  # synchronises the stream: the copies above are complete
        for k, v in host.items():
            out[k] = v.numpy()
    return out
