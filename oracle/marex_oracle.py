"""
CPU oracle for the marEx ``preprocess_data`` detection hot path.

*** TEST INFRASTRUCTURE ONLY ***  Nothing under ``marex_b200/`` may import this
module.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, as the checker or as the
timed CPU baseline -- never as the product path.

It is a plain numpy restatement of the arithmetic in the reference
``/root/reference/marEx/detect.py`` (every function cites the lines it
follows).  The reference itself cannot be imported in this image (xarray, dask,
flox and xhistogram are not installed), so:

* PARITY PINNED for the pure-numpy leaves that can be AST-extracted from the
  reference and executed here: ``_rolling_histogram_quantile``
  (detect.py:2465-2559), the nested ``_doy_percentiles`` (detect.py:1936-1942),
  ``add_decimal_year`` (detect.py:2031-2058) and ``_get_preprocessing_steps``
  (detect.py:844-888).  ``tests/golden/make_golden.py`` executes those leaves
  and stores their outputs; ``tests/test_oracle_golden.py`` checks this file
  against them bit-for-bit.
* PARITY UNPINNED for the stages whose arithmetic lives in third-party
  libraries that are absent here (xarray ``rolling().mean()``, flox
  ``nanmean``/``count``, xhistogram ``histogram``, xarray ``quantile``/``dot``):
  they are restated from the libraries' documented semantics.  Where the
  reference's floating-point summation order is unknowable (rolling mean,
  grouped nanmean, dot products) the oracle accumulates in float64 and rounds
  once to float32, which is within 1 ulp(f32) of any float32 summation order's
  exact target.

Array conventions: time-major ``x[T, N]`` float32 with ``N = ny*nx`` flattened
row-major (lat-major), or ``x[T, ny, nx]``; the functions flatten internally.
"""

from __future__ import annotations

import warnings
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

NDOY = 366


# --------------------------------------------------------------------------- #
# Calendar helpers
# --------------------------------------------------------------------------- #
def calendar_tables(time) -> Tuple[np.ndarray, np.ndarray]:
    """year[T] (int32), dayofyear[T] (int16, 1..366) of a datetime64 axis.

    Follows ``da[time].dt.year`` / ``.dt.dayofyear`` (detect.py:1605-1606):
    proleptic Gregorian calendar.
    """
    t = np.asarray(time).astype("datetime64[D]")
    y = t.astype("datetime64[Y]")
    year = y.astype(np.int64) + 1970
    doy = (t - y.astype("datetime64[D]")).astype(np.int64) + 1
    return year.astype(np.int32), doy.astype(np.int16)


def decimal_year(time) -> np.ndarray:
    """year + days_elapsed / days_in_year, float64 (detect.py:2031-2058)."""
    t = np.asarray(time).astype("datetime64[D]")
    y = t.astype("datetime64[Y]")
    start = y.astype("datetime64[D]")
    nxt = (y + 1).astype("datetime64[D]")
    elapsed = (t - start).astype(np.int64)
    duration = (nxt - start).astype(np.int64)
    return (y.astype(np.int64) + 1970) + elapsed / duration


def _flat(x: np.ndarray) -> Tuple[np.ndarray, Tuple[int, ...]]:
    x = np.asarray(x)
    return x.reshape(x.shape[0], -1), x.shape[1:]


# --------------------------------------------------------------------------- #
# (a) shifting baseline
# --------------------------------------------------------------------------- #
def smooth_centered(x: np.ndarray, S: int) -> np.ndarray:
    """``x.rolling(time=S, center=True).mean().astype(float32)`` (detect.py:1810-1812).

    xarray semantics: ``min_periods=None`` -> a full window of S valid samples
    is required, so the result is NaN when any sample of the window is NaN or
    the window sticks out of ``[0, T)``.  The window of output ``t`` is
    ``[t - S//2, t - S//2 + S - 1]``.  float64 accumulate, one rounding to f32.
    """
    x2, tail = _flat(x)
    T, N = x2.shape
    out = np.full((T, N), np.nan, dtype=np.float32)
    if S > T:
        return out.reshape((T,) + tail)
    xf = x2.astype(np.float64)
    isn = np.isnan(xf)
    ispi = xf == np.inf
    isni = xf == -np.inf
    fin = np.where(isn | ispi | isni, 0.0, xf)

    def wsum(a):
        c = np.concatenate([np.zeros((1, N), dtype=a.dtype), np.cumsum(a, axis=0)], axis=0)
        return c[S:] - c[:-S]  # (T-S+1, N): window starting at row i

    # exact window sums: cumsum of f64 over f32 data is exact enough only for
    # short series; use a blocked direct summation instead to stay exact.
    sums = np.zeros((T - S + 1, N), dtype=np.float64)
    for k in range(S):
        sums += fin[k : k + T - S + 1]
    n_nan = wsum(isn.astype(np.int32))
    n_pi = wsum(ispi.astype(np.int32))
    n_ni = wsum(isni.astype(np.int32))
    res = sums / S
    res = np.where(n_pi > 0, np.inf, res)
    res = np.where(n_ni > 0, -np.inf, res)
    res = np.where((n_pi > 0) & (n_ni > 0), np.nan, res)
    res = np.where(n_nan > 0, np.nan, res)
    off = S // 2
    out[off : off + T - S + 1] = res.astype(np.float32)
    return out.reshape((T,) + tail)


def _nanmean_rows(v: np.ndarray) -> np.ndarray:
    """nanmean over axis 0 in float64, NaN where no valid sample (flox ``nanmean``
    with ``fill_value=np.nan``, detect.py:1659-1669).  inf propagates as in IEEE."""
    v = v.astype(np.float64)
    isn = np.isnan(v)
    cnt = (~isn).sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        sm = np.where(isn, 0.0, v).sum(axis=0)
        return np.where(cnt > 0, sm / np.maximum(cnt, 1), np.nan)


def rolling_climatology(s: np.ndarray, year: np.ndarray, doy: np.ndarray, W: int) -> np.ndarray:
    """Per-time-step rolling climatology (detect.py:1511-1688).

    ``clim[t] = nanmean{ s[u] : year[t]-W <= year[u] <= year[t]-1, doy[u] == doy[t] }``
    for target years ``>= min(year) + W`` (detect.py:1631-1634), NaN elsewhere.
    Returns float32 ``(T, ...)``.
    """
    s2, tail = _flat(s)
    T, N = s2.shape
    year = np.asarray(year).astype(np.int64)
    doy = np.asarray(doy).astype(np.int64)
    out = np.full((T, N), np.nan, dtype=np.float32)
    min_year = int(year.min())
    for Ty in np.unique(year):
        if Ty < min_year + W:
            continue
        tgt = year == Ty
        contrib = (year >= Ty - W) & (year <= Ty - 1)
        for d in np.unique(doy[tgt]):
            idx = np.nonzero(contrib & (doy == d))[0]
            if idx.size == 0:
                continue
            out[tgt & (doy == d)] = _nanmean_rows(s2[idx]).astype(np.float32)
    return out.reshape((T,) + tail)


def anomaly_shifting_baseline(x, year, doy, W: int = 15, S: int = 21, trim: bool = True):
    """``_compute_anomaly_shifting_baseline`` (detect.py:1819-1850) + the trim of
    ``preprocess_data`` (detect.py:615-641).

    Returns ``(anom[T_out, ...] f32, mask[...] bool, keep[T] bool)``.
    """
    x = np.asarray(x, dtype=np.float32)
    s = smooth_centered(x, S)
    clim = rolling_climatology(s, year, doy, W)
    with np.errstate(invalid="ignore"):
        anom = (x - clim).astype(np.float32)
    mask = np.isfinite(x[0])
    year = np.asarray(year)
    keep = year >= int(year.min()) + W
    if trim:
        anom = anom[keep]
    return anom, mask, keep


# --------------------------------------------------------------------------- #
# (a') fixed baseline, (a'') detrend + fixed baseline
# --------------------------------------------------------------------------- #
def daily_climatology(x, year, doy, reference_period: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """Per-doy nanmean, float32 ``(366, ...)`` (detect.py:2334-2373).  Rows of days
    of year that never occur stay NaN."""
    x2, tail = _flat(x)
    year = np.asarray(year).astype(np.int64)
    doy = np.asarray(doy).astype(np.int64)
    sel = np.ones(len(year), dtype=bool)
    if reference_period is not None:
        a, b = reference_period
        if a > b:
            raise ValueError("Invalid reference_period")
        sel = (year >= a) & (year <= b)
        if not sel.any():
            raise ValueError("No data found in reference_period")
    clim = np.full((NDOY, x2.shape[1]), np.nan, dtype=np.float32)
    for d in np.unique(doy[sel]):
        clim[d - 1] = _nanmean_rows(x2[sel & (doy == d)]).astype(np.float32)
    return clim.reshape((NDOY,) + tail)


def anomaly_fixed_baseline(x, year, doy, reference_period=None):
    """``_compute_anomaly_fixed_baseline`` (detect.py:2299-2397): returns (anom f32, mask)."""
    x = np.asarray(x, dtype=np.float32)
    clim = daily_climatology(x, year, doy, reference_period)
    d = np.asarray(doy).astype(np.int64) - 1
    with np.errstate(invalid="ignore"):
        anom = (x - clim[d]).astype(np.float32)
    return anom, np.isfinite(x[0])


def detrend_model(time, detrend_orders: Sequence[int], remove_harmonics: bool = False):
    """Design matrix ``model`` (K, T) float64 and its pseudo-inverse ``pmodel`` (T, K)
    (detect.py:2139-2169).  ``time``: datetime64 axis, or a float64 decimal-year array (model calendars)."""
    dy = np.asarray(time) if np.asarray(time).dtype.kind == "f" else decimal_year(time)
    comps = [np.ones(len(dy))]
    centered = dy - np.mean(dy)
    for order in detrend_orders:
        comps.append(centered**order)
    if remove_harmonics:
        comps.extend([np.sin(2 * np.pi * dy), np.cos(2 * np.pi * dy), np.sin(4 * np.pi * dy), np.cos(4 * np.pi * dy)])
    model = np.array(comps)
    for i in range(1, model.shape[0]):
        model[i] = model[i] - np.mean(model[i]) * model[0]
    pmodel = np.linalg.pinv(model)
    return model, pmodel


def detrend(x, time, detrend_orders=(1,), force_zero_mean=True, remove_harmonics=False) -> np.ndarray:
    """``_compute_anomaly_detrended`` without std normalisation (detect.py:2128-2224)."""
    x = np.asarray(x, dtype=np.float32)
    x2, tail = _flat(x)
    model, pmodel = detrend_model(time, list(detrend_orders), remove_harmonics)
    with np.errstate(invalid="ignore", over="ignore"):
        coef = pmodel.T @ x2.astype(np.float64)  # (K, N)  detect.py:2206
        fit = (model.T @ coef).astype(np.float32)  # detect.py:2220
        xd = (x2 - fit).astype(np.float32)
        if force_zero_mean:  # detect.py:2223-2224 (xarray mean skips NaN)
            xd = (xd - _nanmean_rows(xd).astype(np.float32)).astype(np.float32)
    return xd.reshape(x.shape)


def std_normalise(xd, doy, window: int = 30):
    """The ``std_normalise`` branch of ``_compute_anomaly_detrended`` (detect.py:2257-2293): per-day-of-year
    population std (flox ``std``, float32 result), squared, centred ``window``-day rolling mean with annual
    wrap (window [d-15, d+14] for 30), sqrt, values <= 1e-10 -> NaN, division by day of year.
    Returns ``dat_stn`` like ``xd`` and ``STD`` as ``(N, 366)`` float32.  PARITY UNPINNED (flox / bottleneck
    summation order): float64 accumulation rounded once."""
    x2, tail = _flat(np.asarray(xd, dtype=np.float32))
    T, N = x2.shape
    d0 = np.asarray(doy).astype(np.int64) - 1
    std_day = np.full((NDOY, N), np.nan, dtype=np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        for d in range(NDOY):
            rows = np.nonzero(d0 == d)[0]
            if rows.size:
                std_day[d] = np.std(x2[rows].astype(np.float64), axis=0).astype(np.float32)
        sq = (std_day * std_day).astype(np.float32)
        half = window // 2
        out = np.empty((NDOY, N), dtype=np.float32)
        for d in range(NDOY):
            idx = (np.arange(d - half, d - half + window)) % NDOY
            out[d] = np.sqrt(np.mean(sq[idx].astype(np.float64), axis=0).astype(np.float32))
        out = np.where(out > np.float32(1e-10), out, np.float32(np.nan)).astype(np.float32)
        stn = (x2 / out[d0]).astype(np.float32)
    return stn.reshape(np.asarray(xd).shape), np.ascontiguousarray(out.T)


def anomaly_detrend_fixed_baseline(x, time, year, doy, detrend_orders=(1,), force_zero_mean=True, reference_period=None):
    """``_compute_anomaly_detrend_fixed_baseline`` (detect.py:2400-2462)."""
    xd = detrend(x, time, detrend_orders, force_zero_mean, remove_harmonics=False)
    return anomaly_fixed_baseline(xd, year, doy, reference_period)


# --------------------------------------------------------------------------- #
# (b-approx) hobday histogram thresholds
# --------------------------------------------------------------------------- #
def hobday_bins(precision: float = 0.01, max_anomaly: float = 5.0):
    """float32 bin edges / centres of the 2-D histogram path (detect.py:2601-2608)."""
    edges = np.concatenate(
        [[-np.inf], np.arange(-precision, max_anomaly + precision, precision, dtype=np.float32)], dtype=np.float32
    )
    centers = (edges[1:] + edges[:-1]) / 2
    centers[0] = 0.0
    return edges, centers.astype(np.float32)


def digitize(a: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """``np.digitize(a, edges) - 1`` as uint16 (detect.py:2622-2631).  NaN and
    ``a >= edges[-1]`` map to ``len(edges)-1`` (outside expected_groups -> not counted)."""
    return (np.digitize(a, edges) - 1).astype(np.uint16)


def doy_bin_counts(bins: np.ndarray, doy: np.ndarray, nb: int) -> np.ndarray:
    """flox ``count`` by (dayofyear, bin) -> ``h[N, 366, nb]`` int32 (detect.py:2638-2648)."""
    b2, _ = _flat(bins)
    T, N = b2.shape
    d0 = np.asarray(doy).astype(np.int64) - 1
    b = b2.astype(np.int64)
    flat = (np.arange(N, dtype=np.int64)[None, :] * NDOY + d0[:, None]) * nb + b
    return np.bincount(flat[b < nb], minlength=N * NDOY * nb).astype(np.int32).reshape(N, NDOY, nb)


def pool_spatial(h: np.ndarray, ny: int, nx: int, ws: int) -> np.ndarray:
    """``ws x ws`` box sum of per-cell histograms: periodic in lon, truncated in lat
    (``min_periods=1``) (detect.py:2651-2668).  ``h[N, ...]`` with ``N = ny*nx``."""
    if ws is None or ws <= 1:
        return h
    p = ws // 2
    g = h.reshape((ny, nx) + h.shape[1:]).astype(np.int64)
    gx = np.zeros_like(g)
    for dx in range(-p, p + 1):
        gx += np.roll(g, -dx, axis=1)
    gy = np.zeros_like(gx)
    for dy in range(-p, p + 1):
        lo, hi = max(0, -dy), min(ny, ny - dy)
        if hi > lo:
            gy[lo:hi] += gx[lo + dy : hi + dy]
    return gy.reshape(h.shape)


def rolling_histogram_quantile(hist: np.ndarray, w: int, q: float, centers: np.ndarray) -> np.ndarray:
    """Restatement of ``_rolling_histogram_quantile`` (detect.py:2465-2559), vectorised
    over leading axes: ``hist[..., 366, nb]`` -> ``thr[..., 366]`` float32."""
    n_doy, n_bins = hist.shape[-2:]
    pad = w // 2
    hist = np.asarray(hist)
    H = np.zeros(hist.shape, dtype=np.float64 if hist.dtype.kind == "f" else np.int64)
    for k in range(-pad, pad + 1):  # detect.py:2494-2500 (wrap pad + window sum)
        H += np.roll(hist, -k, axis=-2)
    cum = np.cumsum(H, axis=-1, dtype=np.int32)  # detect.py:2510
    total = cum[..., -1]
    pos = q * total  # float64, detect.py:2516
    iu = (cum > pos[..., None]).argmax(axis=-1)  # searchsorted(side="right"), detect.py:2527
    none_gt = ~(cum > pos[..., None]).any(axis=-1)
    iu = np.where(none_gt, n_bins, iu)
    iu = np.where(total <= 0, 0, iu)
    iu = np.clip(iu, 0, n_bins - 1)
    il = np.maximum(0, iu - 1)
    cl = np.take_along_axis(cum, il[..., None], axis=-1)[..., 0]
    cu = np.take_along_axis(cum, iu[..., None], axis=-1)[..., 0]
    centers = np.asarray(centers)
    bl = centers[il]
    bu = centers[iu]
    diff = cu - cl
    safe = np.where(diff > 1e-10, diff, 1.0)
    frac = np.where(diff > 1e-10, (pos - cl) / safe, 0.5)
    thr = bl + frac * (bu - bl)  # (bu - bl) in the dtype of ``centers`` (f32), detect.py:2550
    thr = np.where(total > 0, thr, np.nan)
    thr = np.where((iu == 0) & (total > 0), centers[0], thr)
    return thr.astype(np.float32)


def hobday_thresholds_approx(
    anom,
    doy,
    q: float,
    w: int = 11,
    ws: Optional[int] = None,
    grid: Optional[Tuple[int, int]] = None,
    precision: float = 0.01,
    max_anomaly: float = 5.0,
    warn: bool = False,
) -> np.ndarray:
    """``_compute_histogram_quantile_2d`` (detect.py:2562-2734) -> ``thr[N, 366]`` f32.

    ``grid=(ny, nx)`` and ``ws`` select the spatial pooling; ``q = percentile/100``.
    """
    a2, _ = _flat(np.asarray(anom, dtype=np.float32))
    edges, centers = hobday_bins(precision, max_anomaly)
    nb = len(edges) - 1
    h = doy_bin_counts(digitize(a2, edges), doy, nb)
    if ws is not None and ws > 1:
        ny, nx = grid
        h = pool_spatial(h, ny, nx, ws)
    thr = rolling_histogram_quantile(h, w, q, centers)
    thr[np.isnan(a2[0])] = np.nan  # detect.py:2704-2705
    upper, lower = edges[-2], edges[3]
    with np.errstate(invalid="ignore"):
        if warn and (thr > upper).any():
            warnings.warn("Quantile values exceed expected range", UserWarning, stacklevel=2)
        low = thr < lower
    if low.any():
        if warn:
            warnings.warn("Quantile values below expected range in some locations", UserWarning, stacklevel=2)
        thr = np.where(low, lower, thr).astype(np.float32)  # detect.py:2732
    return thr


# --------------------------------------------------------------------------- #
# (b-exact) hobday exact thresholds
# --------------------------------------------------------------------------- #
def doy_window_masks(doy, w: int) -> np.ndarray:
    """366 boolean time masks of the +-w//2 doy window, wrap 366 (detect.py:1925-1934)."""
    doy = np.asarray(doy).astype(np.int64)
    half = w // 2
    masks = np.zeros((NDOY, len(doy)), dtype=bool)
    for d in range(1, NDOY + 1):
        for off in range(-half, half + 1):
            masks[d - 1] |= doy == ((d - 1 + off) % NDOY) + 1
    return masks


def hobday_thresholds_exact(anom, doy, percentile: float, w: int = 11) -> np.ndarray:
    """Exact branch of ``_identify_extremes_hobday`` (detect.py:1921-1956) -> ``thr[366, N]`` f32.
    Calls ``np.nanpercentile`` itself, exactly as the reference does."""
    a2, _ = _flat(np.asarray(anom, dtype=np.float32))
    data = np.ascontiguousarray(a2.T)  # (N, T): core dim last, like apply_ufunc
    masks = doy_window_masks(doy, w)
    res = np.full((data.shape[0], NDOY), np.nan, dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for i in range(NDOY):
            if masks[i].any():
                res[:, i] = np.nanpercentile(data[:, masks[i]], percentile, axis=-1)
    return np.ascontiguousarray(res.T)


# --------------------------------------------------------------------------- #
# (b-global) constant-in-time thresholds
# --------------------------------------------------------------------------- #
def global_bins(precision: float = 0.01, max_anomaly: float = 5.0):
    """float64 edges / centres of the 1-D histogram path (detect.py:2770-2784)."""
    edges = np.concatenate([[-np.inf], np.arange(-precision, max_anomaly + precision, precision)])
    centers = (edges[1:] + edges[:-1]) / 2
    centers[0] = 0.0
    return edges, centers


def global_hist_counts(a2: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """xhistogram semantics (detect.py:2775): left-closed bins, last bin right-closed,
    NaN and out-of-range dropped.  ``a2[T, N]`` -> ``hist[N, nb]`` int64."""
    nb = len(edges) - 1
    af = a2.astype(np.float64)
    idx = np.searchsorted(edges[:-1], af, side="right") - 1
    with np.errstate(invalid="ignore"):
        ok = ~np.isnan(af) & (af <= edges[-1]) & (idx >= 0)
    hist = np.zeros((a2.shape[1], nb), dtype=np.int64)
    cols = np.broadcast_to(np.arange(a2.shape[1]), a2.shape)
    np.add.at(hist, (cols[ok], idx[ok]), 1)
    return hist


def global_threshold_approx(anom, q: float, precision: float = 0.01, max_anomaly: float = 5.0, warn: bool = False):
    """``_compute_histogram_quantile_1d`` (detect.py:2737-2865) -> ``thr[N]`` float64."""
    a2, _ = _flat(np.asarray(anom, dtype=np.float32))
    edges, centers = global_bins(precision, max_anomaly)
    nb = len(centers)
    hist = global_hist_counts(a2, edges)
    eps = 1e-10
    hist_sum = hist.sum(axis=1) + 1e-10
    pdf = hist / hist_sum[:, None]
    cdf = np.cumsum(pdf, axis=1)
    rows = np.arange(hist.shape[0])
    iu = (cdf >= (q - eps)).argmax(axis=1)
    ib = np.where(iu - 1 > 0, iu - 1, 0)
    target = cdf[rows, ib]
    il = (cdf > target[:, None]).argmax(axis=1)
    il = np.where(il < 0, 0, np.where(il > nb - 2, nb - 2, il))
    iu = np.where(iu < 1, 1, np.where(iu > nb - 1, nb - 1, iu))
    cl, cu = cdf[rows, il], cdf[rows, iu]
    bl, bu = centers[il], centers[iu]
    denom = cu - cl
    exact = np.fabs(cl - q) < eps
    zero = np.fabs(denom) <= eps
    frac = (q - cl) / np.where(np.fabs(denom) > eps, denom, 1.0)
    thr = bl + frac * (bu - bl)
    thr = np.where(exact, bl, thr)
    thr = np.where(zero & ~exact, (bl + bu) / 2, thr)
    thr = np.where(np.isnan(a2).any(axis=0), np.nan, thr)
    upper, lower = edges[-2], edges[3]
    with np.errstate(invalid="ignore"):
        if warn and (thr > upper).any():
            warnings.warn("Quantile values exceed expected range", UserWarning, stacklevel=2)
        low = thr < lower
    if low.any():
        if warn:
            warnings.warn("Quantile values below expected range in some locations", UserWarning, stacklevel=2)
        thr = np.where(low, lower, thr)
    return thr


def global_threshold_exact(anom, percentile: float) -> np.ndarray:
    """``da.quantile(p/100, dim=time)`` (detect.py:2899): nanquantile with a float64 q -> float64."""
    a2, _ = _flat(np.asarray(anom, dtype=np.float32))
    # numpy quirk: np.nanquantile(..., axis) goes through apply_along_axis, whose output dtype is
    # taken from the FIRST slice; an all-NaN first cell returns a float32 NaN and silently demotes
    # every other cell's float64 result to float32.  The clean float64 semantics are restated here
    # by keeping all-NaN cells out of the call.
    valid = ~np.isnan(a2).all(axis=0)
    out = np.full(a2.shape[1], np.nan, dtype=np.float64)
    if valid.any():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            out[valid] = np.nanquantile(a2[:, valid], np.asarray([percentile / 100.0]), axis=0)[0]
    return out


# --------------------------------------------------------------------------- #
# (c) compare
# --------------------------------------------------------------------------- #
def compare_hobday(anom, doy, thr_doy_major) -> np.ndarray:
    """``anom[t] >= thr[doy[t]]`` (detect.py:2001-2004); ``thr_doy_major[366, N]``."""
    a2, _ = _flat(np.asarray(anom))
    d = np.asarray(doy).astype(np.int64) - 1
    with np.errstate(invalid="ignore"):
        return (a2 >= thr_doy_major[d]).reshape(np.asarray(anom).shape)


def compare_global(anom, thr) -> np.ndarray:
    """``anom >= thr`` with a float64 threshold (detect.py:2915)."""
    a2, _ = _flat(np.asarray(anom))
    with np.errstate(invalid="ignore"):
        return (a2.astype(np.float64) >= np.asarray(thr, dtype=np.float64)[None, :]).reshape(np.asarray(anom).shape)


def pack_bits_time_major(events: np.ndarray) -> np.ndarray:
    """Bit-pack ``events[T, N]`` along cells, little-endian bit order: bit ``c & 31`` of
    word ``c >> 5``; rows padded to whole 32-bit words."""
    e2, _ = _flat(np.asarray(events, dtype=bool))
    T, N = e2.shape
    nw = (N + 31) // 32
    pad = np.zeros((T, nw * 32), dtype=bool)
    pad[:, :N] = e2
    return np.packbits(pad.reshape(T, nw, 32), axis=-1, bitorder="little").view(np.uint32).reshape(T, nw)


# --------------------------------------------------------------------------- #
# validation + pipeline
# --------------------------------------------------------------------------- #
def validate_data_values(x) -> Dict[str, int]:
    """The numbers ``_validate_data_values`` (detect.py:205-279) bases its errors on."""
    x2, _ = _flat(np.asarray(x))
    mask0 = np.isfinite(x2[0])
    invalid = (~np.isfinite(x2)).sum(axis=0)
    inv_ocean = np.where(mask0, invalid, 0)
    return {
        "any_valid": bool(mask0.any()),
        "total_invalid_in_ocean": int(inv_ocean.sum()),
        "locations_affected": int((inv_ocean > 0).sum()),
        "max_invalid": int(inv_ocean.max()) if inv_ocean.size else 0,
        "total_ocean_locations": int(mask0.sum()),
    }


def preprocess(
    x,
    time,
    method_anomaly: str = "shifting_baseline",
    method_extreme: str = "hobday_extreme",
    threshold_percentile: float = 95,
    window_year_baseline: int = 15,
    smooth_days_baseline: int = 21,
    window_days_hobday: int = 11,
    window_spatial_hobday: Optional[int] = None,
    detrend_orders=(1,),
    force_zero_mean: bool = True,
    reference_period=None,
    method_percentile: str = "approximate",
    precision: float = 0.01,
    max_anomaly: float = 5.0,
    std_normalise_flag: bool = False,
    year_doy=None,
) -> Dict[str, np.ndarray]:
    """Array-level ``preprocess_data`` (detect.py:287-841).  ``x[T, ny, nx]`` (gridded)
    or ``x[T, ncells]`` (unstructured).  Thresholds are returned in the reference's
    layouts: hobday-approx ``(..space, 366)``, hobday-exact ``(366, ..space)``,
    global ``(..space)`` float64."""
    x = np.asarray(x, dtype=np.float32)
    space = x.shape[1:]
    gridded = len(space) == 2
    if year_doy is not None:  # model calendar (noleap / 360_day ...): (year, doy, decimal_year) given, `time` = labels
        year, doy, dy_model = (np.asarray(v) for v in year_doy)
        time_fit = dy_model.astype(np.float64)
    else:
        year, doy = calendar_tables(time)
        time_fit = time
    if method_anomaly == "shifting_baseline":
        anom, mask, keep = anomaly_shifting_baseline(x, year, doy, window_year_baseline, smooth_days_baseline)
        doy_o, time_o = doy[keep], np.asarray(time)[keep]
    elif method_anomaly == "fixed_baseline":
        anom, mask = anomaly_fixed_baseline(x, year, doy, reference_period)
        doy_o, time_o = doy, np.asarray(time)
    elif method_anomaly == "detrend_fixed_baseline":
        anom, mask = anomaly_detrend_fixed_baseline(x, time_fit, year, doy, detrend_orders, force_zero_mean, reference_period)
        doy_o, time_o = doy, np.asarray(time)
    elif method_anomaly == "detrend_harmonic":  # detect.py:2061-2296 with std_normalise=False
        anom = detrend(x, time_fit, detrend_orders, force_zero_mean, remove_harmonics=True)
        mask = np.isfinite(_flat(x)[0][0])  # detect.py:2228: first step of the raw field
        doy_o, time_o = doy, np.asarray(time)
    else:
        raise ValueError(method_anomaly)
    a2 = anom.reshape(anom.shape[0], -1)
    q = threshold_percentile / 100.0
    if method_extreme == "hobday_extreme":
        if method_percentile == "exact":
            thr_dm = hobday_thresholds_exact(a2, doy_o, threshold_percentile, window_days_hobday)
            thr_out = thr_dm.reshape((NDOY,) + space)
        else:
            ws = window_spatial_hobday
            if ws is None and gridded:
                ws = 5  # detect.py:1451-1452
            thr_cm = hobday_thresholds_approx(
                a2, doy_o, q, window_days_hobday, ws, space if gridded else None, precision, max_anomaly
            )
            thr_dm = np.ascontiguousarray(thr_cm.T)
            thr_out = thr_cm.reshape(space + (NDOY,))
        events = compare_hobday(a2, doy_o, thr_dm)
    elif method_extreme == "global_extreme":
        if method_percentile == "exact":
            thr = global_threshold_exact(a2, threshold_percentile)
        else:
            thr = global_threshold_approx(a2, q, precision, max_anomaly)
        thr_out = thr.reshape(space)
        events = compare_global(a2, thr)
    else:
        raise ValueError(method_extreme)
    out = {
        "dat_anomaly": anom,
        "mask": mask.reshape(space),
        "thresholds": thr_out,
        "extreme_events": events.reshape(anom.shape),
        "time": time_o,
    }
    if std_normalise_flag and method_anomaly == "detrend_harmonic":  # detect.py:686-715
        stn, std = std_normalise(anom, doy_o)
        sub = preprocess_from_anomaly(stn, doy_o, method_extreme, threshold_percentile, window_days_hobday,
                                      window_spatial_hobday, method_percentile, precision, max_anomaly)
        out.update({"dat_stn": stn, "STD": std.reshape(space + (NDOY,)), "extreme_events_stn": sub[0], "thresholds_stn": sub[1]})
    return out


def preprocess_from_anomaly(anom, doy_o, method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday,
                            method_percentile, precision, max_anomaly):
    """``identify_extremes`` (detect.py:1119-1503) on an anomaly field: (events, thresholds in the reference's layout)."""
    anom = np.asarray(anom, dtype=np.float32)
    space = anom.shape[1:]
    gridded = len(space) == 2
    a2 = anom.reshape(anom.shape[0], -1)
    q = threshold_percentile / 100.0
    if method_extreme == "hobday_extreme":
        if method_percentile == "exact":
            thr_dm = hobday_thresholds_exact(a2, doy_o, threshold_percentile, window_days_hobday)
            thr_out = thr_dm.reshape((NDOY,) + space)
        else:
            ws = window_spatial_hobday
            if ws is None and gridded:
                ws = 5
            thr_cm = hobday_thresholds_approx(a2, doy_o, q, window_days_hobday, ws, space if gridded else None, precision, max_anomaly)
            thr_dm = np.ascontiguousarray(thr_cm.T)
            thr_out = thr_cm.reshape(space + (NDOY,))
        events = compare_hobday(a2, doy_o, thr_dm)
    else:
        thr = global_threshold_exact(a2, threshold_percentile) if method_percentile == "exact" else global_threshold_approx(a2, q, precision, max_anomaly)
        thr_out = thr.reshape(space)
        events = compare_global(a2, thr)
    return events.reshape(anom.shape), thr_out
