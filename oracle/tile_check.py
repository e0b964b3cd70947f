"""TEST INFRASTRUCTURE (like everything under oracle/): checks a small tile cut from a full-size result of the CUDA
pipeline against the numpy restatement of the reference.  Used by tests/test_zz_fullsize_gpu.py and by bench.py's
post-run parity check; never by the product path.

At the benchmark sizes the oracle cannot redo the whole field, but gridpoints are independent except for the ws x ws
pooling of the approximate Hobday thresholds, so a tile is a self-contained problem:

  * anomalies and mask of the tile: the oracle's anomaly stage on the tile of the INPUT (within 1e-5 of the field scale,
    the north-star bar; NaN pattern and mask identical),
  * thresholds and events: the oracle's threshold + compare stages on the tile of the CUDA ANOMALIES (stage-wise parity,
    SURVEY.md F7), bit for bit on the tile interior -- gridpoints whose pooling neighbourhood lies inside the tile
    (every gridpoint when nothing is pooled).  A tile that straddles the longitude seam checks the periodic wrap.
"""
from typing import Any, Dict

import numpy as np

from . import marex_oracle as mo

_DEFAULTS = dict(
    method_anomaly="shifting_baseline", method_extreme="hobday_extreme", threshold_percentile=95, window_year_baseline=15,
    smooth_days_baseline=21, window_days_hobday=11, window_spatial_hobday=None, detrend_orders=(1,), force_zero_mean=True,
    reference_period=None, method_percentile="approximate", precision=0.01, max_anomaly=5.0,
)  # fmt: skip  (preprocess_data defaults, detect.py:287-313)


def check_tile(x_tile: np.ndarray, time: np.ndarray, got: Dict[str, Any], **kw) -> int:
    """``x_tile`` (T, h, w) or (T, n): the tile of the input.  ``got``: the matching tiles of the CUDA result --
    ``dat_anomaly`` (T_out, ...tile), ``thresholds`` in the reference's layout for the method ((...tile, 366),
    (366, ...tile) or (...tile)), ``extreme_events`` (T_out, ...tile) bool, ``mask`` (...tile) bool.  ``kw``: the
    ``preprocess_data`` arguments that differ from the defaults.  Returns the number of threshold values compared."""
    cfg = dict(_DEFAULTS, **kw)
    x_tile = np.asarray(x_tile, dtype=np.float32)
    space = x_tile.shape[1:]
    gridded = len(space) == 2
    year, doy = mo.calendar_tables(time)
    ma = cfg["method_anomaly"]
    if ma == "shifting_baseline":
        ref_anom, ref_mask, keep = mo.anomaly_shifting_baseline(x_tile, year, doy, cfg["window_year_baseline"], cfg["smooth_days_baseline"])
    elif ma == "fixed_baseline":
        ref_anom, ref_mask = mo.anomaly_fixed_baseline(x_tile, year, doy, cfg["reference_period"])
        keep = np.ones(len(doy), bool)
    elif ma == "detrend_fixed_baseline":
        ref_anom, ref_mask = mo.anomaly_detrend_fixed_baseline(x_tile, time, year, doy, cfg["detrend_orders"], cfg["force_zero_mean"], cfg["reference_period"])
        keep = np.ones(len(doy), bool)
    else:
        raise ValueError(ma)
    anom = np.asarray(got["dat_anomaly"], dtype=np.float32)
    np.testing.assert_array_equal(np.asarray(got["mask"]).astype(bool), np.asarray(ref_mask).reshape(space))
    np.testing.assert_array_equal(np.isnan(anom), np.isnan(ref_anom.reshape(anom.shape)))
    scale = float(np.nanmax(np.abs(x_tile))) if np.isfinite(x_tile).any() else 1.0
    np.testing.assert_allclose(anom, ref_anom.reshape(anom.shape), rtol=0, atol=1e-5 * scale, equal_nan=True)

    # thresholds and events from the SAME anomalies
    doy_out = doy[keep]
    hobday = cfg["method_extreme"] == "hobday_extreme"
    approx = cfg["method_percentile"] == "approximate"
    ws = cfg["window_spatial_hobday"]
    if ws is None and gridded and hobday and approx:
        ws = 5
    half = (ws // 2) if (gridded and hobday and approx and ws) else 0
    ev_ref, thr_ref = mo.preprocess_from_anomaly(
        anom, doy_out, cfg["method_extreme"], cfg["threshold_percentile"], cfg["window_days_hobday"],
        cfg["window_spatial_hobday"], cfg["method_percentile"], cfg["precision"], cfg["max_anomaly"],
    )  # fmt: skip
    inner = tuple(slice(half, n - half) for n in space) if half else tuple(slice(None) for _ in space)
    thr_got = np.asarray(got["thresholds"])
    if hobday and not approx:  # (366, ...space)
        g, r = thr_got[(slice(None),) + inner], thr_ref[(slice(None),) + inner]
    else:  # (...space, 366) or (...space)
        g, r = thr_got[inner], thr_ref[inner]
    g, r = np.ascontiguousarray(g), np.ascontiguousarray(r)
    assert g.dtype == r.dtype, (g.dtype, r.dtype)
    np.testing.assert_array_equal(np.isnan(g), np.isnan(r))
    ok = ~np.isnan(r)
    bits = np.uint32 if g.dtype == np.float32 else np.uint64
    np.testing.assert_array_equal(g[ok].view(bits), r[ok].view(bits))
    ev_got = np.asarray(got["extreme_events"]).astype(bool)
    np.testing.assert_array_equal(ev_got[(slice(None),) + inner], ev_ref.astype(bool)[(slice(None),) + inner])
    return int(ok.sum())
