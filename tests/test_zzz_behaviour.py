"""The reference's own behavioural bars for the threshold stage, on the reference's own test inputs (numpy's legacy
seeded generator reproduces them exactly), run THROUGH THE CUDA KERNELS:

  * N(0,1) field, p90, ws = 1, window 11 / 21 / 41 days: every threshold finite, mean within 0.015 of 1.2816, day-to-day
    variation of a cell below 1 (/root/reference/tests/test_detect_helpers.py:641-690);
  * uniform field with a doubled December, precision 0.05, p90, 41-day window: the histogram quantile within three bin
    widths of the exact percentile at the days and cells the reference samples (:524-599);
  * p80 / p95 / p99 frequency of the detected events over the ocean cells (tests/conftest.py:168-231 bar: the requested
    fraction +- max(0.5 %, 20 % relative)), with the NaN column the reference tests inject, for a 3 x 4 grid and
    window_year_baseline = 2 as in /root/reference/tests/test_error_handling.py:512-542.

Each bar is written once against an "engine" and run twice: with the numpy oracle here on the CPU (which pins the bar's
numbers and the test's own logic), and with the CUDA library on the GPU."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import marex_oracle as mo  # noqa: E402


class OracleEngine:
    name = "oracle"

    def approx(self, a2, doy, grid, p, w, ws, precision=0.01, max_anomaly=5.0):
        return mo.hobday_thresholds_approx(a2, doy, p / 100.0, w, ws, grid, precision, max_anomaly)  # (N, 366)

    def exact(self, a2, doy, grid, p, w):
        return mo.hobday_thresholds_exact(a2, doy, p, w)  # (366, N)

    def preprocess(self, x, time, **kw):
        return mo.preprocess(x, time, **kw)


class CudaEngine:
    name = "cuda"

    def __init__(self):
        import torch

        if not torch.cuda.is_available():
            pytest.skip("needs a CUDA device")
        import marex_b200

        self.torch, self.mb = torch, marex_b200

    def _call(self, a2, doy, grid, *args, **kw):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return self.mb.identify_extremes_arrays(self.torch.from_numpy(np.ascontiguousarray(a2)).cuda(), doy, grid, "hobday_extreme", *args, **kw)

    def approx(self, a2, doy, grid, p, w, ws, precision=0.01, max_anomaly=5.0):
        res = self._call(a2, doy, grid, p, w, ws, "approximate", precision, max_anomaly)
        return res["thresholds"].cpu().numpy().reshape(-1, 366)

    def exact(self, a2, doy, grid, p, w):
        res = self._call(a2, doy, grid, p, w, None, method_percentile="exact")
        return res["thresholds"].cpu().numpy().reshape(366, -1)

    def preprocess(self, x, time, **kw):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return self.mb.preprocess_arrays(x, time, **kw)


ENGINES = [pytest.param(OracleEngine, id="oracle"), pytest.param(CudaEngine, id="cuda", marks=pytest.mark.gpu)]


def _daily(start, periods):
    return np.datetime64(start) + np.arange(periods).astype("timedelta64[D]")


@pytest.mark.parametrize("engine", ENGINES)
def test_normal_field_p90_window_sizes(engine):
    eng = engine()
    np.random.seed(456)  # the reference test's generator and shapes
    time = _daily("2019-01-01", 365 * 8)
    a = np.random.normal(0, 1, (len(time), 2, 2)).astype(np.float32)
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    for w in (11, 21, 41):
        thr = eng.approx(a2, doy, (2, 2), 90.0, w, 1)
        assert thr.shape == (4, 366)
        assert np.isfinite(thr).all()
        assert abs(float(thr.mean()) - 1.2816) < 0.015, (w, float(thr.mean()))
        assert float(np.std(np.diff(thr[0]))) < 1.0


@pytest.mark.parametrize("engine", ENGINES)
def test_histogram_quantile_within_three_bins_of_exact(engine):
    eng = engine()
    np.random.seed(42)
    lat, lon = 4, 5
    time = _daily("2020-01-01", 5 * 365)
    a = np.random.uniform(0, 1.5, (len(time), lat, lon))
    december = (time.astype("datetime64[M]").astype(int) % 12) == 11
    a[december] = a[december] * 2 + 1
    a = a.astype(np.float32)
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    precision = 0.05
    hist = eng.approx(a2, doy, (lat, lon), 90.0, 41, 1, precision, 5.0).reshape(lat, lon, 366)
    exact = eng.exact(a2, doy, (lat, lon), 90.0, 41).reshape(366, lat, lon)
    for lon_idx in range(lon):
        for day in (1, 200, 365):
            h, e = hist[2, lon_idx, day - 1], exact[day - 1, 2, lon_idx]
            assert abs(h - e) <= 3 * precision, (lon_idx, day, h, e)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("p", [80, 95, 99])
def test_percentile_frequency_on_a_3x4_grid(engine, p):
    eng = engine()
    rng = np.random.default_rng(7)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64("2000-01-01"))
    T = len(time)
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    x = (15 + 4 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((T, 3, 4))).astype(np.float32)
    x[:, 1, 1] = np.nan  # the NaN column of the reference tests
    r = eng.preprocess(x, time, threshold_percentile=p, window_year_baseline=2, smooth_days_baseline=11)
    ev, mask = np.asarray(r["extreme_events"]), np.asarray(r["mask"])
    assert ev.dtype == bool and mask.dtype == bool
    assert not mask[1, 1] and mask.sum() == 11 and not ev[:, 1, 1].any()
    want = 1 - p / 100.0
    freq = float(ev[:, mask].mean())
    assert abs(freq - want) <= max(0.005, 0.2 * want) + 0.004, (p, freq)  # (+ the sampling error of 11 cells x 8 years)
