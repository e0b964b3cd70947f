"""GPU parity tests: every stage of the CUDA path, called through the C-ABI (ctypes binding in
marex_b200/_lib.py), against the numpy oracle on the same seeded inputs.

Bars (north_star): histogram counts / masks / thresholds-from-identical-anomalies BIT-EXACT;
anomalies and end-to-end thresholds within 1e-5 relative of the field scale in float32.
"""
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import marex_oracle as mo  # noqa: E402


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import marex_b200

    return marex_b200


def _field(T0="1990-01-01", T1="2006-01-01", ny=7, nx=37, seed=0, kelvin=False):
    rng = np.random.default_rng(seed)
    time = np.arange(np.datetime64(T0), np.datetime64(T1))
    T = len(time)
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    amp = rng.uniform(0.5, 6, (ny, nx))
    ph = rng.uniform(0, 1, (ny, nx))
    x = 15 + amp * np.cos(2 * np.pi * (frac[:, None, None] - ph)) + 0.02 * np.arange(T)[:, None, None] / 365.25
    ar = np.zeros((ny, nx))
    noise = np.empty((T, ny, nx))
    for t in range(T):
        ar = 0.9 * ar + 0.26 * rng.standard_normal((ny, nx))
        noise[t] = ar
    x = (x + noise + (273.15 if kelvin else 0)).astype(np.float32)
    f = x.reshape(T, -1)  # view
    f[:, nx + 1 if ny > 1 else 1] = np.nan  # land column (reference tests inject one)
    f[:, 7] = np.float32(2.5)  # constant cell -> anomaly exactly 0 -> clamp rule
    return x, time


def _canon(a):
    """Bit pattern with every NaN mapped to one canonical payload (CUDA's NaN is 0x7fffffff, numpy's 0x7fc00000)."""
    a = np.ascontiguousarray(a)
    u = a.view(np.uint32 if a.dtype == np.float32 else np.uint64).copy()
    u[np.isnan(a)] = 0
    u[a == 0] = 0  # -0.0 == +0.0: which of two tied zeros a selection returns is unspecified in numpy too
    return u


def _ulp_equal(a, b):
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(_canon(a), _canon(b))


def _frac_bits_differ(a, b):
    return (_canon(a) != _canon(b)).mean()


# ---------------------------------------------------------------- (a) anomalies
@pytest.mark.parametrize("W,S", [(5, 11), (3, 1), (4, 6), (15, 21)])
def test_shifting_baseline_anomaly(W, S):
    mb = _cuda()
    x, time = _field(T1="2010-03-05" if W == 15 else "2001-07-01")
    if W == 15:
        time = np.arange(np.datetime64("1982-01-01"), np.datetime64("1982-01-01") + len(time))
    year, doy = mo.calendar_tables(time)
    ref, mask, keep = mo.anomaly_shifting_baseline(x, year, doy, W, S)
    cal = mb.detect.build_calendar(time)
    xd, space = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S)
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["keep"], keep)
    np.testing.assert_array_equal(res["mask"].cpu().numpy().reshape(mask.shape), mask)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    # tolerance: 1e-5 relative to the field scale (|x| ~ 30) -> 3e-4; in practice both sides round one
    # float64 result to float32, so they agree to the last bit almost everywhere.
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5 * 30, equal_nan=True)
    assert _frac_bits_differ(got, ref) < 1e-3


@pytest.mark.skipif(os.environ.get("MAREX_TEST_EXPERIMENTAL") != "1", reason="round-2 experiment: set MAREX_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("W,S,kelvin", [(5, 11, False), (15, 21, True)])
def test_shifting_baseline_anomaly_float32_sums(monkeypatch, W, S, kelvin):
    """MAREX_SHIFT_ACC=f32: float32 window sums + Kahan ring sum in the TMA kernel.  Not bit-identical to the oracle's
    float64 sums, but far inside the north-star tolerance (tests/test_f32_accumulation_study.py: 2e-7 of the field scale)."""
    mb = _cuda()
    monkeypatch.setenv("MAREX_SHIFT_ACC", "f32")
    x, time = _field(T1="2031-03-05" if W == 15 else "2001-07-01", kelvin=kelvin)
    year, doy = mo.calendar_tables(time)
    ref, mask, keep = mo.anomaly_shifting_baseline(x, year, doy, W, S)
    cal = mb.detect.build_calendar(time)
    xd, _space = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S)
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["keep"], keep)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    scale = float(np.nanmax(np.abs(x)))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6 * scale, equal_nan=True)  # ten times inside the 1e-5 bar


@pytest.mark.skipif(os.environ.get("MAREX_TEST_EXPERIMENTAL") != "1", reason="round-2 experiment: set MAREX_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("W,S", [(5, 11), (15, 21)])
def test_shifting_baseline_anomaly_lean_variant(monkeypatch, W, S):
    """MAREX_SHIFT_LEAN=1 (leap-year bit mask, 32-bit row arithmetic) changes no arithmetic: same bars as the default kernel,
    and bit-identical to it."""
    mb = _cuda()
    x, time = _field(T1="2010-03-05" if W == 15 else "2001-07-01")
    if W == 15:
        time = np.arange(np.datetime64("1982-01-01"), np.datetime64("1982-01-01") + len(time))
    year, doy = mo.calendar_tables(time)
    ref, _mask, _keep = mo.anomaly_shifting_baseline(x, year, doy, W, S)
    cal = mb.detect.build_calendar(time)
    xd, _space = mb.detect._to_device_field(x, "cuda")
    base = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S)["dat_anomaly"].cpu().numpy()
    monkeypatch.setenv("MAREX_SHIFT_LEAN", "1")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S)
    got = res["dat_anomaly"].cpu().numpy()
    np.testing.assert_array_equal(_canon(got), _canon(base))
    assert _frac_bits_differ(got.reshape(ref.shape), ref) < 1e-3


def test_shifting_baseline_nonfinite_and_gaps():
    """NaN / inf bookkeeping in the running sums, missing days and a missing year."""
    mb = _cuda()
    x, time = _field(T1="2002-01-01", ny=3, nx=33, seed=3)
    x[100:140, 0, 2] = np.nan
    x[2000, 0, 3] = np.inf
    x[2500, 0, 4] = -np.inf
    x[0, 0, 6] = np.nan  # masked cell with later finite data
    sel = np.ones(len(time), bool)
    sel[400:430] = False  # a gap of days
    sel[(time >= np.datetime64("1995-01-01")) & (time < np.datetime64("1996-01-01"))] = False  # a missing year
    x, time = x[sel], time[sel]
    year, doy = mo.calendar_tables(time)
    W, S = 4, 7
    ref, mask, keep = mo.anomaly_shifting_baseline(x, year, doy, W, S)
    cal = mb.detect.build_calendar(time)
    xd, _ = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S, validate=False)
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["keep"], keep)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=0, atol=3e-4)


def test_rolling_climatology_modes():
    mb = _cuda()
    x, time = _field(T1="1999-01-01", ny=2, nx=33)
    year, doy = mo.calendar_tables(time)
    for W, S in [(3, 1), (3, 9)]:
        ref = mo.rolling_climatology(mo.smooth_centered(x, S), year, doy, W)
        got = mb.rolling_climatology_arrays(x, time, W, S).cpu().numpy()
        np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(got, ref, rtol=0, atol=3e-4, equal_nan=True)


@pytest.mark.parametrize("period", [None, (1992, 1997)])
def test_fixed_baseline_anomaly(period):
    mb = _cuda()
    x, time = _field(T1="2000-01-01")
    year, doy = mo.calendar_tables(time)
    ref, mask = mo.anomaly_fixed_baseline(x, year, doy, period)
    cal = mb.detect.build_calendar(time)
    xd, _ = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "fixed_baseline", reference_period=period)
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["mask"].cpu().numpy().reshape(mask.shape), mask)
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-4, equal_nan=True)
    assert _frac_bits_differ(got, ref) < 1e-3


@pytest.mark.parametrize("orders,fzm,period", [([1], True, None), ([1, 2], False, None), ([1, 2, 3], True, (1991, 1995))])
def test_detrend_fixed_baseline_anomaly(orders, fzm, period):
    mb = _cuda()
    x, time = _field(T1="1999-06-01")
    year, doy = mo.calendar_tables(time)
    ref, mask = mo.anomaly_detrend_fixed_baseline(x, time, year, doy, orders, fzm, period)
    cal = mb.detect.build_calendar(time)
    xd, _ = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(
        xd, cal, "detrend_fixed_baseline", detrend_orders=orders, force_zero_mean=fzm, reference_period=period
    )
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["mask"].cpu().numpy().reshape(mask.shape), mask)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-4, equal_nan=True)


# ---------------------------------------------------------------- (b) thresholds on identical anomalies
def _anoms(seed=1, ny=6, nx=35, T1="2003-01-01", scale=1.0):
    rng = np.random.default_rng(seed)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64(T1))
    a = (rng.standard_normal((len(time), ny, nx)) * rng.uniform(0.2, 2.0, (ny, nx)) * scale).astype(np.float32)
    f = a.reshape(len(time), -1)  # view: special cells by flat index (works for any ny)
    f[:, 0] = np.nan
    f[:, 9] = 0.0  # constant zero -> threshold clamped to edges[3]
    f[:, 10] = -1.0  # all negative -> bin 0
    f[::3, 20] = 7.0  # values above max_anomaly are dropped
    f[5:9, 30] = np.nan  # NaNs after the first step: dropped from counts, no mask
    return a, time


def test_digitize_bit_exact():
    mb = _cuda()
    edges, _ = mo.hobday_bins()
    rng = np.random.default_rng(5)
    v = np.concatenate(
        [
            edges[1:].astype(np.float32),
            np.nextafter(edges[1:], np.float32(np.inf)),
            np.nextafter(edges[1:], np.float32(-np.inf)),
            rng.uniform(-1, 6, 20000).astype(np.float32),
            np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5.0, 4.9999995], np.float32),
        ]
    ).astype(np.float32)
    N = 33
    v = np.resize(v, (len(v) // N + 1, N)).astype(np.float32)
    a = torch.from_numpy(v).cuda()
    bins = torch.empty(v.shape, dtype=torch.uint16, device="cuda")
    D = mb.detect
    e_d = D._up(edges, np.float32, a.device)
    mb._lib.call("marex_digitize_f32", D._p(a), v.shape[0], N, N, D._p(e_d), len(edges), D._p(bins), N, D._stream())
    np.testing.assert_array_equal(bins.cpu().numpy(), mo.digitize(v, edges))


@pytest.mark.parametrize("ws,w,p", [(None, 11, 95), (1, 5, 90), (5, 11, 95), (3, 31, 99), (5, 3, 60)])
def test_hobday_approx_thresholds_bit_exact(ws, w, p):
    mb = _cuda()
    a, time = _anoms()
    _, doy = mo.calendar_tables(time)
    ny, nx = a.shape[1:]
    a2 = a.reshape(len(time), -1)
    eff_ws = 5 if ws is None else ws
    ref = mo.hobday_thresholds_approx(a2, doy, p / 100.0, w, eff_ws, (ny, nx))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, (ny, nx), "hobday_extreme", p, w, ws)
    got = res["thresholds"].cpu().numpy().reshape(-1, 366)
    _ulp_equal(got, ref)
    ev = mo.compare_hobday(a2, doy, np.ascontiguousarray(ref.T))
    np.testing.assert_array_equal(res["extreme_events"].cpu().numpy(), ev)
    assert int(res["count"]) == int(ev.sum())


def test_hobday_approx_unstructured_bit_exact():
    mb = _cuda()
    a, time = _anoms(ny=1, nx=70)
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    ref = mo.hobday_thresholds_approx(a2, doy, 0.95, 11, None, None)
    with pytest.warns(UserWarning):
        res = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, None, "hobday_extreme", 95, 11, None, want_bits=True)
    _ulp_equal(res["thresholds"].cpu().numpy(), ref)
    ev = mo.compare_hobday(a2, doy, np.ascontiguousarray(ref.T))
    np.testing.assert_array_equal(res["bits"].cpu().numpy().view(np.uint32), mo.pack_bits_time_major(ev))


@pytest.mark.parametrize("w,p", [(11, 95), (5, 99.5), (1, 50), (11, 10)])
def test_hobday_exact_thresholds_bit_exact(w, p):
    mb = _cuda()
    a, time = _anoms(T1="2001-01-01")
    a[:, 3, 3] = np.round(a[:, 3, 3], 1)  # heavy ties
    a[10, 3, 4] = np.inf
    a[20, 3, 5] = -np.inf
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    ref = mo.hobday_thresholds_exact(a2, doy, p, w)
    res = mb.identify_extremes_arrays(
        torch.from_numpy(a2).cuda(), doy, a.shape[1:], "hobday_extreme", p, w, None, method_percentile="exact"
    )
    got = res["thresholds"].cpu().numpy().reshape(366, -1)
    _ulp_equal(got, ref)
    np.testing.assert_array_equal(res["extreme_events"].cpu().numpy(), mo.compare_hobday(a2, doy, ref))


def test_hobday_exact_matches_reference_golden(golden_dir):
    """Against outputs of the reference's own _doy_percentiles (detect.py:1936-1942)."""
    mb = _cuda()
    g = np.load(os.path.join(golden_dir, "ref_doy_percentiles.npz"))
    data, doy = g["data"], g["doy"]
    a = torch.from_numpy(np.ascontiguousarray(data.T)).cuda()
    for i, key in enumerate(g["keys"]):
        w, p = str(key).split("|")
        res = mb.identify_extremes_arrays(a, doy, None, "hobday_extreme", float(p), int(w), None, method_percentile="exact")
        _ulp_equal(res["thresholds"].cpu().numpy().T, g[f"out_{i}"])


@pytest.mark.parametrize("p", [95, 60, 99])
def test_global_approx_thresholds_bit_exact(p):
    mb = _cuda()
    a, time = _anoms()
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    ref = mo.global_threshold_approx(a2, p / 100.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, a.shape[1:], "global_extreme", p)
    got = res["thresholds"].cpu().numpy().reshape(-1)
    _ulp_equal(got, ref)
    np.testing.assert_array_equal(res["extreme_events"].cpu().numpy(), mo.compare_global(a2, ref))


@pytest.mark.parametrize("p", [95, 50, 99.9])
def test_global_exact_thresholds_bit_exact(p):
    mb = _cuda()
    a, time = _anoms()
    a[:, 3, 3] = np.round(a[:, 3, 3], 1)
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    ref = mo.global_threshold_exact(a2, p)
    res = mb.identify_extremes_arrays(
        torch.from_numpy(a2).cuda(), doy, a.shape[1:], "global_extreme", p, method_percentile="exact"
    )
    got = res["thresholds"].cpu().numpy().reshape(-1)
    _ulp_equal(got, ref)


def test_rolling_histogram_quantile_reference_golden(golden_dir):
    """Feed per-cell (366 x 502) histograms of the reference golden set through the CUDA kernel by
    synthesising a bins array that realises exactly those counts, and compare with the outputs of
    the reference's own _rolling_histogram_quantile (detect.py:2465-2559)."""
    mb = _cuda()
    g = np.load(os.path.join(golden_dir, "ref_rolling_hist_quantile.npz"))
    names = [n for n in g["hist_names"] if not str(n).startswith("pooled")]
    hists = [g[f"hist_{n}"].astype(np.int64) for n in names]
    nb = hists[0].shape[1]
    n_rows = max(int(h.sum(axis=1).max()) for h in hists)
    N = len(hists)
    # row r of doy d holds, for cell c, the r-th sample of that (cell, doy) or the "dropped" bin nb
    bins = np.full((n_rows * 366, N), nb, dtype=np.uint16)
    doy = np.repeat(np.arange(1, 367), n_rows).astype(np.int16)
    for c, h in enumerate(hists):
        for d in range(366):
            vals = np.repeat(np.arange(nb), h[d])
            bins[d * n_rows : d * n_rows + len(vals), c] = vals
    D = mb.detect
    dev = torch.device("cuda")
    ptr, rows = D.doy_csr(doy)
    _, centers = mo.hobday_bins()
    bins_d = torch.from_numpy(bins.view(np.int16)).cuda()
    row0 = torch.zeros(N, dtype=torch.float32, device=dev)
    ptr_d, rows_d, cen_d = D._up(ptr, np.int32, dev), D._up(rows, np.int32, dev), D._up(centers, np.float32, dev)
    for w in (3, 11, 31):
        for q in (0.6, 0.9, 0.95, 0.99):
            thr = torch.empty((366, N), dtype=torch.float32, device=dev)
            stats = torch.empty(2, dtype=torch.float32, device=dev)
            mb._lib.call(
                "marex_hobday_thresholds_hist", D._p(bins_d), bins.shape[0], 1, N, N, D._p(ptr_d),
                D._p(rows_d), D.max_window_rows(ptr, w), D._p(cen_d), nb, w, 1,
                float(q), D._p(row0), float("-inf"), D._p(thr), D._p(stats), D._stream(),
            )  # fmt: skip
            got = thr.cpu().numpy()
            keys = list(g["case_keys"])
            for c, n in enumerate(names):
                ref = g[f"out_{keys.index(f'{n}|{w}|{q}')}"]
                _ulp_equal(got[:, c], ref)


# ---------------------------------------------------------------- end to end
@pytest.mark.parametrize(
    "kw",
    [
        dict(),
        dict(method_percentile="exact"),
        dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme"),
        dict(method_anomaly="fixed_baseline", method_extreme="hobday_extreme", window_spatial_hobday=3),
        dict(method_anomaly="fixed_baseline", method_extreme="global_extreme", method_percentile="exact"),
    ],
)
def test_preprocess_real_data_gridded(golden_dir, kw):
    """The reference's own OSTIA fixture subset (Kelvin SST, 40 years) end to end."""
    mb = _cuda()
    g = np.load(os.path.join(golden_dir, "sst_gridded_subset.npz"))
    x, time = g["sst"], g["time"]
    ref = mo.preprocess(x, time, **kw)
    got = mb.preprocess_arrays(x, time, **kw)
    assert got["dat_anomaly"].dtype == np.float32 and got["extreme_events"].dtype == bool
    assert got["thresholds"].dtype == ref["thresholds"].dtype and got["thresholds"].shape == ref["thresholds"].shape
    np.testing.assert_array_equal(got["time"], ref["time"])
    np.testing.assert_array_equal(got["mask"], ref["mask"])
    # anomalies: 1e-5 relative to the field scale (|SST| ~ 300 K -> 3e-3)
    np.testing.assert_allclose(got["dat_anomaly"], ref["dat_anomaly"], rtol=0, atol=1e-5 * 300, equal_nan=True)
    # thresholds: one count moved across a 0.01 bin edge shifts an interpolated threshold by <= ~1 bin
    np.testing.assert_allclose(got["thresholds"], ref["thresholds"], rtol=0, atol=0.011, equal_nan=True)
    # masks may differ only where |anomaly - threshold| is within the float tolerance
    diff = got["extreme_events"] != ref["extreme_events"]
    assert diff.mean() < 2e-3
    m = got["mask"]
    freq = got["extreme_events"][:, m].mean()
    assert 0.04 < freq < 0.06  # assert_percentile_frequency of the reference tests (5% +- 20%)


def test_preprocess_real_data_unstructured(golden_dir):
    mb = _cuda()
    g = np.load(os.path.join(golden_dir, "sst_unstructured_subset.npz"))
    x, time = g["sst"], g["time"]
    for kw in (dict(), dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme")):
        ref = mo.preprocess(x, time, **kw)
        got = mb.preprocess_arrays(x, time, **kw)
        np.testing.assert_array_equal(got["mask"], ref["mask"])
        np.testing.assert_allclose(got["dat_anomaly"], ref["dat_anomaly"], rtol=0, atol=3e-4, equal_nan=True)
        np.testing.assert_allclose(got["thresholds"], ref["thresholds"], rtol=0, atol=0.011, equal_nan=True)
        assert (got["extreme_events"] != ref["extreme_events"]).mean() < 2e-3


def test_stagewise_exact_from_gpu_anomalies(golden_dir):
    """Counts/masks are bit-exact when both sides start from the SAME anomalies."""
    mb = _cuda()
    g = np.load(os.path.join(golden_dir, "sst_gridded_subset.npz"))
    x, time = g["sst"], g["time"]
    got = mb.preprocess_arrays(x, time)
    a = got["dat_anomaly"]
    _, doy = mo.calendar_tables(got["time"])
    thr = mo.hobday_thresholds_approx(a.reshape(a.shape[0], -1), doy, 0.95, 11, 5, a.shape[1:])
    _ulp_equal(got["thresholds"].reshape(-1, 366), thr)
    ev = mo.compare_hobday(a.reshape(a.shape[0], -1), doy, np.ascontiguousarray(thr.T))
    np.testing.assert_array_equal(got["extreme_events"].reshape(a.shape[0], -1), ev)


def test_threshold_range_warnings():
    """UserWarnings of detect.py:2711-2730: constant-zero cell -> below range; huge anomalies -> above."""
    mb = _cuda()
    a, time = _anoms(ny=1, nx=40)
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    with pytest.warns(UserWarning, match="below expected range"):
        mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, None, "hobday_extreme", 95, 11, None)
    big = (a2 * 0 + 4.995).astype(np.float32)
    with pytest.warns(UserWarning, match="exceed expected range"):
        mb.identify_extremes_arrays(torch.from_numpy(big).cuda(), doy, None, "hobday_extreme", 95, 11, None)


def test_data_validation_errors():
    mb = _cuda()
    x, time = _field(T1="1998-01-01", ny=2, nx=33)
    bad = x.copy()
    bad[500:510, 0, 4] = np.nan
    with pytest.raises(mb.DataValidationError, match=r"Dataset contains 10 invalid values in 1 ocean locations"):
        mb.preprocess_arrays(bad, time, window_year_baseline=3)
    with pytest.raises(mb.DataValidationError, match="no valid"):
        mb.preprocess_arrays(np.full_like(x, np.nan), time, window_year_baseline=3)
    with pytest.raises(mb.DataValidationError, match="Insufficient data for shifting_baseline"):
        mb.preprocess_arrays(x, time, window_year_baseline=15)


# ---------------------------------------------------------------- banded pooled kernel: rebuilds, fall-back, edges
def _hetero_anoms(ny=13, nx=70, T1="2001-01-01", seed=11):
    """Anomalies whose spread varies strongly with season and position, so that the thresholds of a
    tile drift through (and beyond) one band: exercises re-centring and the full-range fall-back."""
    rng = np.random.default_rng(seed)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64(T1))
    T = len(time)
    _, doy = mo.calendar_tables(time)
    season = 0.25 + 1.0 * (1 + np.cos(2 * np.pi * doy / 366.0))  # 0.25 .. 2.25
    space = np.linspace(0.3, 2.2, nx)[None, :] * np.linspace(1.0, 1.6, ny)[:, None]
    a = rng.standard_normal((T, ny, nx)) * season[:, None, None] * space[None]
    a = a.astype(np.float32)
    f = a.reshape(T, -1)
    f[:, 3] = np.nan
    f[:, nx * 5 + 9] = np.nan
    f[:, nx * 2 + 4] = 0.0  # constant cell -> clamp
    f[::3, nx * 7 + 20] = 9.0  # values beyond the last edge are dropped
    f[0, nx * 9 + 30] = np.nan  # NaN on the first day -> masked threshold, but later samples still pool
    return a, time, doy


@pytest.mark.parametrize(
    "env,ws,w,p",
    [
        ({}, 5, 11, 95),
        ({"MAREX_POOL_K": "64"}, 5, 11, 95),
        ({"MAREX_POOL_K": "64", "MAREX_POOL_MARGIN": "0"}, 3, 5, 90),
        ({"MAREX_POOL_K": "128", "MAREX_POOL_TY": "3"}, 7, 11, 99),
        ({"MAREX_POOL_FORCE_FAIL": "1"}, 5, 11, 95),
        ({"MAREX_POOL_TY": "1"}, 5, 31, 80),
    ],
)
def test_banded_pooled_kernel_bit_exact(monkeypatch, env, ws, w, p):
    mb = _cuda()
    for k in ("MAREX_POOL_K", "MAREX_POOL_MARGIN", "MAREX_POOL_TY", "MAREX_POOL_FORCE_FAIL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    a, time, doy = _hetero_anoms()
    ny, nx = a.shape[1:]
    ref = mo.hobday_thresholds_approx(a.reshape(len(time), -1), doy, p / 100.0, w, ws, (ny, nx))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = mb.identify_extremes_arrays(
            torch.from_numpy(a.reshape(len(time), -1)).cuda(), doy, (ny, nx), "hobday_extreme", p, w, ws
        )
    got = res["thresholds"].cpu().numpy().reshape(-1, 366)
    _ulp_equal(got, ref)
    np.testing.assert_array_equal(
        res["extreme_events"].cpu().numpy(), mo.compare_hobday(a.reshape(len(time), -1), doy, np.ascontiguousarray(ref.T))
    )


def test_digitize_ffff_matches_numpy_digitize():
    """The pooled path digitizes internally (invalid class coded 0x7FFF); its counts feed the same
    thresholds as np.digitize - checked here through a band wide enough that nothing is pooled away."""
    mb = _cuda()
    rng = np.random.default_rng(3)
    time = np.arange(np.datetime64("1995-01-01"), np.datetime64("2001-01-01"))
    _, doy = mo.calendar_tables(time)
    a = (rng.standard_normal((len(time), 6, 33)) * 2.5).astype(np.float32)  # odd N: scalar digitize path
    edges, _ = mo.hobday_bins()
    a[5, 2, 7] = edges[200]  # exactly on an edge
    a[6, 2, 7] = np.nextafter(edges[200], np.float32(-np.inf))
    a[7, 2, 7] = edges[-1]
    a[8, 2, 7] = np.inf
    a[9, 2, 7] = -np.inf
    ref = mo.hobday_thresholds_approx(a.reshape(len(time), -1), doy, 0.9, 5, 3, (6, 33))
    res = mb.identify_extremes_arrays(torch.from_numpy(a.reshape(len(time), -1)).cuda(), doy, (6, 33), "hobday_extreme", 90, 5, 3)
    _ulp_equal(res["thresholds"].cpu().numpy().reshape(-1, 366), ref)


# ---------------------------------------------------------------- streamed host path == one-piece path
@pytest.mark.parametrize(
    "kw,unstructured",
    [
        (dict(window_year_baseline=4, smooth_days_baseline=9, window_days_hobday=5), False),
        (dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme", detrend_orders=[1, 2]), False),
        (dict(method_anomaly="fixed_baseline", method_percentile="exact", window_days_hobday=7), True),
        (dict(window_year_baseline=3, smooth_days_baseline=5, method_extreme="global_extreme", method_percentile="exact"), True),
    ],
)
def test_streamed_host_path_matches_one_piece(kw, unstructured):
    """Chunked, copy/compute-overlapped processing of a host array (latitude bands with the pooling
    halo re-loaded, or cell ranges) must reproduce the one-piece result bit for bit."""
    mb = _cuda()
    x, time = _field(T1="2000-01-01", ny=23, nx=36, seed=4)
    if unstructured:
        x = x.reshape(len(time), -1)[:, : 23 * 36 // 32 * 32 + 5]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = mb.preprocess_arrays(x, time, chunks=1, **kw)
        many = mb.preprocess_arrays(x, time, chunks=4, output="pinned", **kw)
    assert many["chunks"] >= 2 and "chunks" not in one
    for k in ("dat_anomaly", "thresholds"):
        _ulp_equal(np.asarray(many[k]), np.asarray(one[k]))
    np.testing.assert_array_equal(many["mask"], one["mask"])
    np.testing.assert_array_equal(many["extreme_events"], one["extreme_events"])
    assert many["extreme_count"] == one["extreme_count"] == int(one["extreme_events"].sum())
    assert many["thresholds_layout"] == one["thresholds_layout"] and many["attrs"] == one["attrs"]
    np.testing.assert_array_equal(many["time"], one["time"])


def test_streamed_host_path_validation_is_global():
    """A band that is all land must not raise; invalid values in an ocean cell of any band must."""
    mb = _cuda()
    x, time = _field(T1="1998-01-01", ny=12, nx=36, seed=6)
    x[:, :4, :] = np.nan  # the first chunk is entirely land
    kw = dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = mb.preprocess_arrays(x, time, chunks=3, **kw)
    assert not out["mask"][:4].any() and out["mask"][4:].any()
    x[700, 9, 5] = np.nan
    with pytest.raises(mb.DataValidationError, match="invalid values in 1 ocean locations"):
        mb.preprocess_arrays(x, time, chunks=3, **kw)


@pytest.mark.parametrize("p,T1,nx", [(95, "1995-06-24", 36), (80, "1992-09-27", 40), (97.5, "2001-01-01", 64), (99.9, "1994-02-08", 33)])
def test_global_approx_fast_path_ties_and_edges(p, T1, nx):
    """Fast global-histogram path: series lengths for which q*S is an integer (decided by the 1e-10
    slack, deferred to the float64-order kernel), quantile in the first / last bins, dropped and NaN
    samples, aligned (float4 compare) and unaligned widths."""
    mb = _cuda()
    a, time = _anoms(T1=T1, ny=3, nx=nx, seed=8)
    f = a.reshape(len(time), -1)
    f[:, 11] = 4.995  # quantile in the last bins
    f[:, 12] = np.linspace(-3, 4.99, len(time), dtype=np.float32)
    f[:, 13] = 5.0  # exactly the last edge: right-closed last bin
    f[::2, 14] = 6.0  # half of the samples out of range
    _, doy = mo.calendar_tables(time)
    ref = mo.global_threshold_approx(f, p / 100.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = mb.identify_extremes_arrays(torch.from_numpy(f).cuda(), doy, a.shape[1:], "global_extreme", p, want_bits=(nx % 32 == 0))
    _ulp_equal(res["thresholds"].cpu().numpy().reshape(-1), ref)
    ev = mo.compare_global(f, ref)
    np.testing.assert_array_equal(res["extreme_events"].cpu().numpy(), ev)
    assert int(res["count"]) == int(ev.sum())
    if nx % 32 == 0:
        np.testing.assert_array_equal(res["bits"].cpu().numpy().view(np.uint32), mo.pack_bits_time_major(ev))


# ---------------------------------------------------------------- detrend_harmonic (without std_normalise)
@pytest.mark.parametrize("orders,fzm,nx", [([1], True, 36), ([1, 2], False, 37), ([1, 2, 3], True, 40)])
def test_detrend_harmonic_anomaly(orders, fzm, nx):
    """Polynomial + annual / semi-annual harmonic fit (detect.py:2143-2224): same fit kernels with 4 more
    model columns; the detrended series is the anomaly, the mask comes from the raw first step."""
    mb = _cuda()
    x, time = _field(T1="1999-01-01", ny=5, nx=nx, seed=9)
    ref = mo.detrend(x, time, orders, fzm, remove_harmonics=True)
    cal = mb.detect.build_calendar(time)
    xd, _ = mb.detect._to_device_field(x, "cuda")
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "detrend_harmonic", detrend_orders=orders, force_zero_mean=fzm)
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5 * 30, equal_nan=True)
    assert _frac_bits_differ(got, ref) < 2e-2  # float64 dot products in a different order than numpy's BLAS
    np.testing.assert_array_equal(res["mask"].cpu().numpy().reshape(x.shape[1:]), np.isfinite(x[0]))


@pytest.mark.parametrize("extreme,chunks", [("hobday_extreme", 1), ("global_extreme", 3)])
def test_preprocess_detrend_harmonic(extreme, chunks):
    mb = _cuda()
    x, time = _field(T1="2000-01-01", ny=8, nx=36, seed=10)
    kw = dict(method_anomaly="detrend_harmonic", method_extreme=extreme, window_days_hobday=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = mb.preprocess_arrays(x, time, chunks=chunks, **kw)
    ref = mo.preprocess(x, time, **kw)
    np.testing.assert_allclose(got["dat_anomaly"], ref["dat_anomaly"], rtol=0, atol=3e-4, equal_nan=True)
    np.testing.assert_array_equal(got["mask"], ref["mask"])
    assert got["attrs"]["method_anomaly"] == "detrend_harmonic" and got["attrs"]["std_normalise"] is False
    assert got["attrs"]["preprocessing_steps"][0] == "Removed polynomial trend orders=[1] & seasonal cycle"
    # thresholds / events from the SAME anomalies must agree bit for bit (stage-wise parity)
    a2 = np.asarray(got["dat_anomaly"]).reshape(len(time), -1)
    _, doy = mo.calendar_tables(time)
    if extreme == "hobday_extreme":
        thr = mo.hobday_thresholds_approx(a2, doy, 0.95, 5, 5, x.shape[1:])
        _ulp_equal(np.asarray(got["thresholds"]).reshape(-1, 366), thr)
        ev = mo.compare_hobday(a2, doy, np.ascontiguousarray(thr.T))
    else:
        thr = mo.global_threshold_approx(a2, 0.95)
        _ulp_equal(np.asarray(got["thresholds"]).reshape(-1), thr)
        ev = mo.compare_global(a2, thr)
    np.testing.assert_array_equal(np.asarray(got["extreme_events"]).reshape(len(time), -1), ev)


# ---------------------------------------------------------------- std_normalise (detect.py:2257-2293, 686-715)
@pytest.mark.parametrize("extreme", ["global_extreme", "hobday_extreme"])
def test_std_normalise(extreme):
    mb = _cuda()
    x, time = _field(T1="2001-01-01", ny=6, nx=36, seed=12)
    kw = dict(method_anomaly="detrend_harmonic", method_extreme=extreme, window_days_hobday=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = mb.preprocess_arrays(x, time, std_normalise=True, **kw)
        ref = mo.preprocess(x, time, std_normalise_flag=True, **kw)
    for k in ("dat_stn", "STD", "extreme_events_stn", "thresholds_stn"):
        assert k in got, k
    assert got["attrs"]["std_normalise"] is True
    assert got["attrs"]["preprocessing_steps"][1] == "Normalised by 30-day rolling STD"
    # stage-wise: from the GPU's own detrended anomalies, STD and dat_stn within float32 rounding of the oracle
    _, doy = mo.calendar_tables(time)
    stn_ref, std_ref = mo.std_normalise(np.asarray(got["dat_anomaly"]), doy)
    np.testing.assert_array_equal(np.isnan(got["STD"].reshape(-1, 366)), np.isnan(std_ref))
    np.testing.assert_allclose(got["STD"].reshape(-1, 366), std_ref, rtol=2e-6, atol=0, equal_nan=True)
    np.testing.assert_allclose(got["dat_stn"], stn_ref, rtol=4e-6, atol=0, equal_nan=True)
    # thresholds / events of the standardised field: bit-exact from identical dat_stn
    ev_ref, thr_ref = mo.preprocess_from_anomaly(np.asarray(got["dat_stn"]), doy, extreme, 95, 5, None, "approximate", 0.01, 5.0)
    _ulp_equal(np.asarray(got["thresholds_stn"]), thr_ref)
    np.testing.assert_array_equal(np.asarray(got["extreme_events_stn"]), ev_ref)
    # and the end-to-end oracle agrees to tolerance
    np.testing.assert_allclose(got["dat_stn"], ref["dat_stn"], rtol=0, atol=2e-3, equal_nan=True)
    # std_normalise is ignored for the other baselines, as upstream
    other = mb.preprocess_arrays(x, time, std_normalise=True, method_anomaly="fixed_baseline", method_extreme="global_extreme")
    assert "dat_stn" not in other


def test_banded_kernel_long_series_many_rows_per_day():
    """More rows per day of year than the kernel stages / prefetches in one go (34 years: 34 rows > the 26
    prefetched and the 32 staged), through the un-trimmed fixed baseline."""
    mb = _cuda()
    rng = np.random.default_rng(21)
    time = np.arange(np.datetime64("1980-01-01"), np.datetime64("2014-01-01"))
    a = (rng.standard_normal((len(time), 5, 34)) * rng.uniform(0.3, 1.5, (5, 34))).astype(np.float32)
    a[:, 2, 5] = np.nan
    _, doy = mo.calendar_tables(time)
    a2 = a.reshape(len(time), -1)
    for ws, w in ((5, 11), (3, 5)):
        ref = mo.hobday_thresholds_approx(a2, doy, 0.95, w, ws, a.shape[1:])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, a.shape[1:], "hobday_extreme", 95, w, ws)
        _ulp_equal(res["thresholds"].cpu().numpy().reshape(-1, 366), ref)
        np.testing.assert_array_equal(
            res["extreme_events"].cpu().numpy(), mo.compare_hobday(a2, doy, np.ascontiguousarray(ref.T))
        )
    # exact path with the same row counts (window kept in shared memory: 11 x 34 rows)
    ref_e = mo.hobday_thresholds_exact(a2, doy, 95.0, 11)
    res_e = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, a.shape[1:], "hobday_extreme", 95, 11, None, "exact")
    _ulp_equal(res_e["thresholds"].cpu().numpy().reshape(366, -1), ref_e)


def test_preprocess_zarr_roundtrip(tmp_path):
    """zarr v2 store -> pinned host field -> pipeline -> zarr v2 group: same result as the array API."""
    mb = _cuda()
    from marex_b200 import io_zarr as zio

    x, time = _field(T1="1998-01-01", ny=6, nx=36, seed=13)
    src, dst = str(tmp_path / "in.zarr"), str(tmp_path / "out.zarr")
    zio.write_array(src, "sst", x.astype(np.float64), (40, 6, 36), ["time", "lat", "lon"])  # float64 store, cast on read
    days = (time - np.datetime64("1981-01-01")).astype(np.int64)
    zio.write_array(src, "time", (days * 86400).astype(np.int64), (100,), ["time"],
                    {"units": "seconds since 1981-01-01", "calendar": "proleptic_gregorian"})  # fmt: skip
    kw = dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = zio.preprocess_zarr(src, "sst", out_store=dst, **kw)
        ref = mb.preprocess_arrays(x, time, **kw)
    for k in ("dat_anomaly", "thresholds"):
        _ulp_equal(np.asarray(got[k]), np.asarray(ref[k]))
    np.testing.assert_array_equal(got["extreme_events"], ref["extreme_events"])
    back = zio.read_array(dst, "extreme_events")
    assert back.dtype == np.bool_
    np.testing.assert_array_equal(back, ref["extreme_events"])
    _ulp_equal(zio.read_array(dst, "thresholds"), np.asarray(ref["thresholds"]))
    _, t_back, dims = zio.read_field(dst, "dat_anomaly", pinned=False)
    assert dims == ["time", "lat", "lon"]
    np.testing.assert_array_equal(t_back.astype("datetime64[D]"), ref["time"])


@pytest.mark.parametrize("kw", [dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5),
                                dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme")])
def test_model_calendar_noleap(kw):
    """A CF `noleap` axis: generic table-driven kernels, same arithmetic as the oracle on the
    (year, day-of-year, decimal-year) tables."""
    mb = _cuda()
    from marex_b200 import calendar as cal

    rng = np.random.default_rng(31)
    T = 9 * 365
    mt = cal.model_time_from_cf(np.arange(T), "days since 1990-01-01", "noleap")
    frac = (mt.doy - 1) / 365.0
    x = (12 + 3 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((T, 5, 36))).astype(np.float32)
    x[:, 1, 1] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = mb.preprocess_arrays(x, mt, **kw)
        ref = mo.preprocess(x, np.arange(T), year_doy=(mt.year, mt.doy, mt.decimal_year), **kw)
    np.testing.assert_array_equal(np.isnan(got["dat_anomaly"]), np.isnan(ref["dat_anomaly"]))
    np.testing.assert_allclose(got["dat_anomaly"], ref["dat_anomaly"], rtol=0, atol=3e-4, equal_nan=True)
    np.testing.assert_array_equal(got["mask"], ref["mask"])
    np.testing.assert_array_equal(got["time"], ref["time"])
    # stage-wise: thresholds and events from the GPU's own anomalies, bit for bit
    keep_doy = mt.doy[np.isin(np.arange(T), got["time"])]
    ev, thr = mo.preprocess_from_anomaly(np.asarray(got["dat_anomaly"]), keep_doy, kw.get("method_extreme", "hobday_extreme"),
                                         95, kw.get("window_days_hobday", 11), None, "approximate", 0.01, 5.0)  # fmt: skip
    _ulp_equal(np.asarray(got["thresholds"]), thr)
    np.testing.assert_array_equal(np.asarray(got["extreme_events"]), ev)
