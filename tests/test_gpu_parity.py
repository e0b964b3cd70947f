"""GPU parity tests: every stage of the CUDA path, called through the C-ABI (ctypes binding in
marex_b200/_lib.py), against the numpy oracle on the same seeded inputs.

Bars (north_star): histogram counts / masks / thresholds-from-identical-anomalies BIT-EXACT;
anomalies and end-to-end thresholds within 1e-5 relative of the field scale in float32.
"""
import functools
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
@pytest.fixture
def tune():
    """Pins tuning / test knobs of the library for one test (marex_tune), released afterwards."""
    import marex_b200

    pinned = []

    def _tune(**kw):
        pinned.extend(kw)
        marex_b200._lib.tune(**kw)

    yield _tune
    marex_b200._lib.tune(**{k: None for k in pinned})


def _trace(mb):
    calls = []
    mb._lib.TRACE = calls.append
    return calls


def _shift_case(mb, x, time, W, S, expect_daily=True, validate=True, atol_scale=1e-6):
    """The anomaly stage against the oracle.  Bars: NaN pattern, mask and trim identical; values within
    ``atol_scale`` x the field scale (float32 block sums + Kahan ring; the north-star bar is 1e-5)."""
    year, doy = mo.calendar_tables(time)
    ref, mask, keep = mo.anomaly_shifting_baseline(x, year, doy, W, S)
    cal = mb.detect.build_calendar(time)
    xd, _space = mb.detect._to_device_field(x, "cuda")
    calls = _trace(mb)
    try:
        res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S, validate=validate)
    finally:
        mb._lib.TRACE = None
    assert ("marex_shift_anomaly_daily_f32" in calls) == expect_daily, calls
    got = res["dat_anomaly"].cpu().numpy().reshape(ref.shape)
    np.testing.assert_array_equal(res["keep"], keep)
    np.testing.assert_array_equal(res["mask"].cpu().numpy().reshape(mask.shape), mask)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    scale = float(np.nanmax(np.abs(np.where(np.isfinite(x), x, np.nan))))
    np.testing.assert_allclose(got[fin], ref[fin], rtol=0, atol=atol_scale * scale)
    return got, ref


@pytest.mark.parametrize("W,S", [(5, 11), (3, 1), (4, 6), (15, 21)])
def test_shifting_baseline_anomaly_generic_kernel(W, S):
    """N = 7 * 37 is not a multiple of 4: the generic table-driven kernel (float64 sums, one rounding)."""
    mb = _cuda()
    x, time = _field(T1="2010-03-05" if W == 15 else "2001-07-01")
    if W == 15:
        time = np.arange(np.datetime64("1982-01-01"), np.datetime64("1982-01-01") + len(time))
    got, ref = _shift_case(mb, x, time, W, S, expect_daily=False)
    assert _frac_bits_differ(got, ref) < 1e-3


DAILY_SHAPES = [
    # (T0, T1, ny, nx, W, S, knobs)
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {}),                          # default shape, series ends mid-year
    ("1990-03-17", "2003-01-01", 8, 36, 3, 21, {}),                          # starts mid-year (doy0 != 1)
    ("1990-03-17", "2002-11-05", 4, 64, 1, 1, {}),                           # W = 1, S = 1
    ("1990-01-01", "1999-02-11", 4, 64, 3, 6, {}),                           # even S
    ("1982-01-01", "2002-03-05", 8, 36, 15, 21, {}),                         # the benchmark's windows
    ("1982-03-17", "2002-03-05", 4, 64, 15, 21, {}),                         # ... starting mid-year (constants-folded instantiation)
    ("1982-01-01", "2002-03-05", 8, 36, 15, 21, {"shift_generic": 1}),       # ... through the instantiation with run-time windows
    ("1988-02-29", "2001-03-01", 8, 36, 2, 3, {}),                           # starts on a leap day, W = 2
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {"shift_v": 1, "shift_r": 4}),
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {"shift_v": 2, "shift_r": 4}),
    ("1990-03-17", "2001-07-01", 8, 36, 5, 11, {"shift_v": 2, "shift_r": 4, "shift_cps": 1}),
    ("1990-03-17", "2001-07-01", 8, 36, 5, 11, {"shift_v": 2, "shift_r": 2, "shift_cps": 1}),
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {"shift_v": 1, "shift_r": 1}),
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {"shift_v": 1, "shift_r": 2, "shift_nw": 3}),
    ("1990-01-01", "2001-07-01", 8, 36, 5, 11, {"shift_f64": 1}),
    ("1982-01-01", "2002-03-05", 4, 36, 15, 21, {"shift_f64": 1, "shift_v": 2, "shift_r": 4}),
]


@pytest.mark.parametrize("T0,T1,ny,nx,W,S,knobs", DAILY_SHAPES)
@pytest.mark.parametrize("kelvin", [False, True])
def test_shifting_baseline_anomaly_daily_kernel(tune, T0, T1, ny, nx, W, S, knobs, kelvin):
    """The default (TMA-staged) kernel: N % 4 == 0.  Land column, a constant cell, and gridpoints that mix finite
    values with NaN / +-inf (they go through the fix-up list), for every kernel shape."""
    mb = _cuda()
    tune(**knobs)
    x, time = _field(T0=T0, T1=T1, ny=ny, nx=nx, seed=11, kelvin=kelvin)
    f = x.reshape(len(time), -1)
    f[100:140, 2] = np.nan          # a NaN gap
    f[2000, 3] = np.inf
    f[2500, 4] = -np.inf
    f[0, 6] = np.nan                # masked cell (first day NaN) with later finite data
    f[-1, 9] = np.nan               # last day
    f[:, 12] = np.nan               # a second land cell inside the same thread's vector
    got, ref = _shift_case(mb, x, time, W, S, expect_daily=True, validate=False)
    if knobs.get("shift_f64"):  # float64 sums, one rounding: what the oracle does
        fin = np.isfinite(ref)
        assert _frac_bits_differ(got[fin], ref[fin]) < 1e-3


@pytest.mark.parametrize("mode", ["anomaly", "climatology"])
def test_constants_folded_instantiation_is_bit_identical(tune, mode):
    """S = 21, W = 15 run an instantiation with the windows and the strip length as compile-time constants: same
    arithmetic in the same order as the run-time instantiation (shift_generic = 1), bit for bit."""
    mb = _cuda()
    x, time = _field(T0="1982-01-01", T1="2003-05-09", ny=8, nx=36, seed=4)
    f = x.reshape(len(time), -1)
    f[:, 5] = np.nan
    f[700:760, 8] = np.nan
    cal = mb.detect.build_calendar(time)
    xd, _space = mb.detect._to_device_field(x, "cuda")

    def run():
        if mode == "anomaly":
            r = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", 15, 21, validate=False)
            return r["dat_anomaly"].cpu().numpy()
        return mb.rolling_climatology_arrays(xd, time, 15, 21).cpu().numpy()

    a = run()
    tune(shift_generic=1)
    b = run()
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


def test_shifting_baseline_falls_back_to_generic_kernel(tune):
    """Windows that do not fit the staged kernel (W > 31) take the generic kernel instead of failing."""
    mb = _cuda()
    x, time = _field(T0="1960-01-01", T1="1995-01-01", ny=2, nx=4, seed=2)
    _shift_case(mb, x, time, 33, 5, expect_daily=False)


def test_shifting_baseline_nonfinite_and_gaps():
    """NaN / inf bookkeeping in the running sums, missing days and a missing year (generic kernel: gappy axis)."""
    mb = _cuda()
    x, time = _field(T1="2002-01-01", ny=3, nx=36, seed=3)
    x[100:140, 0, 2] = np.nan
    x[2000, 0, 3] = np.inf
    x[2500, 0, 4] = -np.inf
    x[0, 0, 6] = np.nan  # masked cell with later finite data
    sel = np.ones(len(time), bool)
    sel[400:430] = False  # a gap of days
    sel[(time >= np.datetime64("1995-01-01")) & (time < np.datetime64("1996-01-01"))] = False  # a missing year
    x, time = x[sel], time[sel]
    _shift_case(mb, x, time, 4, 7, expect_daily=False, validate=False, atol_scale=1e-5)


@pytest.mark.parametrize("ny,nx,T0", [(2, 33, "1990-01-01"), (2, 36, "1990-01-01"), (4, 64, "1990-03-17")])
def test_rolling_climatology_modes(ny, nx, T0):
    """mode 1 (the public rolling_climatology / smoothed_rolling_climatology) on the generic (N = 66) and the
    staged kernel (N % 4 == 0), with a mid-year start and a mixed finite / NaN gridpoint."""
    mb = _cuda()
    x, time = _field(T0=T0, T1="1999-06-01", ny=ny, nx=nx)
    x.reshape(len(time), -1)[50:60, 5] = np.nan
    year, doy = mo.calendar_tables(time)
    for W, S in [(3, 1), (3, 9), (1, 21)]:
        ref = mo.rolling_climatology(mo.smooth_centered(x, S), year, doy, W)
        calls = _trace(mb)
        try:
            got = mb.rolling_climatology_arrays(x, time, W, S).cpu().numpy()
        finally:
            mb._lib.TRACE = None
        assert ("marex_shift_anomaly_daily_f32" in calls) == ((ny * nx) % 4 == 0)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6 * 30, equal_nan=True)


def test_fused_bin_codes_match_digitize():
    """The bin codes the shifting-baseline kernel writes next to its anomalies (day-of-year major, one slot per
    output year) are np.digitize of those very anomalies, slot for slot; slots without a row hold the invalid code.
    Covers the fix-up gridpoints (re-digitized) and a series that ends mid-year."""
    mb = _cuda()
    x, time = _field(T0="1990-03-17", T1="2001-07-01", ny=8, nx=36, seed=5)
    f = x.reshape(len(time), -1)
    f[300:340, 2] = np.nan
    f[:, 40] += 8.0 * (np.arange(len(time)) % 7 == 0)  # anomalies beyond the last edge
    W, S = 4, 11
    cal = mb.detect.build_calendar(time)
    xd, _ = mb.detect._to_device_field(x, "cuda")
    edges, _ = mo.hobday_bins()
    res = mb.compute_normalised_anomaly_arrays(xd, cal, "shifting_baseline", W, S, validate=False, hobday_edges=edges)
    assert res["bins"] is not None
    anom = res["dat_anomaly"].cpu().numpy()
    keep = res["keep"]
    NY, slot_row = mb.calendar.doy_slots(cal.doy[keep], cal.year[keep])
    bins = res["bins"].cpu().numpy().view(np.uint16)[:, : anom.shape[1]]
    assert bins.shape[0] == 366 * NY
    want = np.full(bins.shape, 0x7FFF, dtype=np.uint16)
    dig = mo.digitize(anom, edges).astype(np.uint16)
    dig[dig >= len(edges) - 1] = 0x7FFF
    have = slot_row >= 0
    want[have] = dig[slot_row[have]]
    np.testing.assert_array_equal(bins, want)


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


@pytest.mark.parametrize("table", ["reference", "coarse", "slightly_irregular", "irregular"])
def test_digitize_bit_exact(table):
    """np.digitize - 1 for the reference's edge table (arithmetic fast path + look-ups near the edges), for another
    precision, and for tables that deviate from uniform spacing by 0.5 % (wide margin tier) and by 30 % of a step (walked)."""
    mb = _cuda()
    edges, _ = mo.hobday_bins() if table != "coarse" else mo.hobday_bins(0.05, 3.0)
    rng = np.random.default_rng(5)
    if table.endswith("irregular"):
        jit = rng.uniform(-1, 1, len(edges) - 1).astype(np.float32) * np.float32(0.00005 if table.startswith("slightly") else 0.003)
        edges = edges.copy()
        edges[1:] = np.sort(edges[1:] + jit)
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
    D = mb.detect
    e_d = D._up(edges, np.float32, a.device)
    # slots in a scrambled order, some without a row
    rng2 = np.random.default_rng(1)
    slot_row = np.concatenate([rng2.permutation(v.shape[0]), np.full(5, -1)]).astype(np.int32)
    rng2.shuffle(slot_row)
    s_d = D._up(slot_row, np.int32, a.device)
    bins = torch.empty((len(slot_row), N), dtype=torch.uint16, device="cuda")
    mb._lib.call("marex_digitize_doy_f32", D._p(a), N, N, D._p(s_d), len(slot_row), D._p(e_d), len(edges), D._p(bins), N, D._stream())
    want = mo.digitize(v, edges).astype(np.uint16)
    want[want >= len(edges) - 1] = 0x7FFF  # the kernels' code for "not counted"
    want = np.where((slot_row >= 0)[:, None], want[np.maximum(slot_row, 0)], np.uint16(0x7FFF))
    np.testing.assert_array_equal(bins.cpu().numpy().view(np.uint16), want)


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


@functools.lru_cache(maxsize=None)
def _exact_case(kind):
    from test_exact_queue_host import _field

    a, doy = _field(11, 25, 203, kind)  # 6 full groups of 32 gridpoints and a ragged one
    return a, doy, mo.hobday_thresholds_exact(a, doy, 95.0, 11)


@pytest.mark.parametrize("kind", ["plain", "seasonal", "special"])
@pytest.mark.parametrize("mode", ["queue", "all_listed", "half_listed", "histogram"])
def test_hobday_exact_queue_kernel_bit_exact(tune, kind, mode):
    """The queue kernel (exact_queue.cuh) on 25 years of windows with ties, gaps, infinities and a moving threshold; the
    groups it gives up on (here: forced) recomputed by the histogram kernel in list mode; and the histogram kernel alone."""
    mb = _cuda()
    a, doy, ref = _exact_case(kind)
    tune(exact_queue=0 if mode == "histogram" else 1, exact_force_fail={"all_listed": 1, "half_listed": 2}.get(mode, 0))
    n0 = mb._lib.launch_count()
    res = mb.identify_extremes_arrays(torch.from_numpy(a).cuda(), doy, None, "hobday_extreme", 95.0, 11, None, method_percentile="exact")
    got = res["thresholds"].cpu().numpy().reshape(366, -1)
    launched = mb._lib.launch_count() - n0
    _ulp_equal(got, ref)
    assert (np.isnan(got) == np.isnan(ref)).all()
    np.testing.assert_array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    listed = int(res["exact_scratch"][:1].view(torch.int32).item())  # [0] of the queue kernel's list of groups it gave up on
    groups = (a.shape[1] + 31) // 32
    if mode == "histogram":
        assert launched >= 3  # range pass (2 kernels) + histogram kernel (+ compare)
    elif mode == "all_listed":
        assert listed == groups
    elif mode == "half_listed":
        assert listed >= groups // 2
    elif kind != "special":
        assert listed == 0
    else:
        assert listed <= 1  # the group with the columns that are mostly infinities may be listed


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
        ({"pool_k": 64}, 5, 11, 95),
        ({"pool_k": 64, "pool_margin": 0}, 3, 5, 90),
        ({"pool_k": 128, "pool_ty": 3}, 7, 11, 99),
        ({"pool_force_fail": 1}, 5, 11, 95),
        ({"pool_ty": 1}, 5, 31, 80),
        ({"pool_ring": 1}, 5, 11, 95),
        ({"pool_ring": 1, "pool_tma": 0}, 7, 11, 99),
        ({"pool_ring": 1, "pool_force_fail": 1}, 3, 5, 90),
    ],
)
def test_banded_pooled_kernel_bit_exact(tune, env, ws, w, p):
    mb = _cuda()
    tune(**env)
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


@pytest.mark.parametrize("generic", [0, 1])
def test_banded_kernel_with_rows_per_day_folded(tune, generic):
    """15 rows per day of year select the band kernel instantiation with that count as a compile-time constant (no
    per-sample guards); pool_generic = 1 runs the same field through the run-time instantiation."""
    mb = _cuda()
    tune(pool_generic=generic)
    a, time, doy = _hetero_anoms(T1="2005-01-01", seed=12)
    ny, nx = a.shape[1:]
    a2 = a.reshape(len(time), -1)
    ref = mo.hobday_thresholds_approx(a2, doy, 0.95, 11, 5, (ny, nx))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = mb.identify_extremes_arrays(torch.from_numpy(a2).cuda(), doy, (ny, nx), "hobday_extreme", 95, 11, 5)
    _ulp_equal(res["thresholds"].cpu().numpy().reshape(-1, 366), ref)


def test_pooled_path_digitize_matches_numpy_digitize():
    """The pooled path digitizes with the invalid class coded 0x7FFF; its counts feed the same thresholds as
    np.digitize (samples exactly on an edge, one ulp below it, on the last edge, +-inf)."""
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
        # ragged cell count; a multiple of 4 for the shifting baseline so that every chunk takes the same (staged,
        # float32-sum) kernel as the one-piece call -- the generic kernel rounds float64 sums and differs in the last bit
        shifting = kw.get("method_anomaly", "shifting_baseline") == "shifting_baseline"
        x = np.ascontiguousarray(x.reshape(len(time), -1)[:, : 23 * 36 // 32 * 32 + (4 if shifting else 5)])
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


# ---------------------------------------------------------------- device argument
def test_non_current_device_is_honoured():
    """``device="cuda:1"`` (or tensors living on cuda:1) while cuda:0 is current: launches, streams and scratch buffers
    must follow the data.  Needs two GPUs."""
    mb = _cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from oracle import track_oracle as to

    x, time = _field(T1="2000-01-01", ny=6, nx=36, seed=8)
    kw = dict(window_year_baseline=4, smooth_days_baseline=9, window_days_hobday=5)
    torch.cuda.set_device(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = mb.preprocess_arrays(x, time, device="cuda:0", **kw)
        got = mb.preprocess_arrays(x, time, device="cuda:1", **kw)                       # host array, explicit device
        got_t = mb.preprocess_arrays(torch.from_numpy(x).to("cuda:1"), time, output="torch", **kw)  # tensor on cuda:1
    assert torch.cuda.current_device() == 0
    assert got_t["dat_anomaly"].device == torch.device("cuda:1")
    for k in ("dat_anomaly", "thresholds"):
        _ulp_equal(got[k], ref[k])
        _ulp_equal(got_t[k].cpu().numpy(), ref[k])
    np.testing.assert_array_equal(got["extreme_events"], ref["extreme_events"])
    ev, ocean = ref["extreme_events"].astype(bool), ref["mask"].astype(bool)
    filled = mb.MaskFiller(ocean, 2, 2, device="cuda:1").run(ev)
    np.testing.assert_array_equal(filled, to.stage1(ev, ocean, 2, 2))


def test_ring_kernel_tma_path_runs_and_is_bit_exact(tune):
    """The TMA-staged ring kernel on grids whose rows are TMA-eligible (nx % 8 == 0): homogeneous variance with a seasonal
    drift, so tiles keep their band (with re-centres) instead of falling back.  The kernel's progress markers (a
    page-locked host buffer, debug knob pool_dbg_ptr) prove that TMA tiles ran through all 366 days."""
    mb = _cuda()
    rng = np.random.default_rng(5)
    for ny, nx, T1, w in ((24, 72, "1998-01-01", 5), (13, 360, "2001-01-01", 11)):
        time = np.arange(np.datetime64("1990-01-01"), np.datetime64(T1))
        year, doy = mo.calendar_tables(time)
        season = 0.5 + 0.25 * np.cos(2 * np.pi * doy / 366.0)
        a = (rng.standard_normal((len(time), ny, nx)) * season[:, None, None]).astype(np.float32)
        f = a.reshape(len(time), -1)
        f[:, 5] = np.nan
        f[:, nx * 3 + 40] = np.nan
        f[::4, nx * 6 + 50] = 9.0
        dbg = torch.zeros(8 * 256, dtype=torch.int32).pin_memory()
        tune(pool_dbg_ptr=dbg.data_ptr(), pool_ring=1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = mb.identify_extremes_arrays(torch.from_numpy(f.copy()).cuda(), doy, (ny, nx), "hobday_extreme", 95, w, 5, year=year)
            torch.cuda.synchronize()
        tune(pool_dbg_ptr=None)
        m = dbg.view(-1, 8).numpy()
        ran = (m[:, 0] == 2) & (m[:, 1] == 1) & (m[:, 5] == 365)
        assert ran.any(), m[:16]
        ref = mo.hobday_thresholds_approx(f, doy, 0.95, w, 5, (ny, nx))
        _ulp_equal(res["thresholds"].cpu().numpy().reshape(-1, 366), ref)
        np.testing.assert_array_equal(res["extreme_events"].cpu().numpy(), mo.compare_hobday(f, doy, np.ascontiguousarray(ref.T)))


def test_synthetic_field_numpy_twin_matches_the_cuda_generator():
    """bench.py's CPU arms cut their tiles from the SAME synthetic field as the GPU arm through a numpy twin of the
    counter-based generator: identical land cells, values equal up to the float32 fast-math intrinsics of the kernel."""
    _cuda()
    from marex_b200 import synthetic

    time = synthetic.daily_time_axis("1991-01-01", "1993-03-01")
    ny, nx = 48, 80
    dev = synthetic.synth_sst(time, (ny, nx), rows=(8, 40), seed=2).cpu().numpy()
    cells = (np.arange(8, 40)[:, None] * nx + np.arange(nx)[None, :])
    host = synthetic.synth_sst_numpy(time, (ny, nx), cells, seed=2)
    np.testing.assert_array_equal(np.isnan(dev), np.isnan(host))
    assert 0.1 < np.isnan(host[0]).mean() < 0.5
    np.testing.assert_allclose(dev, host, rtol=0, atol=2e-3, equal_nan=True)


def test_pinned_outputs_do_not_alias_unless_asked():
    """``output="pinned"`` results own their page-locked buffers; only ``"pinned_reuse"`` hands out the cached ones."""
    mb = _cuda()
    x, time = _field(T1="1998-01-01", ny=4, nx=36, seed=9)
    kw = dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = mb.preprocess_arrays(x, time, output="pinned", **kw)
        keep = a["dat_anomaly"].copy()
        b = mb.preprocess_arrays(x + np.float32(1.0), time, output="pinned", **kw)
        c = mb.preprocess_arrays(x, time, output="pinned_reuse", **kw)
        d = mb.preprocess_arrays(x, time, output="pinned_reuse", **kw)
    assert not np.shares_memory(a["dat_anomaly"], b["dat_anomaly"])
    np.testing.assert_array_equal(a["dat_anomaly"], keep)
    assert np.shares_memory(c["dat_anomaly"], d["dat_anomaly"])
    with pytest.raises(ValueError, match="output must be"):
        mb.preprocess_arrays(x, time, output="host", **kw)
