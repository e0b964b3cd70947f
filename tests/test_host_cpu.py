"""CPU-only tests (`-m "not gpu"`): the C-ABI library loads and exports every symbol that
include/marex_b200.h declares, the host-side mirror of the reference's validation / calendar /
attrs logic behaves like detect.py, the oracle obeys the reference's own loose test bars, and
the latitude-band sharding (world_size 2, gloo) reproduces the unsharded result.

No kernel is launched here (there is no GPU on the build box)."""
import ctypes
import json
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import marex_oracle as mo  # noqa: E402


# ------------------------------------------------------------------ C-ABI
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "marex_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(marex_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_what_the_binding_uses():
    from marex_b200 import _lib

    assert sorted(_lib.EXPORTED) == _declared_symbols()


def test_library_builds_loads_and_exports_every_declared_symbol():
    from marex_b200 import _build, _lib

    path = _build.build(force=False)
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/marex_b200.h but not exported"
    assert _lib.load().marex_version() >= 100
    assert _lib.load().marex_last_error() == b""
    assert _lib.launch_count() == 0  # nothing was launched by loading


def test_library_is_sm100a_with_tma():
    """The shipped cubin targets sm_100a and the staged kernels use the bulk-copy (TMA) engine."""
    from marex_b200 import _build

    path = _build.build(force=False)
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
    # SASS evidence of the TMA path of the shifting-baseline kernel: UTMALDG (cp.async.bulk.tensor) and the
    # mbarrier transaction-count arrive (B200_PROFILING.md "What proves a Blackwell-native kernel")
    obj = os.path.join(os.path.dirname(path), "build", "shift_daily.o")
    if not os.path.exists(obj):
        _build.build(force=True)
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "UTMALDG.2D" in sass and "SYNCS.ARRIVE.TRANS64" in sass


def test_library_holds_the_constants_folded_instantiations():
    """The default windows run instantiations with S / W / strip length, rows per day of year and the day-of-year window
    as compile-time constants; losing one silently would cost 10-14 % of its kernel's time."""
    from marex_b200 import _build

    path = _build.build(force=False)
    out = subprocess.run(["cuobjdump", "--dump-elf-symbols", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    for mangled in (
        "shift_daily_kernelILi1ELi4ELi2ELi0EfLb0ELi21ELi15ELi48E",  # anomaly, mode 0
        "shift_daily_kernelILi1ELi4ELi2ELi1EfLb0ELi21ELi15ELi48E",  # climatology, mode 1
        "hobday_band_kernelILi2ELi64ELi16ELi25E",
        "hobday_band_kernelILi2ELi64ELi16ELi15E",
        "hobday_exact_queue_kernelILi64ELi11E",
    ):
        assert mangled in out.stdout, mangled


def test_no_product_import_of_the_oracle():
    """The product path must not route through the oracle (test infrastructure only)."""
    pkg = os.path.join(ROOT, "marex_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


def test_product_fails_loudly_without_cuda():
    import torch

    import marex_b200

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    time = np.arange(np.datetime64("2000-01-01"), np.datetime64("2004-01-01"))
    x = np.zeros((len(time), 2, 3), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        marex_b200.preprocess_arrays(x, time, window_year_baseline=2)


# ------------------------------------------------------------------ calendar / host tables
def test_calendar_tables_and_csr():
    from marex_b200 import calendar as cal

    time = np.arange(np.datetime64("1999-12-30"), np.datetime64("2001-01-03"))
    c = cal.build_calendar(time)
    y, d = mo.calendar_tables(time)
    np.testing.assert_array_equal(c.year, y)
    np.testing.assert_array_equal(c.doy, d)
    assert list(c.year_val) == [1999, 2000, 2001] and c.tidx.shape == (3 * 366,)
    assert c.tidx[0 * 366 + 363] == 0 and c.tidx[1 * 366 + 365] == 2 + 365  # 1999-12-30 ; 2000-12-31 (leap)
    assert (c.tidx[2 * 366 + 2 :] == -1).all()
    ptr, rows = cal.doy_csr(c.doy)
    assert ptr[-1] == len(time)
    for dd in (1, 2, 60, 365, 366):
        np.testing.assert_array_equal(np.sort(rows[ptr[dd - 1] : ptr[dd]]), np.nonzero(d == dd)[0])
    assert cal.max_window_rows(ptr, 11) == int(max(sum((d == ((k - 1 + o) % 366) + 1).sum() for o in range(-5, 6)) for k in range(1, 367)))
    np.testing.assert_array_equal(cal.decimal_year(time), mo.decimal_year(time))
    out_row, keep = cal.shifting_out_rows(c, 1)
    assert keep.sum() == (y >= 2000).sum() and out_row[keep][0] == 0 and (out_row[~keep] == -1).all()
    with pytest.raises(NotImplementedError):
        cal.build_calendar(np.array(["2000-01-01T00", "2000-01-01T12"], dtype="datetime64[h]"))


def test_model_calendars_noleap_360_all_leap():
    """year / day-of-year tables for CF model calendars (what `.dt.year/.dt.dayofyear` give on a
    cftime index), decimal year, and the kernels' tables built from them."""
    from marex_b200 import calendar as cal

    mt = cal.model_time_from_cf(np.arange(0, 800), "days since 2000-01-01", "noleap")
    assert mt.days_in_year == 365 and mt.year[0] == 2000 and mt.doy[0] == 1
    assert (mt.year[364], mt.doy[364], mt.year[365], mt.doy[365]) == (2000, 365, 2001, 1)
    assert mt.doy.max() == 365 and mt.year[-1] == 2002 and mt.doy[-1] == 800 - 2 * 365
    m360 = cal.model_time_from_cf(np.array([0.0, 29.5, 30.0, 359.0, 360.0]) * 24, "hours since 1850-02-01 00:00:00", "360_day")
    assert list(m360.doy) == [31, 60, 61, 30, 31] and list(m360.year) == [1850, 1850, 1850, 1851, 1851]
    mal = cal.model_time_from_cf(np.array([59, 60, 365, 366]), "days since 1999-01-01", "all_leap")
    assert list(mal.doy) == [60, 61, 366, 1] and list(mal.year) == [1999, 1999, 1999, 2000]
    np.testing.assert_allclose(mt.decimal_year[[0, 365]], [2000.0, 2001.0])
    c = cal.build_calendar(mt)
    assert not c.is_daily and c.n_years == 3 and c.T == 800 and c.model_dy is not None
    assert c.tidx[0] == 0 and c.tidx[366 + 0] == 365 and c.tidx[365] == -1  # doy 366 never occurs on noleap
    np.testing.assert_array_equal(c.decimal_year, mt.decimal_year)
    with pytest.raises(NotImplementedError):
        cal.model_time_from_cf([0], "days since 2000-01-01", "julian")
    with pytest.raises(NotImplementedError):  # two samples on one model day
        cal.build_calendar(cal.model_time_from_cf([0, 12], "hours since 2000-01-01", "noleap"))


def test_bin_tables_match_oracle_and_reference_expression():
    from marex_b200 import detect as d

    e, c = d.hobday_bins()
    eo, co = mo.hobday_bins()
    np.testing.assert_array_equal(e.view(np.uint32), eo.view(np.uint32))
    np.testing.assert_array_equal(c.view(np.uint32), co.view(np.uint32))
    e64, c64 = d.global_bins()
    eo64, co64 = mo.global_bins()
    np.testing.assert_array_equal(e64, eo64)
    np.testing.assert_array_equal(c64, co64)
    assert e.dtype == np.float32 and e64.dtype == np.float64 and len(e) == 503


def test_detrend_model_is_the_reference_design_matrix():
    from marex_b200 import detect as d

    time = np.arange(np.datetime64("1990-01-01"), np.datetime64("1996-01-01"))
    M, P = d.detrend_model(time, [1, 2])
    assert M.shape == (3, len(time)) and P.shape == (len(time), 3)
    np.testing.assert_allclose(M[0], 1.0)
    np.testing.assert_allclose(M[1:].mean(axis=1), 0.0, atol=1e-9)  # detect.py:2165-2166
    np.testing.assert_allclose(P.T @ M.T, np.eye(3), atol=1e-9)


# ------------------------------------------------------------------ configuration rules (reference regexes)
def test_configuration_errors_follow_the_reference_messages():
    from marex_b200 import ConfigurationError
    from marex_b200 import detect as d

    rc = d.resolve_extreme_config
    assert rc("hobday_extreme", 95, 11, None, "approximate", 0.01, 5.0, True) == 5  # detect.py:1451-1452
    assert rc("hobday_extreme", 95, 11, None, "approximate", 0.01, 5.0, False) is None
    assert rc("hobday_extreme", 95, 11, 3, "approximate", 0.01, 5.0, True) == 3
    cases = [
        (("hobday_extreme", 95, 11, None, "bogus", 0.01, 5.0, True), "Unknown method_percentile"),
        (("hobday_extreme", 95, 11, None, "exact", 0.02, 5.0, True), "'precision' cannot be used"),
        (("hobday_extreme", 95, 11, None, "exact", 0.01, 4.0, True), "'max_anomaly' cannot be used"),
        (("hobday_extreme", 50, 11, None, "approximate", 0.01, 5.0, True), "not supported with method_percentile='approximate'"),
        (("hobday_extreme", 95, 11, 5, "approximate", 0.01, 5.0, False), "not supported for unstructured grids"),
        (("global_extreme", 95, 11, 5, "approximate", 0.01, 5.0, True), "can only be used with method_extreme='hobday_extreme'"),
        (("hobday_extreme", 95, 11, 5, "exact", 0.01, 5.0, True), "not supported with method_percentile='exact'"),
        (("hobday_extreme", 95, 10, None, "approximate", 0.01, 5.0, True), "window_days_hobday must be an odd number"),
        (("hobday_extreme", 95, 11, 4, "approximate", 0.01, 5.0, True), "window_spatial_hobday must be an odd number"),
        (("nope", 95, 11, None, "approximate", 0.01, 5.0, True), "Unknown extreme method"),
    ]
    for args, msg in cases:
        with pytest.raises(ConfigurationError, match=re.escape(msg)):
            rc(*args)
    with pytest.raises(ConfigurationError, match="Unknown anomaly method"):
        d.validate_anomaly_method("nope")
    with pytest.raises(ConfigurationError, match="reference_period is not supported"):
        d.validate_reference_period_method((1990, 2000), "shifting_baseline")
    with pytest.raises(ConfigurationError, match="detrend_orders cannot be empty"):
        d.validate_detrend_orders([])
    with pytest.raises(ConfigurationError, match="Invalid polynomial orders"):
        d.validate_detrend_orders([1, 0])
    years = np.arange(1990, 2000)
    with pytest.raises(ConfigurationError, match="start year"):
        d.validate_reference_period((1995, 1991), years)
    with pytest.raises(ConfigurationError, match="No data found in reference_period"):
        d.validate_reference_period((1800, 1801), years)
    np.testing.assert_array_equal(d.validate_reference_period((1992, 1993), years), [2, 3])


def test_detrend_harmonic_model_and_attrs():
    """detrend_harmonic: 4 harmonic columns after the polynomial ones (detect.py:2150-2159); attrs of detect.py:751-758."""
    from marex_b200 import detect as d

    time = np.arange(np.datetime64("1990-01-01"), np.datetime64("1994-01-01"))
    M, P = d.detrend_model(time, [1], remove_harmonics=True)
    Mo, Po = mo.detrend_model(time, [1], True)
    assert M.shape == (6, len(time))
    np.testing.assert_array_equal(M, Mo)
    np.testing.assert_array_equal(P, Po)
    d.validate_anomaly_method("detrend_harmonic")
    a = d._dataset_attrs("detrend_harmonic", "global_extreme", 95, False, [1], 15, 21, 11, None, None, True, "approximate", 0.01, 5.0)
    assert a["detrend_orders"] == [1] and a["force_zero_mean"] is True and a["std_normalise"] is False
    assert a["preprocessing_steps"] == ["Removed polynomial trend orders=[1] & seasonal cycle", "Global percentile threshold applied to all days"]


def test_preprocessing_steps_match_reference_golden(golden_dir):
    from marex_b200 import detect as d

    cases = json.load(open(os.path.join(golden_dir, "ref_preprocessing_steps.json")))
    assert len(cases) > 5
    for c in cases:
        kw = dict(c["args"])
        if kw.get("reference_period") is not None:
            kw["reference_period"] = tuple(kw["reference_period"])
        assert d.get_preprocessing_steps(**kw) == c["out"], kw


def test_dims_coords_inference_duck_typed():
    from marex_b200 import DataValidationError, xr_api

    class DA:
        def __init__(self, dims, coords):
            self.dims, self.coords = dims, coords

    da = DA(("time", "lat", "lon"), {"time": 0, "lat": 0, "lon": 0})
    dims, coords = xr_api._infer_dims_coords(da, None, None)
    assert dims == {"time": "time", "x": "lon", "y": "lat"} and coords == dims
    with pytest.raises(DataValidationError, match="Missing required dimensions"):
        xr_api._infer_dims_coords(da, {"x": "xx", "y": "lat"}, None)
    with pytest.raises(DataValidationError, match="Missing required coordinates"):
        xr_api._infer_dims_coords(da, None, {"x": "nlon", "y": "lat"})
    un = DA(("time", "ncells"), {"time": 0, "lat": 0, "lon": 0})
    with pytest.raises(DataValidationError, match="must be explicitly specified for unstructured"):
        xr_api._infer_dims_coords(un, {"x": "ncells"}, None)
    dims, coords = xr_api._infer_dims_coords(un, {"x": "ncells"}, {"x": "lon", "y": "lat"})
    assert "y" not in dims and coords["time"] == "time"
    with pytest.raises(KeyError):  # accidental upstream behaviour pinned by tests/test_error_handling.py:171-181
        xr_api._space_dims({"time": "time", "y": "lat"}) and xr_api._infer_dims_coords(da, {"y": "lat"}, None)[0]["x"]


# ------------------------------------------------------------------ the oracle against the reference's own (loose) bars
def test_oracle_hist_quantile_vs_numpy_percentile_reference_bars():
    """tests/test_detect_helpers.py:171-278 (1-D, atol 0.05) and :590-599, :676-686 (2-D: approx vs
    exact within 3 bin widths; N(0,1) p90 mean 1.2816 +- 0.015)."""
    rng = np.random.default_rng(42)
    time = np.arange(np.datetime64("2000-01-01"), np.datetime64("2012-01-01"))
    year, doy = mo.calendar_tables(time)
    a = rng.standard_normal((len(time), 30)).astype(np.float32)
    thr1 = mo.global_threshold_approx(a, 0.95)
    np.testing.assert_allclose(thr1, np.percentile(a.astype(np.float64), 95, axis=0), atol=0.05)
    approx = mo.hobday_thresholds_approx(a, doy, 0.90, 31, 1, None)  # (N, 366)
    exact = mo.hobday_thresholds_exact(a, doy, 90.0, 31)  # (366, N)
    assert approx.shape == (30, 366) and exact.shape == (366, 30)
    assert np.nanmax(np.abs(approx - exact.T)) < 0.2  # sampling noise dominates at 341 samples per window
    assert abs(float(np.nanmean(approx)) - 1.2816) < 0.03


def test_oracle_extreme_frequency_bounds():
    """tests/conftest.py:168-231 `assert_percentile_frequency`: 5 % +- max(0.5 %, 20 % relative)."""
    rng = np.random.default_rng(1)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64("2002-01-01"))
    T = len(time)
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    x = (15 + 4 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((T, 4, 6))).astype(np.float32)
    x[:, 1, 1] = np.nan  # NaN column injected by the reference tests (test_gridded_preprocessing.py:22-25)
    for kw in (dict(), dict(method_extreme="global_extreme"), dict(method_anomaly="detrend_fixed_baseline")):
        r = mo.preprocess(x, time, window_year_baseline=4, smooth_days_baseline=11, **kw)
        ev, mask = r["extreme_events"], r["mask"]
        assert ev.dtype == bool and mask.dtype == bool and not mask[1, 1] and mask.sum() == 23
        freq = ev[:, mask].mean()
        assert abs(freq - 0.05) < 0.012, (kw, freq)
        assert not ev[:, 1, 1].any()


def test_oracle_chunking_invariance_is_sharding_invariance():
    """The reference pins identical extreme counts across time chunkings (tests/test_integration.py:176-226);
    the analogue here is that a latitude-band split with halo reproduces the unsharded result bit for bit."""
    from marex_b200 import sharding

    rng = np.random.default_rng(5)
    time = np.arange(np.datetime64("1995-01-01"), np.datetime64("2003-01-01"))
    x = (10 + rng.standard_normal((len(time), 9, 7))).astype(np.float32)
    kw = dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5)
    full = mo.preprocess(x, time, **kw)
    parts = []
    for rank in range(3):
        out = sharding.preprocess_sharded(lambda lo, hi: x[:, lo:hi], time, 9, rank, 3, mo.preprocess, **kw)
        parts.append(out)
    thr = np.concatenate([p["thresholds"] for p in parts], axis=0)
    ev = np.concatenate([p["extreme_events"] for p in parts], axis=1)
    np.testing.assert_array_equal(thr.view(np.uint32), full["thresholds"].view(np.uint32))
    np.testing.assert_array_equal(ev, full["extreme_events"])
    assert [p["rows"] for p in parts] == [(0, 3), (3, 6), (6, 9)]


def test_lat_band_and_cell_range_partition():
    from marex_b200 import sharding

    for ny, world in [(720, 8), (10, 3), (5, 5), (1801, 8)]:
        cover = []
        for r in range(world):
            lo, hi, llo, lhi = sharding.lat_band(ny, world, r, 2)
            assert llo == max(0, lo - 2) and lhi == min(ny, hi + 2)
            cover += list(range(lo, hi))
        assert cover == list(range(ny))
    assert [sharding.cell_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.effective_halo("hobday_extreme", "approximate", None, True) == 2
    assert sharding.effective_halo("hobday_extreme", "exact", None, True) == 0
    assert sharding.effective_halo("global_extreme", "approximate", None, True) == 0
    assert sharding.effective_halo("hobday_extreme", "approximate", 7, True) == 3
    assert sharding.effective_halo("hobday_extreme", "approximate", None, False) == 0


def test_streamed_chunk_bounds_cover_the_field_once():
    """Chunks of the streamed host path: contiguous, disjoint, complete, interior boundaries aligned."""
    from marex_b200 import detect as d

    for n_units, n_chunks, align in [(720, 10, 1), (7, 16, 1), (1 << 20, 8, 32), (1000, 3, 32), (33, 4, 32), (5, 1, 1)]:
        b = d._chunk_bounds(n_units, n_chunks, align)
        assert b[0][0] == 0 and b[-1][1] == n_units and len(b) <= max(1, n_chunks)
        for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
            assert a1 == b0 and a0 < a1 and b0 % align == 0
        assert all(hi > lo for lo, hi in b)


# ------------------------------------------------------------------ world_size-2 gloo run of the N > 1 path
_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MAREX_ROOT"])
from marex_b200 import sharding
from oracle import marex_oracle as mo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
time = np.arange(np.datetime64("1996-01-01"), np.datetime64("2003-01-01"))
x = (12 + rng.standard_normal((len(time), 7, 6))).astype(np.float32)
x[:, 2, 3] = np.nan
kw = dict(window_year_baseline=3, smooth_days_baseline=5, window_days_hobday=5)
out = sharding.preprocess_sharded(lambda lo, hi: x[:, lo:hi], time, 7, rank, world, mo.preprocess,
                                  gather=sharding.dist_gather, **kw)
cnt = torch.tensor([int(out["extreme_events"].sum())]); dist.all_reduce(cnt)
eq = sharding.dist_gather(np.arange(6, dtype=np.float32).reshape(3, 2) + 10 * rank, 0)  # equal shards: one collective
assert eq.shape == (3 * world, 2) and all(eq[3 * r, 0] == 10 * r for r in range(world))
eq1 = sharding.dist_gather(np.arange(6, dtype=np.float32).reshape(2, 3) + 10 * rank, 1)
assert eq1.shape == (2, 3 * world) and all(eq1[0, 3 * r] == 10 * r for r in range(world))
if rank == 0:
    full = mo.preprocess(x, time, **kw)
    assert np.array_equal(out["thresholds_global"].view(np.uint32), full["thresholds"].view(np.uint32))
    assert np.array_equal(out["mask_global"], full["mask"])
    assert int(cnt) == int(full["extreme_events"].sum())
    print("GLOO_OK", int(cnt))
dist.destroy_process_group()
"""


def test_sharded_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MAREX_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GLOO_OK" in res.stdout


def test_exception_rendering_matches_reference_layout():
    """Message layout and default codes of marEx/exceptions.py (the reference's tests regex-match the text)."""
    from marex_b200 import ConfigurationError, DataValidationError, ProcessingError, create_data_validation_error

    e = ConfigurationError("Unknown x", details="d", suggestions=["a", "b"], context={"k": 1})
    assert str(e) == "Unknown x\nDetails: d\nContext: k=1\nSuggestions:\n  - a\n  - b\nError Code: CONFIGURATION_ERROR"
    assert str(DataValidationError("m")) == "m\nError Code: DATA_VALIDATION"
    assert ProcessingError("m").error_code == "PROCESSING_ERROR"
    v = create_data_validation_error("bad", data_info={"n": 3}, details="dd")
    assert isinstance(v, DataValidationError) and v.context == {"n": 3} and "Context: n=3" in str(v)
    v.add_suggestion("s")
    v.add_context("z", 1)
    assert v.suggestions == ["s"] and v.context["z"] == 1
