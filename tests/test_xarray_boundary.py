"""The xarray boundary of the drop-in (``marex_b200.xr_api``): the reference's public signatures, DataArray in and
Dataset out (detect.py:287-313, 679-683, 718-828), executed through ``tests/_fake_xarray.py`` because neither xarray nor
dask is installed in this image.  Mirrors what the reference's own tests check at this boundary:

* output variables, dimension order, dtypes, coords, attrs -- tests/test_gridded_preprocessing.py:49-140,
  tests/test_unstructured_preprocessing.py:57-140,
* custom dimension / coordinate names -- tests/test_gridded_preprocessing.py:389-524,
* the input contract of ``marEx.tracker`` (bool, dask-backed, time-chunked; track.py:411-418, 579-591, 640-668) --
  tests/test_integration.py:37-172,
* the error contract -- tests/test_error_handling.py:50-57 (non-dask), 171-260, 331-376 (dims / coords).
"""
import sys
import warnings

import numpy as np
import pytest

from oracle import marex_oracle as mo

import _fake_xarray as fx  # tests/ is on sys.path (conftest.py)


@pytest.fixture
def xr(monkeypatch):
    """``import xarray`` inside the adapter resolves to the stand-in (unless the real package exists)."""
    try:
        import xarray  # noqa: F401

        pytest.skip("real xarray present: this file exercises the stand-in")
    except ImportError:
        pass
    monkeypatch.setitem(sys.modules, "xarray", fx)
    return fx


def _sst(T0="1990-01-01", T1="2001-01-01", ny=6, nx=40, seed=0):
    rng = np.random.default_rng(seed)
    time = np.arange(np.datetime64(T0), np.datetime64(T1))
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    x = (15 + 4 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((len(time), ny, nx))).astype(np.float32)
    if ny > 1:
        x[:, 1, 1] = np.nan  # the NaN column the reference's tests inject (tests/test_gridded_preprocessing.py:22-25)
    return x, time


def _gridded_da(x, time, names=("time", "lat", "lon"), chunk=30):
    t, y, xn = names
    da = fx.DataArray(
        x, dims=names,
        coords={t: fx.DataArray(time, dims=(t,), attrs={"calendar": "proleptic_gregorian", "units": "days since 1970-01-01"}),
                y: np.linspace(35, 40, x.shape[1]), xn: np.linspace(-40, -30, x.shape[2])},
    )  # fmt: skip
    return da.chunk({t: chunk})


# ---------------------------------------------------------------- errors that need no device
def test_non_dask_input_is_rejected(xr):
    import marex_b200

    x, time = _sst(T1="1992-01-01")
    da = fx.DataArray(x, dims=("time", "lat", "lon"), coords={"time": time, "lat": np.arange(6.0), "lon": np.arange(40.0)})
    with pytest.raises(marex_b200.DataValidationError, match="Input DataArray must be Dask-backed"):
        marex_b200.preprocess_data(da)


def test_missing_dimensions_and_coordinates(xr):
    import marex_b200

    x, time = _sst(T1="1992-01-01")
    da = _gridded_da(x, time, names=("time", "latitude", "longitude"))
    with pytest.raises(marex_b200.DataValidationError, match="Missing required dimensions"):
        marex_b200.preprocess_data(da)
    with pytest.raises(marex_b200.DataValidationError, match="Missing required coordinates"):
        marex_b200.preprocess_data(da, dimensions={"time": "time", "x": "longitude", "y": "latitude"},
                                   coordinates={"time": "time", "x": "lon", "y": "lat"})  # fmt: skip
    flat = fx.DataArray(x.reshape(len(time), -1), dims=("time", "ncells"), coords={"time": time}).chunk({"time": 30})
    with pytest.raises(marex_b200.DataValidationError, match="Coordinates parameter must be explicitly specified"):
        marex_b200.preprocess_data(flat, dimensions={"time": "time", "x": "ncells"})
    with pytest.raises(KeyError):  # the accidental behaviour upstream pins (tests/test_error_handling.py:171-181)
        marex_b200.preprocess_data(da, dimensions={"y": "latitude"})


def test_reference_period_with_the_wrong_method_is_a_configuration_error(xr):
    import marex_b200

    x, time = _sst(T1="1992-01-01")
    with pytest.raises(marex_b200.ConfigurationError, match="reference_period"):
        marex_b200.preprocess_data(_gridded_da(x, time), method_anomaly="shifting_baseline", reference_period=(1990, 1991))


def test_chunking_must_produce_dask_backed_variables(xr, monkeypatch):
    from marex_b200 import xr_api

    x, time = _sst(T1="1990-03-01")
    da = fx.DataArray(x, dims=("time", "lat", "lon"))
    monkeypatch.setattr(fx.DataArray, "chunk", lambda self, chunks: self)  # a chunk() that silently does nothing
    with pytest.raises(RuntimeError, match="dask is required"):
        xr_api._chunked(da, {"time": 10})


def test_dask_blocks_are_streamed_into_the_pinned_buffer(xr):
    """``preprocess_data`` never materialises the whole field as one pageable array: every time chunk is computed and
    copied on its own."""
    from marex_b200 import xr_api

    x, time = _sst(T1="1990-07-01")
    da = _gridded_da(x, time, chunk=25).transpose("lat", "time", "lon")  # time need not be the first axis
    buf = xr_api._pinned_field(da, {"time": "time", "x": "lon", "y": "lat"})
    np.testing.assert_array_equal(buf.numpy(), x)
    log = da.data.log
    assert len(log) == -(-len(time) // 25) and max(s[0] for s in log) == 25, log  # one materialisation per time chunk


# ---------------------------------------------------------------- the Dataset, on the GPU
def _cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import marex_b200

    return marex_b200


GRIDDED_CASES = [
    (dict(), "doy_last", np.float32),
    (dict(method_percentile="exact"), "doy_first", np.float32),
    (dict(method_extreme="global_extreme"), "space", np.float64),
    (dict(method_anomaly="detrend_fixed_baseline", method_extreme="global_extreme", method_percentile="exact"), "space", np.float64),
    (dict(method_anomaly="fixed_baseline", reference_period=(1992, 1998)), "doy_last", np.float32),
    (dict(method_anomaly="detrend_harmonic", std_normalise=True), "doy_last", np.float32),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw,layout,thr_dtype", GRIDDED_CASES)
@pytest.mark.parametrize("names", [("time", "lat", "lon"), ("t", "y_dim", "x_dim")])
def test_preprocess_data_dataset_contract(xr, kw, layout, thr_dtype, names):
    mb = _cuda()
    x, time = _sst()
    t, y, xn = names
    da = _gridded_da(x, time, names=names)
    extra = {} if names[0] == "time" else dict(dimensions={"time": t, "x": xn, "y": y}, coordinates={"time": t, "x": xn, "y": y})
    small = dict(window_year_baseline=4, smooth_days_baseline=9, window_days_hobday=5, dask_chunks={"time": 40})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ds = mb.preprocess_data(da, **small, **kw, **extra)
        ref = mo.preprocess(x, time, window_year_baseline=4, smooth_days_baseline=9, window_days_hobday=5,
                            **{k: v for k, v in kw.items() if k != "std_normalise"})  # fmt: skip
    T_out = len(ref["time"])
    # variables, dimension order, dtypes (detect.py:682-683, 1956, 2682-2693, 2899)
    assert ds["dat_anomaly"].dims == (t, y, xn) and ds["dat_anomaly"].dtype == np.float32 and ds["dat_anomaly"].shape == (T_out, 6, 40)
    assert ds["extreme_events"].dims == (t, y, xn) and ds["extreme_events"].dtype == np.bool_
    assert ds["mask"].dims == (y, xn) and ds["mask"].dtype == np.bool_
    want_dims = {"doy_last": (y, xn, "dayofyear"), "doy_first": ("dayofyear", y, xn), "space": (y, xn)}[layout]
    assert ds["thresholds"].dims == want_dims and ds["thresholds"].dtype == thr_dtype
    if layout != "space":
        np.testing.assert_array_equal(ds["dayofyear"].values, np.arange(1, 367))
    # what marEx.tracker demands of its input (track.py:411-418, 579-591, 640-668)
    for v in ("dat_anomaly", "extreme_events"):
        assert ds[v].chunks is not None and ds[v].chunks[0][0] == 40 and len(ds[v].chunks[1]) == 1 and len(ds[v].chunks[2]) == 1
    assert ds["thresholds"].chunks is None and ds["mask"].chunks is None  # computed (numpy-backed), detect.py:817-828
    assert ds["mask"].values.any() and not ds["mask"].values[1, 1]
    # coordinates travel, trimmed like the data; CF attrs are stripped from time (detect.py:803-808)
    np.testing.assert_array_equal(ds[t].values, ref["time"])
    assert "calendar" not in ds[t].attrs and "units" not in ds[t].attrs
    np.testing.assert_array_equal(ds[y].values, da[y].values)
    # attrs (detect.py:731-783)
    a = ds.attrs
    assert a["method_anomaly"] == kw.get("method_anomaly", "shifting_baseline") and a["method_extreme"] == kw.get("method_extreme", "hobday_extreme")
    assert a["threshold_percentile"] == 95 and a["method_percentile"] == kw.get("method_percentile", "approximate")
    assert isinstance(a["preprocessing_steps"], list) and a["precision"] == 0.01 and a["max_anomaly"] == 5.0
    if a["method_anomaly"] == "shifting_baseline":
        assert a["window_year_baseline"] == 4 and a["smooth_days_baseline"] == 9
    if "reference_period" in kw:
        assert a["reference_period"] == [1992, 1998]
    if a["method_extreme"] == "hobday_extreme":
        assert a["window_days_hobday"] == 5
    # values: the array level is checked bit by bit elsewhere; here the assembled Dataset against the oracle pipeline
    np.testing.assert_allclose(ds["dat_anomaly"].values, ref["dat_anomaly"], rtol=0, atol=3e-4, equal_nan=True)
    np.testing.assert_array_equal(ds["mask"].values, ref["mask"])
    freq = ds["extreme_events"].values[:, ds["mask"].values].mean()
    assert 0.03 < freq < 0.08, freq  # assert_percentile_frequency of the reference's conftest (5 % +- 20 %, min 0.5 %)
    if kw.get("std_normalise"):
        assert ds["dat_stn"].dims == (t, y, xn) and ds["STD"].dims == (y, xn, "dayofyear")
        assert ds["extreme_events_stn"].dtype == np.bool_ and ds["thresholds_stn"].dims == want_dims


@pytest.mark.gpu
def test_preprocess_data_unstructured_with_neighbours_and_areas(xr):
    """Unstructured input (no ``y`` dimension): explicit coordinates, ``neighbours`` / ``cell_areas`` pass through as
    int32 / float32 (detect.py:718-728), no spatial pooling."""
    mb = _cuda()
    x, time = _sst(ny=1, nx=64)
    x = x[:, 0, :]
    x[:, 2] = np.nan
    nc = x.shape[1]
    da = fx.DataArray(
        x, dims=("time", "ncells"),
        coords={"time": time, "lon": fx.DataArray(np.linspace(0, 5, nc), dims=("ncells",)),
                "lat": fx.DataArray(np.linspace(39, 40, nc), dims=("ncells",))},
    ).chunk({"time": 50})  # fmt: skip
    nbr = fx.DataArray(np.tile(np.arange(nc), (3, 1)).astype(np.int64), dims=("nv", "ncells"), coords={"nv": np.arange(3)})
    areas = fx.DataArray(np.ones(nc, np.float64), dims=("ncells",))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ds = mb.preprocess_data(da, window_year_baseline=4, smooth_days_baseline=9, window_days_hobday=5,
                                dimensions={"time": "time", "x": "ncells"}, coordinates={"time": "time", "x": "lon", "y": "lat"},
                                neighbours=nbr, cell_areas=areas)  # fmt: skip
    assert ds["dat_anomaly"].dims == ("time", "ncells") and ds["thresholds"].dims == ("ncells", "dayofyear")
    assert ds["neighbours"].dtype == np.int32 and ds["neighbours"].dims == ("nv", "ncells") and "nv" in ds.coords
    assert ds["cell_areas"].dtype == np.float32
    assert ds["extreme_events"].chunks is not None and not ds["mask"].values[2]


@pytest.mark.gpu
def test_sibling_exports(xr):
    """compute_normalised_anomaly / identify_extremes / rolling_climatology / smoothed_rolling_climatology
    (marEx/__init__.py:36-42)."""
    mb = _cuda()
    x, time = _sst(T1="1999-01-01")
    da = _gridded_da(x, time)
    year, doy = mo.calendar_tables(time)
    # untrimmed anomalies, NaN in the first W years (detect.py:1086 returns the full series; preprocess_data trims)
    ds = mb.compute_normalised_anomaly(da, window_year_baseline=3, smooth_days_baseline=5)
    assert ds["dat_anomaly"].dims == ("time", "lat", "lon") and ds["dat_anomaly"].shape == x.shape
    ref, mask, keep = mo.anomaly_shifting_baseline(x, year, doy, 3, 5)
    got = ds["dat_anomaly"].values
    assert np.isnan(got[~keep]).all()
    np.testing.assert_allclose(got[keep], ref, rtol=0, atol=3e-4, equal_nan=True)
    np.testing.assert_array_equal(ds["mask"].values, mask)
    # numpy-backed input: the TypeError upstream's tests pin (tests/test_error_handling.py:59-67)
    with pytest.raises(TypeError, match="NoneType"):
        mb.compute_normalised_anomaly(fx.DataArray(x, dims=("time", "lat", "lon"), coords=dict(da.coords)))
    # identify_extremes -> (extremes, thresholds)
    anom = fx.DataArray(np.where(keep[:, None, None], got, 0).astype(np.float32)[keep], dims=("time", "lat", "lon"),
                        coords={"time": time[keep], "lat": da["lat"].values, "lon": da["lon"].values}).chunk({"time": 30})  # fmt: skip
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, thr = mb.identify_extremes(anom, method_extreme="hobday_extreme", window_days_hobday=5)
    assert ev.dims == ("time", "lat", "lon") and ev.dtype == np.bool_ and thr.dims == ("lat", "lon", "dayofyear")
    ev_ref, thr_ref = mo.preprocess_from_anomaly(anom.values, doy[keep], "hobday_extreme", 95, 5, None, "approximate", 0.01, 5.0)
    np.testing.assert_array_equal(ev.values, ev_ref)
    ok = ~np.isnan(thr_ref)
    np.testing.assert_array_equal(thr.values[ok].view(np.uint32), thr_ref[ok].view(np.uint32))
    with pytest.raises(mb.ConfigurationError):
        mb.identify_extremes(anom, method_extreme="hobday_extreme", window_days_hobday=4)
    # climatologies
    for fn, S in ((mb.rolling_climatology, 1), (mb.smoothed_rolling_climatology, 9)):
        clim = fn(da, 3) if S == 1 else fn(da, 3, S)
        want = mo.rolling_climatology(mo.smooth_centered(x, S), year, doy, 3)
        assert clim.dims == ("time", "lat", "lon")
        np.testing.assert_allclose(clim.values, want, rtol=0, atol=3e-4, equal_nan=True)
