"""
xarray-facing drop-in for the reference's public functions on this path
(marEx/__init__.py:36-42): same names, same keyword arguments, same errors, DataArray in and
Dataset / DataArray out, so ``marEx.tracker`` consumes the result unchanged.

xarray (and dask, for the chunked outputs ``marEx.tracker`` insists on, track.py:411-418) are
imported lazily: the array-level API in ``detect.py`` works without them.  The dimension /
coordinate inference and its error messages follow detect.py:53-202 and only rely on the
``.dims`` / ``.coords`` duck type, which keeps them testable without xarray.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import detect as _d
from .exceptions import create_data_validation_error

logger = logging.getLogger("marex_b200")


def _require_names(kind: str, wanted: Dict[str, str], present) -> None:
    """detect.py:53-128: every name the caller mapped must exist on the DataArray (``kind`` = dimensions / coordinates)."""
    present = list(present)
    missing = [f"'{actual}' (for {concept})" for concept, actual in wanted.items() if actual not in present]
    if missing:
        raise create_data_validation_error(
            f"Missing required {kind}: {', '.join(missing)}",
            details=f"the DataArray has {kind} {present}",
            suggestions=[f"pass {kind}={{...}} with the names your data uses"],
            data_info={f"missing_{kind}": missing, f"available_{kind}": present, f"provided_{kind}": wanted},
        )


def _infer_dims_coords(da, dimensions, coordinates) -> Tuple[Dict[str, str], Dict[str, str]]:
    """detect.py:131-202: defaults (time / lat / lon), "time" added when omitted, no "y" means an unstructured mesh,
    which must name its coordinates explicitly."""
    if dimensions is None:
        dimensions = {"time": "time", "x": "lon", "y": "lat"}
    if "time" not in dimensions:
        dimensions = {"time": "time", **dimensions}
    if coordinates is None:
        if "y" not in dimensions:
            raise create_data_validation_error(
                "Coordinates parameter must be explicitly specified for unstructured data",
                details="a mesh has one spatial dimension, so the x / y coordinate variables cannot be guessed from the dimension names",
                suggestions=["e.g. coordinates={'time': 'time', 'x': 'lon', 'y': 'lat'}"],
                data_info={"data_structure": "unstructured (2D)", "dimensions": dimensions},
            )
        coordinates = dimensions.copy()
    elif "time" not in coordinates:
        coordinates = {"time": dimensions.get("time", "time"), **coordinates}
    _require_names("dimensions", dimensions, da.dims)
    _require_names("coordinates", coordinates, da.coords.keys())
    return dimensions, coordinates


def _is_dask(da) -> bool:
    return hasattr(da.data, "dask") or type(da.data).__module__.startswith("dask")


def _space_dims(dimensions: Dict[str, str]) -> List[str]:
    """[y,] x.  A mapping without "x" raises the bare ``KeyError: 'x'`` upstream's tests pin
    (tests/test_error_handling.py:171-181)."""
    return ([dimensions["y"]] if "y" in dimensions else []) + [dimensions["x"]]


def _host_field(da, dimensions) -> np.ndarray:
    """(time, [y,] x)-ordered numpy view of the DataArray (computes a dask-backed array in one piece: used by the
    small sibling functions; ``preprocess_data`` streams, see ``_pinned_field``)."""
    order = [dimensions["time"]] + _space_dims(dimensions)
    return np.asarray(da.transpose(*order).values)


def _pinned_field(da, dimensions):
    """The dask-backed DataArray as a page-locked float32 host tensor (time, [y,] x), filled one TIME CHUNK at a time:
    every dask block is computed, cast (``da.astype(np.float32)``, detect.py:600) and copied straight into the buffer
    the host->device stream of ``preprocess_arrays`` reads from, so a 60 GB field is never materialised as a second,
    pageable numpy array."""
    import torch

    tdim = dimensions["time"]
    dat = da.transpose(tdim, *_space_dims(dimensions))
    shape = tuple(int(n) for n in dat.shape)
    buf = torch.empty(shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
    view = buf.numpy()
    t_chunks = dat.chunks[0] if dat.chunks is not None else (shape[0],)
    t0 = 0
    for n in t_chunks:
        view[t0 : t0 + n] = np.asarray(dat.isel({tdim: slice(t0, t0 + n)}).values, dtype=np.float32)
        t0 += n
    return buf


def _chunked(obj, chunks):
    """``ds.chunk(...)`` of detect.py:786-792: ``marEx.tracker`` rejects numpy-backed input (track.py:411-418), so a
    failure to produce dask-backed variables is an error, not something to paper over."""
    out = obj.chunk(chunks)
    if getattr(out, "chunks", None) is None:
        raise RuntimeError("DataArray.chunk() returned a numpy-backed variable: dask is required for the Dataset marEx.tracker consumes")
    return out


def preprocess_data(
    da,
    method_anomaly="shifting_baseline",
    method_extreme="hobday_extreme",
    threshold_percentile=95,
    window_year_baseline=15,
    smooth_days_baseline=21,
    window_days_hobday=11,
    window_spatial_hobday=None,
    std_normalise=False,
    detrend_orders=None,
    force_zero_mean=True,
    reference_period=None,
    method_percentile="approximate",
    precision=0.01,
    max_anomaly=5.0,
    dask_chunks=None,
    dimensions=None,
    coordinates=None,
    neighbours=None,
    cell_areas=None,
    use_temp_checkpoints=False,
    verbose=None,
    quiet=None,
    device=None,
):
    """Drop-in for ``marEx.preprocess_data`` (detect.py:287-841) on the hot-path methods.

    Same signature (plus ``device``).  ``use_temp_checkpoints`` is accepted and ignored: it only
    exists upstream to cut dask graphs (helper.py:642-777) and there is no graph here.
    Returns an ``xr.Dataset`` with ``dat_anomaly`` (float32), ``mask`` (bool), ``extreme_events``
    (bool) and ``thresholds`` laid out as upstream, plus the same attrs.
    """
    import xarray as xr

    if detrend_orders is None:
        detrend_orders = [1]
    if dask_chunks is None:
        dask_chunks = {"time": 25}
    if verbose:
        logging.getLogger("marex_b200").setLevel(logging.DEBUG)
    elif quiet:
        logging.getLogger("marex_b200").setLevel(logging.WARNING)
    dimensions, coordinates = _infer_dims_coords(da, dimensions, coordinates)
    if not _is_dask(da):  # detect.py:557-568
        raise create_data_validation_error(
            "Input DataArray must be Dask-backed",
            details="the reference only accepts chunked input; this drop-in keeps the contract and reads it block by block",
            suggestions=["da.chunk({'time': 30}), or open the file with chunks={'time': 30}"],
            data_info={"data_type": type(da.data).__name__, "shape": da.shape},
        )
    _d.validate_reference_period_method(reference_period, method_anomaly)
    tdim = dimensions["time"]
    sdims = _space_dims(dimensions)
    gridded = "y" in dimensions
    res = _d.preprocess_arrays(
        _pinned_field(da, dimensions), da[coordinates["time"]].values, method_anomaly, method_extreme,
        threshold_percentile, window_year_baseline, smooth_days_baseline, window_days_hobday, window_spatial_hobday,
        std_normalise, detrend_orders, force_zero_mean, reference_period, method_percentile, precision, max_anomaly,
        device=device, output="numpy", gridded=gridded,
    )  # fmt: skip
    keep_time = res["time"]
    tsel = np.isin(da[coordinates["time"]].values.astype("datetime64[D]"), keep_time)
    base = da.isel({tdim: tsel}).transpose(tdim, *sdims)
    coords = {k: v for k, v in base.coords.items()}
    ds = xr.Dataset(coords=coords)
    ds["dat_anomaly"] = ((tdim, *sdims), res["dat_anomaly"])
    ds["mask"] = (tuple(sdims), res["mask"])
    ds["extreme_events"] = ((tdim, *sdims), res["extreme_events"])
    if res["thresholds_layout"] == "doy_last":
        ds["thresholds"] = ((*sdims, "dayofyear"), res["thresholds"])
        ds = ds.assign_coords(dayofyear=np.arange(1, 367, dtype=np.int32))
    elif res["thresholds_layout"] == "doy_first":
        ds["thresholds"] = (("dayofyear", *sdims), res["thresholds"])
        ds = ds.assign_coords(dayofyear=np.arange(1, 367))
    else:
        ds["thresholds"] = (tuple(sdims), res["thresholds"])
    if "dat_stn" in res:  # std_normalise (detect.py:686-715, 2290-2293)
        ds["dat_stn"] = ((tdim, *sdims), res["dat_stn"])
        ds["STD"] = ((*sdims, "dayofyear"), res["STD"])
        ds["extreme_events_stn"] = ((tdim, *sdims), res["extreme_events_stn"])
        lay = res["thresholds_layout"]
        tdims_ = (*sdims, "dayofyear") if lay == "doy_last" else (("dayofyear", *sdims) if lay == "doy_first" else tuple(sdims))
        ds["thresholds_stn"] = (tdims_, res["thresholds_stn"])
        if "dayofyear" not in ds.coords:
            ds = ds.assign_coords(dayofyear=np.arange(1, 367))
    if neighbours is not None:  # detect.py:718-723
        ds["neighbours"] = neighbours.astype(np.int32)
        if "nv" in neighbours.dims:
            ds = ds.assign_coords(nv=neighbours.nv)
    if cell_areas is not None:  # detect.py:725-728
        ds["cell_areas"] = cell_areas.astype(np.float32)
    ds.attrs.update(res["attrs"])
    tcoord = coordinates["time"]
    for key in ("calendar", "units"):  # detect.py:803-808
        if key in ds[tcoord].attrs:
            del ds[tcoord].attrs[key]
    # detect.py:786-792, 817-828: time-chunked dask variables, computed thresholds / mask / coords
    time_chunks = dask_chunks.get(tdim, dask_chunks.get("time", 10))
    chunk = {d: -1 for d in sdims}
    chunk[tdim] = time_chunks
    for var in ("dat_anomaly", "extreme_events"):
        ds[var] = _chunked(ds[var], chunk)
    return ds


def compute_normalised_anomaly(
    da,
    method_anomaly="shifting_baseline",
    dimensions=None,
    coordinates=None,
    window_year_baseline=15,
    smooth_days_baseline=21,
    std_normalise=False,
    detrend_orders=None,
    force_zero_mean=True,
    reference_period=None,
    use_temp_checkpoints=False,
    verbose=None,
    quiet=None,
    device=None,
):
    """Drop-in for ``marEx.compute_normalised_anomaly`` (detect.py:891-1116): returns a Dataset
    with ``dat_anomaly`` (NOT trimmed, as upstream: NaN in the first years for shifting_baseline)
    and ``mask``."""
    import torch
    import xarray as xr

    dimensions, coordinates = _infer_dims_coords(da, dimensions, coordinates)
    _d.validate_reference_period_method(reference_period, method_anomaly)
    if da.chunks is None:  # upstream: TypeError from da.chunks[0] (detect.py:2180), pinned by its tests
        raise TypeError("'NoneType' object is not subscriptable")
    tdim, sdims = dimensions["time"], _space_dims(dimensions)
    x = _host_field(da, dimensions).astype(np.float32)
    time = da[coordinates["time"]].values
    dev = _d._device(device)
    cal = _d.build_calendar(time)
    x_dev, space = _d._to_device_field(x, dev)
    res = _d.compute_normalised_anomaly_arrays(
        x_dev, cal, method_anomaly, window_year_baseline, smooth_days_baseline, detrend_orders, force_zero_mean,
        reference_period, validate=False, in_place=(method_anomaly != "shifting_baseline"),
        std_normalise=bool(std_normalise) and method_anomaly == "detrend_harmonic",
    )  # fmt: skip
    anom = res["dat_anomaly"]
    if method_anomaly == "shifting_baseline":  # upstream returns the untrimmed series
        full = torch.full((cal.T, anom.shape[1]), float("nan"), dtype=torch.float32, device=dev)
        full[torch.from_numpy(np.nonzero(res["keep"])[0]).to(dev)] = anom
        anom = full
    base = da.transpose(tdim, *sdims)
    ds = xr.Dataset(coords={k: v for k, v in base.coords.items()})
    ds["dat_anomaly"] = ((tdim, *sdims), anom.reshape((cal.T,) + space).cpu().numpy())
    ds["mask"] = (tuple(sdims), res["mask"].reshape(space).cpu().numpy())
    if "dat_stn" in res:  # detect.py:2290-2293
        ds["dat_stn"] = ((tdim, *sdims), res["dat_stn"].reshape((cal.T,) + space).cpu().numpy())
        ds["STD"] = ((*sdims, "dayofyear"), res["STD"].reshape(space + (366,)).cpu().numpy())
        ds = ds.assign_coords(dayofyear=np.arange(1, 367))
    return ds


def identify_extremes(
    da,
    method_extreme="global_extreme",
    threshold_percentile=95,
    dimensions=None,
    coordinates=None,
    window_days_hobday=11,
    window_spatial_hobday=None,
    method_percentile="approximate",
    precision=0.01,
    max_anomaly=5.0,
    use_temp_checkpoints=False,
    verbose=None,
    quiet=None,
    device=None,
):
    """Drop-in for ``marEx.identify_extremes`` (detect.py:1119-1503): ``(extremes, thresholds)``."""
    import xarray as xr

    dimensions, coordinates = _infer_dims_coords(da, dimensions, coordinates)
    tdim, sdims = dimensions["time"], _space_dims(dimensions)
    gridded = "y" in dimensions and dimensions["y"] in da.dims
    # configuration errors first, before any data is touched (same order as upstream)
    _d.resolve_extreme_config(
        method_extreme, threshold_percentile, window_days_hobday, window_spatial_hobday, method_percentile, precision,
        max_anomaly, gridded, dimensions, list(da.dims),
    )  # fmt: skip
    x = _host_field(da, dimensions).astype(np.float32)
    time = da[coordinates["time"]].values
    dev = _d._device(device)
    a_dev, space = _d._to_device_field(x, dev)
    _, doy = _d.build_calendar(time).year, _d.build_calendar(time).doy
    res = _d.identify_extremes_arrays(
        a_dev, doy, tuple(space) if gridded else None, method_extreme, threshold_percentile, window_days_hobday,
        window_spatial_hobday, method_percentile, precision, max_anomaly,
        n_years=int(np.unique(_d.build_calendar(time).year).size),
    )  # fmt: skip
    base = da.transpose(tdim, *sdims)
    ev = xr.DataArray(res["extreme_events"].reshape(x.shape).cpu().numpy(), dims=(tdim, *sdims), coords=base.coords)
    thr = res["thresholds"].cpu().numpy()
    scoords = {d: base.coords[d] for d in sdims if d in base.coords}
    if res["thresholds_layout"] == "doy_last":
        thr_da = xr.DataArray(thr, dims=(*sdims, "dayofyear"), coords={**scoords, "dayofyear": np.arange(1, 367)})
    elif res["thresholds_layout"] == "doy_first":
        thr_da = xr.DataArray(thr, dims=("dayofyear", *sdims), coords={**scoords, "dayofyear": np.arange(1, 367)})
    else:
        thr_da = xr.DataArray(thr, dims=tuple(sdims), coords=scoords)
    return ev, thr_da


def rolling_climatology(da, window_year_baseline=15, dimensions=None, coordinates=None, use_temp_checkpoints=False, device=None):
    """Drop-in for ``marEx.rolling_climatology`` (detect.py:1511-1688)."""
    return _rolling(da, window_year_baseline, 1, dimensions, coordinates, device)


def smoothed_rolling_climatology(
    da, window_year_baseline=15, smooth_days_baseline=21, dimensions=None, coordinates=None, use_temp_checkpoints=False, device=None
):
    """Drop-in for ``marEx.smoothed_rolling_climatology`` (detect.py:1691-1816)."""
    return _rolling(da, window_year_baseline, smooth_days_baseline, dimensions, coordinates, device)


def _rolling(da, W, S, dimensions, coordinates, device):
    import xarray as xr

    dimensions, coordinates = _infer_dims_coords(da, dimensions, coordinates)
    tdim, sdims = dimensions["time"], _space_dims(dimensions)
    x = _host_field(da, dimensions)
    clim = _d.rolling_climatology_arrays(x, da[coordinates["time"]].values, W, S, device=device)
    base = da.transpose(tdim, *sdims)
    return xr.DataArray(clim.cpu().numpy(), dims=(tdim, *sdims), coords=base.coords)
