"""
Generate the committed golden fixtures in tests/golden/.

Run in the BUILD container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

What is pinned here
-------------------
The reference (`marEx/detect.py`) cannot be imported in this image (xarray, dask, flox
and xhistogram are missing), but four pure-numpy/pandas leaves can be AST-extracted
from the read-only source and executed unchanged.  Nothing is copied into the repo:
the function bodies are compiled from `/root/reference/marEx/detect.py` at run time
and only their INPUTS and OUTPUTS are stored.

  ref_rolling_hist_quantile.npz   _rolling_histogram_quantile   detect.py:2465-2559
  ref_doy_percentiles.npz         nested _doy_percentiles        detect.py:1936-1942
                                  (+ the doy_masks loop           detect.py:1925-1934)
  ref_decimal_year.npz            add_decimal_year               detect.py:2031-2058
  ref_preprocessing_steps.json    _get_preprocessing_steps       detect.py:844-888

  sst_gridded_subset.npz / sst_unstructured_subset.npz
      small real-data INPUT slices decoded from the reference's own test fixtures
      (tests/data/sst_gridded.zarr, sst_unstructured.zarr) with the NaN column the
      reference tests inject (tests/test_gridded_preprocessing.py:22-25).
"""
import ast
import json
import os
import sys
import textwrap

import numpy as np
import pandas as pd
from numpy.lib.stride_tricks import sliding_window_view
from numpy.typing import NDArray

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference/marEx/detect.py"


def _extract(name: str, parent: str = None):
    """Compile function `name` (optionally nested inside `parent`) from the reference source."""
    src = open(REF).read()
    tree = ast.parse(src)
    scope = tree.body
    if parent is not None:
        (pnode,) = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == parent]
        scope = list(ast.walk(pnode))
    (node,) = [n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name]
    code = textwrap.dedent(ast.get_source_segment(src, node))
    ns = {
        "np": np,
        "pd": pd,
        "sliding_window_view": sliding_window_view,
        "NDArray": NDArray,
        "List": list,
        "Optional": object,
        "Tuple": tuple,
        "xr": type("xr", (), {"DataArray": object}),
    }
    from typing import List, Optional, Tuple

    ns.update(List=List, Optional=Optional, Tuple=Tuple)
    exec(compile(code, f"<reference:{name}>", "exec"), ns)
    return ns[name]


def golden_rolling_hist_quantile():
    f = _extract("_rolling_histogram_quantile")
    rng = np.random.default_rng(20260101)
    edges = np.concatenate([[-np.inf], np.arange(-0.01, 5.01, 0.01, dtype=np.float32)], dtype=np.float32)
    centers = (edges[1:] + edges[:-1]) / 2
    centers[0] = 0.0
    centers = centers.astype(np.float32)
    nb = len(centers)
    cases = {}

    def synth(n_years, scale, sparse=False, leap_every=4):
        h = np.zeros((366, nb), dtype=np.uint16)
        for d in range(366):
            ny = n_years if d < 365 else n_years // leap_every
            a = rng.normal(0.0, scale, size=ny).astype(np.float32)
            if sparse:
                a = a[: max(1, ny // 6)]
            b = np.digitize(a, edges) - 1
            for v in b[b < nb]:
                h[d, v] += 1
        return h

    hists = {
        "n25_s1": synth(25, 1.0),
        "n40_s03": synth(40, 0.3),
        "n8_sparse": synth(8, 2.0, sparse=True),
        "n25_s3_overflow": synth(25, 3.0),
        "all_negative": np.zeros((366, nb), dtype=np.uint16),
        "constant_zero": np.zeros((366, nb), dtype=np.uint16),
        "empty": np.zeros((366, nb), dtype=np.uint16),
    }
    hists["all_negative"][:, 0] = 25
    hists["constant_zero"][:, 2] = 25
    hists["constant_zero"][365, 2] = 6
    pooled = sum(synth(25, 1.0).astype(np.float64) for _ in range(25))  # float counts, like after 5x5 pooling
    hists["pooled25_float"] = pooled
    for name, h in hists.items():
        for w in (3, 11, 31):
            for q in (0.6, 0.9, 0.95, 0.99):
                cases[f"{name}|{w}|{q}"] = f(h, w, q, centers)
    np.savez_compressed(
        os.path.join(HERE, "ref_rolling_hist_quantile.npz"),
        centers=centers,
        hist_names=np.array(list(hists.keys())),
        **{f"hist_{k}": v for k, v in hists.items()},
        case_keys=np.array(list(cases.keys())),
        **{f"out_{i}": v for i, v in enumerate(cases.values())},
    )
    print("rolling_hist_quantile:", len(cases), "cases")


def golden_doy_percentiles():
    f = _extract("_doy_percentiles", parent="_identify_extremes_hobday")
    rng = np.random.default_rng(7)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64("2002-03-15"))
    y = time.astype("datetime64[Y]")
    doy = ((time - y.astype("datetime64[D]")).astype(int) + 1).astype(np.int64)
    T = len(time)
    data = (rng.standard_normal((7, T)) * np.linspace(0.2, 3, 7)[:, None]).astype(np.float32)
    data[2] = np.nan  # land cell
    data[3, ::17] = np.nan  # gappy cell
    data[4] = np.float32(0.25)  # constant cell
    data[5] = np.round(data[5], 1)  # many duplicates
    out = {}
    for w in (1, 5, 11):
        half = w // 2
        masks = []
        for d in range(1, 367):  # detect.py:1929-1934, restated (loop only)
            m = np.zeros(T, dtype=bool)
            for off in range(-half, half + 1):
                m |= doy == ((d - 1 + off) % 366) + 1
            masks.append(m)
        for p in (50, 90, 95, 99.5):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                out[f"{w}|{p}"] = f(data, masks, p)
    np.savez_compressed(
        os.path.join(HERE, "ref_doy_percentiles.npz"),
        data=data,
        doy=doy.astype(np.int16),
        keys=np.array(list(out.keys())),
        **{f"out_{i}": v for i, v in enumerate(out.values())},
    )
    print("doy_percentiles:", len(out), "cases")


class _FakeDA:
    """Just enough of a DataArray for add_decimal_year: ``da[coord]`` and ``assign_coords``."""

    def __init__(self, time):
        self._t = time

    def __getitem__(self, k):
        return self._t

    def assign_coords(self, **kw):
        return kw


def golden_decimal_year():
    f = _extract("add_decimal_year")
    time = np.concatenate(
        [
            np.arange(np.datetime64("1982-01-01"), np.datetime64("1985-01-03")),
            np.arange(np.datetime64("1999-12-25"), np.datetime64("2001-01-05")),
            np.array(["1900-03-01", "2100-12-31", "2024-02-29"], dtype="datetime64[D]"),
        ]
    )
    res = f(_FakeDA(time.astype("datetime64[ns]")), dim="time")
    dim, dy = res["decimal_year"]
    np.savez_compressed(os.path.join(HERE, "ref_decimal_year.npz"), time=time, decimal_year=np.asarray(dy, dtype=np.float64))
    print("decimal_year:", len(time), "dates")


def golden_preprocessing_steps():
    f = _extract("_get_preprocessing_steps")
    cases = []
    for ma in ("shifting_baseline", "fixed_baseline", "detrend_fixed_baseline", "detrend_harmonic"):
        for me in ("hobday_extreme", "global_extreme"):
            for ws in (None, 5):
                for rp in (None, (1990, 2010)):
                    for stdn in (False, True):
                        args = dict(
                            method_anomaly=ma,
                            method_extreme=me,
                            std_normalise=stdn,
                            detrend_orders=[1, 2],
                            window_year_baseline=15,
                            smooth_days_baseline=21,
                            window_days_hobday=11,
                            window_spatial_hobday=ws,
                            reference_period=rp,
                        )
                        cases.append({"args": args, "out": f(**args)})
    json.dump(cases, open(os.path.join(HERE, "ref_preprocessing_steps.json"), "w"), indent=0)
    print("preprocessing_steps:", len(cases), "cases")


def real_data_subsets():
    from _zarr_blosc import read_zarr_array

    base = "/root/reference/tests/data"
    x = read_zarr_array(f"{base}/sst_gridded.zarr/to")  # (14611, 20, 40) f32, Kelvin
    lat = read_zarr_array(f"{base}/sst_gridded.zarr/lat")
    lon = read_zarr_array(f"{base}/sst_gridded.zarr/lon")
    t = read_zarr_array(f"{base}/sst_gridded.zarr/time")
    time = (np.datetime64("1981-01-01T00:00:00") + t.astype("timedelta64[s]")).astype("datetime64[D]")
    sub = x[:, 4:10, 10:18].copy()  # 6 x 8 cells
    sub[:, 1, 1] = np.nan  # the NaN column the reference tests inject
    np.savez_compressed(os.path.join(HERE, "sst_gridded_subset.npz"), sst=sub, time=time, lat=lat[4:10], lon=lon[10:18])
    u = read_zarr_array(f"{base}/sst_unstructured.zarr/to")  # (14611, 405)
    usub = u[:, :40].copy()
    usub[:, 2] = np.nan
    np.savez_compressed(os.path.join(HERE, "sst_unstructured_subset.npz"), sst=usub, time=time)
    print("real data:", sub.shape, usub.shape)


if __name__ == "__main__":
    golden_rolling_hist_quantile()
    golden_doy_percentiles()
    golden_decimal_year()
    golden_preprocessing_steps()
    real_data_subsets()
