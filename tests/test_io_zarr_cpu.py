"""CPU tests of the zarr v2 I/O edge (marex_b200/io_zarr.py)."""
import json
import os
import struct
import zlib

import numpy as np
import pytest

from marex_b200 import io_zarr as zio

REF_DATA = "/root/reference/tests/data"


def test_raw_roundtrip_and_partial_chunks(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((17, 5, 7)).astype(np.float32)
    a[3, 2, 1] = np.nan
    store = str(tmp_path / "g.zarr")
    zio.write_array(store, "to", a, (4, 5, 3), ["time", "lat", "lon"])
    meta = zio.array_meta(store, "to")
    assert meta["shape"] == [17, 5, 7] and meta["chunks"] == [4, 5, 3] and meta["compressor"] is None
    assert meta["attrs"]["_ARRAY_DIMENSIONS"] == ["time", "lat", "lon"]
    np.testing.assert_array_equal(zio.read_array(store, "to"), a)
    # conversion on read into a caller-provided buffer (float64 store -> float32 field)
    zio.write_array(store, "d", a.astype(np.float64), (17, 5, 7), ["time", "lat", "lon"])
    out = np.empty(a.shape, np.float32)
    zio.read_array(store, "d", out=out)
    np.testing.assert_array_equal(out, a)
    b = rng.integers(0, 2, (9, 4)).astype(bool)
    zio.write_array(store, "ev", b, (2, 4), ["time", "ncells"])
    got = zio.read_array(store, "ev")
    assert got.dtype == np.bool_
    np.testing.assert_array_equal(got, b)
    # a missing chunk reads as the fill value
    os.remove(os.path.join(store, "to", "1.0.2"))
    r = zio.read_array(store, "to")
    assert np.isnan(r[4:8, :, 6]).all() and np.array_equal(r[:4], a[:4], equal_nan=True)


def _blosc_frame(payload: bytes, typesize: int, shuffle: bool, codec: str) -> bytes:
    """A one-block, unsplit blosc v1 frame (test helper: the product only decodes)."""
    import pyarrow as pa

    body = payload
    if shuffle:
        arr = np.frombuffer(payload, np.uint8)
        ne = arr.size // typesize
        body = arr[: ne * typesize].reshape(ne, typesize).T.reshape(-1).tobytes() + arr[ne * typesize :].tobytes()
    if codec == "lz4":
        comp, fmt = pa.Codec("lz4_raw").compress(body, asbytes=True), 1
    elif codec == "zstd":
        comp, fmt = pa.Codec("zstd").compress(body, asbytes=True), 4
    else:
        comp, fmt = zlib.compress(body), 3
    flags = (fmt << 5) | 0x10 | (0x1 if shuffle else 0)  # dont-split
    stream = struct.pack("<i", len(comp)) + comp
    header = struct.pack("<BBBBIII", 2, 1, flags, typesize, len(payload), len(payload), 16 + 4 + len(stream))
    return header + struct.pack("<i", 20) + stream


@pytest.mark.parametrize("codec,shuffle", [("lz4", True), ("lz4", False), ("zstd", True), ("zlib", True)])
def test_blosc_frames(codec, shuffle):
    rng = np.random.default_rng(1)
    a = np.round(rng.standard_normal(3000), 1).astype(np.float32)
    frame = _blosc_frame(a.tobytes(), 4, shuffle, codec)
    np.testing.assert_array_equal(np.frombuffer(zio.blosc_decompress(frame), np.float32), a)
    memcpy = struct.pack("<BBBBIII", 2, 1, 0x2, 4, a.nbytes, a.nbytes, a.nbytes + 16) + a.tobytes()
    np.testing.assert_array_equal(np.frombuffer(zio.blosc_decompress(memcpy), np.float32), a)


def test_cf_time_decoding():
    t = zio.decode_cf_time(np.array([0, 86400, 86400 * 366], np.int32), "seconds since 1981-01-01", "proleptic_gregorian")
    assert list(t.astype("datetime64[D]").astype(str)) == ["1981-01-01", "1981-01-02", "1982-01-02"]
    t = zio.decode_cf_time(np.array([0.5, 1.0]), "days since 2000-01-01 00:00:00")
    assert str(t[0]) == "2000-01-01T12:00:00.000000000" and str(t[1].astype("datetime64[D]")) == "2000-01-02"
    with pytest.raises(NotImplementedError):
        zio.decode_cf_time(np.array([0]), "days since 2000-01-01", "noleap")


def test_result_group_layout(tmp_path):
    time = np.arange(np.datetime64("2000-01-01"), np.datetime64("2000-03-01"))
    res = {
        "dat_anomaly": np.zeros((len(time), 3, 4), np.float32), "extreme_events": np.zeros((len(time), 3, 4), bool),
        "mask": np.ones((3, 4), bool), "thresholds": np.ones((3, 4, 366), np.float32), "thresholds_layout": "doy_last",
        "time": time, "attrs": {"method_anomaly": "shifting_baseline", "preprocessing_steps": ["a", "b"], "reference_period": (1, 2)},
    }  # fmt: skip
    store = str(tmp_path / "out.zarr")
    zio.write_result(store, res, ["lat", "lon"])
    assert json.load(open(os.path.join(store, ".zattrs")))["reference_period"] == [1, 2]
    m = zio.array_meta(store, "dat_anomaly")
    assert m["chunks"] == [25, 3, 4] and m["attrs"]["_ARRAY_DIMENSIONS"] == ["time", "lat", "lon"]
    assert zio.array_meta(store, "thresholds")["attrs"]["_ARRAY_DIMENSIONS"] == ["lat", "lon", "dayofyear"]
    x, t, dims = zio.read_field(store, "dat_anomaly", pinned=False)
    assert dims == ["time", "lat", "lon"] and np.array_equal(t.astype("datetime64[D]"), time)
    np.testing.assert_array_equal(zio.read_array(store, "extreme_events"), res["extreme_events"])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DATA, "sst_gridded.zarr")), reason="reference fixtures not present")
def test_reads_the_reference_fixture_like_the_golden_subset(golden_dir):
    """The reference's own blosc(lz4, shuffle) fixture decodes to the values stored in tests/golden (which were
    produced by the independent decoder of tests/golden/_zarr_blosc.py)."""
    g = np.load(os.path.join(golden_dir, "sst_gridded_subset.npz"))
    x, time, dims = zio.read_field(os.path.join(REF_DATA, "sst_gridded.zarr"), "to", pinned=False)
    assert dims == ["time", "lat", "lon"] and x.shape == (14611, 20, 40) and x.dtype == np.float32
    assert str(time[0].astype("datetime64[D]")) == "1982-01-01" and str(time[-1].astype("datetime64[D]")) == "2022-01-01"
    store = os.path.join(REF_DATA, "sst_gridded.zarr")
    sub = np.asarray(x)[:, 4:10, 10:18].copy()  # the slice tests/golden/make_golden.py stored (with its injected NaN column)
    sub[:, 1, 1] = np.nan
    np.testing.assert_array_equal(sub, g["sst"])
    np.testing.assert_array_equal(time.astype("datetime64[D]"), g["time"])
    # coordinates are blosc(zstd, BIT-shuffle) in that store
    lat, lon = zio.read_array(store, "lat"), zio.read_array(store, "lon")
    np.testing.assert_allclose(lat, 35.125 + 0.25 * np.arange(20))
    np.testing.assert_allclose(lon, -39.875 + 0.25 * np.arange(40))


def test_cf_mask_and_scale_is_applied_like_xarray(tmp_path):
    """A packed, masked variable (int16 SST with scale_factor / add_offset and a land sentinel; a float variable with a
    1e20 missing_value): ``read_field`` decodes it as ``xr.open_zarr(mask_and_scale=True)`` would, so land is NaN and
    not a finite ocean value (the reference reads its input through xarray)."""
    store = str(tmp_path / "packed.zarr")
    os.makedirs(store)
    json.dump({"zarr_format": 2}, open(os.path.join(store, ".zgroup"), "w"))
    T, ny, nx = 7, 5, 9
    rng = np.random.default_rng(0)
    sst = rng.uniform(-1.8, 30, (T, ny, nx))
    scale, offset, fv = 0.01, 15.0, -32768
    packed = np.round((sst - offset) / scale).astype(np.int16)
    packed[:, 1, 2] = fv  # land
    zio.write_array(store, "sst", packed, (3, ny, nx), ["time", "lat", "lon"],
                        {"scale_factor": scale, "add_offset": offset, "_FillValue": fv})  # fmt: skip
    flt = sst.astype(np.float32)
    flt[:, 3, 4] = np.float32(1e20)
    zio.write_array(store, "sst_f", flt, (4, ny, nx), ["time", "lat", "lon"], {"missing_value": 1e20})
    zio.write_array(store, "time", np.arange(T, dtype=np.int64), (T,), ["time"], {"units": "days since 2000-01-01"})
    x, time, dims = zio.read_field(store, "sst", pinned=False)
    assert x.dtype == np.float32 and dims == ["time", "lat", "lon"] and time[0] == np.datetime64("2000-01-01")
    want = (packed.astype(np.float64) * scale + offset).astype(np.float32)
    want[:, 1, 2] = np.nan
    np.testing.assert_array_equal(np.isnan(x), np.isnan(want))
    np.testing.assert_array_equal(x[~np.isnan(want)], want[~np.isnan(want)])
    xf, _, _ = zio.read_field(store, "sst_f", pinned=False)
    assert np.isnan(xf[:, 3, 4]).all() and np.isfinite(np.delete(xf.reshape(T, -1), 3 * nx + 4, axis=1)).all()
    raw = zio.read_array(store, "sst")  # the raw stored values stay available
    assert raw.dtype == np.int16 and raw[0, 1, 2] == fv
