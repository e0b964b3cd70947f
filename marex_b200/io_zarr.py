"""
Zarr v2 I/O edge for the detection path: read a (time, lat, lon) /
(time, ncells) variable and its time coordinate from a zarr v2 directory store straight into a
page-locked host buffer laid out as the kernels want it (time-major float32), run
``preprocess_arrays`` on it, and write the outputs back as a zarr v2 group.

The reference reads and writes its data through xarray + zarr + numcodecs
(``examples/batch jobs/run_detect.py:54-82``; its own fixtures are zarr v2 with blosc(lz4,
byte-shuffle) chunks, ``tests/data/*.zarr/.zmetadata``).  None of those packages is installed
here, so this module implements the small part of the format the path needs:

* metadata: ``.zarray`` (shape, chunks, dtype, C order, fill_value), ``.zattrs``
  (``_ARRAY_DIMENSIONS``, CF ``units`` / ``calendar`` of the time coordinate);
* codecs: ``null`` (raw), ``blosc`` v1 frames (memcpy, lz4, zstd, zlib; byte and bit shuffle; split and
  unsplit blocks), ``zlib`` / ``gzip``, ``zstd``, ``lz4`` (numcodecs framing) -- lz4 and zstd
  streams are inflated with pyarrow's codecs;
* writing: raw (``compressor: null``) C-order chunks, which every zarr reader accepts.

This is host-side plumbing: nothing here launches a kernel.
"""
from __future__ import annotations

import json
import os
import struct
import zlib
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# codecs
# --------------------------------------------------------------------------------------
def _pa_inflate(codec: str, buf: bytes, n: int) -> bytes:
    import pyarrow as pa

    return pa.Codec(codec).decompress(buf, decompressed_size=n).to_pybytes()


def _unshuffle(block: bytes, typesize: int) -> bytes:
    a = np.frombuffer(block, dtype=np.uint8)
    ne = a.size // typesize
    body = a[: ne * typesize].reshape(typesize, ne).T.reshape(-1)
    return body.tobytes() + a[ne * typesize :].tobytes()


def _unbitshuffle(block: bytes, typesize: int) -> bytes:
    """Inverse of blosc's bit-shuffle: the block stores, for every bit of an element (byte k, bit j), one row
    of n/8 bytes holding that bit of 8 consecutive elements (element 8m + i in bit i of byte m); elements
    beyond a multiple of 8 are stored unchanged."""
    a = np.frombuffer(block, dtype=np.uint8)
    ne = a.size // typesize
    ne8 = ne - ne % 8
    if ne8 == 0:
        return block
    bits = np.unpackbits(a[: ne8 * typesize].reshape(typesize * 8, ne8 // 8), axis=1, bitorder="little")
    bits = bits.reshape(typesize, 8, ne8).transpose(2, 0, 1)  # [element][byte][bit]
    body = np.packbits(bits, axis=2, bitorder="little").reshape(-1)
    return body.tobytes() + a[ne8 * typesize :].tobytes()


def blosc_decompress(buf: bytes) -> bytes:
    """One blosc v1 frame: 16-byte header, block-start table, per-block (optionally split) streams."""
    ver, verlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack("<BBBBIII", buf[:16])
    if flags & 0x2:  # memcpy frame
        return bytes(buf[16 : 16 + nbytes])
    bitshuffle = bool(flags & 0x4)
    fmt = flags >> 5
    inflate = {
        0: lambda b, n: _blosclz_unsupported(),
        1: lambda b, n: _pa_inflate("lz4_raw", b, n),
        3: lambda b, n: zlib.decompress(b),
        4: lambda b, n: _pa_inflate("zstd", b, n),
    }.get(fmt)
    if inflate is None:
        raise NotImplementedError(f"blosc compressor format {fmt} is not supported")
    doshuffle, dont_split = bool(flags & 0x1), bool(flags & 0x10)
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = struct.unpack(f"<{nblocks}i", buf[16 : 16 + 4 * nblocks])
    out = bytearray()
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split) and typesize <= 16 and bsize // typesize >= 128 and not leftover
        nstreams = typesize if split else 1
        neblock = bsize // nstreams
        p = bstarts[b]
        blk = bytearray()
        for _ in range(nstreams):
            (cb,) = struct.unpack("<i", buf[p : p + 4])
            p += 4
            blk += buf[p : p + cb] if cb == neblock else inflate(bytes(buf[p : p + cb]), neblock)
            p += cb
        if bitshuffle:
            out += _unbitshuffle(bytes(blk), typesize)
        elif doshuffle and typesize > 1:
            out += _unshuffle(bytes(blk), typesize)
        else:
            out += blk
    return bytes(out[:nbytes])


def _blosclz_unsupported():
    raise NotImplementedError("blosc 'blosclz' streams are not supported (re-encode with lz4 / zstd / zlib)")


def _decode_chunk(raw: bytes, compressor: Optional[Dict[str, Any]], nbytes: int) -> bytes:
    if compressor is None:
        return raw
    cid = compressor.get("id")
    if cid == "blosc":
        return blosc_decompress(raw)
    if cid in ("zlib", "gzip"):
        return zlib.decompress(raw, 15 + 32)
    if cid == "zstd":
        return _pa_inflate("zstd", raw, nbytes)
    if cid == "lz4":  # numcodecs.LZ4: 4-byte little-endian uncompressed size, then a raw lz4 block
        (n,) = struct.unpack("<i", raw[:4])
        return _pa_inflate("lz4_raw", raw[4:], n)
    raise NotImplementedError(f"zarr compressor '{cid}' is not supported")


# --------------------------------------------------------------------------------------
# reading
# --------------------------------------------------------------------------------------
def array_meta(store: str, name: str) -> Dict[str, Any]:
    """``.zarray`` + ``.zattrs`` of one array of a zarr v2 directory store."""
    path = os.path.join(store, name)
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    if meta.get("zarr_format") != 2:
        raise NotImplementedError("only zarr format 2 is supported")
    if meta.get("order", "C") != "C":
        raise NotImplementedError("only C-ordered chunks are supported")
    if meta.get("filters"):
        raise NotImplementedError("zarr filters are not supported")
    attrs_p = os.path.join(path, ".zattrs")
    meta["attrs"] = json.load(open(attrs_p)) if os.path.exists(attrs_p) else {}
    meta["path"] = path
    meta["dimension_separator"] = meta.get("dimension_separator", ".")
    return meta


def _fill_value(meta: Dict[str, Any], dt: np.dtype):
    fv = meta.get("fill_value")
    if fv is None:
        return np.nan if dt.kind == "f" else 0
    if isinstance(fv, str):
        return {"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(fv, 0)
    return fv


def _cf_decoder(attrs: Dict[str, Any], dt: np.dtype):
    """xarray's ``mask_and_scale=True`` (what ``xr.open_zarr`` applies before marEx sees the data): values equal to
    ``_FillValue`` / ``missing_value`` become NaN, then ``x * scale_factor + add_offset``.  Returns None when the
    variable carries none of these attributes."""
    fills = []
    for key in ("_FillValue", "missing_value"):
        v = attrs.get(key)
        if v is None:
            continue
        for item in (v if isinstance(v, (list, tuple)) else [v]):
            if isinstance(item, str):
                item = {"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(item, item)
            fills.append(np.array(item).astype(dt))
    scale, offset = attrs.get("scale_factor"), attrs.get("add_offset")
    if not fills and scale is None and offset is None:
        return None

    def decode(blk: np.ndarray) -> np.ndarray:
        bad = np.zeros(blk.shape, dtype=bool)
        for fv in fills:
            bad |= np.isnan(blk) if (fv.dtype.kind == "f" and np.isnan(fv)) else (blk == fv)
        val = blk.astype(np.float64 if dt.itemsize > 4 or scale is not None or offset is not None else np.float32)
        if scale is not None:
            val = val * np.float64(scale)
        if offset is not None:
            val = val + np.float64(offset)
        val[bad] = np.nan
        return val

    return decode


def read_array(store: str, name: str, out: Optional[np.ndarray] = None, dtype=None, decode_cf: bool = False) -> np.ndarray:
    """Read a whole array.  ``out`` (e.g. the numpy view of a page-locked torch tensor) receives the data,
    converted to its dtype chunk by chunk, so a float64 store lands as the float32 field the kernels read
    (``da.astype(np.float32)``, detect.py:600) without a second full-size copy.  ``decode_cf`` applies the variable's
    ``_FillValue`` / ``missing_value`` / ``scale_factor`` / ``add_offset`` attributes chunk by chunk (packed int16 SST,
    land sentinels), as xarray does when it opens the store."""
    meta = array_meta(store, name)
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dt = np.dtype(meta["dtype"])
    if out is None:
        out = np.empty(shape, dtype=dtype or dt)
    if tuple(out.shape) != shape:
        raise ValueError(f"out has shape {out.shape}, the array has {shape}")
    fill = _fill_value(meta, dt)
    decode = _cf_decoder(meta["attrs"], dt) if decode_cf else None
    if decode is not None and out.dtype.kind != "f":
        raise ValueError("decode_cf needs a floating-point destination")
    sep = meta["dimension_separator"]
    grid = [(s + c - 1) // c for s, c in zip(shape, chunks)] if shape else []
    chunk_bytes = int(np.prod(chunks)) * dt.itemsize if shape else dt.itemsize
    for idx in np.ndindex(*grid) if shape else [()]:
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        f = os.path.join(meta["path"], sep.join(str(i) for i in idx) if idx else "0")
        if not os.path.exists(f):
            out[sl] = decode(np.full((1,) * len(shape), fill, dtype=dt)) if decode is not None else fill
            continue
        with open(f, "rb") as fh:
            raw = _decode_chunk(fh.read(), meta.get("compressor"), chunk_bytes)
        blk = np.frombuffer(raw, dtype=dt).reshape(chunks if shape else ())
        blk = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        out[sl] = decode(blk) if decode is not None else blk
    return out


_UNIT_NS = {"seconds": 10**9, "second": 10**9, "minutes": 60 * 10**9, "hours": 3600 * 10**9, "hour": 3600 * 10**9,
            "days": 86400 * 10**9, "day": 86400 * 10**9, "milliseconds": 10**6, "microseconds": 10**3, "nanoseconds": 1}  # fmt: skip


def decode_cf_time(values: np.ndarray, units: str, calendar: str = "standard") -> np.ndarray:
    """CF ``"<unit> since <epoch>"`` numbers -> ``datetime64[ns]`` (standard / proleptic Gregorian only; the
    kernels' calendar tables assume it)."""
    if calendar not in ("standard", "gregorian", "proleptic_gregorian"):
        raise NotImplementedError(f"calendar '{calendar}' is not supported (standard / proleptic_gregorian only)")
    unit, _, epoch = units.partition(" since ")
    if unit.strip() not in _UNIT_NS or not epoch:
        raise ValueError(f"cannot parse time units '{units}'")
    t0 = np.datetime64(epoch.strip().replace(" ", "T").rstrip("Z"), "ns")
    v = np.asarray(values)
    if v.dtype.kind == "f":
        ns = np.round(v.astype(np.float64) * _UNIT_NS[unit.strip()]).astype(np.int64)
    else:
        ns = v.astype(np.int64) * _UNIT_NS[unit.strip()]
    return t0 + ns.astype("timedelta64[ns]")


def read_field(store: str, var: str, time_name: Optional[str] = None, pinned: bool = True):
    """``(x, time, dims)``: the variable as a time-major float32 array (a numpy view of a page-locked torch
    tensor when ``pinned`` and torch is available) and its decoded time axis.  The variable's first
    ``_ARRAY_DIMENSIONS`` entry must be its time dimension (transpose upstream otherwise)."""
    meta = array_meta(store, var)
    dims = list(meta["attrs"].get("_ARRAY_DIMENSIONS", []))
    tname = time_name or (dims[0] if dims else "time")
    if dims and dims[0] != tname:
        raise NotImplementedError(f"'{var}' has dimensions {dims}: time ('{tname}') must come first")
    shape = tuple(meta["shape"])
    holder = None
    if pinned:
        try:
            import torch

            holder = torch.empty(shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
            out = holder.numpy()
        except Exception:
            out = np.empty(shape, dtype=np.float32)
    else:
        out = np.empty(shape, dtype=np.float32)
    read_array(store, var, out=out, decode_cf=True)
    tmeta = array_meta(store, tname)
    tvals = read_array(store, tname)
    units = tmeta["attrs"].get("units")
    time = decode_cf_time(tvals, units, tmeta["attrs"].get("calendar", "standard")) if units else tvals.astype("datetime64[ns]")
    return (holder if holder is not None else out), time, dims


# --------------------------------------------------------------------------------------
# writing
# --------------------------------------------------------------------------------------
def write_array(store: str, name: str, a: np.ndarray, chunks: Sequence[int], dims: Sequence[str],
                attrs: Optional[Dict[str, Any]] = None) -> None:  # fmt: skip
    """Write ``a`` as a zarr v2 array with raw (uncompressed) C-order chunks."""
    a = np.asarray(a)
    chunks = tuple(int(min(c, s)) if s else 1 for c, s in zip(chunks, a.shape))
    path = os.path.join(store, name)
    os.makedirs(path, exist_ok=True)
    dt = a.dtype
    dstr = "|b1" if dt == np.bool_ else (dt.str if dt.itemsize > 1 else "|" + dt.str[1:])
    fill = "NaN" if dt.kind == "f" else (False if dt == np.bool_ else 0)
    meta = {"chunks": list(chunks), "compressor": None, "dtype": dstr, "fill_value": fill, "filters": None, "order": "C",
            "shape": list(a.shape), "zarr_format": 2}  # fmt: skip
    json.dump(meta, open(os.path.join(path, ".zarray"), "w"), indent=1)
    json.dump({"_ARRAY_DIMENSIONS": list(dims), **(attrs or {})}, open(os.path.join(path, ".zattrs"), "w"), indent=1)
    grid = [(s + c - 1) // c for s, c in zip(a.shape, chunks)]
    for idx in np.ndindex(*grid):
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, a.shape))
        blk = np.zeros(chunks, dtype=dt) if dt.kind != "f" else np.full(chunks, np.nan, dtype=dt)
        blk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
        with open(os.path.join(path, ".".join(str(i) for i in idx)), "wb") as fh:
            fh.write(np.ascontiguousarray(blk).tobytes())


def write_result(store: str, res: Dict[str, Any], space_dims: Sequence[str], time_chunk: int = 25,
                 time_name: str = "time") -> None:  # fmt: skip
    """The output of ``preprocess_arrays`` as a zarr v2 group: ``dat_anomaly`` / ``extreme_events`` chunked
    ``(time_chunk, all space)`` like the reference's ``dask_chunks={"time": 25}`` (detect.py:532-535, 786-792),
    ``thresholds`` / ``mask`` in one chunk, time as days since 1970-01-01, attrs in the group's ``.zattrs``."""
    os.makedirs(store, exist_ok=True)
    json.dump({"zarr_format": 2}, open(os.path.join(store, ".zgroup"), "w"))
    attrs = {k: (v if not isinstance(v, tuple) else list(v)) for k, v in res.get("attrs", {}).items()}
    json.dump(attrs, open(os.path.join(store, ".zattrs"), "w"), indent=1)
    sd = list(space_dims)
    t = np.asarray(res["time"]).astype("datetime64[D]").astype(np.int64)
    write_array(store, time_name, t, (len(t),), [time_name], {"units": "days since 1970-01-01", "calendar": "proleptic_gregorian"})
    nt = np.asarray(res["dat_anomaly"]).shape[0]
    tc = (min(time_chunk, nt),) + tuple(np.asarray(res["dat_anomaly"]).shape[1:])
    write_array(store, "dat_anomaly", np.asarray(res["dat_anomaly"]), tc, [time_name] + sd)
    if "extreme_events" in res:
        write_array(store, "extreme_events", np.asarray(res["extreme_events"]), tc, [time_name] + sd)
    write_array(store, "mask", np.asarray(res["mask"]), np.asarray(res["mask"]).shape, sd)
    thr = np.asarray(res["thresholds"])
    lay = res.get("thresholds_layout", "space")
    tdims = sd + ["dayofyear"] if lay == "doy_last" else (["dayofyear"] + sd if lay == "doy_first" else sd)
    write_array(store, "thresholds", thr, thr.shape, tdims)
    if lay != "space":
        write_array(store, "dayofyear", np.arange(1, 367, dtype=np.int32), (366,), ["dayofyear"])


def preprocess_zarr(store: str, var: str, out_store: Optional[str] = None, time_name: Optional[str] = None, **kwargs):
    """Read ``var`` from a zarr v2 store into page-locked memory, run the detection pipeline
    (``preprocess_arrays``, streamed through the GPU in spatial chunks) and optionally write the result."""
    from .detect import preprocess_arrays

    x, time, dims = read_field(store, var, time_name)
    res = preprocess_arrays(x, time.astype("datetime64[D]"), **kwargs)
    if out_store is not None:
        write_result(out_store, res, dims[1:] if dims else [f"dim_{i}" for i in range(np.asarray(res["mask"]).ndim)],
                     time_name=dims[0] if dims else "time")  # fmt: skip
    return res
