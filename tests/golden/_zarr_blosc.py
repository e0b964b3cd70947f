"""Minimal zarr-v2 + blosc1 (lz4, byte-shuffle) chunk reader used ONLY by make_golden.py
to read the reference's own test fixture (/root/reference/tests/data/*.zarr).
No zarr/numcodecs in this image, so: 16-byte blosc header, per-block start table,
per-block (optionally split) lz4 streams decoded with pyarrow's ``lz4_raw`` codec."""
import json
import os
import struct

import numpy as np
import pyarrow as pa


def _inflate(buf: bytes, n: int, fmt: int) -> bytes:
    codec = {1: "lz4_raw", 4: "zstd"}[fmt]  # blosc flags bits 5-7: 1 = lz4, 4 = zstd
    return pa.Codec(codec).decompress(buf, decompressed_size=n).to_pybytes()


def blosc_decompress(buf: bytes) -> bytes:
    ver, verlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack("<BBBBIII", buf[:16])
    if flags & 0x2:  # memcpy
        return buf[16 : 16 + nbytes]
    doshuffle = bool(flags & 0x1)
    dont_split = bool(flags & 0x10)
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
            blk += buf[p : p + cb] if cb == neblock else _inflate(buf[p : p + cb], neblock, flags >> 5)
            p += cb
        if doshuffle and typesize > 1:
            a = np.frombuffer(bytes(blk), dtype=np.uint8)
            ne = bsize // typesize
            body = a[: ne * typesize].reshape(typesize, ne).T.reshape(-1)
            blk = bytearray(body.tobytes() + a[ne * typesize :].tobytes())
        out += blk
    return bytes(out[:nbytes])


def read_zarr_array(path: str) -> np.ndarray:
    meta = json.load(open(os.path.join(path, ".zarray")))
    shape, chunks = meta["shape"], meta["chunks"]
    dt = np.dtype(meta["dtype"])
    out = np.full(shape, np.nan if dt.kind == "f" else 0, dtype=dt)
    grid = [(s + c - 1) // c for s, c in zip(shape, chunks)]
    for idx in np.ndindex(*grid):
        f = os.path.join(path, ".".join(str(i) for i in idx))
        if not os.path.exists(f):
            continue
        raw = blosc_decompress(open(f, "rb").read())
        blk = np.frombuffer(raw, dtype=dt).reshape(chunks)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out
