"""CPU tests of tracker stage 1 (marEx/track.py:1520-1726).

1. The oracle (oracle/track_oracle.py) against scipy.ndimage's own binary_closing / binary_opening, i.e. against the
   reference's non-dask branch (track.py:1646-1660), and against the committed golden vectors.
2. The per-word code of marex_b200/csrc/morph_core.cuh -- the functions the CUDA kernels call -- compiled for the HOST
   (tests/morph_host.cu) and driven by the product's own host logic (marex_b200/track.py with its C-ABI calls
   redirected to the harness), bit for bit against the oracle.  The GPU tests (tests/test_track_gpu.py) then run the
   same cases through the real kernels.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy import ndimage

from oracle import track_oracle as to

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def events_field(T=9, ny=20, nx=45, seed=0, density=0.35, noise=0.03):
    """Blobby random events + ocean mask with a land block and a land column."""
    rng = np.random.default_rng(seed)
    raw = rng.random((T, ny, nx))
    sm = ndimage.uniform_filter(raw, size=(3, 3, 3), mode="wrap")
    ev = sm > np.quantile(sm, 1 - density)
    ev ^= rng.random((T, ny, nx)) < noise  # salt and pepper: small holes and specks
    mask = np.ones((ny, nx), dtype=bool)
    mask[ny // 3 : ny // 3 + 4, nx // 4 : nx // 4 + 6] = False
    mask[:, nx - 3] = False
    mask[0, :5] = False
    return ev & mask, mask


def mesh(n_side=12, seed=0):
    """A triangulated periodic-free patch: cell ids of a (n_side x 2 n_side) strip of triangles, 3 neighbours each,
    -1 at the boundary (the reference drops negative neighbours, track.py:1099-1101)."""
    ncol = 2 * n_side
    N = n_side * ncol
    nb = -np.ones((3, N), dtype=np.int32)
    for r in range(n_side):
        for c in range(ncol):
            i = r * ncol + c
            if c > 0:
                nb[0, i] = i - 1
            if c < ncol - 1:
                nb[1, i] = i + 1
            up = c % 2 == 0  # upward triangles touch the row below, downward ones the row above
            rr = r + 1 if up else r - 1
            cc = c + 1 if up else c - 1
            if 0 <= rr < n_side and 0 <= cc < ncol:
                nb[2, i] = rr * ncol + cc
    rng = np.random.default_rng(seed)
    perm = rng.permutation(N).astype(np.int32)  # unstructured meshes are not stored in raster order
    inv = np.empty_like(perm)
    inv[perm] = np.arange(N, dtype=np.int32)
    nb2 = np.where(nb >= 0, inv[np.clip(nb, 0, None)], -1).astype(np.int32)
    out = np.empty_like(nb2)
    out[:, inv] = nb2
    return np.ascontiguousarray(out)


# ------------------------------------------------------------------------------------------------ oracle pins
@pytest.mark.parametrize("R", [1, 2, 3, 5])
@pytest.mark.parametrize("regional", [False, True])
def test_oracle_equals_reference_scipy_branch(R, regional):
    """track.py:1646-1660: np.pad(diameter) -> scipy binary_closing -> binary_opening -> unpad, per time step."""
    ev, mask = events_field(seed=R)
    got = to.fill_holes(ev, mask, R, regional)
    se = to.disk(R)
    d = 2 * R
    for t in range(ev.shape[0]):
        padded = np.pad(ev[t], ((d, d), (d, d)), mode="edge" if regional else "wrap")
        s2 = ndimage.binary_opening(ndimage.binary_closing(padded, se, iterations=1), se, iterations=1)
        np.testing.assert_array_equal(got[t], s2[d:-d, d:-d] & mask)


def test_oracle_time_closing_is_scipy_binary_closing():
    ev, mask = events_field(T=15, seed=3)
    k = 3
    padded = np.pad(ev, ((k, k), (0, 0), (0, 0)))
    ref = ndimage.binary_closing(padded, np.ones((k, 1, 1), bool))[k:-k]
    np.testing.assert_array_equal(to.fill_time_gaps(ev, mask, 0, 2), ref & mask)
    # a one- or two-step gap is closed, a three-step gap is not (T_fill = 2)
    col = np.zeros((12, 1, 1), bool)
    col[[1, 3, 6, 10], 0, 0] = True
    out = to.fill_time_gaps(col, np.ones((1, 1), bool), 0, 2)[:, 0, 0]
    np.testing.assert_array_equal(out, [0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 0])


def test_oracle_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "track_stage1.npz"))
    ev, mask = g["events"], g["mask"]
    np.testing.assert_array_equal(to.fill_holes(ev, mask, 3), g["fill_holes_R3"])
    np.testing.assert_array_equal(to.stage1(ev, mask, 4, 2), g["stage1_R4_T2"])
    np.testing.assert_array_equal(to.stage1(ev, mask, 3, 4, True), g["stage1_R3_T4_regional"])
    np.testing.assert_array_equal(to.stage1_unstructured(g["events_u"], g["mask_u"], g["neighbours"], 2, 2), g["stage1_u_R2_T2"])


# ------------------------------------------------------------------------------------------------ host harness
@pytest.fixture(scope="session")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("morph") / "libmorph_host.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-o", out, os.path.join(HERE, "morph_host.cu")],
                   check=True, capture_output=True)  # fmt: skip
    lib = ctypes.CDLL(out)
    from marex_b200 import _lib

    for name, (argtypes, restype) in _lib._SIGNATURES.items():
        if name.startswith("marex_morph_"):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return lib


@pytest.fixture()
def host_track(harness, monkeypatch):
    """marex_b200.track with its device set to the CPU and every C-ABI call redirected to the host harness."""
    import torch

    from marex_b200 import track

    def call(name, *args):
        assert name.startswith("marex_morph_"), name
        rc = getattr(harness, name)(*args)
        assert rc == 0, (name, rc)

    monkeypatch.setattr(track, "_device", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(track, "_call", call)
    return track


def test_h1_window_is_exact_for_32_steps(harness):
    rng = np.random.default_rng(0)
    for steps in (1, 7, 31, 32):
        bits = rng.random(96) < 0.05  # words w-1, w, w+1; whatever lies beyond them cannot reach word w in <= 32 steps
        ref = ndimage.binary_dilation(bits, np.ones(2 * steps + 1, bool))[32:64]
        words = np.packbits(bits, bitorder="little").view(np.uint32).copy()
        harness.host_h1_steps(words.ctypes.data_as(ctypes.c_void_p), steps)
        got = np.unpackbits(words.view(np.uint8), bitorder="little")[32:64].astype(bool)
        np.testing.assert_array_equal(got, ref)


GRID_CASES = [
    # (T, ny, nx, R, T_fill, regional, density, noise): densities chosen so that the filled field is neither empty nor full
    (5, 20, 45, 1, 2, False, 0.35, 0.03),
    (5, 20, 45, 3, 2, False, 0.12, 0.03),
    (4, 24, 70, 5, 4, False, 0.03, 0.004),  # two words per row + tail, R//2 = 2
    (4, 24, 64, 4, 2, True, 0.12, 0.03),  # word-aligned rows, edge padding
    (3, 48, 97, 8, 2, False, 0.02, 0.001),  # the usual 0.25-degree radius
    (6, 12, 100, 2, 0, False, 0.35, 0.03),  # T_fill = 0: fill_time_gaps is the identity
    (6, 12, 31, 0, 2, False, 0.35, 0.03),  # R_fill = 0: only the mask and the temporal closing
    (3, 9, 10, 1, 2, False, 0.4, 0.03),  # rows shorter than a word: one flattened word spans several rows
    (3, 10, 9, 1, 2, True, 0.25, 0.02),
    (2, 70, 80, 17, 2, True, 0.004, 0.0),  # pad = 34 > 32: whole words of replicated edge cells
    (2, 70, 80, 17, 2, False, 0.004, 0.0),  # ... and of wrapped cells
]


@pytest.mark.parametrize("separable", [False, True, "disk3", "disk4", "disk4_small_tiles", "pack"])
@pytest.mark.parametrize("T,ny,nx,R,T_fill,regional,density,noise", GRID_CASES)
def test_host_word_code_gridded(host_track, monkeypatch, T, ny, nx, R, T_fill, regional, density, noise, separable):
    ev, mask = events_field(T, ny, nx, seed=R + nx, density=density, noise=noise)
    if separable == "pack":  # bool bytes -> bits first (MAREX_MORPH_PACK=1)
        monkeypatch.setenv("MAREX_MORPH_PACK", "1")
        separable = False
    if separable == "disk3":  # the third variant of the direct disk pass (MAREX_MORPH_DISK=3)
        monkeypatch.setenv("MAREX_MORPH_DISK", "3")
        separable = False
    elif separable in ("disk4", "disk4_small_tiles"):  # the shared-memory tile variant
        monkeypatch.setenv("MAREX_MORPH_DISK", "4")
        if separable == "disk4_small_tiles":  # 4- or 8-row tiles: several tiles per time step, a ragged last tile
            nlev = max(1, host_track._lib.load().marex_morph_disk_levels(R))
            monkeypatch.setenv("MORPH_HOST_TILE_BUDGET", str((1 + nlev) * (8 + 2 * R) * ((nx + 4 * R + 31) // 32) * 4))
        separable = False
    if separable:  # a scratch of two time steps' level buffers: the chunk loop runs several times, with a ragged tail
        nlev = max(1, host_track._lib.load().marex_morph_disk_levels(R))
        monkeypatch.setattr(host_track, "SEPARABLE_SCRATCH_BYTES", 2 * 4 * nlev * (ny + 4 * R) * ((nx + 4 * R + 31) // 32))
    f = host_track.MaskFiller(mask, R, T_fill, regional, separable=separable)
    ref_h = to.fill_holes(ev, mask, R, regional)
    got_h = f.fill_holes(ev)
    np.testing.assert_array_equal(got_h, ref_h)
    assert f.last_count == int(ref_h.sum())
    assert 0.02 < ref_h.mean() < 0.9 * mask.mean()  # a saturated field would test nothing
    ref_t = to.fill_time_gaps(ref_h, mask, R, T_fill, regional)
    np.testing.assert_array_equal(f.fill_time_gaps(ref_h), ref_t)
    np.testing.assert_array_equal(f.run(ev), ref_t)
    assert f.last_count == int(ref_t.sum())
    # packed in / packed out (the layout of marex_compare_*)
    import torch

    flat = ev.reshape(T, -1)
    nw = (flat.shape[1] + 31) // 32
    padded = np.zeros((T, nw * 32), bool)
    padded[:, : flat.shape[1]] = flat
    bits = np.packbits(padded, axis=1, bitorder="little").view(np.uint32).view(np.int32)
    out_bits = f.run(from_bits=(torch.from_numpy(bits.copy()), T), packed=True).numpy()
    ref_bits = np.zeros((T, nw * 32), bool)
    ref_bits[:, : flat.shape[1]] = ref_t.reshape(T, -1)
    np.testing.assert_array_equal(out_bits.view(np.uint32), np.packbits(ref_bits, axis=1, bitorder="little").view(np.uint32))


@pytest.mark.parametrize("pack", [False, True])
@pytest.mark.parametrize("T,R,T_fill", [(7, 1, 2), (40, 2, 2), (70, 3, 4), (33, 0, 2), (9, 2, 0)])
def test_host_word_code_unstructured(host_track, monkeypatch, T, R, T_fill, pack):
    if pack:
        monkeypatch.setenv("MAREX_MORPH_PACK", "1")
    nb = mesh(10, seed=T)
    N = nb.shape[1]
    rng = np.random.default_rng(T)
    mask = rng.random(N) > 0.15
    ev = (rng.random((T, N)) < 0.3) & mask
    ev[:, :5] |= rng.random((T, 5)) < 0.5
    f = host_track.MaskFiller(mask, R, T_fill, neighbours=nb)
    ref_h = to.fill_holes_unstructured(ev, mask, nb, R)
    np.testing.assert_array_equal(f.fill_holes(ev), ref_h)
    ref_t = to.fill_time_gaps_unstructured(ref_h, mask, nb, R, T_fill)
    np.testing.assert_array_equal(f.fill_time_gaps(ref_h), ref_t)
    np.testing.assert_array_equal(f.run(ev), ref_t)
    assert f.last_count == int(ref_t.sum())


def test_validation_messages(host_track):
    from marex_b200.exceptions import ConfigurationError, DataValidationError

    mask = np.ones((8, 8), bool)
    with pytest.raises(ConfigurationError, match="T_fill must be even for temporal symmetry"):
        host_track.MaskFiller(mask, 2, 3)
    with pytest.raises(ConfigurationError, match="outside the range"):
        host_track.MaskFiller(mask, 40, 2)
    with pytest.raises(DataValidationError, match="does not match the mask"):
        host_track.MaskFiller(mask, 1, 2).fill_holes(np.zeros((3, 8, 9), bool))
    with pytest.raises(DataValidationError, match="smaller than the padding"):
        host_track.MaskFiller(mask, 5, 2).fill_holes(np.zeros((3, 8, 8), bool))
    with pytest.raises(NotImplementedError):
        host_track.MaskFiller(np.ones(4, bool), 1, 2, regional_mode=True, neighbours=-np.ones((3, 4), np.int32))


def test_product_track_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from marex_b200 import track

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        track.MaskFiller(np.ones((8, 8), bool), 2, 2)


def test_api_forms_and_input_kinds(host_track):
    """R_fill override, functional forms, torch / integer inputs, packed output shape, from_bits validation."""
    import torch

    from marex_b200.exceptions import ConfigurationError, DataValidationError

    ev, mask = events_field(5, 20, 45, seed=9, density=0.12)
    f = host_track.MaskFiller(mask, 3, 2)
    np.testing.assert_array_equal(f.fill_holes(ev, R_fill=1), to.fill_holes(ev, mask, 1))  # fill_time_gaps calls it with R_fill // 2
    np.testing.assert_array_equal(host_track.fill_holes(ev, mask, 3), to.fill_holes(ev, mask, 3))
    np.testing.assert_array_equal(host_track.fill_time_gaps(ev, mask, 3, 2), to.fill_time_gaps(ev, mask, 3, 2))
    got = f.run(torch.from_numpy(ev.astype(np.int32) * 7))  # any integer dtype counts as "!= 0"; a tensor in gives a tensor out
    assert isinstance(got, torch.Tensor) and got.dtype == torch.bool and tuple(got.shape) == ev.shape
    np.testing.assert_array_equal(got.numpy(), to.stage1(ev, mask, 3, 2))
    packed = f.run(ev, packed=True)
    assert packed.dtype == np.int32 and packed.shape == (5, (20 * 45 + 31) // 32)
    with pytest.raises(ConfigurationError, match="outside the range"):
        f.fill_holes(ev, R_fill=33)
    with pytest.raises(DataValidationError, match="from_bits expects"):
        f.run(from_bits=(torch.zeros((5, 3), dtype=torch.int32), 5))  # too few words for 900 cells
    with pytest.raises(DataValidationError, match="from_bits expects"):
        f.run(from_bits=(torch.zeros((5, 29), dtype=torch.float32), 5))
    with pytest.raises(DataValidationError, match="from_bits expects"):
        f.run(from_bits=(torch.zeros((5, 29), dtype=torch.int32), 6))  # more time steps than rows
    # T_fill = 0: fill_time_gaps hands the data back untouched, even outside the mask (track.py:1692-1693)
    f0 = host_track.MaskFiller(mask, 3, 0)
    raw = np.ones_like(ev)
    np.testing.assert_array_equal(f0.fill_time_gaps(raw), raw)


# ------------------------------------------------------------------ world_size-2 gloo run of the time-sharded stage 1
_SHARD_WORKER = r"""
import ctypes, os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MAREX_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MAREX_ROOT"], "tests"))
from marex_b200 import _lib, track
from oracle import track_oracle as to
from test_track_cpu import events_field
lib = ctypes.CDLL(os.environ["MORPH_HOST_LIB"])
for name, (argtypes, restype) in _lib._SIGNATURES.items():
    if name.startswith("marex_morph_"):
        fn = getattr(lib, name); fn.argtypes, fn.restype = argtypes, restype
def call(name, *args):
    assert getattr(lib, name)(*args) == 0, name
track._device = lambda device=None: torch.device("cpu")   # the kernels' per-word code on the host (tests/morph_host.cu)
track._call = call
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
T, ny, nx, R, T_fill = 11, 16, 64, 2, 4
ev, mask = events_field(T, ny, nx, seed=4, density=0.2)
rows = ny // world
band = ev[:, rank * rows:(rank + 1) * rows].reshape(T, -1)
bits = torch.from_numpy(np.packbits(band, axis=1, bitorder="little").view(np.uint32).view(np.int32).copy())
ref = to.stage1(ev, mask, R, T_fill)
out, (lo, hi) = track.stage1_time_sharded(bits, T, rows, mask, R, T_fill, packed=False)
assert np.array_equal(out.numpy(), ref[lo:hi]), (rank, lo, hi)
outp, _ = track.stage1_time_sharded(bits, T, rows, mask, R, T_fill, packed=True)
flat = np.unpackbits(outp.numpy().view(np.uint8), axis=1, bitorder="little")[:, : ny * nx].astype(bool)
assert np.array_equal(flat.reshape(hi - lo, ny, nx), ref[lo:hi])
n = torch.tensor([hi - lo]); dist.all_reduce(n)
if rank == 0:
    assert int(n) == T
    print("STAGE1_GLOO_OK", lo, hi)
dist.destroy_process_group()
"""


def test_time_blocks_cover_the_series():
    from marex_b200.track import time_blocks

    for T, world, halo in [(11, 2, 4), (9131, 8, 2), (5, 8, 2), (100, 3, 0)]:
        b = time_blocks(T, world, halo)
        assert b[0][0] == 0 and b[-1][1] == T and len(b) == world
        for (lo, hi, llo, lhi), nxt in zip(b, b[1:] + [None]):
            assert llo == max(0, lo - halo) and lhi == min(T, hi + halo) and lo <= hi
            if nxt is not None:
                assert hi == nxt[0]


def test_stage1_time_sharded_world2_gloo(harness, tmp_path):
    import socket
    import sys

    script = tmp_path / "worker.py"
    script.write_text(_SHARD_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MAREX_ROOT=ROOT, MORPH_HOST_LIB=harness._name, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "STAGE1_GLOO_OK" in res.stdout


def test_pack_word_aligned_and_ragged(harness):
    """morph_pack_word: the 16-byte path (multiply-gather) and the byte path agree with numpy's packbits."""
    rng = np.random.default_rng(1)
    for N in (64, 96, 70, 31, 1):
        T = 3
        pitch = 112  # a multiple of 16: rows of the aligned case start on 16-byte boundaries
        buf = np.zeros(T * pitch + 16, np.uint8)
        off = (-buf.ctypes.data) % 16
        ev = buf[off : off + T * pitch].reshape(T, pitch)
        ev[:, :N] = rng.random((T, N)) < 0.4
        nw = (N + 31) // 32
        bits = np.zeros((T, nw), np.uint32)
        rc = harness.marex_morph_pack_u8(ev.ctypes.data_as(ctypes.c_void_p), T, N, pitch, bits.ctypes.data_as(ctypes.c_void_p), nw, None)
        assert rc == 0
        padded = np.zeros((T, nw * 32), bool)
        padded[:, :N] = ev[:, :N] != 0
        np.testing.assert_array_equal(bits, np.packbits(padded, axis=1, bitorder="little").view(np.uint32))
