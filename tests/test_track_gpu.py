"""GPU parity tests of tracker stage 1: marex_b200.track.MaskFiller -> C-ABI -> morph.cu kernels,
bit for bit against the scipy-based oracle (oracle/track_oracle.py) and the golden vectors produced by the
reference's own nested functions (tests/golden/make_track_golden.py)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import track_oracle as to  # noqa: E402
from test_track_cpu import GRID_CASES, events_field, mesh  # noqa: E402


def _track():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from marex_b200 import track

    return track


def _pack(flat):
    T, N = flat.shape
    nw = (N + 31) // 32
    padded = np.zeros((T, nw * 32), bool)
    padded[:, :N] = flat
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32)


@pytest.mark.parametrize("separable", [False, True])
@pytest.mark.parametrize("T,ny,nx,R,T_fill,regional,density,noise", GRID_CASES)
def test_gridded_fill(monkeypatch, T, ny, nx, R, T_fill, regional, density, noise, separable):
    track = _track()
    ev, mask = events_field(T, ny, nx, seed=R + nx, density=density, noise=noise)
    if separable:  # scratch for two time steps: several chunks with a ragged tail
        nlev = max(1, track._lib.load().marex_morph_disk_levels(R))
        monkeypatch.setattr(track, "SEPARABLE_SCRATCH_BYTES", 2 * 4 * nlev * (ny + 4 * R) * ((nx + 4 * R + 31) // 32))
    f = track.MaskFiller(mask, R, T_fill, regional, separable=separable)
    ref_h = to.fill_holes(ev, mask, R, regional)
    np.testing.assert_array_equal(f.fill_holes(ev), ref_h)
    assert f.last_count == int(ref_h.sum())
    ref_t = to.fill_time_gaps(ref_h, mask, R, T_fill, regional)
    np.testing.assert_array_equal(f.fill_time_gaps(ref_h), ref_t)
    got = f.run(torch.from_numpy(ev).cuda())  # tensor in -> CUDA tensor out
    assert got.is_cuda and got.dtype == torch.bool
    np.testing.assert_array_equal(got.cpu().numpy(), ref_t)
    bits = torch.from_numpy(_pack(ev.reshape(T, -1)).view(np.int32).copy()).cuda()
    out_bits = f.run(from_bits=(bits, T), packed=True)
    np.testing.assert_array_equal(out_bits.cpu().numpy().view(np.uint32), _pack(ref_t.reshape(T, -1)))
    assert f.last_count == int(ref_t.sum())


@pytest.mark.parametrize("T,R,T_fill", [(7, 1, 2), (40, 2, 2), (70, 3, 4), (33, 0, 2), (9, 2, 0)])
def test_unstructured_fill(T, R, T_fill):
    track = _track()
    nb = mesh(10, seed=T)
    N = nb.shape[1]
    rng = np.random.default_rng(T)
    mask = rng.random(N) > 0.15
    ev = (rng.random((T, N)) < 0.3) & mask
    f = track.MaskFiller(mask, R, T_fill, neighbours=nb)
    ref_h = to.fill_holes_unstructured(ev, mask, nb, R)
    np.testing.assert_array_equal(f.fill_holes(ev), ref_h)
    ref_t = to.fill_time_gaps_unstructured(ref_h, mask, nb, R, T_fill)
    np.testing.assert_array_equal(f.fill_time_gaps(ref_h), ref_t)
    np.testing.assert_array_equal(f.run(ev), ref_t)
    assert f.last_count == int(ref_t.sum())
    out_bits = f.run(ev, packed=True)
    np.testing.assert_array_equal(out_bits.view(np.uint32), _pack(ref_t))


def test_golden_reference_functions(golden_dir):
    """Outputs of the reference's own `binary_open_close` / `sparse_bool_power` (AST-extracted, track.py:1549-1582,
    1646-1660, 5423-5470)."""
    track = _track()
    g = np.load(os.path.join(golden_dir, "track_stage1.npz"))
    ev, mask = g["events"], g["mask"]
    np.testing.assert_array_equal(track.MaskFiller(mask, 3, 2).fill_holes(ev), g["fill_holes_R3"])
    np.testing.assert_array_equal(track.MaskFiller(mask, 4, 2).run(ev), g["stage1_R4_T2"])
    np.testing.assert_array_equal(track.MaskFiller(mask, 3, 4, regional_mode=True).run(ev), g["stage1_R3_T4_regional"])
    f = track.MaskFiller(g["mask_u"], 2, 2, neighbours=g["neighbours"])
    np.testing.assert_array_equal(f.run(g["events_u"]), g["stage1_u_R2_T2"])


def test_quarter_degree_slices():
    """Full 0.25-degree time steps (720 x 1440, 45 words per row, R_fill = 8): many CTAs, rows of several words."""
    track = _track()
    T, ny, nx = 4, 720, 1440
    ev, mask = events_field(T, ny, nx, seed=5, density=0.02, noise=0.0005)
    ref = to.stage1(ev, mask, 8, 2)
    assert 0.02 < ref.mean() < 0.9
    np.testing.assert_array_equal(track.MaskFiller(mask, 8, 2, separable=False).run(ev), ref)
    np.testing.assert_array_equal(track.MaskFiller(mask, 8, 2, separable=True).run(ev), ref)


def test_stage1_consumes_the_packed_mask_of_preprocess():
    """preprocess_arrays -> extreme_events -> stage 1, against the oracle on the same events."""
    track = _track()
    import marex_b200

    rng = np.random.default_rng(1)
    time = np.arange(np.datetime64("2000-01-01"), np.datetime64("2009-01-01"))
    T, ny, nx = len(time), 16, 40
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    x = (15 + 5 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((T, ny, nx))).astype(np.float32)
    x[:, 3:6, 10:14] = np.nan
    res = marex_b200.preprocess_arrays(x, time, window_year_baseline=4, smooth_days_baseline=11, window_days_hobday=5,
                                       output="torch", want_bits=True)  # fmt: skip
    ev = res["extreme_events"].cpu().numpy().astype(bool)
    mask = res["mask"].cpu().numpy().astype(bool).reshape(ny, nx)
    ref = to.stage1(ev, mask, 2, 2)
    f = track.MaskFiller(mask, 2, 2)
    got = f.run(from_bits=(res["bits"], ev.shape[0]))  # the bit-packed mask straight from marex_compare_hobday
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    np.testing.assert_array_equal(f.run(res["extreme_events"]).cpu().numpy(), ref)


@pytest.fixture
def tune():
    import marex_b200

    pinned = []

    def _tune(**kw):
        pinned.extend(kw)
        marex_b200._lib.tune(**kw)

    yield _tune
    marex_b200._lib.tune(**{k: None for k in pinned})


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("T,ny,nx,R,T_fill,regional,density,noise", GRID_CASES)
def test_gridded_fill_direct_disk_variants(tune, T, ny, nx, R, T_fill, regional, density, noise, variant):
    """The default disk pass is the shared-memory tile kernel (morph_disk = 4, measured fastest); the two direct
    kernels (2: first version, 3: branch-free borders + funnel-shift widening, also the fall-back when a tile does not
    fit shared memory) give the same bits."""
    track = _track()
    tune(morph_disk=variant)
    ev, mask = events_field(T, ny, nx, seed=R + nx, density=density, noise=noise)
    np.testing.assert_array_equal(track.MaskFiller(mask, R, T_fill, regional).run(ev), to.stage1(ev, mask, R, T_fill, regional))


@pytest.mark.parametrize("variant", [2, 3, 4])
def test_quarter_degree_slices_disk_variants(tune, variant):
    track = _track()
    tune(morph_disk=variant)
    ev, mask = events_field(4, 720, 1440, seed=5, density=0.02, noise=0.0005)
    np.testing.assert_array_equal(track.MaskFiller(mask, 8, 2).run(ev), to.stage1(ev, mask, 8, 2))


@pytest.mark.parametrize("pack", ["1", "0"])
def test_bool_input_with_and_without_the_pack_kernel(monkeypatch, pack):
    """Bool bytes -> flattened bits (a word per thread, the default) before the word-gather kernels, or the lane-per-cell
    __ballot_sync pad (MAREX_MORPH_PACK=0)."""
    track = _track()
    monkeypatch.setenv("MAREX_MORPH_PACK", pack)
    for T, ny, nx, R in ((5, 20, 45, 3), (3, 48, 96, 4), (2, 720, 1440, 8)):
        ev, mask = events_field(T, ny, nx, seed=R, density=0.05 if R > 4 else 0.12, noise=0.002)
        np.testing.assert_array_equal(track.MaskFiller(mask, R, 2).run(ev), to.stage1(ev, mask, R, 2))
