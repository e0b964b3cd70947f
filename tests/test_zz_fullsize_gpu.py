"""Full-size checks (BASELINE.json configs[0] = 1 degree, 40 years; configs[1] = 0.25 degree with MAREX_TEST_FULLSIZE_025=1):
at these sizes the oracle cannot redo the whole field, so the tests use what the domain offers that does not depend on size

  * a checksum of checksums: the extreme count, the bool events and the bit-packed mask agree,
  * land stays land: events only on ocean cells, thresholds NaN exactly on land,
  * the p95 frequency bars of the reference's own tests (tests/conftest.py:168-231),
  * longitude-roll equivariance, bit for bit (gridpoints are independent and the 5 x 5 pooling wraps in longitude),
  * 12 x 12 tiles cut from the full result against the oracle: anomalies within 1e-5 of the field scale, and - from the SAME
    anomalies - thresholds and events of the tile interior (which only see cells of the tile) bit for bit.

The file sorts last so that a problem here cannot hide the stage-wise parity tests under `pytest -x`."""
import os

import numpy as np
import pytest

from oracle import marex_oracle as mo

W, S, WD, WS, P = 15, 21, 11, 5, 95  # preprocess_data defaults (detect.py:287-313)


def check_tile(x_tile, time, anom_tile, thr_tile, ev_tile, mask_tile):
    """One (T, h, w) tile of the input against the matching tiles of a full-field result: ``anom_tile`` (T_out, h, w),
    ``thr_tile`` (h, w, 366), ``ev_tile`` (T_out, h, w) bool, ``mask_tile`` (h, w) bool.  Returns the number of
    threshold values compared."""
    year, doy = mo.calendar_tables(time)
    ref_anom, ref_mask, keep = mo.anomaly_shifting_baseline(x_tile, year, doy, W, S)
    np.testing.assert_array_equal(mask_tile, ref_mask)
    np.testing.assert_array_equal(np.isnan(anom_tile), np.isnan(ref_anom))
    scale = float(np.nanmax(np.abs(x_tile))) if np.isfinite(x_tile).any() else 1.0
    np.testing.assert_allclose(anom_tile, ref_anom, rtol=0, atol=1e-5 * scale, equal_nan=True)
    # thresholds and events from the SAME anomalies; the interior of the tile pools over tile cells only
    h, w = anom_tile.shape[1:]
    half = WS // 2
    doy_out = doy[keep]
    a2 = np.ascontiguousarray(anom_tile.reshape(anom_tile.shape[0], -1))
    thr_ref = mo.hobday_thresholds_approx(a2, doy_out, P / 100.0, WD, WS, (h, w)).reshape(h, w, 366)
    inner = (slice(half, h - half), slice(half, w - half))
    got, ref = thr_tile[inner], thr_ref[inner]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_array_equal(np.ascontiguousarray(got)[ok].view(np.uint32), np.ascontiguousarray(ref)[ok].view(np.uint32))
    with np.errstate(invalid="ignore"):
        ev_ref = anom_tile >= np.moveaxis(thr_ref, -1, 0)[doy_out - 1]
    np.testing.assert_array_equal(ev_tile[(slice(None),) + inner], ev_ref[(slice(None),) + inner])
    return int(ok.sum())


def test_tile_check_is_consistent_with_the_oracle_pipeline():
    """CPU self-check of `check_tile` (layouts, interior rule): a tile of an oracle result of a small field passes, and a
    corrupted threshold is caught."""
    rng = np.random.default_rng(3)
    time = np.arange(np.datetime64("1996-01-01"), np.datetime64("2014-01-01"))
    T, ny, nx = len(time), 14, 20
    frac = (time - time.astype("datetime64[Y]")).astype(float) / 365.25
    x = (12 + 4 * np.cos(2 * np.pi * frac)[:, None, None] + rng.standard_normal((T, ny, nx))).astype(np.float32)
    x[:, 4:6, 7:9] = np.nan
    res = mo.preprocess(x, time)
    thr = np.asarray(res["thresholds"]).reshape(ny, nx, 366)
    y0, x0, n = 1, 3, 12
    sl = (slice(y0, y0 + n), slice(x0, x0 + n))
    args = (x[(slice(None),) + sl], time, res["dat_anomaly"][(slice(None),) + sl], thr[sl],
            res["extreme_events"][(slice(None),) + sl].astype(bool), res["mask"][sl].astype(bool))  # fmt: skip
    assert check_tile(*args) > 0
    bad = thr[sl].copy()
    bad[5, 5, 100] = np.nextafter(bad[5, 5, 100], np.float32(10))
    with pytest.raises(AssertionError):
        check_tile(args[0], time, args[2], bad, args[4], args[5])


SIZES = [(180, 360)] + ([(720, 1440)] if os.environ.get("MAREX_TEST_FULLSIZE_025") == "1" else [])


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx", SIZES)
def test_full_size_properties(ny, nx):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import marex_b200
    from marex_b200 import synthetic

    time = synthetic.daily_time_axis("1982-01-01", "2022-01-01")
    x = synthetic.synth_sst(time, (ny, nx), seed=2)
    res = marex_b200.preprocess_arrays(x, time, output="torch", want_bits=True)
    assert res["thresholds_layout"] == "doy_last"  # (lat, lon, dayofyear), the reference's layout of the approximate path
    T_out = res["dat_anomaly"].shape[0]
    N = ny * nx
    events = res["extreme_events"].reshape(T_out, N)
    mask = res["mask"].reshape(N).bool()

    # checksum of checksums
    n_bytes = int(events.sum(dtype=torch.int64).item())
    words = res["bits"].reshape(T_out, -1).to(torch.int64) & 0xFFFFFFFF
    n_bits = 0
    for k in range(32):
        n_bits += int(((words >> k) & 1).sum().item())
    assert int(res["extreme_count"]) == n_bytes == n_bits
    # land stays land
    assert not bool(events[:, ~mask].any())
    thr = res["thresholds"].reshape(N, 366)
    land_nan = torch.isnan(thr).all(dim=1)
    assert bool((land_nan == ~mask).all())
    assert 0 < int(mask.sum()) < N
    # p95 frequency bar of the reference tests (0.03 .. 0.08 over ocean cells)
    freq = n_bytes / (T_out * int(mask.sum()))
    assert 0.03 < freq < 0.08, freq

    # tiles against the oracle
    compared = 0
    n = 12
    for y0, x0 in ((ny // 7, nx // 5), (ny // 2 - 3, nx // 2 + 11), (ny - n - 4, nx - n - 9)):
        sl = (slice(y0, y0 + n), slice(x0, x0 + n))
        tsl = (slice(None),) + sl
        compared += check_tile(
            x[tsl].cpu().numpy(), time, res["dat_anomaly"][tsl].cpu().numpy(), thr.reshape(ny, nx, 366)[sl].cpu().numpy(),
            res["extreme_events"].reshape(T_out, ny, nx)[tsl].cpu().numpy().astype(bool), mask.reshape(ny, nx)[sl].cpu().numpy(),
        )  # fmt: skip
    assert compared > 0

    if (ny, nx) == (180, 360):  # longitude-roll equivariance (needs a second copy of the field: 1-degree size only)
        shift = 37
        thr0 = thr.reshape(ny, nx, 366).clone()
        ev0 = res["extreme_events"].reshape(T_out, ny, nx).clone()
        res2 = marex_b200.preprocess_arrays(torch.roll(x, shifts=shift, dims=2), time, output="torch")
        thr2 = res2["thresholds"].reshape(ny, nx, 366)
        a, b = torch.roll(thr0, shifts=shift, dims=1), thr2
        assert bool((torch.isnan(a) == torch.isnan(b)).all())
        assert bool((a.view(torch.int32)[~torch.isnan(a)] == b.view(torch.int32)[~torch.isnan(b)]).all())
        assert bool((torch.roll(ev0, shifts=shift, dims=2) == res2["extreme_events"].reshape(T_out, ny, nx)).all())
