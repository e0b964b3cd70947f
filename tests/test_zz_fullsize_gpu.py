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

from oracle.tile_check import check_tile as _check_tile  # noqa: E402


def check_tile(x_tile, time, anom_tile, thr_tile, ev_tile, mask_tile, **kw):
    return _check_tile(x_tile, time, dict(dat_anomaly=anom_tile, thresholds=thr_tile, extreme_events=ev_tile, mask=mask_tile), **kw)


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


# 1 degree x 40 years (BASELINE configs[0]); a 48-row band of the 0.25 degree grid at full width (the 52 tile columns and
# the longitude seam of configs[1] at 1/15 of the memory); the whole 0.25 degree field with MAREX_TEST_FULLSIZE_025=1
SIZES = [(180, 360), (48, 1440)] + ([(720, 1440)] if os.environ.get("MAREX_TEST_FULLSIZE_025") == "1" else [])


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

    # tiles against the oracle; the last one straddles the longitude seam (its interior pools across the wrap)
    compared = 0
    n = 12
    thr3 = thr.reshape(ny, nx, 366)
    ev3 = res["extreme_events"].reshape(T_out, ny, nx)
    for y0, x0 in ((ny // 7, nx // 5), (ny // 2 - 3, nx // 2 + 11), (ny - n - 4, nx - n - 9), (ny // 3, nx - n // 2)):
        rows = slice(y0, y0 + n)
        cols = torch.arange(x0, x0 + n, device=x.device) % nx
        cut = lambda a, lead: a[(slice(None),) * lead + (rows,)].index_select(lead + 1, cols).cpu().numpy()  # noqa: E731
        compared += check_tile(cut(x, 1), time, cut(res["dat_anomaly"], 1), cut(thr3, 0), cut(ev3, 1).astype(bool),
                               cut(mask.reshape(ny, nx), 0))  # fmt: skip
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
