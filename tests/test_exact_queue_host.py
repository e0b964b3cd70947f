"""The lane algorithm of the exact day-of-year percentile kernel (marex_b200/csrc/exact_queue.cuh: per-gridpoint
first-in first-out queue of the samples above a pivot) compiled for the host with a one-lane environment
(tests/host/exact_queue_host.cpp) and checked against np.nanpercentile, the call the reference makes
(detect.py:1936-1942).  Covers what the heuristics have to survive: ties at and around the threshold, NaN gaps,
infinities, constant series, very short windows, leap-day-only days of year, pivot drift over the year."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.marex_oracle as mo  # noqa: E402
from marex_b200.calendar import doy_csr  # noqa: E402


@pytest.fixture(scope="module")
def xq(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("xq") / "xq_host.so")
    subprocess.check_call(
        ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "marex_b200", "csrc"),
         os.path.join(ROOT, "tests", "host", "exact_queue_host.cpp"), "-o", so]
    )  # fmt: skip
    lib = ctypes.CDLL(so)
    lib.xq_host.restype = ctypes.c_int
    lib.xq_kk_max_host.restype = ctypes.c_int
    lib.xq_kk_max_host.argtypes = [ctypes.c_int, ctypes.c_float]
    return lib


def _run(lib, a2, doy, w, p, Q):
    T, N = a2.shape
    ptr, rows = doy_csr(doy)
    a2 = np.ascontiguousarray(a2, dtype=np.float32)
    thr = np.full((366, N), -777.0, dtype=np.float32)
    failed = np.zeros(N, dtype=np.int32)
    stats = np.zeros((N, 366, 6), dtype=np.int16)  # per gridpoint and day: raises, lowerings, rebuilds, bracketing passes, selection passes
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    nfail = lib.xq_host(vp(a2), ctypes.c_int64(N), ctypes.c_int64(N), vp(ptr), vp(rows), ctypes.c_int(w), ctypes.c_float(p),
                        ctypes.c_int(Q), vp(thr), vp(failed), vp(stats))  # fmt: skip
    assert nfail >= 0
    return thr, failed.astype(bool), stats


def _field(seed, years, N, kind):
    rng = np.random.default_rng(seed)
    time = np.arange(np.datetime64("1990-01-01"), np.datetime64(f"{1990 + years}-01-01"))
    T = len(time)
    a = (rng.standard_normal((T, N)) * rng.uniform(0.2, 2.0, N)).astype(np.float32)
    if kind == "seasonal":  # thresholds that move over the year: the pivot has to follow
        a += (1.5 * np.sin(2 * np.pi * np.arange(T) / 365.25))[:, None].astype(np.float32)
    if kind == "special":
        a[:, 0] = np.nan
        a[:, 1] = 0.0
        a[:, 2] = np.round(a[:, 2], 1)  # heavy ties everywhere
        a[:, 3] = np.where(a[:, 3] > 0.3, np.float32(0.3), a[:, 3])  # the upper tail is one block of equal values
        a[:, 4] = np.where(a[:, 4] < 0.0, np.float32(0.0), a[:, 4])  # sea-ice like: half the samples are exactly 0
        a[::7, 5] = np.nan
        a[100:900, 6] = np.nan  # a long gap: windows with few or no samples
        a[10, 7] = np.inf
        a[20, 7] = -np.inf
        a[::5, 8] = np.inf
        a[::2, 9] = -np.inf
        a[:, 10] = np.float32(-3.25)
        a[:, 11] = np.where(rng.uniform(size=T) < 0.9, np.nan, a[:, 11])  # sparse
        a[:, 12] = (a[:, 12] * 1e30).astype(np.float32)
        a[:, 13] = (a[:, 13] * 1e-30).astype(np.float32)
        a[:, 14] = np.round(a[:, 14] * 2) / 2  # ties in blocks of ~50 samples
        a[:, 15] = np.where(a[:, 15] < 0.8, np.float32(-1.5), a[:, 15])  # most samples are one block at the minimum
    _, doy = mo.calendar_tables(time)
    return a, doy


def _same(got, ref):
    assert got.shape == ref.shape
    both_nan = np.isnan(got) & np.isnan(ref)
    eq = (got == ref) | both_nan
    assert eq.all(), (np.argwhere(~eq)[:5], got[~eq][:5], ref[~eq][:5])


@pytest.mark.parametrize("kind", ["plain", "seasonal", "special"])
@pytest.mark.parametrize("w,p,years,Q", [(11, 95.0, 25, 64), (11, 90.0, 25, 64), (5, 99.5, 13, 64), (1, 50.0, 13, 64), (11, 50.0, 9, 128),
                                         (31, 99.0, 25, 64), (3, 0.0, 9, 64), (11, 100.0, 25, 64), (11, 95.0, 25, 16)])  # fmt: skip
def test_queue_algorithm_matches_nanpercentile(xq, kind, w, p, years, Q):
    a, doy = _field(3, years, 24, kind)
    ref = mo.hobday_thresholds_exact(a, doy, p, w)
    thr, failed, stats = _run(xq, a, doy, w, p, Q)
    ok = ~failed
    _same(thr[:, ok], ref[:, ok])
    # the dispatch rule of the host wrapper: queues this size are only used when kk + ROOM fits
    rows = int(np.bincount(doy - 1, minlength=366).max()) * w
    fits = xq.xq_kk_max_host(rows, p) + 26 <= Q - 4
    if fits and kind != "special":
        assert not failed.any()
    if fits and kind == "special":
        assert failed.sum() <= 2, np.nonzero(failed)[0]  # columns that are mostly infinities may give up


def _warp_cost(ev):
    """What a warp of 32 adjacent gridpoints executes per day of year: a raise / lowering / rebuild runs when ANY lane
    needs it, the selection runs as many passes as the slowest lane."""
    N = ev.shape[0] // 32 * 32
    g = ev[:N].reshape(-1, 32, 366, 6)
    return dict(raises=(g[..., 0].max(1) > 0).mean(), lowerings=(g[..., 1].max(1) > 0).mean(), rebuilds=(g[:, :, 1:, 2].max(1) > 0).mean(),
                bracket_passes=g[:, :, 1:, 3].max(1).mean(), select_passes=g[..., 4].max(1).mean(), select_passes_lane=g[..., 4].mean(), select_samples=g[..., 5].max(1).mean())  # fmt: skip


def test_queue_work_per_day_is_small(xq):
    """The point of the design: per warp and day about two selection passes, and a window re-read every few days."""
    a, doy = _field(5, 25, 64, "plain")
    thr, failed, ev = _run(xq, a, doy, 11, 95.0, 64)
    assert not failed.any()
    c = _warp_cost(ev)
    print(c)
    assert c["select_passes"] < 4.0
    assert c["lowerings"] + c["rebuilds"] < 0.5

