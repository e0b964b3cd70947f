"""ctypes binding of libmarex_b200.so (include/marex_b200.h).  There is NO fallback: if the
library is missing or a call fails, the caller gets an exception."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_longlong, c_uint64, c_void_p

from .exceptions import ProcessingError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmarex_b200.so")

_P = c_void_p
_SIGNATURES = {
    "marex_version": ([], ctypes.c_int),
    "marex_last_error": ([], c_char_p),
    "marex_launch_count": ([], c_longlong),
    "marex_tune": ([c_char_p, c_longlong, c_int32], ctypes.c_int),
    "marex_shift_anomaly_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, c_int32, c_int32, c_int32, _P, c_int32, _P, c_int64, _P, _P, _P], ctypes.c_int),
    "marex_shift_anomaly_daily_f32": ([_P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, _P, c_int64, _P, _P, _P, c_int32, _P, c_int64, _P], ctypes.c_int),
    "marex_shift_anomaly_fixup_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, c_int32, c_int32, c_int32, _P, c_int32, _P, c_int64, _P, _P, _P, _P, c_int32, _P, c_int64, c_int64, c_int32, _P], ctypes.c_int),
    "marex_doy_climatology_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P], ctypes.c_int),
    "marex_sub_doy_climatology_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P, _P, _P], ctypes.c_int),
    "marex_detrend_coef_f64": ([_P, c_int64, c_int64, c_int64, _P, c_int32, _P, _P, _P, _P], ctypes.c_int),
    "marex_detrend_apply_f32": ([_P, c_int64, c_int64, c_int64, _P, c_int32, _P, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_doy_std_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, _P, _P], ctypes.c_int),
    "marex_doy_rolling_rms_f32": ([_P, c_int64, c_int32, _P, _P], ctypes.c_int),
    "marex_div_doy_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, _P, c_int64, _P], ctypes.c_int),
    "marex_digitize_doy_f32": ([_P, c_int64, c_int64, _P, c_int64, _P, c_int32, _P, c_int64, _P], ctypes.c_int),
    "marex_hobday_thresholds_hist": ([_P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_int32, _P, c_int32, c_int32, c_int32, c_double, _P, c_float, _P, _P, _P], ctypes.c_int),
    "marex_hobday_pooled_workspace_bytes": ([c_int64, c_int64], c_int64),
    "marex_hobday_thresholds_pooled_bins": ([_P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_int32, c_int32, c_int32, c_double, _P, c_float, _P, _P, _P, c_int64, _P], ctypes.c_int),
    "marex_hobday_thresholds_exact_f32": ([_P, c_int64, c_int64, c_int64, _P, _P, c_int32, c_int32, c_int32, c_float, _P, _P, _P], ctypes.c_int),
    "marex_global_threshold_hist_f64": ([_P, c_int64, c_int64, c_int64, _P, _P, c_int32, c_double, c_double, _P, _P, _P], ctypes.c_int),
    "marex_global_threshold_hist_fast_f64": ([_P, c_int64, c_int64, c_int64, _P, _P, c_float, _P, c_int32, c_double, c_double, _P, _P, _P, _P], ctypes.c_int),
    "marex_global_threshold_exact_f64": ([_P, c_int64, c_int64, c_int64, c_double, _P, _P], ctypes.c_int),
    "marex_compare_hobday": ([_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_compare_hobday_bins": ([_P, c_int64, c_int64, _P, _P, c_int64, c_int64, _P, c_int64, _P, c_int32, _P, c_int64, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_memcpy2d_async": ([_P, c_int64, _P, c_int64, c_int64, c_int64, c_int32, _P], ctypes.c_int),
    "marex_compare_global": ([_P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_morph_slab_words": ([c_int64, c_int64, c_int32], c_int64),
    "marex_morph_pad_bits": ([_P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, c_int64, c_int32, c_int32, _P, _P], ctypes.c_int),
    "marex_morph_disk": ([_P, _P, c_int64, c_int64, c_int64, c_int32, c_int32, _P], ctypes.c_int),
    "marex_morph_disk_levels": ([c_int32], c_int32),
    "marex_morph_disk_sep": ([_P, _P, c_int64, c_int64, c_int64, c_int32, c_int32, _P, c_int64, _P], ctypes.c_int),
    "marex_morph_time": ([_P, c_int64, c_int64, _P, c_int64, c_int32, c_int32, c_int32, _P], ctypes.c_int),
    "marex_morph_extract": ([_P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_morph_pack_u8": ([_P, c_int64, c_int64, c_int64, _P, c_int64, _P], ctypes.c_int),
    "marex_morph_tpack_words": ([c_int64], c_int64),
    "marex_morph_tpack": ([_P, _P, c_int64, c_int64, c_int64, _P, _P], ctypes.c_int),
    "marex_morph_nbr": ([_P, _P, c_int64, c_int64, _P, c_int32, _P, c_int32, c_int32, _P], ctypes.c_int),
    "marex_morph_tshift": ([_P, _P, c_int64, c_int64, c_int32, c_int32, c_int32, _P], ctypes.c_int),
    "marex_morph_tunpack": ([_P, c_int64, c_int64, _P, _P, c_int64, _P, c_int64, _P, _P], ctypes.c_int),
    "marex_transpose_f32": ([_P, c_int64, c_int64, _P, _P], ctypes.c_int),
    "marex_synth_sst_f32": ([_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, c_uint64, c_float, _P], ctypes.c_int),
}
EXPORTED = tuple(_SIGNATURES)

ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3  # MAREX_ERR_* of include/marex_b200.h

_lib = None
TRACE = None  # optional callable(name) invoked after every kernel-launching call (bench.py stage timing)


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ProcessingError(
                "libmarex_b200.so is not built",
                details=f"expected {LIB_PATH}",
                suggestions=["run `python -c 'import __graft_entry__ as g; g.build()'` or `python marex_b200/_build.py`"],
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise ProcessingError on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.marex_last_error().decode("utf-8", "replace")
        raise ProcessingError(f"{name} failed (code {rc})", details=msg, context={"code": int(rc)})
    if TRACE is not None:
        TRACE(name)


def tune(**knobs) -> None:
    """Pin (value) or release (None) tuning / test knobs of the library, e.g. ``tune(shift_v=4, pool_k=None)``."""
    lib = load()
    for key, value in knobs.items():
        lib.marex_tune(key.encode(), 0 if value is None else int(value), 0 if value is None else 1)


def launch_count() -> int:
    return int(load().marex_launch_count())
