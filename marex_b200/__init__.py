"""marex_b200 -- B200-native (sm_100a) implementation of the marEx detection hot path.

Public API mirrors ``marEx/__init__.py:36-42`` for this path: ``preprocess_data``,
``compute_normalised_anomaly``, ``identify_extremes``, ``rolling_climatology``,
``smoothed_rolling_climatology`` (xarray in/out, ``marex_b200/xr_api.py``), plus the
array-level functions of ``marex_b200/detect.py`` and stage 1 of ``marEx.tracker``
(``fill_holes`` / ``fill_time_gaps``, ``marex_b200/track.py``).
"""
from .detect import (
    compute_normalised_anomaly_arrays,
    identify_extremes_arrays,
    preprocess_arrays,
    release_host_buffers,
    rolling_climatology_arrays,
)
from . import track  # noqa: F401  (tracker stage 1: fill_holes / fill_time_gaps on bit-packed masks)
from .exceptions import ConfigurationError, DataValidationError, MarExError, ProcessingError, create_data_validation_error
from .track import MaskFiller, fill_holes, fill_time_gaps
from .xr_api import (
    compute_normalised_anomaly,
    identify_extremes,
    preprocess_data,
    rolling_climatology,
    smoothed_rolling_climatology,
)

__version__ = "0.1.0"
__all__ = [
    "preprocess_data",
    "compute_normalised_anomaly",
    "identify_extremes",
    "rolling_climatology",
    "smoothed_rolling_climatology",
    "preprocess_arrays",
    "compute_normalised_anomaly_arrays",
    "identify_extremes_arrays",
    "rolling_climatology_arrays",
    "release_host_buffers",
    "MaskFiller",
    "fill_holes",
    "fill_time_gaps",
    "MarExError",
    "ConfigurationError",
    "DataValidationError",
    "ProcessingError",
    "create_data_validation_error",
]
