"""TEST INFRASTRUCTURE ONLY -- CPU oracle of tracker stage 1 (marEx/track.py fill_holes :1520-1669 and
fill_time_gaps :1671-1726).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

The reference delegates the arithmetic to ``dask_image.ndmorph.binary_closing / binary_opening`` (track.py:38-39,
1630-1634, 1710; dask_image is unpinned and absent from this image), which are documented as, and implemented by,
``scipy.ndimage.binary_dilation`` / ``binary_erosion`` applied chunk-wise with an overlap of the structure's
half-size: closing = erosion(dilation(x)), opening = dilation(erosion(x)), ``border_value=0``, ``origin=0``.  The
reference itself states the equivalence in its non-dask branch (track.py:1646-1660: ``scipy.ndimage.binary_closing``
then ``binary_opening`` on the same padded slice).  scipy IS installed here, so this oracle calls scipy directly:
parity pinned against the reference's own dependency (tests/test_track_cpu.py also checks the composed
erosion(dilation) form against scipy's binary_closing / binary_opening).

The unstructured branch (track.py:1543-1607, sparse_bool_power :5423-5470) is restated with plain numpy loops.
"""
import numpy as np
from scipy import ndimage


def disk(R_fill: int) -> np.ndarray:
    """track.py:1613-1616."""
    y, x = np.ogrid[-R_fill : R_fill + 1, -R_fill : R_fill + 1]
    return (x**2 + y**2) < (R_fill**2) + 1


def _closing(a, structure):
    return ndimage.binary_erosion(ndimage.binary_dilation(a, structure=structure), structure=structure)


def _opening(a, structure):
    return ndimage.binary_dilation(ndimage.binary_erosion(a, structure=structure), structure=structure)


def fill_holes(data_bin: np.ndarray, mask: np.ndarray, R_fill: int, regional_mode: bool = False) -> np.ndarray:
    """Gridded branch, track.py:1609-1667.  data_bin (T, ny, nx) bool, mask (ny, nx) bool."""
    data_bin = np.asarray(data_bin, dtype=bool)
    if R_fill > 0:
        se = disk(R_fill)[np.newaxis, :, :]
        d = 2 * R_fill
        padded = np.pad(data_bin, ((0, 0), (d, d), (d, d)), mode="edge" if regional_mode else "wrap")
        padded = _opening(_closing(padded, se), se)
        data_bin = padded[:, d:-d, d:-d]
    return np.where(mask[np.newaxis], data_bin, False)


def fill_time_gaps(data_bin: np.ndarray, mask: np.ndarray, R_fill: int, T_fill: int, regional_mode: bool = False) -> np.ndarray:
    """track.py:1671-1726 (gridded): temporal closing with T_fill + 1 ones on a False-padded axis, then
    fill_holes with R_fill // 2."""
    if T_fill == 0:
        return np.asarray(data_bin, dtype=bool)
    k = T_fill + 1
    padded = np.pad(np.asarray(data_bin, dtype=bool), ((k, k), (0, 0), (0, 0)), mode="constant", constant_values=False)
    closed = _closing(padded, np.ones((k, 1, 1), dtype=bool))[k:-k]
    return fill_holes(closed, mask, R_fill // 2, regional_mode)


def stage1(data_bin, mask, R_fill, T_fill, regional_mode=False):
    """tracker.run_preprocess order, track.py:1288-1297."""
    return fill_time_gaps(fill_holes(data_bin, mask, R_fill, regional_mode), mask, R_fill, T_fill, regional_mode)


# ------------------------------------------------------------------------------------------ unstructured
def sparse_dilate(vec: np.ndarray, neighbours: np.ndarray, exponent: int) -> np.ndarray:
    """(neighbours + identity)^exponent applied to vec (T, ncells) bool; neighbours (nv, ncells) int, negative = none.
    track.py:1093-1115 (matrix), :5423-5470 (power)."""
    res = np.asarray(vec, dtype=bool).copy()
    for _ in range(exponent):
        nxt = res.copy()
        for j in range(neighbours.shape[0]):
            col = neighbours[j]
            ok = col >= 0
            nxt[:, ok] |= res[:, col[ok]]
        res = nxt
    return res


def fill_holes_unstructured(data_bin: np.ndarray, mask: np.ndarray, neighbours: np.ndarray, R_fill: int) -> np.ndarray:
    """track.py:1549-1582 (note: no final masking in this branch; land cells are SET, not cleared)."""
    b = sparse_dilate(data_bin, neighbours, R_fill)
    b[:, ~mask] = True
    b = ~sparse_dilate(~b, neighbours, R_fill)
    b[:, ~mask] = True
    b = ~sparse_dilate(~b, neighbours, R_fill)
    return sparse_dilate(b, neighbours, R_fill)


def fill_time_gaps_unstructured(data_bin, mask, neighbours, R_fill, T_fill):
    if T_fill == 0:
        return np.asarray(data_bin, dtype=bool)
    k = T_fill + 1
    padded = np.pad(np.asarray(data_bin, dtype=bool), ((k, k), (0, 0)), mode="constant", constant_values=False)
    closed = _closing(padded, np.ones((k, 1), dtype=bool))[k:-k]
    return fill_holes_unstructured(closed, mask, neighbours, R_fill // 2)


def stage1_unstructured(data_bin, mask, neighbours, R_fill, T_fill):
    return fill_time_gaps_unstructured(fill_holes_unstructured(data_bin, mask, neighbours, R_fill), mask, neighbours, R_fill, T_fill)
