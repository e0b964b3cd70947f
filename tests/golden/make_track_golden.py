"""
Generate tests/golden/track_stage1.npz (tracker stage 1).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_track_golden.py

`marEx/track.py` cannot be imported here (dask, dask_image, xarray, numba, skimage are missing), but the arithmetic of
fill_holes lives in nested / module-level functions that only need numpy and scipy.  They are AST-extracted from the
read-only source and executed unchanged (nothing is copied into the repo; only INPUTS and OUTPUTS are stored):

  gridded       `binary_open_close` nested in tracker.fill_holes, the non-dask branch     track.py:1646-1660
                (np.pad -> scipy.ndimage.binary_closing -> binary_opening -> unpad), applied per time step as its
                apply_ufunc(vectorize=True) does, followed by `data_bin.where(self.mask, other=False)` track.py:1667
  unstructured  `binary_open_close` nested in the unstructured branch                     track.py:1549-1582
                with the module-level `sparse_bool_power`                                 track.py:5423-5470
                (its numba decorator and prange are replaced by plain Python / range)

The temporal closing of fill_time_gaps (track.py:1695-1719) is `dask_image.ndmorph.binary_closing` on the False-padded
time axis; the golden uses scipy.ndimage.binary_closing with the same structure (what dask_image applies per chunk).
"""
import ast
import os
import sys
import textwrap

import numpy as np
from numpy.typing import NDArray
from scipy.ndimage import binary_closing, binary_opening

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/marEx/track.py"
SRC = open(REF).read()
TREE = ast.parse(SRC)


def _nested(defs_named: str, must_contain: str):
    """Source of the nested FunctionDef `defs_named` whose body mentions `must_contain`."""
    hits = [n for n in ast.walk(TREE) if isinstance(n, ast.FunctionDef) and n.name == defs_named]
    hits = [n for n in hits if must_contain in ast.get_source_segment(SRC, n)]
    assert len(hits) == 1, (defs_named, must_contain, len(hits))
    return hits[0]


def _compile(node, ns, strip_decorators=True):
    if strip_decorators:
        node.decorator_list = []
    code = textwrap.dedent(ast.get_source_segment(SRC, node))
    # get_source_segment starts at `def`, decorators are not part of it
    exec(compile(code, f"<reference:{node.name}:{node.lineno}>", "exec"), ns)
    return ns[node.name]


def reference_sparse_bool_power():
    (node,) = [n for n in TREE.body if isinstance(n, ast.FunctionDef) and n.name == "sparse_bool_power"]
    return _compile(node, {"np": np, "NDArray": NDArray, "prange": range})


def reference_fill_holes_gridded(events, mask, R_fill, regional_mode):
    node = _nested("binary_open_close", "binary_closing(bitmap_binary_padded")
    y, x = np.ogrid[-R_fill : R_fill + 1, -R_fill : R_fill + 1]  # closure variables, as built at track.py:1613-1617
    ns = {"np": np, "NDArray": NDArray, "binary_closing": binary_closing, "binary_opening": binary_opening,
          "diameter": 2 * R_fill, "se_kernel": (x**2 + y**2) < (R_fill**2) + 1, "mode": "wrap" if not regional_mode else "edge"}  # fmt: skip
    f = _compile(node, ns)
    out = np.stack([f(events[t]) for t in range(events.shape[0])]) if R_fill > 0 else events.copy()
    return np.where(mask[None], out, False)


def reference_fill_holes_unstructured(events, mask, neighbours, R_fill):
    from scipy.sparse import coo_matrix, csr_matrix, eye

    # the matrix of tracker._build_sparse_dilation_matrix (track.py:1093-1115), with numpy in place of jnp
    ncells = neighbours.shape[1]
    row = np.repeat(np.arange(ncells), neighbours.shape[0])
    col = neighbours.T.flatten()
    ok = col >= 0
    m = csr_matrix(coo_matrix((np.ones(ok.sum(), dtype=bool), (row[ok], col[ok])), shape=(ncells, ncells)))
    m = m + eye(ncells, dtype=bool, format="csr")
    node = _nested("binary_open_close", "sparse_bool_power(bitmap_binary")
    f = _compile(node, {"np": np, "NDArray": NDArray, "sparse_bool_power": reference_sparse_bool_power(), "R_fill": R_fill})
    return f(events.copy(), m.data, m.indices, m.indptr, mask)


def time_closing(events, T_fill):
    k = T_fill + 1
    pad = [(k, k)] + [(0, 0)] * (events.ndim - 1)
    se = np.ones((k,) + (1,) * (events.ndim - 1), dtype=bool)
    return binary_closing(np.pad(events, pad, mode="constant", constant_values=False), structure=se)[k:-k]


def stage1_gridded(ev, mask, R, T_fill, regional=False):
    h = reference_fill_holes_gridded(ev, mask, R, regional)
    return reference_fill_holes_gridded(time_closing(h, T_fill), mask, R // 2, regional)


def stage1_unstructured(ev, mask, nb, R, T_fill):
    h = reference_fill_holes_unstructured(ev, mask, nb, R)
    return reference_fill_holes_unstructured(time_closing(h, T_fill), mask, nb, R // 2)


def main():
    from test_track_cpu import events_field, mesh

    ev, mask = events_field(T=8, ny=30, nx=50, seed=11)
    nb = mesh(8, seed=5)
    rng = np.random.default_rng(17)
    mask_u = rng.random(nb.shape[1]) > 0.15
    ev_u = (rng.random((40, nb.shape[1])) < 0.3) & mask_u
    out = dict(
        events=ev, mask=mask, events_u=ev_u, mask_u=mask_u, neighbours=nb,
        fill_holes_R3=reference_fill_holes_gridded(ev, mask, 3, False),
        stage1_R4_T2=stage1_gridded(ev, mask, 4, 2),
        stage1_R3_T4_regional=stage1_gridded(ev, mask, 3, 4, True),
        stage1_u_R2_T2=stage1_unstructured(ev_u, mask_u, nb, 2, 2),
    )  # fmt: skip
    path = os.path.join(HERE, "track_stage1.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, int(v.sum())) for k, v in out.items() if v.dtype == bool})


if __name__ == "__main__":
    main()
