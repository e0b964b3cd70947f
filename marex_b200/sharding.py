"""
Multi-GPU: one process per GPU, gridpoints sharded spatially, no collective in the data path.

Gridpoints are independent in every stage except the ws x ws histogram pooling (a radius
ws//2 stencil), so a gridded field is cut into LATITUDE BANDS (full longitude rows keep the
periodic wrap local) and each rank loads ``halo = ws//2`` extra rows on its interior edges and
recomputes them instead of exchanging histograms.  ``torch.distributed`` is used
only to gather the small per-shard outputs (thresholds, mask, extreme counts).
"""
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np


def lat_band(ny: int, world: int, rank: int, halo: int) -> Tuple[int, int, int, int]:
    """Rows owned by ``rank`` and rows it must load: (own_lo, own_hi, load_lo, load_hi).
    Bands differ by at most one row; the halo is truncated at the true grid edge, exactly where
    the reference truncates its pooling window (detect.py:2666)."""
    base, rem = divmod(ny, world)
    own_lo = rank * base + min(rank, rem)
    own_hi = own_lo + base + (1 if rank < rem else 0)
    return own_lo, own_hi, max(0, own_lo - halo), min(ny, own_hi + halo)


def cell_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous ncells range of an unstructured mesh (no halo: no pooling on unstructured data)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def effective_halo(method_extreme: str, method_percentile: str, window_spatial_hobday: Optional[int], gridded: bool) -> int:
    """Rows of overlap the configuration needs (default 5x5 pooling on gridded data, detect.py:1451)."""
    if not gridded or method_extreme != "hobday_extreme" or method_percentile != "approximate":
        return 0
    ws = 5 if window_spatial_hobday is None else int(window_spatial_hobday)
    return ws // 2


def crop_owned(res: Dict[str, Any], own: Tuple[int, int], load: Tuple[int, int]) -> Dict[str, Any]:
    """Drop the halo rows from a per-shard result of ``preprocess_arrays`` (gridded)."""
    a, b = own[0] - load[0], own[1] - load[0]
    out = dict(res)
    out["dat_anomaly"] = res["dat_anomaly"][:, a:b]
    out["mask"] = res["mask"][a:b]
    if "extreme_events" in res:
        out["extreme_events"] = res["extreme_events"][:, a:b]
    lay = res.get("thresholds_layout", "doy_last")
    thr = res["thresholds"]
    out["thresholds"] = thr[:, a:b] if lay == "doy_first" else thr[a:b]
    return out


def preprocess_sharded(
    load_rows: Callable[[int, int], Any],
    time,
    ny: int,
    rank: int,
    world: int,
    compute: Callable[..., Dict[str, Any]],
    gather: Optional[Callable[[Any, int], Any]] = None,
    **kwargs,
) -> Dict[str, Any]:
    """Run the pipeline on this rank's latitude band.

    ``load_rows(lo, hi)`` returns the (T, hi-lo, nx) slab (host or device); ``compute`` is
    ``marex_b200.preprocess_arrays`` (tests substitute the numpy oracle to exercise the
    partition/halo/gather logic on CPU with gloo); ``gather(x, axis)`` concatenates a per-rank
    array over ranks along ``axis`` (see ``dist_gather``) or is None to keep results sharded."""
    halo = effective_halo(
        kwargs.get("method_extreme", "hobday_extreme"),
        kwargs.get("method_percentile", "approximate"),
        kwargs.get("window_spatial_hobday"),
        True,
    )
    own_lo, own_hi, load_lo, load_hi = lat_band(ny, world, rank, halo)
    res = compute(load_rows(load_lo, load_hi), time, **kwargs)
    out = crop_owned(res, (own_lo, own_hi), (load_lo, load_hi))
    out["rows"] = (own_lo, own_hi)
    if gather is not None:
        lay = out.get("thresholds_layout", "doy_last")
        out["thresholds_global"] = gather(out["thresholds"], 1 if lay == "doy_first" else 0)
        out["mask_global"] = gather(out["mask"], 0)
    return out


def dist_gather(x, axis: int, equal: bool = False, async_op: bool = False):
    """All-gather a per-rank array with unequal extents along ``axis`` through torch.distributed
    (NCCL for CUDA tensors over NVLink; gloo for CPU tensors in the tests).  ``equal=True`` promises
    equal extents on all ranks and skips the size exchange (a host round trip); with ``async_op``
    it returns ``(work, result)`` so the collective overlaps whatever is launched next."""
    import torch
    import torch.distributed as dist

    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    was_bool = t.dtype == torch.bool
    if was_bool:
        t = t.to(torch.uint8)
    t = t.movedim(axis, 0).contiguous()
    world = dist.get_world_size()
    if equal:
        full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        work = dist.all_gather_into_tensor(full, t, async_op=async_op)
        res = full.movedim(0, axis) if axis != 0 else full
        if async_op:
            return work, (res, t)  # the caller keeps `t` alive until the work has completed
        if was_bool:
            res = res.to(torch.bool)
        return res if isinstance(x, torch.Tensor) else res.numpy()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    if min(sizes) == mx:  # equal shards (the usual case): one collective straight into the result, no padding / concat
        full = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t)
        full = full.movedim(0, axis) if axis != 0 else full
        if was_bool:
            full = full.to(torch.bool)
        return full if isinstance(x, torch.Tensor) else full.numpy()
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0).movedim(0, axis)
    if was_bool:
        full = full.to(torch.bool)
    return full if isinstance(x, torch.Tensor) else full.numpy()
