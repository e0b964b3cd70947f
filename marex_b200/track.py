"""Tracker stage 1 on the GPU: ``fill_holes`` and ``fill_time_gaps`` of ``marEx.tracker``
(marEx/track.py:1520-1669, 1671-1726), the first consumers of the ``extreme_events`` / ``mask`` pair that
``preprocess_data`` returns (order of use: track.py:1288-1297).

The reference runs ``dask_image.ndmorph.binary_closing`` / ``binary_opening`` (scipy.ndimage under dask) on
bool arrays padded by ``2 * R_fill`` cells; here every time step is a BIT-PACKED slab (32 cells per word) and the
disk dilation / erosion work on whole words (``marex_b200/csrc/morph.cu``).  The results are the reference's
bit for bit, including what its padding does near the borders (the opening's erosion reads cells that the
closing's zero border has already cleared; reproduced by doing literally the same four passes on the same
padded array).  Unstructured meshes (track.py:1543-1607) keep 32 TIME steps of a cell in a word, so that one
neighbour gather of the sparse "neighbours + identity" dilation serves 32 days.

There is no CPU fallback: every step is a call into libmarex_b200.so on CUDA buffers.
"""
import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .exceptions import ConfigurationError, DataValidationError

MAX_R_FILL = 32  # the word-level disk kernel looks one word to either side
MAX_T_FILL = 32
SEPARABLE_SCRATCH_BYTES = 48 << 20  # level buffers of one chunk of time steps: well inside the 126 MB L2


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("marex_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _on_self_device(fn):
    """Run a MaskFiller method with the filler's CUDA device made current (launches, streams and scratch allocations
    must agree with the device the mask lives on, whatever the caller's current device is)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if self.device.type != "cuda":
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapper


def _stream(dev: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else None


def _call(name: str, *args) -> None:
    _lib.call(name, *args)


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Source:
    """Cells of every time step as the C-ABI describes a source (include/marex_b200.h, tracker stage 1)."""

    def __init__(self, tensor: torch.Tensor, is_bits: bool, t_pitch: int, row_stride: int, origin: int):
        self.tensor, self.is_bits, self.t_pitch, self.row_stride, self.origin = tensor, is_bits, t_pitch, row_stride, origin

    def args(self, mask_bits: Optional[torch.Tensor]):
        return (None if self.is_bits else _p(self.tensor), _p(self.tensor) if self.is_bits else None, self.t_pitch,
                self.row_stride, self.origin, _p(mask_bits))  # fmt: skip


class MaskFiller:
    """Stage 1 of ``marEx.tracker`` with the tracker's own parameters (track.py:323-347, 396-402):

    mask           ocean mask, (ny, nx) bool for a gridded field or (ncells,) for an unstructured mesh
    R_fill         radius of the disk (gridded, cells) / number of neighbour hops (unstructured); ``int(R_fill)``
    T_fill         largest temporal gap that is closed; must be even (track.py:704-709)
    regional_mode  pad with the edge value instead of wrapping periodically (track.py:1617)
    neighbours     unstructured only: (nv, ncells) int32, 0-based, negative = no neighbour (track.py:1095-1101)
    separable      gridded only: run the disk passes in separable form (same bits, fewer instructions); default from
                   the environment variable MAREX_MORPH_SEPARABLE

    ``fill_holes`` / ``fill_time_gaps`` take and return (time, ...space) bool arrays: a numpy array in gives a numpy
    array out, a torch tensor (any device) gives a CUDA bool tensor.  ``packed=True`` returns the flattened bit mask
    instead (int32 [T, ceil(N / 32)], bit ``c & 31`` of word ``c >> 5``); ``from_bits=(bits, T)`` accepts that layout
    as the input (what ``marex_compare_*`` writes).  ``last_count`` holds the number of True cells of the last result.
    """

    def __init__(self, mask, R_fill, T_fill: int = 2, regional_mode: bool = False, neighbours=None, device=None,
                 separable: Optional[bool] = None):
        self.separable = (os.environ.get("MAREX_MORPH_SEPARABLE", "0") == "1") if separable is None else bool(separable)
        self.pack_input = os.environ.get("MAREX_MORPH_PACK", "1") == "1"  # measured 7.3 vs 9.3 ms per 1,024 days
        self.R_fill = int(R_fill)
        self.T_fill = T_fill
        self.regional_mode = bool(regional_mode)
        if self.T_fill % 2 != 0:
            raise ConfigurationError(
                "T_fill must be even for temporal symmetry",
                details=f"Provided T_fill={self.T_fill} is odd",
                suggestions=["Use even values: 2, 4, 6, 8, etc."],
                context={"provided_value": self.T_fill, "requirement": "even number"},
            )
        if not 0 <= self.R_fill <= MAX_R_FILL or not 0 <= self.T_fill <= MAX_T_FILL:
            raise ConfigurationError(
                "R_fill / T_fill outside the range of the bit-packed kernels",
                details=f"R_fill={self.R_fill} (0..{MAX_R_FILL}), T_fill={self.T_fill} (0..{MAX_T_FILL})",
            )
        self.device = _device(device)
        mask = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask).astype(bool)
        self.unstructured = neighbours is not None
        if self.unstructured:
            if self.regional_mode:
                raise NotImplementedError("regional_mode is not yet implemented for unstructured grids")  # track.py:501-502
            if mask.ndim != 1:
                raise DataValidationError("unstructured mask must be 1-D (ncells,)", details=f"got shape {mask.shape}")
            nb = np.ascontiguousarray(np.asarray(neighbours.cpu() if isinstance(neighbours, torch.Tensor) else neighbours), dtype=np.int32)
            if nb.ndim != 2 or nb.shape[1] != mask.shape[0]:
                raise DataValidationError("neighbours must be (nv, ncells)", details=f"got {nb.shape} for {mask.shape[0]} cells")
            if nb.max(initial=-1) >= mask.shape[0]:
                raise DataValidationError("neighbour index out of range", details=f"max {int(nb.max())} for {mask.shape[0]} cells")
            self.space: Tuple[int, ...] = (mask.shape[0],)
            self.nv = int(nb.shape[0])
            self.neighbours = torch.from_numpy(nb).to(self.device)
        else:
            if mask.ndim != 2:
                raise DataValidationError("gridded mask must be 2-D (lat, lon)", details=f"got shape {mask.shape}")
            self.space = (int(mask.shape[0]), int(mask.shape[1]))
        self.N = int(np.prod(self.space))
        self.mask = torch.from_numpy(np.ascontiguousarray(mask.reshape(-1)).view(np.uint8)).to(self.device)
        flat = np.zeros(((self.N + 31) // 32) * 32, dtype=bool)
        flat[: self.N] = mask.reshape(-1)
        # the ocean mask as flattened bits (bit c & 31 of word c >> 5): what the gridded kernels apply a word at a time
        self.mask_bits = torch.from_numpy(np.packbits(flat, bitorder="little").view(np.int32).copy()).to(self.device)
        self.last_count: Optional[int] = None

    # ------------------------------------------------------------------ input / output plumbing
    def _input(self, data_bin, from_bits) -> Tuple[_Source, int, bool]:
        row_stride = self.space[-1]
        if from_bits is not None:
            bits, T = from_bits
            bits = bits.to(self.device)
            if (bits.dtype not in (torch.int32, torch.uint32) or bits.dim() != 2 or bits.shape[1] * 32 < self.N
                    or not bits.is_contiguous() or not 0 < int(T) <= bits.shape[0]):  # fmt: skip
                raise DataValidationError("from_bits expects a contiguous int32 [T, >= ceil(N / 32)] tensor", details=f"got {tuple(bits.shape)} {bits.dtype} for T={T}")
            return _Source(bits, True, int(bits.shape[1]), row_stride, 0), int(T), False
        as_numpy = not isinstance(data_bin, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(data_bin)) if as_numpy else data_bin
        if tuple(t.shape[1:]) != self.space:
            raise DataValidationError("data_bin does not match the mask", details=f"data {tuple(t.shape)} vs mask {self.space}")
        t = t.to(self.device)
        if t.dtype != torch.bool:
            t = t != 0  # any other dtype: "non-zero is True"; bool storage is exactly 0 / 1
        t = t.reshape(t.shape[0], -1).contiguous().view(torch.uint8)
        T = int(t.shape[0])
        if self.pack_input:  # experiment: bool bytes -> flattened bits first (a word per thread), then the word-gather kernels
            nw = (self.N + 31) // 32
            bits = self._words(T, nw)
            _call("marex_morph_pack_u8", _p(t), T, self.N, self.N, _p(bits), nw, _stream(self.device))
            return _Source(bits, True, nw, row_stride, 0), T, as_numpy
        return _Source(t, False, self.N, row_stride, 0), T, as_numpy

    def _output(self, T: int, packed: bool):
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        if packed:
            nw = (self.N + 31) // 32
            bits = torch.empty((T, nw), dtype=torch.int32, device=self.device)
            return None, bits, count
        return torch.empty((T, self.N), dtype=torch.uint8, device=self.device), None, count

    def _finish(self, events, bits, count, T: int, as_numpy: bool):
        self.last_count = int(count.item())
        if bits is not None:
            return bits.cpu().numpy() if as_numpy else bits
        out = events.view(torch.bool).reshape((T,) + self.space)
        return out.cpu().numpy() if as_numpy else out

    def _words(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ gridded primitives
    def _pad(self, src: _Source, mask, T: int, pad: int) -> torch.Tensor:
        ny, nx = self.space
        Hp, Wpw = ny + 2 * pad, (nx + 2 * pad + 31) // 32
        assert Hp * Wpw == _lib.load().marex_morph_slab_words(ny, nx, pad)
        slab = self._words(T, Hp, Wpw)
        _call("marex_morph_pad_bits", *src.args(mask), T, ny, nx, pad, 0 if self.regional_mode else 1, _p(slab), _stream(self.device))
        return slab

    def _interior(self, slab: torch.Tensor, pad: int) -> _Source:
        Hp, Wpw = int(slab.shape[1]), int(slab.shape[2])
        return _Source(slab, True, Hp * Wpw, Wpw * 32, pad * Wpw * 32 + pad)

    def _close_open(self, slab: torch.Tensor, pad: int, R: int) -> torch.Tensor:
        """binary_closing then binary_opening with the disk of radius R (track.py:1630-1634): dilate, erode, erode, dilate.
        ``separable`` runs each pass as marex_morph_disk_sep (rows widened once, then 2R+1 single-word ORs) with level
        buffers for a chunk of time steps that stays in L2; otherwise as the direct marex_morph_disk."""
        T, Hp, Wpw = (int(s) for s in slab.shape)
        Wp = self.space[1] + 2 * pad
        other = torch.empty_like(slab)
        scratch = None
        if self.separable:
            nlev = int(_lib.load().marex_morph_disk_levels(R))
            chunk = max(1, min(T, SEPARABLE_SCRATCH_BYTES // (4 * nlev * Hp * Wpw)))
            scratch = self._words(nlev * chunk * Hp * Wpw)
        for erode in (0, 1, 1, 0):
            if scratch is not None:
                _call("marex_morph_disk_sep", _p(slab), _p(other), T, Hp, Wp, R, erode, _p(scratch), scratch.numel(), _stream(self.device))
            else:
                _call("marex_morph_disk", _p(slab), _p(other), T, Hp, Wp, R, erode, _stream(self.device))
            slab, other = other, slab
        return slab

    def _time_close(self, slab: torch.Tensor) -> torch.Tensor:
        """Temporal closing with T_fill + 1 ones on the False-padded time axis (track.py:1695-1719)."""
        T = int(slab.shape[0])
        words = int(slab.shape[1]) * int(slab.shape[2])
        half = self.T_fill // 2
        K = 2 * half + 1
        dil = self._words(T + 2 * half, slab.shape[1], slab.shape[2])
        _call("marex_morph_time", _p(slab), T, words, _p(dil), T + 2 * half, -2 * half, K, 0, _stream(self.device))
        out = torch.empty_like(slab)
        _call("marex_morph_time", _p(dil), T + 2 * half, words, _p(out), T, 0, K, 1, _stream(self.device))
        return out

    _OCEAN = object()  # default of _extract: apply the ocean mask

    def _extract(self, src: _Source, T: int, packed: bool, as_numpy: bool, mask_bits=_OCEAN):
        """Trim + ``where(mask)`` (track.py:1638-1643, 1667); ``mask_bits=None`` extracts without masking."""
        ny, nx = self.space
        events, bits, count = self._output(T, packed)
        _call("marex_morph_extract", *src.args(self.mask_bits if mask_bits is MaskFiller._OCEAN else mask_bits), T, ny, nx,
              _p(events), self.N, _p(bits), (self.N + 31) // 32, _p(count), _stream(self.device))  # fmt: skip
        return self._finish(events, bits, count, T, as_numpy)

    # ------------------------------------------------------------------ unstructured primitives
    def _tpack(self, src: _Source, T: int) -> torch.Tensor:
        Tw = (T + 31) // 32 + 2
        assert Tw == _lib.load().marex_morph_tpack_words(T)
        packed = self._words(self.N, Tw)
        _call("marex_morph_tpack", None if src.is_bits else _p(src.tensor), _p(src.tensor) if src.is_bits else None,
              src.t_pitch, T, self.N, _p(packed), _stream(self.device))  # fmt: skip
        return packed

    def _nbr_power(self, x: torch.Tensor, T: int, R: int, erode: int, set_land: bool) -> torch.Tensor:
        """sparse_bool_power (track.py:5423-5470): R applications of neighbours + identity; ``set_land`` applies
        ``bitmap[:, ~mask] = True`` (track.py:1566, 1574) to the input of the first one."""
        other = torch.empty_like(x)
        steps = [(self.nv, 1 if (set_land and i == 0) else 0) for i in range(R)]
        if R == 0 and set_land:
            steps = [(0, 1)]  # no hop, only the land cells are set
        for nv, land in steps:
            _call("marex_morph_nbr", _p(x), _p(other), T, self.N, _p(self.neighbours), nv, _p(self.mask), erode, land, _stream(self.device))
            x, other = other, x
        return x

    def _fill_holes_unstructured(self, x: torch.Tensor, T: int, R: int) -> torch.Tensor:
        x = self._nbr_power(x, T, R, 0, False)  # closing: dilation ...
        x = self._nbr_power(x, T, R, 1, True)  # ... land set True, erosion
        x = self._nbr_power(x, T, R, 1, True)  # opening: land set True, erosion ...
        return self._nbr_power(x, T, R, 0, False)  # ... dilation

    def _time_close_unstructured(self, x: torch.Tensor, T: int) -> torch.Tensor:
        half = self.T_fill // 2
        dil, out = torch.empty_like(x), torch.empty_like(x)
        _call("marex_morph_tshift", _p(x), _p(dil), T, self.N, half, 0, 1, _stream(self.device))
        _call("marex_morph_tshift", _p(dil), _p(out), T, self.N, half, 1, 0, _stream(self.device))
        return out

    def _tunpack(self, x: torch.Tensor, T: int, packed: bool, as_numpy: bool):
        events, bits, count = self._output(T, packed)
        _call("marex_morph_tunpack", _p(x), T, self.N, None, _p(events), self.N, _p(bits), (self.N + 31) // 32,
              _p(count), _stream(self.device))  # fmt: skip
        return self._finish(events, bits, count, T, as_numpy)

    # ------------------------------------------------------------------ the reference's methods
    @_on_self_device
    def fill_holes(self, data_bin=None, R_fill: Optional[int] = None, packed: bool = False, from_bits=None):
        """``tracker.fill_holes`` (track.py:1520-1669)."""
        R = self.R_fill if R_fill is None else int(R_fill)
        if not 0 <= R <= MAX_R_FILL:
            raise ConfigurationError("R_fill outside the range of the bit-packed kernels", details=f"R_fill={R} (0..{MAX_R_FILL})")
        src, T, as_numpy = self._input(data_bin, from_bits)
        if self.unstructured:
            return self._tunpack(self._fill_holes_unstructured(self._tpack(src, T), T, R), T, packed, as_numpy)
        if R == 0:
            return self._extract(src, T, packed, as_numpy)  # only the final `where(mask)` (track.py:1667)
        self._check_pad(2 * R)
        slab = self._close_open(self._pad(src, None, T, 2 * R), 2 * R, R)
        return self._extract(self._interior(slab, 2 * R), T, packed, as_numpy)

    @_on_self_device
    def fill_time_gaps(self, data_bin=None, packed: bool = False, from_bits=None):
        """``tracker.fill_time_gaps`` (track.py:1671-1726): temporal closing, then ``fill_holes`` with ``R_fill // 2``."""
        src, T, as_numpy = self._input(data_bin, from_bits)
        if self.unstructured:
            x = self._tpack(src, T)
            if self.T_fill != 0:
                x = self._fill_holes_unstructured(self._time_close_unstructured(x, T), T, self.R_fill // 2)
            return self._tunpack(x, T, packed, as_numpy)
        return self._time_gaps_gridded(src, None, T, packed, as_numpy)

    @_on_self_device
    def run(self, data_bin=None, packed: bool = False, from_bits=None):
        """``fill_time_gaps(fill_holes(data_bin))`` (track.py:1288-1297) without unpacking in between."""
        src, T, as_numpy = self._input(data_bin, from_bits)
        R = self.R_fill
        if self.unstructured:
            x = self._fill_holes_unstructured(self._tpack(src, T), T, R)
            if self.T_fill != 0:
                x = self._fill_holes_unstructured(self._time_close_unstructured(x, T), T, R // 2)
            return self._tunpack(x, T, packed, as_numpy)
        if R == 0:
            return self._time_gaps_gridded(src, self.mask_bits, T, packed, as_numpy)
        self._check_pad(2 * R)
        slab = self._close_open(self._pad(src, None, T, 2 * R), 2 * R, R)
        return self._time_gaps_gridded(self._interior(slab, 2 * R), self.mask_bits, T, packed, as_numpy)

    # ------------------------------------------------------------------
    def _time_gaps_gridded(self, src: _Source, src_mask, T: int, packed: bool, as_numpy: bool):
        """The gridded body of fill_time_gaps on a source; ``src_mask`` applies the pending `where(mask)` of a preceding
        fill_holes while the cells are read.  Padding commutes with the per-cell temporal closing, so the slab is
        padded for the second fill_holes first and closed in time afterwards."""
        if self.T_fill == 0:
            if src_mask is None and not src.is_bits and not packed:  # the reference returns data_bin itself
                events = src.tensor
                self.last_count = int(events.count_nonzero().item())
                out = events.view(torch.bool).reshape((T,) + self.space)
                return out.cpu().numpy() if as_numpy else out
            return self._extract(src, T, packed, as_numpy, mask_bits=src_mask)
        R2 = self.R_fill // 2
        self._check_pad(2 * R2)
        slab = self._time_close(self._pad(src, src_mask, T, 2 * R2))
        if R2 > 0:
            slab = self._close_open(slab, 2 * R2, R2)
        return self._extract(self._interior(slab, 2 * R2), T, packed, as_numpy)

    def _check_pad(self, pad: int) -> None:
        ny, nx = self.space
        if pad > min(ny, nx):
            raise DataValidationError(
                "grid smaller than the padding of the morphological operations",
                details=f"pad = {pad} cells, grid = {ny} x {nx}",
                suggestions=["use a smaller R_fill"],
            )


def fill_holes(data_bin, mask, R_fill, regional_mode: bool = False, neighbours=None, device=None, **kw):
    """Functional form of ``tracker.fill_holes`` (track.py:1520-1669)."""
    return MaskFiller(mask, R_fill, 0, regional_mode, neighbours, device).fill_holes(data_bin, **kw)


def fill_time_gaps(data_bin, mask, R_fill, T_fill: int = 2, regional_mode: bool = False, neighbours=None, device=None, **kw):
    """Functional form of ``tracker.fill_time_gaps`` (track.py:1671-1726)."""
    return MaskFiller(mask, R_fill, T_fill, regional_mode, neighbours, device).fill_time_gaps(data_bin, **kw)


# --------------------------------------------------------------------------------------------------------------
# multi-GPU: space-sharded mask -> time-sharded stage 1
# --------------------------------------------------------------------------------------------------------------
def time_blocks(T: int, world: int, halo: int):
    """Even split of the T time steps over the ranks: [(own_lo, own_hi, load_lo, load_hi)] per rank; the loaded range
    carries ``halo`` extra steps on interior edges (truncated at the series' ends, where the reference pads with False)."""
    out = []
    base, rem = divmod(T, world)
    lo = 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi, max(0, lo - halo), min(T, hi + halo)))
        lo = hi
    return out


def stage1_time_sharded(bits_band: torch.Tensor, T: int, rows: int, mask, R_fill, T_fill: int = 2, regional_mode: bool = False,
                        group=None, packed: bool = True, device=None):
    """Stage 1 after a SPACE-sharded ``preprocess_data`` (marex_b200/sharding.py: every rank owns a latitude band).

    ``fill_holes`` couples every cell with cells up to 6 * R_fill rows away (four disk passes with R_fill, four with
    R_fill // 2), which is most of a band, while ``fill_time_gaps`` only couples a cell with itself +- T_fill steps.  So
    the packed mask is RE-SHARDED BY TIME -- the one real exchange of this stage: every rank sends each peer the peer's
    block of days (plus a T_fill halo) of its own band, packed bits only (1.2 GB in total for 0.25 deg x 25 yr), with
    point-to-point sends over NCCL (NVLink) or gloo -- and each rank then runs the unchanged single-GPU stage on whole
    (lat, lon) fields of its block of days.  Exact: steps beyond the halo cannot reach the owned steps, and at the ends of
    the series the missing halo is the reference's own False padding (track.py:1706).

    bits_band  int32 [T, rows * nx / 32]: this rank's band of the flattened bit mask (``rows * nx`` must be a multiple
               of 32 so that bands concatenate on word boundaries); bands are ordered by rank
    mask       the FULL (ny, nx) ocean mask
    Returns ``(result, (t_lo, t_hi))``: the filled mask of this rank's block of days, packed int32 [t_hi - t_lo, N / 32]
    or bool [t_hi - t_lo, ny, nx].
    """
    import torch.distributed as dist

    mask = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask).astype(bool)
    ny, nx = mask.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local_ok = (rows * nx) % 32 == 0 and bits_band.dim() == 2 and bits_band.shape[0] == T and bits_band.shape[1] == rows * nx // 32
    local_msg = f"got {tuple(bits_band.shape)} for T={T}, rows={rows}, nx={nx}"
    if world == 1:
        if not local_ok:
            raise DataValidationError("bits_band must be int32 [T, rows * nx / 32] with rows * nx a multiple of 32", details=local_msg)
        if rows != ny:
            raise DataValidationError("a single rank must own the whole grid", details=f"rows={rows}, ny={ny}")
        return MaskFiller(mask, R_fill, T_fill, regional_mode, device=device).run(from_bits=(bits_band.contiguous(), T), packed=packed), (0, T)
    # every rank's band size AND the verdict of its local checks travel in the first collective: a rank with a bad
    # input must not raise while its peers block in the exchange -- all ranks raise together
    dev = bits_band.device  # NCCL moves CUDA tensors, gloo CPU tensors: keep the bookkeeping where the data is
    all_info = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_info, torch.tensor([rows, 1 if local_ok else 0], dtype=torch.int64, device=dev), group=group)
    all_rows = [int(r[0].item()) for r in all_info]
    bad = [p for p, r in enumerate(all_info) if int(r[1].item()) == 0]
    if bad:
        raise DataValidationError(
            "bits_band must be int32 [T, rows * nx / 32] with rows * nx a multiple of 32",
            details=f"rejected on rank(s) {bad}" + (f"; this rank: {local_msg}" if rank in bad else ""),
        )
    if sum(all_rows) != ny or any((r * nx) % 32 for r in all_rows):
        raise DataValidationError("the bands of all ranks must tile the grid on word boundaries", details=f"rows per rank {all_rows}, ny={ny}")
    filler = MaskFiller(mask, R_fill, T_fill, regional_mode, device=device)
    blocks = time_blocks(T, world, int(T_fill))
    own_lo, own_hi, load_lo, load_hi = blocks[rank]
    n_load = load_hi - load_lo
    bits_band = bits_band.contiguous()
    recv = [torch.empty((n_load, all_rows[p] * nx // 32), dtype=torch.int32, device=bits_band.device) for p in range(world)]
    ops, keep = [], []
    for q in range(world):
        if q == rank:
            recv[rank].copy_(bits_band[load_lo:load_hi])
            continue
        piece = bits_band[blocks[q][2] : blocks[q][3]].contiguous()
        keep.append(piece)
        ops.append(dist.P2POp(dist.isend, piece, q if group is None else dist.get_global_rank(group, q), group))
        ops.append(dist.P2POp(dist.irecv, recv[q], q if group is None else dist.get_global_rank(group, q), group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    field = torch.cat(recv, dim=1)  # bands are whole words: the flattened bit mask of the full grid, n_load time steps
    out = filler.run(from_bits=(field, n_load), packed=packed)
    return out[own_lo - load_lo : own_hi - load_lo], (own_lo, own_hi)
