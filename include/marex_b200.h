/*
 * marex_b200.h -- C-ABI of libmarex_b200.so: the B200 (sm_100a) implementation of the
 * marEx `preprocess_data` detection hot path.
 *
 * The reference (wienkers/marEx) is pure Python and has no FFI layer; its boundary for this
 * path is the Python signature `marEx.preprocess_data` (marEx/detect.py:287-313).  Each entry
 * point below replaces one arithmetic stage that the reference expresses as an
 * xarray/dask/flox/numpy graph; the stage it replaces is cited as detect.py:<lines>.
 * `marex_b200/detect.py` (the Python mirror of the reference API) binds these with ctypes;
 * INTEGRATION.md shows the stub a marEx maintainer would add.
 *
 * Conventions
 * -----------
 *  - Every pointer is CALLER-OWNED DEVICE memory (cudaMalloc / torch allocation) unless the
 *    comment says "host".  The library allocates nothing that outlives a call.
 *  - Fields are time-major: element (t, c) of a [T, N] field lives at base[t * pitch + c],
 *    pitch >= N in ELEMENTS.  Gridded fields flatten (lat, lon) row-major: c = iy * nx + ix.
 *  - Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy stream).
 *  - Return value: MAREX_OK or a negative MAREX_ERR_* code; marex_last_error() returns the
 *    thread-local message of the last failure on the calling thread.
 *  - Day-of-year thresholds are produced and consumed DOY-MAJOR: thr[(doy-1) * N + c].
 *
 * Calendar tables (built on the host from the time coordinate, see marex_b200/calendar.py;
 * reference: `.dt.year` / `.dt.dayofyear`, detect.py:1605-1606) and uploaded by the caller:
 *  - doy[T]      int16  day of year 1..366 of every row
 *  - year_val[n_years] int32 ascending distinct calendar years present
 *  - tidx[n_years*366] int32 row index of (year i, doy d) at [i*366 + d-1], or -1
 *  - out_row[T]  int32  output row of input row t, or -1 when the row is trimmed
 *  - doy_ptr[367], doy_rows[doy_ptr[366]] int32  CSR list of the rows of each day of year
 *  - slot_row[366*NY] int32  row of (day of year d, k-th output year) at [d*NY + k], or -1: the
 *    day-of-year-major "slot" order of the histogram bin codes
 */
#ifndef MAREX_B200_H
#define MAREX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAREX_OK 0
#define MAREX_ERR_INVALID_ARG (-1)
#define MAREX_ERR_CUDA (-2)
#define MAREX_ERR_UNSUPPORTED (-3)

#define MAREX_NDOY 366

/* Library version (major*10000 + minor*100 + patch). */
int marex_version(void);
/* Message of the last error on this thread ("" if none). */
const char* marex_last_error(void);
/* Number of kernel launches issued by this process so far (bench.py's gpu_launches). */
long long marex_launch_count(void);
/* Tuning / test knobs (kernel shapes, forced fall-backs; never changes a result): set != 0 pins `key` to `value`
 * for this process, set == 0 removes the pin.  An unpinned key falls back to the environment variable
 * MAREX_<KEY IN CAPITALS> (read once), then to the built-in default.  Keys: DESIGN.md section 9. */
int marex_tune(const char* key, long long value, int32_t set);

/* ---- (a) shifting baseline ------------------------------------------------------------
 * smoothed_rolling_climatology + anomaly + trim (detect.py:1511-1688, 1691-1816, 1819-1850,
 * 615-641) and the per-cell numbers _validate_data_values needs (detect.py:205-279).
 *   s[t]     = centred S-day mean of x (NaN unless the full window is valid)
 *   clim     = nanmean of s over the (year, doy) samples of the previous W calendar years
 *   anom     = x - clim for rows with out_row[t] >= 0   (anom row = out_row[t])
 *   mask0[c] = isfinite(x[0, c]);  nonfinite[c] = number of non-finite x[:, c]
 * mode 0 writes the anomaly, mode 1 writes clim itself (marEx.rolling_climatology /
 * smoothed_rolling_climatology, detect.py:1511, 1691); rows of non-target years are not
 * written.  Requires at most one row per (year, doy) (daily data, gaps allowed).  W, S <= 1023. */
int marex_shift_anomaly_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                            const int32_t* tidx, const int32_t* year_val, int32_t n_years,
                            int32_t W, int32_t S, const int32_t* out_row, int32_t mode,
                            float* anom, int64_t anom_pitch,
                            uint8_t* mask0, int32_t* nonfinite, void* stream);

/* Same stage, fast path for a GAP-FREE DAILY proleptic-Gregorian time axis whose row 0 is day
 * `doy0` (1..366) of calendar year `year0` (the kernel derives every calendar table itself).
 * One TMA box load per (64 gridpoints, day-of-year strip, year) stages the rows in shared
 * memory; requires x / out 16-byte aligned, N and the pitches multiples of 4, W <= 31.  Sums are float32
 * (block sums + Kahan-compensated ring sum, within 1e-6 of the field scale of the float64 result;
 * marex_tune("shift_f64", 1, 1) / MAREX_SHIFT_F64=1 selects float64 sums).  Exact for gridpoints whose series is
 * all finite or all NaN; `nonfinite` receives the per-gridpoint count of non-finite inputs and
 * marex_shift_anomaly_fixup_f32 must follow to recompute gridpoints with 0 < count < T.
 * Output row of input row t is t - (first row of year0 + W) for mode 0, t for mode 1.
 * Fused np.digitize (detect.py:2622-2631), mode 0 only: when `bins` is non-NULL the kernel also writes the
 * histogram bin code of every anomaly it stores (edges[n_edges] float32, edges[0] = -inf; NaN and
 * a >= edges[n_edges-1] -> 0x7FFF) into the DAY-OF-YEAR-MAJOR bin array
 *     bins[(doy0based * NY + k) * bins_pitch + c],  k = year index among the NY output years,
 * and 0x7FFF into the slots of (day, year) pairs that have no row (day 366 of non-leap years, days after the
 * end of the series): the layout marex_hobday_thresholds_pooled_bins / marex_compare_hobday_bins consume.
 * MAREX_ERR_UNSUPPORTED when the windows do not fit shared memory (use marex_shift_anomaly_f32). */
int marex_shift_anomaly_daily_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                                  int32_t year0, int32_t doy0, int32_t W, int32_t S, int32_t mode,
                                  float* out, int64_t out_pitch,
                                  uint8_t* mask0, int32_t* nonfinite,
                                  const float* edges, int32_t n_edges, uint16_t* bins, int64_t bins_pitch,
                                  void* stream);
/* Recomputes, with the generic kernel of marex_shift_anomaly_f32 (same tables), the gridpoints
 * whose `nonfinite` count is strictly between 0 and T.  `work`: device scratch of N + 1 int32.
 * With `bins` non-NULL (same arguments as above; T_out output rows starting on Jan 1 of `year_first`)
 * the bin codes of those gridpoints are recomputed from the corrected anomalies as well. */
int marex_shift_anomaly_fixup_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                                  const int32_t* tidx, const int32_t* year_val, int32_t n_years,
                                  int32_t W, int32_t S, const int32_t* out_row, int32_t mode,
                                  float* anom, int64_t anom_pitch,
                                  uint8_t* mask0, int32_t* nonfinite, int32_t* work,
                                  const float* edges, int32_t n_edges, uint16_t* bins, int64_t bins_pitch,
                                  int64_t T_out, int32_t year_first, void* stream);

/* ---- (a') fixed baseline ---------------------------------------------------------------
 * Per-day-of-year nanmean (flox nanmean, detect.py:2365-2373) over the rows listed in the CSR
 * (doy_ptr, doy_rows) -- all rows, or only the reference_period's rows (detect.py:2334-2361).
 * If `shift` is non-NULL the averaged value is f32(x - shift[c]) (used after detrending with
 * force_zero_mean).  clim[366, N] float32, NaN for days of year without a valid sample. */
int marex_doy_climatology_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                              const int32_t* doy_ptr, const int32_t* doy_rows,
                              const float* shift, float* clim, void* stream);
/* anom[t] = f32(f32(x[t] - shift) - clim[doy[t]])  (groupby subtraction, detect.py:2377-2379);
 * mask0 = isfinite of the first shifted row (detect.py:2391).  `nonfinite` (optional) counts
 * non-finite x per cell.  x and anom may alias (in-place).  clim == NULL subtracts only `shift`
 * (the force_zero_mean step of detrend_harmonic, detect.py:2223-2224). */
int marex_sub_doy_climatology_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                                  const int16_t* doy, const float* shift, const float* clim,
                                  float* anom, int64_t anom_pitch,
                                  uint8_t* mask0, int32_t* nonfinite, void* stream);

/* ---- (a'') polynomial detrend ----------------------------------------------------------
 * coef[K, N] = P^T x  with P[T, K] = pinv(model) (detect.py:2169, 2206), float64 accumulate.
 * Also the raw-input validation numbers (mask0, nonfinite) as in (a). */
int marex_detrend_coef_f64(const float* x, int64_t T, int64_t N, int64_t pitch,
                           const double* P, int32_t K, double* coef,
                           uint8_t* mask0, int32_t* nonfinite, void* stream);
/* xd[t] = x[t] - f32(sum_k M[k, t] * coef[k])  (detect.py:2220); if `mean` is non-NULL it
 * receives f32(nanmean_t xd) (the force_zero_mean term, detect.py:2223-2224). */
int marex_detrend_apply_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                            const double* M, int32_t K, const double* coef,
                            float* xd, int64_t xd_pitch, float* mean, void* stream);

/* ---- std_normalise of detrend_harmonic (detect.py:2257-2293) -------------------------------
 * sd[366, N] = population standard deviation of the rows of each day of year (flox "std",
 * :2260-2268; NaN for an empty group or a non-finite sample). */
int marex_doy_std_f32(const float* x, int64_t T, int64_t N, int64_t pitch,
                      const int32_t* doy_ptr, const int32_t* doy_rows, float* sd, void* stream);
/* out[366, N] = sqrt of the centred `win`-day rolling mean of sd^2 with annual wrap (:2271-2272),
 * values <= 1e-10 replaced by NaN (:2276).  sd and out must not alias. */
int marex_doy_rolling_rms_f32(const float* sd, int64_t N, int32_t win, float* out, void* stream);
/* out[t] = x[t] / sd[doy[t]]  (:2278). */
int marex_div_doy_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int16_t* doy,
                      const float* sd, float* out, int64_t out_pitch, void* stream);

/* ---- (b) thresholds --------------------------------------------------------------------
 * np.digitize(a, edges) - 1 as uint16 (detect.py:2622-2631) into the DAY-OF-YEAR-MAJOR bin array: slot
 * s = doy0based * NY + k (k = index of the year among the NY output years) receives the codes of input row
 * slot_row[s], or 0x7FFF when slot_row[s] < 0 (no row for that day and year).  NaN and a >= edges[n_edges-1]
 * give 0x7FFF (not counted).  edges[0] must be -inf.  n_slots = 366 * NY. */
int marex_digitize_doy_f32(const float* a, int64_t N, int64_t pitch,
                           const int32_t* slot_row, int64_t n_slots,
                           const float* edges, int32_t n_edges,
                           uint16_t* bins, int64_t bins_pitch, void* stream);

/* Approximate Hobday thresholds: (doy x bin) counts, ws x ws spatial pooling (periodic in x,
 * truncated in y), +-w/2 doy window (wrap 366), count-space interpolated quantile, NaN mask
 * from anom_row0, clamp to lower_bound (_compute_histogram_quantile_2d detect.py:2562-2734,
 * _rolling_histogram_quantile detect.py:2465-2559).  Unstructured: ny = 1, nx = N, ws = 1.
 * `bins` rows are listed per day of year by the CSR (doy_ptr, doy_rows): for the day-of-year-major
 * array doy_ptr[d] = d * NY and doy_rows[j] = j.  Codes >= nb are not counted.
 * thr[366, ny*nx] float32 doy-major.  stats[2] receives {min, max} of the thresholds before
 * the clamp, ignoring NaN (for the reference's two UserWarnings).  w odd, 3 <= w <= 365. */
int marex_hobday_thresholds_hist(const uint16_t* bins, int64_t T, int64_t ny, int64_t nx,
                                 int64_t pitch, const int32_t* doy_ptr, const int32_t* doy_rows,
                                 int32_t max_window_rows, const float* centers, int32_t nb,
                                 int32_t w, int32_t ws, double q, const float* anom_row0,
                                 float lower_bound, float* thr, float* stats, void* stream);

/* Same computation for gridded data with ws in {3, 5, 7} from the day-of-year-major bin array
 * (NY slots per day of year) with the banded warp-cooperative kernels (pool_band.cu); tiles whose
 * thresholds do not fit one band are recomputed with full-range counters.  Requires nx >= 32, nb <= 1024,
 * w * NY <= 65535.  (doy_ptr, doy_rows) as above.  `workspace`: device scratch of at least
 * marex_hobday_pooled_workspace_bytes(ny, nx) bytes.  The NaN mask is row 0 of the anomalies
 * (anom_row0[ny*nx], detect.py:2704-2705). */
int64_t marex_hobday_pooled_workspace_bytes(int64_t ny, int64_t nx);
int marex_hobday_thresholds_pooled_bins(const uint16_t* bins, int64_t NY, int64_t ny, int64_t nx,
                                        int64_t bins_pitch, const int32_t* doy_ptr, const int32_t* doy_rows,
                                        const float* centers, int32_t nb, int32_t w, int32_t ws, double q,
                                        const float* anom_row0, float lower_bound,
                                        float* thr, float* stats, void* workspace, int64_t workspace_bytes,
                                        void* stream);

/* Exact Hobday thresholds: np.nanpercentile (float32 'linear') over the +-w/2 doy window
 * (detect.py:1921-1956).  thr[366, N] doy-major, NaN where the window holds no valid sample.
 * max_window_rows / max_doy_rows = the largest number of rows of any doy window / of any single day of year
 * (doy_ptr differences).  `work`: optional device scratch of 2 * N float32.  With it, and when the samples that decide
 * the percentile (the n - floor(q (n - 1)) largest of a window, plus room) fit a queue of 64 or 128 floats per
 * gridpoint, the queue kernel runs (csrc/exact_queue.cuh) and `work` holds, as int32, the number of groups of 32
 * gridpoints it handed to the histogram kernel ([0]) and their indices; otherwise the histogram kernel runs for every
 * gridpoint (window in shared memory when w * max_doy_rows samples fit) and `work` receives the per-gridpoint value
 * range taken in a separate full-occupancy pass.  The result does not depend on which kernel ran. */
int marex_hobday_thresholds_exact_f32(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                      const int32_t* doy_ptr, const int32_t* doy_rows,
                                      int32_t max_window_rows, int32_t max_doy_rows, int32_t w,
                                      float percentile, float* thr, float* work, void* stream);

/* Global (constant in time) thresholds, approximate: per-cell histogram with float64 edges
 * (last bin right-closed), pdf/cdf in float64 in the reference's order
 * (_compute_histogram_quantile_1d detect.py:2737-2865).  thr[N] float64. */
int marex_global_threshold_hist_f64(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                    const double* edges, const double* centers, int32_t nb,
                                    double q, double lower_bound, double* thr, double* stats,
                                    void* stream);
/* Same result, fast path for T <= 65535: thread = gridpoint, two passes (8-bin blocks, then the bins
 * of the block that holds the rank); gridpoints whose quantile bin is not decided by integer
 * counts alone (see thresholds.cu) are recomputed in the reference's float64 order.
 * edges_up[nb + 1] float32: the smallest float32 >= each float64 edge (edges_up[0] = -inf);
 * e_last_dn: the largest float32 <= edges[nb].  work: device scratch of N + 1 int32. */
int marex_global_threshold_hist_fast_f64(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                         const double* edges, const float* edges_up, float e_last_dn,
                                         const double* centers, int32_t nb, double q,
                                         double lower_bound, double* thr, double* stats,
                                         int32_t* work, void* stream);
/* Global thresholds, exact: np.nanquantile(a, q) with a float64 q ('linear'), float64 result
 * (xarray .quantile, detect.py:2899). */
int marex_global_threshold_exact_f64(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                     double q, double* thr, void* stream);

/* ---- (c) compare ------------------------------------------------------------------------
 * events[t, c] = anom[t, c] >= thr  (NaN on either side -> 0).  hobday: thr[366, N] float32
 * selected by doy[t] (detect.py:2001-2004); global: thr[N] float64, compared in float64
 * (detect.py:2915).  Either output may be NULL: `events` is one byte per gridpoint-day (the
 * bool array the reference returns), `bits` is the bit-packed mask, row t at
 * bits[t * bits_pitch ...], bit (c & 31) of word (c >> 5).  `count` (optional, device
 * uint64, accumulated into) receives the number of extreme gridpoint-days.  When the optional CSR
 * (doy_ptr, doy_rows) of the rows of each day of year is given and N % 4 == 0, the hobday compare
 * runs day-of-year major (each threshold row is read once, 16-byte loads).  thr_pitch (elements,
 * >= N) is the row stride of thr, so a sub-range of gridpoints of a larger field can be compared. */
int marex_compare_hobday(const float* anom, int64_t T, int64_t N, int64_t pitch,
                         const int16_t* doy, const int32_t* doy_ptr, const int32_t* doy_rows,
                         const float* thr, int64_t thr_pitch,
                         uint8_t* events, int64_t events_pitch,
                         uint32_t* bits, int64_t bits_pitch,
                         unsigned long long* count, void* stream);
/* The hobday compare from the day-of-year-major bin codes (2 bytes per sample instead of 4): a sample whose
 * bin lies above / below the bin of its threshold is decided without the anomaly; the float anomaly
 * (anom[row * pitch + c], row = slot_row[s]) is read only for samples in the threshold's own bin or with the
 * invalid code.  Results are identical to marex_compare_hobday.  Needs N % 8 == 0, bins_pitch % 8 == 0,
 * events_pitch % 8 == 0 and 16-byte aligned bases (MAREX_ERR_UNSUPPORTED otherwise). */
int marex_compare_hobday_bins(const uint16_t* bins, int64_t NY, int64_t bins_pitch, const int32_t* slot_row,
                              const float* anom, int64_t pitch, int64_t N,
                              const float* thr, int64_t thr_pitch,
                              const float* edges, int32_t n_edges,
                              uint8_t* events, int64_t events_pitch,
                              uint32_t* bits, int64_t bits_pitch,
                              unsigned long long* count, void* stream);
int marex_compare_global(const float* anom, int64_t T, int64_t N, int64_t pitch,
                         const double* thr,
                         uint8_t* events, int64_t events_pitch,
                         uint32_t* bits, int64_t bits_pitch,
                         unsigned long long* count, void* stream);

/* ---- tracker stage 1 on bit-packed masks -------------------------------
 * fill_holes (marEx/track.py:1520-1669) and fill_time_gaps (track.py:1671-1726), the first consumers
 * of `extreme_events`.  Gridded fields are processed as row-aligned padded BIT SLABS: a padded time step
 * is Hp = ny + 2*pad rows of Wpw = ceil((nx + 2*pad) / 32) words, bit i of word w = padded column 32*w + i
 * (marex_morph_slab_words() words per time step).  A "source" is either bool bytes (src_bytes) or bits
 * (src_bits): cell (y, x) of time step t is element / bit  src_origin + y * src_row_stride + x  of the time
 * step starting at t * src_t_pitch (bytes, or uint32 words).  So [T, N] bool bytes are (events, NULL, pitch,
 * nx, 0), the flattened bits of marex_compare_* are (NULL, bits, bits_pitch, nx, 0), and the interior of a
 * slab is (NULL, slab, Hp*Wpw, Wpw*32, pad*Wpw*32 + pad).  `mask_bits` (optional: the ocean mask as flattened
 * bits, ceil(ny*nx / 32) words, bit c & 31 of word c >> 5) zeroes cells outside the mask at read time
 * (`data_bin.where(self.mask, other=False)`, track.py:1667).  Bits sources are moved a word at a time (two
 * loads + a funnel shift per run of cells); bool bytes go through a lane-per-cell __ballot_sync path.
 *
 *  marex_morph_pad_bits  np.pad of every time step by `pad` cells, mode "wrap" (wrap = 1, track.py:1617)
 *                        or "edge" (wrap = 0, regional_mode), into a slab (track.py:1625).
 *  marex_morph_disk      one binary dilation (erode = 0) or erosion (erode = 1) of every padded time step
 *                        by the disk x^2 + y^2 < R^2 + 1 (track.py:1613-1616) with border_value 0, as
 *                        dask_image.ndmorph / scipy.ndimage do inside binary_closing / binary_opening
 *                        (track.py:1630-1634).  Hp, Wp are the PADDED sizes in cells.  0 <= R <= 32.
 *  marex_morph_disk_sep  the same pass in separable form: pass H widens every input row once, step by step, and stores
 *                        it at the disk's marex_morph_disk_levels(R) distinct non-zero half-widths; pass V ORs
 *                        2R+1 single words per output word.  `scratch` (caller-owned, scratch_words uint32, at
 *                        least levels * Hp * ceil(Wp/32)) holds the level buffers of a CHUNK of time steps; size
 *                        it to stay in L2 (tens of MB) and the H -> V traffic never reaches HBM.  1 <= R <= 32.
 *  marex_morph_time      out[t] = OR / AND over k = 0..K-1 of in[t + off + k], whole slabs of `words` words,
 *                        time steps outside [0, T_in) False: the temporal closing of track.py:1695-1719 is
 *                        a dilation (T_out = T + 2*half, off = -2*half) then an erosion (T_out = T, off = 0)
 *                        with K = T_fill + 1 = 2*half + 1.
 *  marex_morph_pack_u8   bool bytes [T, N] (every byte 0 or 1) -> flattened bits, a word per thread (two 16-byte loads):
 *                        the fast way into the bit layouts for callers that hold the reference's bool array.
 *  marex_morph_extract   source cells -> bool bytes [T, N] and / or flattened bits (+ count of True cells,
 *                        accumulated into a device uint64), the trim of track.py:1638-1643 + the mask.
 *
 * Unstructured meshes (track.py:1543-1607) use CELL-MAJOR, TIME-PACKED words: cell c owns
 * marex_morph_tpack_words(T) = ceil(T/32) + 2 words, word k holds time steps 32*(k-1) .. 32*(k-1)+31 (word 0 and
 * the last word are margins for the temporal closing), so one neighbour gather serves 32 days.
 *  marex_morph_tpack     [T, N] bool bytes or flattened bits -> time-packed.
 *  marex_morph_nbr       one application of the sparse dilation matrix "neighbours + identity"
 *                        (track.py:1093-1115; sparse_bool_power track.py:5423-5470 applies it R_fill times):
 *                        nbr is int32 [nv, N], 0-based, negative = no neighbour.  erode = 1 computes
 *                        ~dilate(~x); set_land = 1 first sets cells outside `mask` ([N] bool BYTES here: one byte per
 *                        cell is what a per-cell gather wants) to True (track.py:1566, 1574).
 *  marex_morph_tshift    dilation / erosion by +-half time steps along the packed time axis; clip = 1 reads only
 *                        real time steps [0, T) (the constant False padding of track.py:1706).
 *  marex_morph_tunpack   time-packed -> bool bytes [T, N] and / or flattened bits (+ optional mask, count). */
int64_t marex_morph_slab_words(int64_t ny, int64_t nx, int32_t pad);
int marex_morph_pad_bits(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch,
                         int64_t src_row_stride, int64_t src_origin, const uint32_t* mask_bits, int64_t T,
                         int64_t ny, int64_t nx, int32_t pad, int32_t wrap, uint32_t* slab, void* stream);
int marex_morph_disk(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R,
                     int32_t erode, void* stream);
int32_t marex_morph_disk_levels(int32_t R);
int marex_morph_disk_sep(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R,
                         int32_t erode, uint32_t* scratch, int64_t scratch_words, void* stream);
int marex_morph_time(const uint32_t* in, int64_t T_in, int64_t words, uint32_t* out, int64_t T_out,
                     int32_t off, int32_t K, int32_t erode, void* stream);
int marex_morph_extract(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch,
                        int64_t src_row_stride, int64_t src_origin, const uint32_t* mask_bits, int64_t T,
                        int64_t ny, int64_t nx, uint8_t* events, int64_t events_pitch, uint32_t* bits,
                        int64_t bits_pitch, unsigned long long* count, void* stream);
int marex_morph_pack_u8(const uint8_t* events, int64_t T, int64_t N, int64_t pitch, uint32_t* bits, int64_t bits_pitch,
                        void* stream);
int64_t marex_morph_tpack_words(int64_t T);
int marex_morph_tpack(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch, int64_t T,
                      int64_t N, uint32_t* packed, void* stream);
int marex_morph_nbr(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, const int32_t* nbr, int32_t nv,
                    const uint8_t* mask, int32_t erode, int32_t set_land, void* stream);
int marex_morph_tshift(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, int32_t half, int32_t erode,
                       int32_t clip, void* stream);
int marex_morph_tunpack(const uint32_t* packed, int64_t T, int64_t N, const uint8_t* mask, uint8_t* events,
                        int64_t events_pitch, uint32_t* bits, int64_t bits_pitch, unsigned long long* count,
                        void* stream);

/* ---- utilities --------------------------------------------------------------------------
 * Strided host<->device copy (cudaMemcpy2DAsync) of `height` rows of `width_bytes`: the streamed
 * host path moves one latitude band of a (time, lat, lon) host array per call. */
int marex_memcpy2d_async(void* dst, int64_t dpitch_bytes, const void* src, int64_t spitch_bytes,
                         int64_t width_bytes, int64_t height, int32_t to_device, void* stream);
/*
 * out[c, r] = in[r, c] : doy-major thr[366, N] -> the reference's (..space, dayofyear) layout. */
int marex_transpose_f32(const float* in, int64_t rows, int64_t cols, float* out, void* stream);

/* Deterministic synthetic SST: seasonal cycle + trend + AR(1) noise, with
 * all-NaN "land" blobs; counter-based so any shard [c0, c0+N) of a global grid of
 * `n_global` cells regenerates identically.  doy_frac[T] = decimal year (float32) host-built. */
int marex_synth_sst_f32(float* x, int64_t T, int64_t N, int64_t pitch, int64_t c0,
                        int64_t ny_global, int64_t nx_global, const float* dec_year,
                        uint64_t seed, float land_fraction, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAREX_B200_H */
