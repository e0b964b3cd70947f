// Approximate Hobday thresholds with ws x ws spatial pooling -- banded, warp-cooperative kernel.
// Reference: _compute_histogram_quantile_2d (detect.py:2562-2734: digitize, (doy x bin) counts,
// pooling :2651-2668) and _rolling_histogram_quantile (detect.py:2465-2559).
//
// Tile = OY x 32 "own" gridpoints (warp = own row, lane = own column), of which the inner
// (OY - 2P) x (32 - 2P) are targets.  Each own gridpoint keeps the histogram of ITS OWN +-w/2
// day-of-year window in shared memory, but only for a BAND of K bins [Blo, Blo + K) that brackets
// the tile's thresholds: a p95 threshold is decided by the top few percent of the samples, so
//   * a sample below the band only counts in NT (valid samples) - no shared-memory traffic,
//   * a sample inside the band is a read-modify-write of one bin counter and one 8-bin block
//     counter of its own gridpoint (private to one thread: no atomics),
//   * a sample above the band only counts in TB (samples >= Blo).
// Per day-of-year step every thread first scans its entering / leaving samples (2 instructions
// for a below-band sample) and compacts the few that matter into a small shared list, then
// applies the list - so the divergent read-modify-write code runs for ~10 % of the samples only.
//
// A target's pooled cumulative count is  sum over its ws x ws neighbours of
// (NT - TB) + blocks below + bins below.  Pooling is separable, and the 32 lanes of a warp are 32
// adjacent own columns of one target row, so one counter row is pooled for the whole warp by
// 2P+1 shared loads (column sum) + 2P shuffles (row sum).  The warp scans block rows, then the
// bin rows of the blocks that hold its lanes' quantiles; each lane keeps what it needs.  All
// counts are integers: the result is bit-exact against the reference's arithmetic.
//
// If a threshold leaves the band the tile rebuilds its counters around the new position (a
// coarse full-range pass over the current window picks the band).  A tile whose thresholds do
// not fit one band is appended to `fail_list` and recomputed by the full-range tile kernel
// (thresholds.cu), so exactness never depends on the band heuristic.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "digitize.cuh"
#include "tma.cuh"

namespace marex {

constexpr int BAND_LCAP = 18;  // slots of the per-thread compacted event list
constexpr int BAND_CAP = 32;   // staged row indices per step and direction

struct BandParams {
  const uint16_t* bins;  // [T][ny*nx], 0x7FFF = invalid (NaN or >= last edge)
  int64_t ny, nx, pitch;
  const int32_t* doy_ptr;
  const int32_t* doy_rows;
  const float* centers;
  int nb, w, margin;
  double q;
  const float* anom_row0;
  float lower_bound;
  float* thr;
  float* stats;
  int32_t* fail_list;        // [0] = count, then (y0, x0) pairs of the tiles this launch gives up on
  const int32_t* tile_list;  // nullptr: tiles from blockIdx (x, y); else 1-D grid over the listed tiles
  int force_fail;
};

template <int P>
__device__ __forceinline__ int pooled_row(const uint16_t* __restrict__ p, int lane) {
  int v = 0;
#pragma unroll
  for (int dy = 0; dy <= 2 * P; ++dy) v += p[dy * 32];
  int s = v;
#pragma unroll
  for (int k = 1; k <= P; ++k) s += __shfl_sync(0xffffffffu, v, lane + k) + __shfl_sync(0xffffffffu, v, lane - k);
  return s;
}

constexpr int BAND_INV = BIN_INV;  // bin code of an invalid sample (NaN or >= last edge)
constexpr int BAND_PRE = 26;       // entering samples prefetched into registers per step

// NYC: rows per day of year as a compile-time constant (0 = from the calendar table).  The day-of-year-major code array
// has exactly NY slots per day of year, so for the record lengths the instantiations exist for (NY = 25: 40 years with a
// 15-year baseline; NY = 15: 30 years) the per-sample guards and the tail loops of the scans compile away.
template <int P, int K, int OY, int NYC = 0>
__global__ void __launch_bounds__(OY * 32, (OY <= 16 && K <= 64) ? 2 : 1) hobday_band_kernel(const BandParams p) {
  constexpr int KB = K / 8;  // blocks in the band
  constexpr int TY = OY - 2 * P, TX = 32 - 2 * P, CS = OY * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  uint16_t* const L0 = reinterpret_cast<uint16_t*>(smem_raw);  // [K][CS]; the coarse pass reuses it as [nblk][CS]
  uint16_t* const L1 = L0 + K * CS;                             // [KB][CS]
  uint16_t* const NTr = L1 + KB * CS;                           // [CS] valid samples in the own window
  uint16_t* const TBr = NTr + CS;                               // [CS] of which >= Blo
  uint16_t* const EV = TBr + CS;                                // [BAND_LCAP][CS]
  long long* const s_off = reinterpret_cast<long long*>(EV + BAND_LCAP * CS);  // [2 parity][2 leave/enter][BAND_CAP]
  int* const s_cnt = reinterpret_cast<int*>(s_off + 4 * BAND_CAP);             // [2][2][2] begin, end in doy_rows
  int* const s_misc = s_cnt + 8;                                               // [0] violation flag, [1] min blk, [2] max blk

  const int nb = p.nb, half = p.w / 2;
  const int nblk = (nb + 7) >> 3;
  const int64_t nx = p.nx, ny = p.ny, N = nx * ny;
  int64_t y0 = (int64_t)blockIdx.y * TY, x0 = (int64_t)blockIdx.x * TX;
  if (p.tile_list) {  // retry launch over the tiles a narrower band gave up on
    if ((int)blockIdx.x >= __ldg(&p.tile_list[0])) return;
    y0 = __ldg(&p.tile_list[1 + 2 * blockIdx.x]);
    x0 = __ldg(&p.tile_list[2 + 2 * blockIdx.x]);
  }

  // ---- own gridpoint of this thread (update role) ----
  const int64_t gy = y0 - P + warp;
  int64_t gx = (x0 - P + lane) % nx;
  if (gx < 0) gx += nx;
  const bool own_valid = gy >= 0 && gy < ny;
  const uint16_t* const col = p.bins + (own_valid ? gy * nx + gx : 0);
  const char* const colb = reinterpret_cast<const char*>(col);
  uint16_t* const myL0 = L0 + tid;
  uint16_t* const myL1 = L1 + tid;
  uint16_t* const ev = EV + tid;
  int NT = 0, TB = 0, Blo = 0;
  bool dead = !own_valid;  // no valid sample in the own window: nothing to keep up to date
  int pre[BAND_PRE];       // entering samples of the next step

  // ---- target of this lane (query role: warps 0..TY-1, lanes P..31-P) ----
  const int64_t ty_g = y0 + warp, tx_g = x0 + lane - P;
  const bool target_live = warp < TY && lane >= P && lane < 32 - P && ty_g < ny && tx_g < nx;
  bool masked = true;
  if (target_live) {
    const float a0 = p.anom_row0[ty_g * nx + tx_g];
    masked = a0 != a0;  // detect.py:2704-2705
  }
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;

  auto stage = [&](int step) {  // byte offsets of the rows leaving / entering the window at `step` (1..365)
    const int par = step & 1;
    const int d_leave = (step - 1 - half + 2 * NDOY) % NDOY, d_enter = (step + half) % NDOY;
    if (tid < 2 * BAND_CAP) {
      const int which = tid / BAND_CAP, u = tid % BAND_CAP;
      const int dd = which ? d_enter : d_leave;
      const int b0 = __ldg(&p.doy_ptr[dd]), b1 = __ldg(&p.doy_ptr[dd + 1]);
      s_off[(par * 2 + which) * BAND_CAP + u] = (b0 + u < b1) ? (long long)__ldg(&p.doy_rows[b0 + u]) * p.pitch * 2 : 0;
      if (u == 0) { s_cnt[(par * 2 + which) * 2] = b0; s_cnt[(par * 2 + which) * 2 + 1] = b1; }
    }
  };
  auto load_at = [&](long long off) -> int { return (int)*reinterpret_cast<const uint16_t*>(colb + off); };
  auto load_row = [&](int j) -> int { return (int)col[(int64_t)__ldg(&p.doy_rows[j]) * p.pitch]; };

  // One sample (relative bin s = v - Blo >= 0, sign +1 / -1) applied to the own counters.
  auto band_apply = [&](int s, int sign) {
    if (s >= BAND_INV - Blo) { NT -= sign; return; }  // invalid sample: only the valid count is corrected
    TB += sign;
    if (s < K) {
      myL0[s * CS] = (uint16_t)(myL0[s * CS] + sign);
      myL1[(s >> 3) * CS] = (uint16_t)(myL1[(s >> 3) * CS] + sign);
    }
  };

  // ---- (re)build the counters of the window centred on day-of-year index d ----
  // returns false when the tile's thresholds do not fit one band (tile goes to the fail list)
  auto for_window = [&](int d, auto&& fn) {
    for (int k = -half; k <= half; ++k) {
      const int dd = ((d + k) % NDOY + NDOY) % NDOY;
      const int b0 = __ldg(&p.doy_ptr[dd]), b1 = __ldg(&p.doy_ptr[dd + 1]);
      for (int j = b0; j < b1; j += 8) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (j + u < b1) ? load_row(j + u) : BAND_INV;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (v[u] != BAND_INV) fn(v[u]);
      }
    }
  };
  auto rebuild = [&](int d) -> bool {
    // 1. coarse pass: 8-bin block counts over the full range, own window
    for (int i = tid; i < (K + KB) * CS; i += OY * 32) L0[i] = 0;  // L0 and L1 are contiguous
    if (tid < 3) s_misc[tid] = tid == 1 ? 0x7fffffff : (tid == 2 ? -1 : 0);
    __syncthreads();
    NT = 0;
    if (own_valid) for_window(d, [&](int v) { ++NT; if (K < 8 * nblk) { myL0[(v >> 3) * CS] = (uint16_t)(myL0[(v >> 3) * CS] + 1); } });
    dead = NT == 0;
    NTr[tid] = (uint16_t)NT;
    __syncthreads();
    if (K >= 8 * nblk) {
      Blo = 0;  // the band covers every bin: no coarse pass, no violation possible
    } else {
      if (warp < TY) {
        const int ntot = pooled_row<P>(NTr + warp * 32 + lane, lane);
        const bool livet = target_live && !masked && ntot > 0;
        const int kk = (int)floor(__dmul_rn(p.q, (double)ntot));
        int run = 0, jb = -1;
        for (int j = 0; j < nblk; ++j) {
          const int pj = pooled_row<P>(L0 + j * CS + warp * 32 + lane, lane);
          if (jb < 0) { if (run + pj > kk) jb = j; else run += pj; }
          if (__all_sync(0xffffffffu, jb >= 0 || !livet)) break;
        }
        if (livet) {
          if (jb < 0) jb = nblk + K;  // rank beyond the last bin (q = 1): not representable in a band
          atomicMin(&s_misc[1], jb);
          atomicMax(&s_misc[2], jb);
        }
      }
      __syncthreads();
      const int jmin = s_misc[1], jmax = s_misc[2];
      if (jmax >= 0) {
        int lo = 8 * jmin - p.margin;
        lo = lo < 0 ? 0 : (lo & ~7);
        if (8 * jmax + 7 >= lo + K || p.force_fail) return false;
        Blo = lo;
      } else {
        Blo = 0;
        if (p.force_fail) return false;
      }
      for (int i = tid; i < (K + KB) * CS; i += OY * 32) L0[i] = 0;
      __syncthreads();
    }
    // 2. band counters
    TB = 0;
    if (!dead) for_window(d, [&](int v) { if (v >= Blo) { const int n0 = NT; band_apply(v - Blo, +1); NT = n0; } });
    TBr[tid] = (uint16_t)TB;
    __syncthreads();
    return true;
  };

  // ---- entering samples of `step` into registers (needs stage(step) + a barrier before) ----
  auto prefetch = [&](int step) {
    if (!own_valid) return;
    const long long* oe = s_off + ((step & 1) * 2 + 1) * BAND_CAP;
    const int ne = NYC ? NYC : s_cnt[((step & 1) * 2 + 1) * 2 + 1] - s_cnt[((step & 1) * 2 + 1) * 2];
#pragma unroll
    for (int u = 0; u < BAND_PRE; ++u) pre[u] = (u < ne) ? load_at(oe[u]) : -1;
  };

  // ---- advance the own window by one day of year (needs prefetch(step) before) ----
  auto advance = [&](int step) {
    if (!own_valid) return;
    const int par = step & 1;
    const long long* ol = s_off + (par * 2 + 0) * BAND_CAP;
    const int bl0 = s_cnt[(par * 2 + 0) * 2], bl1 = s_cnt[(par * 2 + 0) * 2 + 1];
    const int be0 = s_cnt[(par * 2 + 1) * 2], be1 = s_cnt[(par * 2 + 1) * 2 + 1];
    const int nl = NYC ? NYC : bl1 - bl0, ne = NYC ? NYC : be1 - be0;
    constexpr bool SHORT = NYC > 0 && NYC <= BAND_PRE;  // every row of a day fits the prefetch registers / the staged offsets
    // Event list: ev[n_ev] with n_ev a multiple of CS (32-bit index arithmetic, one slot per sample that
    // matters).  Entering and leaving samples are flushed separately, so a list carries one sign.
    int n_ev = 0;
    constexpr int EV_FULL = (BAND_LCAP - 15) * CS;  // a batch of 13 (+ 1 scratch slot) still fits below this
    auto flush = [&](int sign) {
      for (int j = 0; j < n_ev; j += CS) band_apply((int)ev[j], sign);
      n_ev = 0;
    };
    // branch-free scan: the slot is overwritten unless the sample matters (inside / above the band, or invalid)
    auto scan = [&](int v) {
      const int s = v - Blo;
      ev[n_ev] = (uint16_t)s;
      n_ev += (s >= 0) ? CS : 0;
    };
    bool skip_leaving = false;
    if (dead) {  // window without a valid sample: stays so unless a valid sample enters
      int all = BAND_INV;
#pragma unroll
      for (int u = 0; u < BAND_PRE; ++u) all &= pre[u];  // missing slots are -1: neutral
      if (!SHORT) for (int b = be0 + BAND_PRE; b < be1; ++b) all &= load_row(b);
      if (all == BAND_INV) return;
      dead = false;
      skip_leaving = true;  // every leaving sample is invalid: the valid count does not change
    }
    // entering samples (prefetched)
#pragma unroll
    for (int u = 0; u < BAND_PRE; ++u) {
      if (u == 13 && n_ev > EV_FULL) flush(+1);
      if (u < ne) scan(pre[u]);  // warp-uniform guard
    }
    if (!SHORT) for (int b = be0 + BAND_PRE; b < be1; ++b) { if (n_ev > EV_FULL) flush(+1); scan(load_row(b)); }
    flush(+1);
    NT += ne;
    if (!skip_leaving) {
      NT -= nl;
      constexpr int BATCH = 13;
      const int nlc = min(nl, BAND_CAP);
      for (int u0 = 0; u0 < nlc; u0 += BATCH) {
        int vl[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) vl[u] = (u0 + u < nlc) ? load_at(ol[u0 + u]) : 0;
        if (n_ev > EV_FULL) flush(-1);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) if (u0 + u < nlc) scan(vl[u]);
      }
      if (!SHORT) for (int a = bl0 + BAND_CAP; a < bl1; ++a) { if (n_ev > EV_FULL) flush(-1); scan(load_row(a)); }
      flush(-1);
    }
    NTr[tid] = (uint16_t)NT;
    TBr[tid] = (uint16_t)TB;
  };

  // ---- query of one day of year ----
  auto query = [&](int d) {
    constexpr int rowbase = 0;
    const int rowoff = warp * 32 + lane + rowbase;
    const int ntot = pooled_row<P>(NTr + rowoff, lane);
    const int tb = pooled_row<P>(TBr + rowoff, lane);
    const bool livet = target_live && !masked && ntot > 0;
    const double pos = __dmul_rn(p.q, (double)ntot);  // detect.py:2516
    const int kk = (int)floor(pos);                   // cum > pos  <=>  cum >= kk + 1
    const int below = ntot - tb;
    bool viol = livet && below > kk;                  // quantile bin lies below the band
    // block scan
    int run = below, jb = -1;
    bool done = !livet || viol;
    for (int j = 0; j < KB; ++j) {
      if (__all_sync(0xffffffffu, done)) break;
      const int pj = pooled_row<P>(L1 + j * CS + rowoff, lane);
      if (!done) {
        if (run + pj > kk) { jb = j; done = true; }
        else run += pj;
      }
    }
    int iu = -1, cl = 0, h = 0;
    if (livet && !viol && jb < 0) {
      // not found inside the band: either the band is too low, or (band = full range) the rank lies
      // beyond the last bin and the reference clips iu to nb - 1 (detect.py:2530-2532)
      if (K >= 8 * nblk) { iu = nb - 1; }
      else viol = true;
    }
    // bin scan inside the blocks that hold a lane's quantile
    const unsigned want = __ballot_sync(0xffffffffu, jb >= 0);
    if (want) {
      int jlo = jb >= 0 ? jb : KB, jhi = jb;
      for (int o = 16; o; o >>= 1) {
        jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
        jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
      }
      for (int j = jlo; j <= jhi; ++j) {
        if (!__any_sync(0xffffffffu, jb == j)) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int pb = pooled_row<P>(L0 + (8 * j + b) * CS + rowoff, lane);
          if (jb == j && iu < 0) {
            if (run + pb > kk) { iu = Blo + 8 * j + b; cl = run; h = pb; }
            else run += pb;
          }
        }
      }
    }
    if (iu == nb - 1 && jb < 0 && livet && !viol) {  // clipped rank (q = 1): cl = cum[nb - 2], h = hist[nb - 1]
      h = pooled_row<P>(L0 + (nb - 1 - Blo) * CS + rowoff, lane);
      cl = ntot - h;
    }
    if (__any_sync(0xffffffffu, viol)) {
      if (lane == 0) s_misc[0] = 1;
      return;
    }
    float res = CUDART_NAN_F;
    if (livet) {
      if (iu == 0) {
        res = __ldg(&p.centers[0]);  // detect.py:2557
      } else {
        const float bl = __ldg(&p.centers[iu - 1]), bu = __ldg(&p.centers[iu]);
        const double frac = (h > 0) ? __ddiv_rn(pos - (double)cl, (double)h) : 0.5;       // detect.py:2545-2547
        res = (float)__dadd_rn((double)bl, __dmul_rn(frac, (double)__fsub_rn(bu, bl)));   // detect.py:2550 (no FMA)
      }
      vmin = fminf(vmin, res);
      vmax = fmaxf(vmax, res);
      if (res < p.lower_bound) res = p.lower_bound;  // detect.py:2722-2732
    }
    if (target_live) p.thr[(int64_t)d * N + ty_g * nx + tx_g] = res;
  };

  auto give_up = [&]() {
    if (tid == 0) {
      const int i = atomicAdd(&p.fail_list[0], 1);
      p.fail_list[1 + 2 * i] = (int)y0;
      p.fail_list[2 + 2 * i] = (int)x0;
    }
  };

  if (!rebuild(0)) { give_up(); return; }
  stage(1);
  __syncthreads();
  prefetch(1);
  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      if (d + 1 < NDOY) stage(d + 1);
      advance(d);
      __syncthreads();
      if (d + 1 < NDOY) prefetch(d + 1);  // in flight while the queries run
    }
    if (warp < TY) query(d);
    __syncthreads();
    if (s_misc[0]) {  // a threshold left the band: re-centre the band on the current window and redo the day
      __syncthreads();
      if (!rebuild(d)) { give_up(); return; }
      if (warp < TY) query(d);
      __syncthreads();
      if (s_misc[0]) { give_up(); return; }
    }
  }
  if (warp < TY && p.stats) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) {
      if (vmin != CUDART_INF_F) atomic_min_f(&p.stats[0], vmin);
      if (vmax != -CUDART_INF_F) atomic_max_f(&p.stats[1], vmax);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Ring variant of the band kernel for the DAY-OF-YEAR-MAJOR bin array (slot s = doy * NY + year): every sample is read
// from global memory ONCE.
//   * the NY rows of the day of year that enters the window are consecutive rows of the bin array: one 2-D TMA box
//     {32 columns, NY slots} (cp.async.bulk.tensor.2d) per own row of the tile stages them in shared memory behind one
//     mbarrier while the tile answers the queries of the previous step (tiles on the longitude seam, or bin arrays whose
//     row pitch is not a multiple of 16 bytes, fill the same stage with plain loads),
//   * a thread classifies its NY entering samples against the band and keeps the few that matter (inside / above the
//     band, or invalid) as 1-byte codes in a per-gridpoint ring of the window's w days: when that day leaves the window
//     w steps later, the thread replays its list with the opposite sign instead of loading and classifying the rows
//     again (the second touch was 55 GB of DRAM traffic and half of the scan instructions of the first band kernel).
// Counters, pooling, queries, re-centring and the fail list are those of hobday_band_kernel.
// ---------------------------------------------------------------------------------------------------------
struct RingParams {
  BandParams b;      // bins = day-of-year-major array, pitch = its row pitch; doy_ptr / doy_rows unused
  int NY;            // slots (years) per day of year
  int use_tma;
  int dbg;           // debug knob pool_dbg (bit 0: no drain before giving up; bit 1: land the staged day before a re-centre)
  volatile int* dbg_ptr;  // debug: page-locked HOST buffer, 8 ints per tile (progress markers that survive a device fault)
};

constexpr int RING_ABOVE = 254, RING_INVALID = 255, RING_ALL_INVALID = 255;
constexpr int RING_BW = 40;  // staged columns per own row: the 32 own columns inside a box that starts on a 16-byte boundary

template <int P, int K, int OY>
__global__ void __launch_bounds__(OY * 32, 1) hobday_ring_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                 const RingParams rp) {
  const BandParams& p = rp.b;
  constexpr int KB = K / 8;
  constexpr int TY = OY - 2 * P, TX = 32 - 2 * P, CS = OY * 32;
  extern __shared__ __align__(1024) unsigned char smem_ring[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int NY = rp.NY, w = p.w, half = p.w / 2;
  // carve-up: TMA stage first (1024-byte aligned), then counters, then the ring
  const int NYP = (NY + 7) & ~7;  // slots per own row of the stage: every own row (NYP x 80 bytes) starts on a 128-byte boundary (TMA)
  // carve-up: mbarrier + flags in the first 128 bytes, then the TMA stage (128-byte aligned), counters, ring
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem_ring);
  int* const s_misc = reinterpret_cast<int*>(smem_ring + 16);                 // [0] violation flag, [1] min blk, [2] max blk
  uint16_t* const stage = reinterpret_cast<uint16_t*>(smem_ring + 128);       // [OY][NYP][RING_BW]
  const size_t stage_bytes = (size_t)NYP * OY * RING_BW * 2;
  uint16_t* const L0 = reinterpret_cast<uint16_t*>(smem_ring + 128 + stage_bytes);  // [K][CS]; the coarse pass reuses it as [nblk][CS]
  uint16_t* const L1 = L0 + K * CS;                                           // [KB][CS]
  uint16_t* const NTr = L1 + KB * CS;                                         // [CS] valid samples in the own window
  uint16_t* const TBr = NTr + CS;                                             // [CS] of which >= Blo
  uint8_t* const ring = reinterpret_cast<uint8_t*>(TBr + CS);                 // [w][NY][CS] codes of the samples that matter
  uint8_t* const rlen = ring + (size_t)w * NY * CS;                           // [w][CS] list lengths (255: every sample invalid)

  const int nb = p.nb;
  const int nblk = (nb + 7) >> 3;
  const int64_t nx = p.nx, ny = p.ny, N = nx * ny;
  int64_t y0 = (int64_t)blockIdx.y * TY, x0 = (int64_t)blockIdx.x * TX;

  // ---- own gridpoint of this thread (update role) ----
  const int64_t gy = y0 - P + warp;
  int64_t gx = (x0 - P + lane) % nx;
  if (gx < 0) gx += nx;
  const bool own_valid = gy >= 0 && gy < ny;
  const uint16_t* const col = p.bins + (own_valid ? gy * nx + gx : 0);
  // TMA boxes start on a 16-byte boundary of global memory (8 codes): the box of an own row is the RING_BW columns from
  // xs0 = (x0 - P) rounded down to a multiple of 8, and must not straddle the longitude seam
  const int64_t xs0 = (x0 - P) & ~(int64_t)7;
  const bool tma = rp.use_tma && x0 - P >= 0 && xs0 + RING_BW <= nx;
  const int xoff = tma ? (int)(x0 - P - xs0) : 0;  // column of lane 0 inside a staged row
  uint16_t* const myL0 = L0 + tid;
  uint16_t* const myL1 = L1 + tid;
  int NT = 0, TB = 0, Blo = 0;
  bool dead = !own_valid;  // no valid sample in the own window: nothing to keep up to date

  // ---- target of this lane (query role: warps 0..TY-1, lanes P..31-P) ----
  const int64_t ty_g = y0 + warp, tx_g = x0 + lane - P;
  const bool target_live = warp < TY && lane >= P && lane < 32 - P && ty_g < ny && tx_g < nx;
  bool masked = true;
  if (target_live) {
    const float a0 = p.anom_row0[ty_g * nx + tx_g];
    masked = a0 != a0;  // detect.py:2704-2705
  }
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  uint32_t phase = 0;

  auto slot_of = [&](int step_doy) { return ((step_doy % w) + w) % w; };  // ring slot of the day that entered at that step

  // One code applied to the own counters.
  auto apply = [&](int code, int sign) {
    if (code == RING_INVALID) { NT -= sign; return; }  // invalid sample: only the valid count is corrected
    TB += sign;
    if (code != RING_ABOVE) {
      const int s = code - 1;
      myL0[s * CS] = (uint16_t)(myL0[s * CS] + sign);
      myL1[(s >> 3) * CS] = (uint16_t)(myL1[(s >> 3) * CS] + sign);
    }
  };
  auto code_of = [&](int v) -> int {  // 0: below the band (valid, not stored)
    const int s = v - Blo;
    const int c = (v == BAND_INV) ? RING_INVALID : (s < K ? s + 1 : RING_ABOVE);
    return s < 0 ? 0 : c;
  };
  auto replay = [&](int slot, int sign) {
    const int n = rlen[slot * CS + tid];
    if (n == RING_ALL_INVALID) return;  // NY invalid samples: the valid count does not change
    NT += sign * NY;
    const uint8_t* r = ring + (size_t)slot * NY * CS + tid;
    for (int j = 0; j < n; ++j) apply((int)r[j * CS], sign);
  };

  // ---- day `dd` of the year (0-based) into ring slot `slot` from global memory (rebuilds) ----
  auto build_slot = [&](int dd, int slot) {
    const uint16_t* src = col + (int64_t)dd * NY * p.pitch;
    uint8_t* r = ring + (size_t)slot * NY * CS + tid;
    int n = 0, all = BAND_INV;
    for (int j = 0; j < NY; j += 5) {
      int v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) v[u] = (j + u < NY) ? (int)src[(int64_t)(j + u) * p.pitch] : BAND_INV;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (j + u >= NY) continue;
        all &= v[u];
        const int c = code_of(v[u]);
        r[n] = (uint8_t)c;
        n += c ? CS : 0;
      }
    }
    n /= CS;
    rlen[slot * CS + tid] = (uint8_t)((all == BAND_INV) ? RING_ALL_INVALID : n);
  };

  // ---- (re)build the counters and the ring of the window centred on step d ----
  // returns false when the tile's thresholds do not fit one band (tile goes to the fail list)
  auto rebuild = [&](int d) -> bool {
    // 1. coarse pass: 8-bin block counts over the full range, own window
    for (int i = tid; i < (K + KB) * CS; i += OY * 32) L0[i] = 0;  // L0 and L1 are contiguous
    if (tid < 3) s_misc[tid] = tid == 1 ? 0x7fffffff : (tid == 2 ? -1 : 0);
    __syncthreads();
    NT = 0;
    if (own_valid) {
      for (int k = -half; k <= half; ++k) {
        const int dd = ((d + k) % NDOY + NDOY) % NDOY;
        const uint16_t* src = col + (int64_t)dd * NY * p.pitch;
        for (int j = 0; j < NY; j += 5) {
          int v[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) v[u] = (j + u < NY) ? (int)src[(int64_t)(j + u) * p.pitch] : BAND_INV;
#pragma unroll
          for (int u = 0; u < 5; ++u)
            if (v[u] != BAND_INV) {
              ++NT;
              if (K < 8 * nblk) myL0[(v[u] >> 3) * CS] = (uint16_t)(myL0[(v[u] >> 3) * CS] + 1);
            }
        }
      }
    }
    dead = NT == 0;
    NTr[tid] = (uint16_t)NT;
    __syncthreads();
    if (K >= 8 * nblk) {
      Blo = 0;  // the band covers every bin: no coarse pass, no violation possible
    } else {
      if (warp < TY) {
        const int ntot = pooled_row<P>(NTr + warp * 32 + lane, lane);
        const bool livet = target_live && !masked && ntot > 0;
        const int kk = (int)floor(__dmul_rn(p.q, (double)ntot));
        int run = 0, jb = -1;
        for (int j = 0; j < nblk; ++j) {
          const int pj = pooled_row<P>(L0 + j * CS + warp * 32 + lane, lane);
          if (jb < 0) { if (run + pj > kk) jb = j; else run += pj; }
          if (__all_sync(0xffffffffu, jb >= 0 || !livet)) break;
        }
        if (livet) {
          if (jb < 0) jb = nblk + K;  // rank beyond the last bin (q = 1): not representable in a band
          atomicMin(&s_misc[1], jb);
          atomicMax(&s_misc[2], jb);
        }
      }
      __syncthreads();
      const int jmin = s_misc[1], jmax = s_misc[2];
      if (jmax >= 0) {
        int lo = 8 * jmin - p.margin;
        lo = lo < 0 ? 0 : (lo & ~7);
        if (8 * jmax + 7 >= lo + K || p.force_fail) return false;
        Blo = lo;
      } else {
        Blo = 0;
        if (p.force_fail) return false;
      }
      for (int i = tid; i < (K + KB) * CS; i += OY * 32) L0[i] = 0;
      __syncthreads();
    }
    // 2. ring lists and band counters of the window's days
    TB = 0;
    NT = 0;
    if (own_valid) {
      for (int k = -half; k <= half; ++k) {
        const int dd = ((d + k) % NDOY + NDOY) % NDOY;
        const int slot = slot_of(d + k);
        build_slot(dd, slot);
        replay(slot, +1);
      }
    }
    dead = NT == 0;
    NTr[tid] = (uint16_t)NT;
    TBr[tid] = (uint16_t)TB;
    __syncthreads();
    return true;
  };

  // ---- the rows of the day entering at `step` into the stage ----
  auto issue = [&](int step) {
    const int dd = (step + half) % NDOY;
    if (tma) {
      if (tid == 0) {  // one 2-D box {32 columns, NY slots} per own row of the grid, all on the same mbarrier
        fence_proxy_async();
        const int oy0 = max(0, (int)(P - y0)), oy1 = min(OY, (int)(ny + P - y0));
        mbar_expect_tx(bar, (uint32_t)(oy1 - oy0) * NY * (RING_BW * 2u));
        for (int oy = oy0; oy < oy1; ++oy)
          tma_load_2d(stage + (size_t)oy * NYP * RING_BW, &tmap, (int)((y0 - P + oy) * nx + xs0), dd * NY, bar);
      }
    } else if (own_valid) {
      const uint16_t* src = col + (int64_t)dd * NY * p.pitch;
      for (int j = 0; j < NY; j += 5) {
        uint16_t v[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) v[u] = (j + u < NY) ? src[(int64_t)(j + u) * p.pitch] : (uint16_t)0;
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if (j + u < NY) stage[(warp * NYP + j + u) * RING_BW + lane] = v[u];
      }
    }
  };

  // ---- advance the own window by one day of year: the day that entered w steps ago leaves, the staged day enters ----
  bool landed = false;  // the staged day has already been waited for
  auto advance = [&](int step) {
    if (tma && !landed) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    landed = false;
    if (!own_valid) return;
    const int slot = slot_of(step + half);
    if (!dead) replay(slot, -1);  // a dead window holds invalid samples only: nothing to take out
    // entering samples: classify, keep the ones that matter
    const uint16_t* sv = stage + (size_t)warp * NYP * RING_BW + xoff + lane;  // [OY][NYP][RING_BW]
    uint8_t* r = ring + (size_t)slot * NY * CS + tid;
    int n = 0, all = BAND_INV;
    for (int j = 0; j < NY; j += 5) {
      int v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) v[u] = (j + u < NY) ? (int)sv[(j + u) * RING_BW] : BAND_INV;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (j + u >= NY) continue;
        all &= v[u];
        const int c = code_of(v[u]);
        r[n] = (uint8_t)c;
        n += c ? CS : 0;
      }
    }
    n /= CS;
    if (all == BAND_INV) {
      rlen[slot * CS + tid] = (uint8_t)RING_ALL_INVALID;
    } else {
      rlen[slot * CS + tid] = (uint8_t)n;
      dead = false;
      NT += NY;
      for (int j = 0; j < n; ++j) apply((int)r[j * CS], +1);
    }
    NTr[tid] = (uint16_t)NT;
    TBr[tid] = (uint16_t)TB;
  };

  // ---- query of one day of year (as hobday_band_kernel) ----
  auto query = [&](int d) {
    const int rowoff = warp * 32 + lane;
    const int ntot = pooled_row<P>(NTr + rowoff, lane);
    const int tb = pooled_row<P>(TBr + rowoff, lane);
    const bool livet = target_live && !masked && ntot > 0;
    const double pos = __dmul_rn(p.q, (double)ntot);  // detect.py:2516
    const int kk = (int)floor(pos);                   // cum > pos  <=>  cum >= kk + 1
    const int below = ntot - tb;
    bool viol = livet && below > kk;                  // quantile bin lies below the band
    int run = below, jb = -1;
    bool done = !livet || viol;
    for (int j = 0; j < KB; ++j) {
      if (__all_sync(0xffffffffu, done)) break;
      const int pj = pooled_row<P>(L1 + j * CS + rowoff, lane);
      if (!done) {
        if (run + pj > kk) { jb = j; done = true; }
        else run += pj;
      }
    }
    int iu = -1, cl = 0, h = 0;
    if (livet && !viol && jb < 0) {
      // not found inside the band: either the band is too low, or (band = full range) the rank lies
      // beyond the last bin and the reference clips iu to nb - 1 (detect.py:2530-2532)
      if (K >= 8 * nblk) { iu = nb - 1; }
      else viol = true;
    }
    const unsigned want = __ballot_sync(0xffffffffu, jb >= 0);
    if (want) {
      int jlo = jb >= 0 ? jb : KB, jhi = jb;
      for (int o = 16; o; o >>= 1) {
        jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
        jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
      }
      for (int j = jlo; j <= jhi; ++j) {
        if (!__any_sync(0xffffffffu, jb == j)) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int pb = pooled_row<P>(L0 + (8 * j + b) * CS + rowoff, lane);
          if (jb == j && iu < 0) {
            if (run + pb > kk) { iu = Blo + 8 * j + b; cl = run; h = pb; }
            else run += pb;
          }
        }
      }
    }
    if (iu == nb - 1 && jb < 0 && livet && !viol) {  // clipped rank (q = 1): cl = cum[nb - 2], h = hist[nb - 1]
      h = pooled_row<P>(L0 + (nb - 1 - Blo) * CS + rowoff, lane);
      cl = ntot - h;
    }
    if (__any_sync(0xffffffffu, viol)) {
      if (lane == 0) s_misc[0] = 1;
      return;
    }
    float res = CUDART_NAN_F;
    if (livet) {
      if (iu == 0) {
        res = __ldg(&p.centers[0]);  // detect.py:2557
      } else {
        const float bl = __ldg(&p.centers[iu - 1]), bu = __ldg(&p.centers[iu]);
        const double frac = (h > 0) ? __ddiv_rn(pos - (double)cl, (double)h) : 0.5;       // detect.py:2545-2547
        res = (float)__dadd_rn((double)bl, __dmul_rn(frac, (double)__fsub_rn(bu, bl)));   // detect.py:2550 (no FMA)
      }
      vmin = fminf(vmin, res);
      vmax = fmaxf(vmax, res);
      if (res < p.lower_bound) res = p.lower_bound;  // detect.py:2722-2732
    }
    if (target_live) p.thr[(int64_t)d * N + ty_g * nx + tx_g] = res;
  };

  auto give_up = [&]() {
    if (tid == 0) {
      const int i = atomicAdd(&p.fail_list[0], 1);
      p.fail_list[1 + 2 * i] = (int)y0;
      p.fail_list[2 + 2 * i] = (int)x0;
    }
  };
  // a TMA load still in flight must land before the CTA exits
  auto drain = [&](bool pending) {
    if (tma && pending && !landed && !(rp.dbg & 1)) mbar_wait(bar, phase);
  };

  volatile int* const mark = rp.dbg_ptr ? rp.dbg_ptr + 8 * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  auto note = [&](int k, int v) {
    if (mark && tid == 0) { mark[k] = v; __threadfence_system(); }
  };
  note(0, 1 + (tma ? 1 : 0));
  if (!rebuild(0)) { note(1, -1); give_up(); return; }
  note(1, 1);
  issue(1);
  note(2, 1);
  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      note(3, d);
      advance(d);
      note(4, d);
      __syncthreads();                 // every thread is done with the stage and the counters are up to date
      if (d + 1 < NDOY) issue(d + 1);  // in flight while the queries run
    }
    if (warp < TY) query(d);
    __syncthreads();
    note(5, d);
    if (s_misc[0]) {  // a threshold left the band: re-centre the band on the current window and redo the day
      __syncthreads();
      if ((rp.dbg & 2) && tma && d + 1 < NDOY && !landed) {
        mbar_wait(bar, phase);
        phase ^= 1u;
        landed = true;
      }
      note(6, d);
      if (!rebuild(d)) { note(7, -d - 1); drain(d + 1 < NDOY); give_up(); return; }
      if (warp < TY) query(d);
      __syncthreads();
      if (s_misc[0]) { note(7, -1000 - d); drain(d + 1 < NDOY); give_up(); return; }
      note(7, d);
    }
  }
  if (warp < TY && p.stats) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) {
      if (vmin != CUDART_INF_F) atomic_min_f(&p.stats[0], vmin);
      if (vmax != -CUDART_INF_F) atomic_max_f(&p.stats[1], vmax);
    }
  }
}

// np.digitize(a, edges) - 1 (detect.py:2622-2631) into the DAY-OF-YEAR-MAJOR bin array the threshold and compare
// kernels walk: slot s = doy * NY + (index of the year among the output years) holds the codes of input row
// slot_row[s], or the invalid code when that (day, year) has no row (slot_row[s] < 0).  Invalid samples (NaN or
// a >= last edge) are coded BIN_INV, which sorts above every band.  Four gridpoints x four slots in flight per thread.
__global__ void __launch_bounds__(256) digitize_doy_kernel(const float* __restrict__ a, int64_t N, int64_t pitch,
                                                           const int32_t* __restrict__ slot_row, int64_t n_slots,
                                                           const float* __restrict__ edges, int n_edges,
                                                           uint16_t* __restrict__ bins, int64_t bins_pitch,
                                                           int slots_per_block) {
  extern __shared__ float s_edges[];
  DigTable dig;
  dig.init_cta(edges, n_edges, s_edges);
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // four gridpoints per thread
  if (c >= N) return;
  const bool quad = c + 3 < N && (pitch & 3) == 0 && (bins_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bins) & 7) == 0;
  const int64_t s0 = (int64_t)blockIdx.y * slots_per_block, s1 = min(n_slots, s0 + slots_per_block);
  const uint32_t inv2 = (uint32_t)BIN_INV | ((uint32_t)BIN_INV << 16);
  if (quad) {
    for (int64_t s = s0; s < s1; s += 4) {
      float4 v[4];
      int64_t row[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        row[u] = (s + u < s1) ? (int64_t)__ldg(&slot_row[s + u]) : -2;
        if (row[u] >= 0) v[u] = __ldcs(reinterpret_cast<const float4*>(a + row[u] * pitch + c));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (row[u] == -2) continue;
        uint2 o = make_uint2(inv2, inv2);
        if (row[u] >= 0) {
          o.x = dig(v[u].x) | (dig(v[u].y) << 16);
          o.y = dig(v[u].z) | (dig(v[u].w) << 16);
        }
        *reinterpret_cast<uint2*>(bins + (s + u) * bins_pitch + c) = o;
      }
    }
  } else {
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t row = __ldg(&slot_row[s]);
      for (int k = 0; k < 4 && c + k < N; ++k)
        bins[s * bins_pitch + c + k] = row >= 0 ? (uint16_t)dig(a[row * pitch + c + k]) : (uint16_t)BIN_INV;
    }
  }
}

// defined in thresholds.cu: the full-range tile kernel over a list of failed band tiles
int launch_pool_tile_list(const uint16_t* bins, int64_t ny, int64_t nx, int64_t pitch, const int32_t* doy_ptr,
                          const int32_t* doy_rows, const float* centers, int nb, int w, int ws, double q,
                          const float* anom_row0, float lower_bound, float* thr, float* stats,
                          const int32_t* fail_list, int max_tiles, int band_ty, int band_tx, cudaStream_t st);

__global__ void init_band_kernel(float* stats, int32_t* list_a, int32_t* list_b) {
  if (stats) { stats[0] = CUDART_INF_F; stats[1] = -CUDART_INF_F; }
  list_a[0] = 0;
  list_b[0] = 0;
}

}  // namespace marex

using namespace marex;

extern "C" int marex_digitize_doy_f32(const float* a, int64_t N, int64_t pitch, const int32_t* slot_row, int64_t n_slots,
                                      const float* edges, int32_t n_edges, uint16_t* bins, int64_t bins_pitch,
                                      void* stream) {
  MAREX_REQUIRE(a && slot_row && edges && bins, "null pointer");
  MAREX_REQUIRE(N > 0 && n_slots > 0 && pitch >= N && bins_pitch >= N, "bad shape");
  MAREX_REQUIRE(n_edges >= 3 && n_edges <= 4096, "n_edges must be in 3..4096");
  const int threads = 256;
  const int64_t bx = ((N + 3) / 4 + threads - 1) / threads;
  int64_t by = (16LL * sm_count() + bx - 1) / bx;
  by = by < 1 ? 1 : (by > n_slots ? n_slots : by);
  if (by > 65535) by = 65535;
  int slots_per_block = (int)((n_slots + by - 1) / by);
  slots_per_block = (slots_per_block + 3) & ~3;
  by = (n_slots + slots_per_block - 1) / slots_per_block;
  digitize_doy_kernel<<<dim3((unsigned)bx, (unsigned)by), threads, n_edges * sizeof(float), (cudaStream_t)stream>>>(
      a, N, pitch, slot_row, n_slots, edges, n_edges, bins, bins_pitch, slots_per_block);
  MAREX_LAUNCH_CHECK("digitize_doy_kernel");
  return MAREX_OK;
}

extern "C" int64_t marex_hobday_pooled_workspace_bytes(int64_t ny, int64_t nx) {
  // two tile lists (count + (y0, x0) pairs) for the band retries
  const int64_t tiles = (ny + 1) * ((nx + 27) / 28 + 1);  // generous upper bound on band tiles
  return 2 * (2 * tiles + 8) * 4 + 512;
}

extern "C" int marex_hobday_thresholds_pooled_bins(const uint16_t* bins, int64_t NY, int64_t ny, int64_t nx,
                                                   int64_t bpitch, const int32_t* doy_ptr, const int32_t* doy_rows,
                                                   const float* centers, int32_t nb, int32_t w, int32_t ws, double q,
                                                   const float* anom_row0, float lower_bound, float* thr, float* stats,
                                                   void* workspace, int64_t workspace_bytes, void* stream) {
  MAREX_REQUIRE(bins && doy_ptr && doy_rows && centers && anom_row0 && thr && workspace, "null pointer");
  MAREX_REQUIRE(NY > 0 && ny > 0 && nx > 0 && bpitch >= ny * nx, "bad shape");
  MAREX_REQUIRE(nb >= 2 && nb <= 1024, "nb must be in 2..1024");
  MAREX_REQUIRE(w >= 3 && w <= 365 && (w & 1), "window_days_hobday must be odd and in 3..365");
  MAREX_REQUIRE(ws == 3 || ws == 5 || ws == 7, "window_spatial_hobday must be 3, 5 or 7 for the pooled kernel");
  MAREX_REQUIRE((int64_t)w * NY <= 65535, "too many rows per day-of-year window for 16-bit counters");
  MAREX_REQUIRE(nx >= 32 && ny <= 65535 * 4, "grid too small or too tall for the band tiles");
  MAREX_REQUIRE(workspace_bytes >= marex_hobday_pooled_workspace_bytes(ny, nx), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = ny * nx;
  (void)N;
  int32_t* fail_list = reinterpret_cast<int32_t*>(workspace);
  const float* anom = anom_row0;
  const int P = ws / 2;
  const int env_k = (int)tune_get("pool_k", 0);
  const int env_ty = (int)tune_get("pool_ty", 0);
  MAREX_REQUIRE(env_k == 0 || env_k == 64 || env_k == 128, "MAREX_POOL_K must be 64 or 128");
  const int TX = 32 - 2 * P;
  // Ring kernel (every sample read once, TMA-staged): 64-bin band, OY = 14 own rows (one 448-thread tile per SM) or 10
  // when the ring of w * NY codes per gridpoint does not fit; otherwise the first band kernel.
  int ring_oy = 0;
  size_t ring_smem = 0;
  // measured on B200 (0.25 deg, 25 output years, w = 11; profiles/r02_band_kernels.json): the ring kernel halves the DRAM
  // traffic (34.7 GB against 75 GB) but one 448-thread tile per SM runs 65.1 ms against 50.5 ms for two 512-thread
  // tiles of the first band kernel, so it is opt-in (marex_tune("pool_ring", 1) / MAREX_POOL_RING=1)
  if (tune_get("pool_ring", 0) && env_k != 128 && !env_ty && nb <= 8 * 64 && NY <= 254) {
    for (int oy : {14, 10}) {
      if (oy - 2 * P < 1) continue;
      const size_t cs = (size_t)oy * 32;
      const size_t need = (size_t)((NY + 7) & ~7LL) * oy * RING_BW * 2 + (size_t)(64 + 8 + 2) * cs * 2 + (size_t)w * NY * cs +
                          (((size_t)w * cs + 15) / 16) * 16 + 128;
      if (need <= 227 * 1024) { ring_oy = oy; ring_smem = need; break; }
    }
  }
  // first band kernel: 12 target rows (16 x 32 own gridpoints, two tiles per SM at 64 registers); pool_ty < 8 selects
  // 3-row tiles (exercised by the tests).  After the ring kernel it retries ring tiles (at most 12 rows) from their origin.
  const int TY = (env_ty && env_ty < 8 && !ring_oy) ? 3 : 12;
  const int OY = TY + 2 * P;
  const int ring_ty = ring_oy ? ring_oy - 2 * P : TY;
  const dim3 grid_all((unsigned)((nx + TX - 1) / TX), (unsigned)((ny + TY - 1) / TY));
  const dim3 grid_ring((unsigned)((nx + TX - 1) / TX), (unsigned)((ny + ring_ty - 1) / ring_ty));
  const int max_tiles = (int)std::max(grid_all.x * grid_all.y, grid_ring.x * grid_ring.y);
  int32_t* list_b = fail_list + 2 * ((ny + 1) * ((nx + 27) / 28 + 1)) + 8;
  init_band_kernel<<<1, 1, 0, st>>>(stats, fail_list, list_b);
  MAREX_LAUNCH_CHECK("init_band_kernel");
  BandParams bp;
  bp.bins = bins; bp.ny = ny; bp.nx = nx; bp.pitch = bpitch;
  bp.doy_ptr = doy_ptr; bp.doy_rows = doy_rows; bp.centers = centers;
  bp.nb = nb; bp.w = w; bp.q = q;
  bp.margin = (int)tune_get("pool_margin", 8);
  bp.anom_row0 = anom; bp.lower_bound = lower_bound; bp.thr = thr; bp.stats = stats;
  bp.force_fail = (int)tune_get("pool_force_fail", 0);
  bp.tile_list = nullptr;
  auto launch_ring = [&](int32_t* fails) -> int {
    RingParams rp;
    rp.b = bp;
    rp.b.fail_list = fails;
    rp.NY = (int)NY;
    rp.dbg = (int)tune_get("pool_dbg", 0);
    rp.dbg_ptr = reinterpret_cast<volatile int*>((uintptr_t)tune_get("pool_dbg_ptr", 0));
    rp.use_tma = (nx % 8) == 0 && (bpitch % 8) == 0 && (reinterpret_cast<uintptr_t>(bins) % 16) == 0 &&
                 tune_get("pool_tma", 1) && NDOY * NY < (1LL << 31) && ny * nx < (1LL << 31);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (rp.use_tma) {
      const int rc = make_tmap_2d(&tmap, bins, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, NDOY * NY, ny * nx, bpitch, (int)NY, RING_BW);
      if (rc) return rc;
    }
#define MAREX_RING(PP, OO)                                                                                       \
  do {                                                                                                           \
    cudaError_t e = cudaFuncSetAttribute(hobday_ring_kernel<PP, 64, OO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)ring_smem);                                                        \
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(hobday_ring)");                              \
    hobday_ring_kernel<PP, 64, OO><<<grid_ring, OO * 32, ring_smem, st>>>(tmap, rp);                             \
  } while (0)
    if (ring_oy == 14) { if (P == 1) MAREX_RING(1, 14); else if (P == 2) MAREX_RING(2, 14); else MAREX_RING(3, 14); }
    else { if (P == 1) MAREX_RING(1, 10); else if (P == 2) MAREX_RING(2, 10); else MAREX_RING(3, 10); }
#undef MAREX_RING
    MAREX_LAUNCH_CHECK("hobday_ring_kernel");
    return MAREX_OK;
  };
  // One launch of the band kernel with K band bins over all tiles (tiles == nullptr) or over a list.
  auto launch_band = [&](int K, const int32_t* tiles, int32_t* fails) -> int {
    if (nb > 8 * K) return fail(MAREX_ERR_UNSUPPORTED, "nb too large for the coarse pass of this band width");
    const size_t smem = (size_t)(K + K / 8 + 2 + BAND_LCAP) * OY * 32 * 2 + 4 * BAND_CAP * 8 + (8 + 4) * sizeof(int);
    if (smem > 227 * 1024) return fail(MAREX_ERR_UNSUPPORTED, "band tile does not fit shared memory");
    bp.tile_list = tiles;
    bp.fail_list = fails;
    const dim3 grid = tiles ? dim3((unsigned)max_tiles, 1) : grid_all;
#define MAREX_BAND(PP, KK, TT)                                                                                 \
  do {                                                                                                         \
    cudaError_t e = cudaFuncSetAttribute(hobday_band_kernel<PP, KK, TT + 2 * PP>,                              \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(hobday_band)");                            \
    hobday_band_kernel<PP, KK, TT + 2 * PP><<<grid, (TT + 2 * PP) * 32, smem, st>>>(bp);                       \
  } while (0)
#define MAREX_BAND_K(PP)                                                                 \
  do {                                                                                   \
    if (K == 64) { if (TY == 12) MAREX_BAND(PP, 64, 12); else MAREX_BAND(PP, 64, 3); }   \
    else { if (TY == 12) MAREX_BAND(PP, 128, 12); else MAREX_BAND(PP, 128, 3); }         \
  } while (0)
#define MAREX_BAND_NY(NN)                                                                                      \
  do {                                                                                                         \
    cudaError_t e = cudaFuncSetAttribute(hobday_band_kernel<2, 64, 16, NN>,                                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(hobday_band)");                            \
    hobday_band_kernel<2, 64, 16, NN><<<grid, 16 * 32, smem, st>>>(bp);                                        \
  } while (0)
    const bool fold = P == 2 && K == 64 && TY == 12 && !tune_get("pool_generic", 0);
    if (fold && NY == 25) MAREX_BAND_NY(25);
    else if (fold && NY == 15) MAREX_BAND_NY(15);
    else if (P == 1) MAREX_BAND_K(1); else if (P == 2) MAREX_BAND_K(2); else MAREX_BAND_K(3);
#undef MAREX_BAND_NY
#undef MAREX_BAND_K
#undef MAREX_BAND
    MAREX_LAUNCH_CHECK("hobday_band_kernel");
    return MAREX_OK;
  };
  // 64-bin band (two tiles per SM) first; tiles whose thresholds spread wider retry with 128 bins;
  // what is left gets full-range counters.  MAREX_POOL_K pins a single width (tuning / tests).
  const int32_t* leftover = fail_list;
  if (ring_oy) {
    int rc = launch_ring(fail_list);
    if (rc) return rc;
    rc = launch_band(128, fail_list, list_b);
    if (rc) return rc;
    leftover = list_b;
  } else if (env_k) {
    const int rc = launch_band(env_k, nullptr, fail_list);
    if (rc) return rc;
  } else if (nb > 8 * 64) {
    const int rc = launch_band(128, nullptr, fail_list);
    if (rc) return rc;
  } else {
    int rc = launch_band(64, nullptr, fail_list);
    if (rc) return rc;
    rc = launch_band(128, fail_list, list_b);
    if (rc) return rc;
    leftover = list_b;
  }
  // tiles whose thresholds did not fit one band: full-range counters
  return launch_pool_tile_list(bins, ny, nx, bpitch, doy_ptr, doy_rows, centers, nb, w, ws, q, anom, lower_bound, thr,
                               stats, leftover, max_tiles, TY, TX, st);
}
