// Threshold kernels: Hobday histogram quantile (own-cell and ws x ws pooled),
// exact Hobday / global percentiles, global histogram quantile.
//
// Common shape: one warp per CTA, lane = gridpoint, a private histogram per lane laid out
// hist[bin][lane] in shared memory (bank = lane, conflict-free up to the 16-bit pairing).
// Counts are integers, so everything up to the final interpolation is bit-exact.
#include <cstdlib>

#include "common.cuh"
#include "exact_queue.cuh"

namespace marex {

// ---------------------------------------------------------------------------------------
// np.digitize(a, edges) - 1  (detect.py:2622-2631).  `edges` (shared memory) is ascending
// with edges[0] = -inf; the near-uniform spacing gives a first guess that the two loops fix
// up against the REAL float32 edge table, so counts match numpy bit for bit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int digitize_f32(float a, const float* __restrict__ edges, int n_edges, float e1,
                                            float inv_step) {
  if (a != a) return n_edges - 1;
  float g = floorf((a - e1) * inv_step) + 1.f;
  g = fminf(fmaxf(g, 0.f), (float)(n_edges - 1));
  int i = (int)g;
  while (i > 0 && a < edges[i]) --i;
  while (i < n_edges - 1 && a >= edges[i + 1]) ++i;
  return i;
}

// ---------------------------------------------------------------------------------------
// Approximate Hobday thresholds (detect.py:2562-2734 + 2465-2559).
//
// One lane = one target gridpoint.  The lane owns the histogram of ALL samples that feed its
// threshold: rows of the +-w/2 day-of-year window x the ws x ws neighbourhood (periodic in x,
// truncated in y).  Advancing the day of year removes one day's rows and adds another's, so the
// histogram, the total N and the running rank state (iu, cl = #samples in bins < iu) are all
// updated incrementally; the quantile bin is re-found by walking iu a few bins.
// ---------------------------------------------------------------------------------------
template <typename CT, bool POOLED>
__global__ void __launch_bounds__(32) hobday_hist_kernel(
    const uint16_t* __restrict__ bins, int64_t ny, int64_t nx, int64_t pitch, const int32_t* __restrict__ doy_ptr,
    const int32_t* __restrict__ doy_rows, const float* __restrict__ centers, int nb, int w, int ws, double q,
    const float* __restrict__ anom_row0, float lower_bound, float* __restrict__ thr, float* __restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  CT* hist = reinterpret_cast<CT*>(smem_raw);  // [nb][32]
  const int lane = threadIdx.x;
  const int64_t xg = (int64_t)blockIdx.x * 32 + lane;
  const int64_t y = blockIdx.y;
  const bool live = xg < nx;
  const int64_t xx = live ? xg : nx - 1;
  const int64_t N = ny * nx;
  for (int b = 0; b < nb; ++b) hist[b * 32 + lane] = 0;
  __syncwarp();
  const int half = w / 2, p = ws / 2;
  int ntot = 0, iu = 0, cl = 0;

  auto apply_doy = [&](int d, int sign) {
    const int b0 = __ldg(&doy_ptr[d]), b1 = __ldg(&doy_ptr[d + 1]);
    if (!POOLED) {  // batches of 8 independent loads (the kernel runs 13 warps per SM: latency is what it waits for)
      const uint16_t* colp = bins + y * nx + xx;
      auto count = [&](int v) {
        if (v < nb) {
          hist[v * 32 + lane] += (CT)sign;
          ntot += sign;
          if (v < iu) cl += sign;
        }
      };
      for (int j = b0; j < b1; j += 8) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (j + u < b1) ? (int)colp[(int64_t)__ldg(&doy_rows[j + u]) * pitch] : 0xFFFF;
#pragma unroll
        for (int u = 0; u < 8; ++u) count(v[u]);
      }
      return;
    }
    for (int j = b0; j < b1; ++j) {
      const uint16_t* row = bins + (int64_t)__ldg(&doy_rows[j]) * pitch;
      if (POOLED) {
        for (int dy = -p; dy <= p; ++dy) {
          const int64_t yy = y + dy;
          if (yy < 0 || yy >= ny) continue;
          const uint16_t* rowy = row + yy * nx;
          for (int dx = -p; dx <= p; ++dx) {
            int64_t xn = (xx + dx) % nx;
            if (xn < 0) xn += nx;
            const int v = rowy[xn];
            if (v < nb) {
              hist[v * 32 + lane] += (CT)sign;
              ntot += sign;
              if (v < iu) cl += sign;
            }
          }
        }
      } else {
        const int v = row[y * nx + xx];
        if (v < nb) {
          hist[v * 32 + lane] += (CT)sign;
          ntot += sign;
          if (v < iu) cl += sign;
        }
      }
    }
  };

  for (int k = -half; k <= half; ++k) apply_doy(((k % NDOY) + NDOY) % NDOY, +1);
  const bool masked = live ? (anom_row0[y * nx + xx] != anom_row0[y * nx + xx]) : true;
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;

  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      apply_doy((d - 1 - half + 2 * NDOY) % NDOY, -1);
      apply_doy((d + half) % NDOY, +1);
    }
    float res = CUDART_NAN_F;
    if (ntot > 0) {
      const double pos = __dmul_rn(q, (double)ntot);   // detect.py:2516
      const int kk = (int)floor(pos);                  // cum > pos  <=>  cum >= kk + 1
      while (cl > kk) { --iu; cl -= (int)hist[iu * 32 + lane]; }
      while (iu < nb - 1 && cl + (int)hist[iu * 32 + lane] <= kk) { cl += (int)hist[iu * 32 + lane]; ++iu; }
      if (iu == 0) {
        res = __ldg(&centers[0]);                      // detect.py:2557
      } else {
        const int h = (int)hist[iu * 32 + lane];
        const float bl = __ldg(&centers[iu - 1]), bu = __ldg(&centers[iu]);
        const double frac = (h > 0) ? __ddiv_rn(pos - (double)cl, (double)h) : 0.5;               // detect.py:2545-2547
        res = (float)__dadd_rn((double)bl, __dmul_rn(frac, (double)__fsub_rn(bu, bl)));           // detect.py:2550 (no FMA)
      }
    }
    if (masked) res = CUDART_NAN_F;                    // detect.py:2704-2705
    if (res == res) { vmin = fminf(vmin, res); vmax = fmaxf(vmax, res); }
    if (res < lower_bound) res = lower_bound;          // detect.py:2722-2732
    if (live) thr[(int64_t)d * N + y * nx + xg] = res;
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if (lane == 0 && stats) {
    if (vmin != CUDART_INF_F) atomic_min_f(&stats[0], vmin);
    if (vmax != -CUDART_INF_F) atomic_max_f(&stats[1], vmax);
  }
}

// ---------------------------------------------------------------------------------------
// Pooled (ws x ws) approximate Hobday thresholds, tiled.
//
// The per-lane version above re-reads every sample ws*ws times.  Here a CTA owns an OY x OX tile
// of gridpoints ("own" cells = targets + a ws/2 halo).  Each own cell keeps the histogram of ITS
// OWN +-w/2 day-of-year window in shared memory, updated incrementally by ONE thread per own
// cell (private 16-bit counters, no atomics; every sample is touched exactly twice: entering
// and leaving the window), at three resolutions: L0 = bins, L1 = blocks of 8 bins, L2 = groups
// of 64 bins.  A target's pooled cumulative count below bin B is the sum over its ws*ws
// neighbours of (B>>6) group + ((B>>3)&7) block + (B&7) bin counters: ~10 counter rows, each
// summed over the neighbours with compile-time offsets, two threads per target.  The quantile
// bin is tracked from the previous day of year and walked.  All counts are integers: bit-exact.
//
// Event phase detail: a leaving and an entering sample are applied as a PAIR -- if they fall in
// the same bin (block, group) the two updates cancel and nothing is touched, otherwise the two
// read-modify-writes hit different addresses and are issued back to back, which halves the
// dependent shared-memory latency chain.
// ---------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(768) hobday_pool_tile_kernel(
    const uint16_t* __restrict__ bins, int64_t ny, int64_t nx, int64_t pitch, const int32_t* __restrict__ doy_ptr,
    const int32_t* __restrict__ doy_rows, const float* __restrict__ centers, int nb, int w, double q,
    const float* __restrict__ anom_row0, float lower_bound, float* __restrict__ thr, float* __restrict__ stats,
    int OY, int OX, int CS, int CR, int LPT, const int32_t* __restrict__ fail_list, int sub_x, int sub_n, int band_ty,
    int band_tx) {
  constexpr int WS = 2 * P + 1, NN = WS * WS;
  // List mode (fail_list != nullptr): the grid is 1-D, sub_n CTAs per failed band tile (pool_band.cu);
  // CTA k recomputes sub-tile k % sub_n of band tile k / sub_n, whose origin is read from the list.
  if (fail_list && (int)(blockIdx.x / sub_n) >= __ldg(&fail_list[0])) return;
  constexpr int CAP = 32;   // slots of the staged per-step row table
  constexpr int MAXQ = (NN + 3) / 4;  // neighbours per query lane when LPT = 4 (LPT = 8 uses fewer)
  extern __shared__ unsigned char smem_raw[];
  const int C = OY * OX;   // own cells; counter rows have stride CS > C, column C is a permanently-zero dummy
  const int nb1 = (nb + 7) >> 3, nb2 = (nb + 63) >> 6;
  uint16_t* L0 = reinterpret_cast<uint16_t*>(smem_raw);  // [nb][CS]  bins >= 1 (bin 0 lives in Z0)
  uint16_t* L1 = L0 + (size_t)nb * CS;                    // [nb1][CS] blocks of 8 bins (block 0 without bin 0)
  uint16_t* L2 = L1 + (size_t)nb1 * CS;                   // [nb2][CS] groups of 64 bins (group 0 without bin 0)
  uint16_t* NT = L2 + (size_t)nb2 * CS;                   // [CS] samples in the own-cell window
  uint16_t* Z0 = NT + CS;                                 // [CS] of which in bin 0
  const int total16 = (nb + nb1 + nb2 + 2) * CS;
  for (int i = threadIdx.x; i < total16; i += blockDim.x) L0[i] = 0;

  const int TY = OY - 2 * P, TX = OX - 2 * P;             // targets per tile
  int64_t y0 = (int64_t)blockIdx.y * TY, x0 = (int64_t)blockIdx.x * TX;
  int64_t y_end = ny, x_end = nx;                         // targets live below these
  if (fail_list) {
    const int it = blockIdx.x / sub_n, sub = blockIdx.x % sub_n;
    const int64_t fy = __ldg(&fail_list[1 + 2 * it]), fx = __ldg(&fail_list[2 + 2 * it]);
    y0 = fy + (int64_t)(sub / sub_x) * TY;
    x0 = fx + (int64_t)(sub % sub_x) * TX;
    y_end = min(ny, fy + band_ty);
    x_end = min(nx, fx + band_tx);
  }
  const int64_t N = ny * nx;
  const int half = w / 2;

  // ---- event roles: three threads per own cell, one per counter level (disjoint arrays, no
  //      atomics, three short dependent chains instead of one long one, 3x the warps) ----
  const int role = threadIdx.x / CR;  // warp-uniform: CR is a multiple of 32
  const int oc = threadIdx.x % CR;
  const bool own_thread = role < 3 && oc < C;
  const int oy = own_thread ? oc / OX : 0, ox = own_thread ? oc % OX : 0;
  const int64_t gy = y0 - P + oy;
  int64_t gx = (x0 - P + ox) % nx;
  if (gx < 0) gx += nx;
  const bool own_valid = own_thread && gy >= 0 && gy < ny;
  const uint16_t* col = bins + (own_valid ? gy * nx + gx : 0);
  const int sh = 3 * role;                                 // key = bin >> sh
  uint16_t* colp = (role == 0 ? L0 : role == 1 ? L1 : L2) + (own_thread ? oc : C);
  int ntot_own = 0, n0_own = 0;                            // maintained by role 0

  // Bin 0 (every anomaly below -precision: about half of all samples) is only counted in Z0.
  auto apply = [&](int vl, int ve) {  // vl leaves the window, ve enters it (0xFFFF = none)
    if (role == 0) {
      ntot_own += (int)(ve < nb) - (int)(vl < nb);
      n0_own += (int)(ve == 0) - (int)(vl == 0);
    }
    const bool al = vl < nb && vl != 0, ae = ve < nb && ve != 0;
    const int kl = vl >> sh, ke = ve >> sh;
    if (al && ae) {
      if (kl != ke) {  // same counter: the two updates cancel
        const int a = colp[kl * CS], b = colp[ke * CS];
        colp[kl * CS] = (uint16_t)(a - 1);
        colp[ke * CS] = (uint16_t)(b + 1);
      }
    } else if (al) {
      colp[kl * CS] = (uint16_t)(colp[kl * CS] - 1);
    } else if (ae) {
      colp[ke * CS] = (uint16_t)(colp[ke * CS] + 1);
    }
  };
  auto load_bin = [&](int j) -> int { return col[(int64_t)__ldg(&doy_rows[j]) * pitch]; };
  // The CTA's shared memory is all counters, so there is no L1: the row lists of the next step
  // are staged into a tiny shared table (one global load per thread) instead of being chased
  // through doy_rows by every thread.
  __shared__ int s_rows[2][2][CAP];   // [step parity][leave / enter][slot] row index or -1
  __shared__ int s_cnt[2][2][2];      // [step parity][leave / enter][begin, end) in doy_rows
  auto stage = [&](int step) {        // rows leaving / entering the window at `step` (1..365)
    const int par = step & 1;
    const int d_leave = (step - 1 - half + 2 * NDOY) % NDOY, d_enter = (step + half) % NDOY;
    if (threadIdx.x < 2 * CAP) {
      const int which = threadIdx.x / CAP, u = threadIdx.x % CAP;
      const int dd = which ? d_enter : d_leave;
      const int b0 = __ldg(&doy_ptr[dd]), b1 = __ldg(&doy_ptr[dd + 1]);
      s_rows[par][which][u] = (b0 + u < b1) ? __ldg(&doy_rows[b0 + u]) : -1;
      if (u == 0) { s_cnt[par][which][0] = b0; s_cnt[par][which][1] = b1; }
    }
  };
  auto publish = [&]() {
    if (role == 0) {
      NT[oc] = (uint16_t)ntot_own;
      Z0[oc] = (uint16_t)n0_own;
    }
  };
  auto warm_l2 = [&](int step) {  // pull the samples entering at `step` into L2 while the queries run
    if (role != 0 || !own_valid) return;
    const int par = step & 1;
#pragma unroll 4
    for (int u = 0; u < CAP; ++u) {
      const int re = s_rows[par][1][u];
      if (re >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (int64_t)re * pitch));
    }
  };
  auto advance = [&](int step) {  // needs stage(step) + a barrier before it
    if (!own_valid) return;
    const int par = step & 1;
    const int nl = s_cnt[par][0][1] - s_cnt[par][0][0], ne = s_cnt[par][1][1] - s_cnt[par][1][0];
    const int npair = min(CAP, max(nl, ne));
    constexpr int BATCH = 14;
    for (int u0 = 0; u0 < npair; u0 += BATCH) {
      int vl[BATCH], ve[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {  // all loads first ...
        const int rl = (u0 + u < CAP) ? s_rows[par][0][u0 + u] : -1, re = (u0 + u < CAP) ? s_rows[par][1][u0 + u] : -1;
        vl[u] = (rl >= 0) ? (int)col[(int64_t)rl * pitch] : 0xFFFF;
        ve[u] = (re >= 0) ? (int)col[(int64_t)re * pitch] : 0xFFFF;
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u) apply(vl[u], ve[u]);  // ... then the shared-memory updates
    }
    for (int a = s_cnt[par][0][0] + CAP; a < s_cnt[par][0][1]; ++a) apply(load_bin(a), 0xFFFF);  // lists longer than CAP
    for (int b = s_cnt[par][1][0] + CAP; b < s_cnt[par][1][1]; ++b) apply(0xFFFF, load_bin(b));
    publish();
  };

  // ---- query role: LPT lanes per target (LPT = 4 or 8), neighbours dealt round-robin ----
  const int nt = TY * TX;
  const int tq = threadIdx.x / LPT, hq = threadIdx.x % LPT;
  const bool q_thread = tq < nt;
  const int tt = q_thread ? tq : 0;
  const int ty = tt / TX, tx = tt % TX;
  const int64_t ty_g = y0 + ty, tx_g = x0 + tx;
  const bool target_live = q_thread && ty_g < y_end && tx_g < x_end;
  const int center = (ty + P) * OX + (tx + P);
  int off[MAXQ];  // this lane's neighbours; unused slots -> dummy zero cell
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int n = LPT * i + hq;
    off[i] = (n < NN) ? center + (n / WS - P) * OX + (n % WS - P) : C;
  }
  const unsigned grp_mask = ((LPT == 8) ? 0xFFu : 0xFu) << (threadIdx.x & 31 & ~(LPT - 1));
  auto grp_sum = [&](int v) {
    v += __shfl_xor_sync(grp_mask, v, 1);
    v += __shfl_xor_sync(grp_mask, v, 2);
    if (LPT == 8) v += __shfl_xor_sync(grp_mask, v, 4);
    return v;
  };
  auto row_sum = [&](const uint16_t* row) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) s += row[off[i]];
    return s;
  };
  auto pooled_cum = [&](int B) {  // pooled samples with bin < B  (B >= 1)
    int s = row_sum(Z0);
    for (int g = 0; g < (B >> 6); ++g) s += row_sum(L2 + g * CS);
    for (int k = (B >> 6) << 3; k < (B >> 3); ++k) s += row_sum(L1 + k * CS);
    for (int b = max(1, (B >> 3) << 3); b < B; ++b) s += row_sum(L0 + b * CS);
    return grp_sum(s);
  };
  auto pooled_bin = [&](int b) { return grp_sum(row_sum(b ? L0 + b * CS : Z0)); };

  __syncthreads();
  if (own_valid) {
    for (int k = -half; k <= half; ++k) {
      const int d = ((k % NDOY) + NDOY) % NDOY;
      const int b0 = __ldg(&doy_ptr[d]), b1 = __ldg(&doy_ptr[d + 1]);
      for (int j = b0; j < b1; ++j) apply(0xFFFF, load_bin(j));
    }
    publish();
  }
  stage(1);
  __syncthreads();

  const bool masked = target_live ? (anom_row0[ty_g * nx + tx_g] != anom_row0[ty_g * nx + tx_g]) : true;
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
  int iu = 1;
  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      if (d + 1 < NDOY) stage(d + 1);
      advance(d);
      __syncthreads();
      if (d + 1 < NDOY) warm_l2(d + 1);
    }
    if (q_thread) {
      float res = CUDART_NAN_F;
      const int ntot = grp_sum(row_sum(NT));
      if (ntot > 0) {
        const double pos = __dmul_rn(q, (double)ntot);
        const int kk = (int)floor(pos);
        int cl = (iu > 0) ? pooled_cum(iu) : 0;
        int h = pooled_bin(iu);
        while (cl > kk) { --iu; h = pooled_bin(iu); cl -= h; }
        while (iu < nb - 1 && cl + h <= kk) { cl += h; ++iu; h = pooled_bin(iu); }
        if (iu == 0) {
          res = __ldg(&centers[0]);
        } else {
          const float bl = __ldg(&centers[iu - 1]), bu = __ldg(&centers[iu]);
          const double frac = (h > 0) ? __ddiv_rn(pos - (double)cl, (double)h) : 0.5;
          res = (float)__dadd_rn((double)bl, __dmul_rn(frac, (double)__fsub_rn(bu, bl)));  // numpy: mul then add, no FMA
        }
      }
      if (masked) res = CUDART_NAN_F;
      if (res == res) { vmin = fminf(vmin, res); vmax = fmaxf(vmax, res); }
      if (res < lower_bound) res = lower_bound;
      if (target_live && hq == 0) thr[(int64_t)d * N + ty_g * nx + tx_g] = res;
    }
    __syncthreads();  // queries done before the next day's updates touch the counters
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0 && stats) {
    if (vmin != CUDART_INF_F) atomic_min_f(&stats[0], vmin);
    if (vmax != -CUDART_INF_F) atomic_max_f(&stats[1], vmax);
  }
}

// ---------------------------------------------------------------------------------------
// Exact order statistics.  A per-lane uniform histogram over [min, max] of the lane's own
// series localises the bin that holds the wanted rank; the few samples of that bin are then
// gathered and ordered exactly.  `SampleIter` abstracts "all samples of the current window".
// ---------------------------------------------------------------------------------------
constexpr int NBX = 256;  // bins of the localising histogram
constexpr int CAND = 8;   // in-bin candidates handled by the fast path

struct BinMap {
  float mn, scale;
  __device__ __forceinline__ int operator()(float v) const {  // monotone non-decreasing in v; NaN excluded by caller
    if (v == CUDART_INF_F) return NBX - 1;
    if (v == -CUDART_INF_F) return 0;
    const float f = (v - mn) * scale;
    const int b = (int)f;
    return b < 0 ? 0 : (b > NBX - 1 ? NBX - 1 : b);
  }
};

// Order statistics of ranks r0 and r1 (r1 == r0 or r0 + 1) among the window's valid samples,
// given that bin `ib` holds rank r0, `cl` samples lie in lower bins and `h` in bin ib.
template <typename ForEach>
__device__ __forceinline__ void select_pair(ForEach&& for_each, const BinMap& bm, int ib, int cl, int h, int r0,
                                            int r1, float& a, float& b) {
  const int j0 = r0 - cl;  // rank inside the bin
  float nextmin = CUDART_INF_F;  // smallest sample in a higher bin
  if (h <= CAND) {
    float cand[CAND] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int m = 0;
    for_each([&](float v) {
      const int bi = bm(v);
      if (bi == ib) {
        int k = m++;
        while (k > 0 && cand[k - 1] > v) { cand[k] = cand[k - 1]; --k; }
        cand[k] = v;
      } else if (bi > ib) {
        nextmin = fminf(nextmin, v);
      }
    });
    a = cand[j0];
    b = (r1 == r0) ? a : ((j0 + 1 < m) ? cand[j0 + 1] : nextmin);
    return;
  }
  // Crowded bin (constant or heavily tied series): walk the distinct values upwards.
  float cur = -CUDART_INF_F;
  bool first = true;
  int seen = 0;
  a = CUDART_NAN_F;
  b = CUDART_NAN_F;
  bool have_a = false;
  while (true) {
    float nxt = CUDART_INF_F;
    int mult = 0;
    bool any = false;
    for_each([&](float v) {
      const int bi = bm(v);
      if (bi == ib) {
        if (first ? true : (v > cur)) {
          if (!any || v < nxt) { nxt = v; mult = 1; any = true; }
          else if (v == nxt) ++mult;
        }
      } else if (bi > ib) {
        nextmin = fminf(nextmin, v);
      }
    });
    if (!any) {  // ran out of the bin: rank r1 lives in a higher bin
      if (!have_a) a = nextmin;
      b = nextmin;
      return;
    }
    first = false;
    cur = nxt;
    if (!have_a && j0 < seen + mult) {
      a = nxt;
      have_a = true;
      if (r1 == r0 || j0 + 1 < seen + mult) { b = nxt; return; }
    } else if (have_a) {
      b = nxt;
      return;
    }
    seen += mult;
  }
}

// numpy 'linear' quantile in float32 (numpy/lib/_function_base_impl.py _quantile/_lerp), as
// np.nanpercentile(float32 data, python scalar) computes it (detect.py:1941).
__device__ __forceinline__ void f32_rank(int n, float qf, int& r0, int& r1, float& g) {
  const float vi = __fmul_rn((float)(n - 1), qf);
  if (vi >= (float)(n - 1)) { r0 = r1 = n - 1; g = 0.f; return; }
  if (vi < 0.f) { r0 = r1 = 0; g = 0.f; return; }
  const float lo = floorf(vi);
  r0 = (int)lo;
  r1 = r0 + 1;
  g = __fsub_rn(vi, lo);
}
__device__ __forceinline__ float f32_lerp(float a, float b, float g) {
  const float diff = __fsub_rn(b, a);
  float r = __fadd_rn(a, __fmul_rn(diff, g));
  if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.f, g)));
  return r;
}

template <typename CT>
__global__ void __launch_bounds__(32) hobday_exact_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                          int64_t pitch, const int32_t* __restrict__ doy_ptr,
                                                          const int32_t* __restrict__ doy_rows, int w, float qf,
                                                          float* __restrict__ thr) {
  extern __shared__ unsigned char smem_raw[];
  CT* hist = reinterpret_cast<CT*>(smem_raw);  // [NBX][32]
  const int lane = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const bool live = c < N;
  const int64_t cc = live ? c : N - 1;
  for (int b = 0; b < NBX; ++b) hist[b * 32 + lane] = 0;
  // value range of the lane's series
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll 4
  for (int64_t t = 0; t < T; ++t) {
    const float v = __ldg(&anom[t * pitch + cc]);
    if (is_finite_f(v)) { mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  BinMap bm;
  bm.mn = (mn <= mx) ? mn : 0.f;
  bm.scale = (mx > mn) ? ((float)NBX / (mx - mn)) : 0.f;
  if (!is_finite_f(bm.scale)) bm.scale = 0.f;
  const int half = w / 2;
  int n = 0, ib = 0, cl = 0;

  auto apply_doy = [&](int d, int sign) {
    const int b0 = __ldg(&doy_ptr[d]), b1 = __ldg(&doy_ptr[d + 1]);
    for (int j = b0; j < b1; ++j) {
      const float v = __ldg(&anom[(int64_t)__ldg(&doy_rows[j]) * pitch + cc]);
      if (v == v) {
        const int bi = bm(v);
        hist[bi * 32 + lane] += (CT)sign;
        n += sign;
        if (bi < ib) cl += sign;
      }
    }
  };
  for (int k = -half; k <= half; ++k) apply_doy(((k % NDOY) + NDOY) % NDOY, +1);

  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      apply_doy((d - 1 - half + 2 * NDOY) % NDOY, -1);
      apply_doy((d + half) % NDOY, +1);
    }
    float res = CUDART_NAN_F;
    if (n > 0) {
      int r0, r1;
      float g;
      f32_rank(n, qf, r0, r1, g);
      while (cl > r0) { --ib; cl -= (int)hist[ib * 32 + lane]; }
      while (ib < NBX - 1 && cl + (int)hist[ib * 32 + lane] <= r0) { cl += (int)hist[ib * 32 + lane]; ++ib; }
      const int h = (int)hist[ib * 32 + lane];
      auto for_each = [&](auto&& fn) {
        for (int k = -half; k <= half; ++k) {
          const int dd = (d + k + NDOY) % NDOY;
          const int b0 = __ldg(&doy_ptr[dd]), b1 = __ldg(&doy_ptr[dd + 1]);
          for (int j = b0; j < b1; ++j) {
            const float v = __ldg(&anom[(int64_t)__ldg(&doy_rows[j]) * pitch + cc]);
            if (v == v) fn(v);
          }
        }
      };
      float a, b;
      select_pair(for_each, bm, ib, cl, h, r0, r1, a, b);
      res = f32_lerp(a, b, g);
    }
    if (live) thr[(int64_t)d * N + c] = res;
  }
}

// Same result with the window's samples kept in shared memory: win[slot][row][lane], one slot per day
// of year of the window (ring over slots), so every sample is loaded from global memory once and
// the per-step candidate scan of select_pair reads shared memory instead of re-gathering ~w * n_years
// strided global values.  Used when the window fits (w * rowcap * 128 B + histogram <= 200 KB).
// Smallest float t with bm(t) >= target (bm is monotone non-decreasing): the analytic guess is fixed up
// by single-ulp steps, so classifying a sample against a bin is two float compares instead of the map.
__device__ __forceinline__ float bin_lower_bound(const BinMap& bm, int target) {
  if (target <= 0) return -CUDART_INF_F;
  if (bm.scale <= 0.f) return CUDART_INF_F;  // degenerate range: every finite sample maps to bin 0
  float t = bm.mn + (float)target / bm.scale;
  if (!(fabsf(t) < CUDART_INF_F)) t = (t > 0.f) ? 3.0e38f : -3.0e38f;
  for (int it = 0; it < 64 && bm(t) >= target; ++it) t = nextafterf(t, -CUDART_INF_F);  // now bm(t) < target (or gave up)
  for (int it = 0; it < 128 && bm(t) < target; ++it) t = nextafterf(t, CUDART_INF_F);   // first t with bm(t) >= target
  return t;
}

template <typename CT>
__global__ void __launch_bounds__(32) hobday_exact_win_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                              int64_t pitch, const int32_t* __restrict__ doy_ptr,
                                                              const int32_t* __restrict__ doy_rows, int w, int rowcap,
                                                              float qf, float* __restrict__ thr,
                                                              const float* __restrict__ minmax,
                                                              const int32_t* __restrict__ warp_list) {
  extern __shared__ unsigned char smem_raw[];
  CT* hist = reinterpret_cast<CT*>(smem_raw);                                     // [NBX][32]
  float* win = reinterpret_cast<float*>(smem_raw + (size_t)NBX * 32 * sizeof(CT)); // [w][rowcap][32]
  int* nrow = reinterpret_cast<int*>(win + (size_t)w * rowcap * 32);              // [w] rows held by each slot
  const int lane = threadIdx.x;
  // all groups of 32 gridpoints, or (list mode) the groups warp_list[1 .. 1 + warp_list[0]) the queue kernel gave up on
  if (warp_list && (int)blockIdx.x >= __ldg(&warp_list[0])) return;
  const int64_t c = (int64_t)(warp_list ? __ldg(&warp_list[1 + blockIdx.x]) : (int)blockIdx.x) * 32 + lane;
  const bool live = c < N;
  const float* col = anom + (live ? c : N - 1);
  for (int b = 0; b < NBX; ++b) hist[b * 32 + lane] = 0;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  if (minmax) {  // finite range of the series, from col_minmax_kernel (this kernel runs 4 warps per SM: a poor place for a full pass)
    mn = minmax[live ? c : N - 1];
    mx = minmax[N + (live ? c : N - 1)];
  } else {
    for (int64_t t0 = 0; t0 < T; t0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (t0 + u < T) ? __ldg(col + (t0 + u) * pitch) : CUDART_NAN_F;
#pragma unroll
      for (int u = 0; u < 8; ++u) if (is_finite_f(v[u])) { mn = fminf(mn, v[u]); mx = fmaxf(mx, v[u]); }
    }
  }
  BinMap bm;
  bm.mn = (mn <= mx) ? mn : 0.f;
  bm.scale = (mx > mn) ? ((float)NBX / (mx - mn)) : 0.f;
  if (!is_finite_f(bm.scale)) bm.scale = 0.f;
  const int half = w / 2;
  int n = 0, ib = 0, cl = 0;

  auto count = [&](float v, int sign) {
    if (v == v) {
      const int bi = bm(v);
      hist[bi * 32 + lane] += (CT)sign;
      n += sign;
      if (bi < ib) cl += sign;
    }
  };
  auto leave_slot = [&](int slot) {
    const int m = nrow[slot];
    const float* ws = win + (size_t)slot * rowcap * 32 + lane;
    for (int j = 0; j < m; ++j) count(ws[j * 32], -1);
  };
  auto enter_slot = [&](int slot, int d) {  // rows of day-of-year index d into `slot`
    const int b0 = __ldg(&doy_ptr[d]), b1 = min(__ldg(&doy_ptr[d + 1]), b0 + rowcap);  // (host guarantees rows <= rowcap)
    float* ws = win + (size_t)slot * rowcap * 32 + lane;
    for (int j0 = b0; j0 < b1; j0 += 13) {
      float v[13];
#pragma unroll
      for (int u = 0; u < 13; ++u) v[u] = (j0 + u < b1) ? __ldg(col + (int64_t)__ldg(&doy_rows[j0 + u]) * pitch) : 0.f;
#pragma unroll
      for (int u = 0; u < 13; ++u) if (j0 + u < b1) { ws[(j0 + u - b0) * 32] = v[u]; count(v[u], +1); }
    }
    __syncwarp();
    if (lane == 0) nrow[slot] = b1 - b0;
    __syncwarp();
  };
  for (int k = 0; k < w; ++k) enter_slot(k, ((k - half) % NDOY + NDOY) % NDOY);

  for (int d = 0; d < NDOY; ++d) {
    if (d > 0) {
      const int slot = (d - 1) % w;  // holds day d - 1 - half, which leaves; day d + half takes its place
      leave_slot(slot);
      enter_slot(slot, (d + half) % NDOY);
    }
    float res = CUDART_NAN_F;
    if (n > 0) {
      int r0, r1;
      float g;
      f32_rank(n, qf, r0, r1, g);
      while (cl > r0) { --ib; cl -= (int)hist[ib * 32 + lane]; }
      while (ib < NBX - 1 && cl + (int)hist[ib * 32 + lane] <= r0) { cl += (int)hist[ib * 32 + lane]; ++ib; }
      const int h = (int)hist[ib * 32 + lane];
      auto for_each = [&](auto&& fn) {
        for (int slot = 0; slot < w; ++slot) {
          const int m = nrow[slot];
          const float* ws = win + (size_t)slot * rowcap * 32 + lane;
          for (int j = 0; j < m; ++j) {
            const float v = ws[j * 32];
            if (v == v) fn(v);
          }
        }
      };
      float a, b;
      // the rank is found among the bin's few samples: classify the window against the bin's value range
      // (two compares per sample) and keep the in-bin samples sorted in registers
      const float tlo = bin_lower_bound(bm, ib);
      const float thi = (ib < NBX - 1) ? bin_lower_bound(bm, ib + 1) : CUDART_INF_F;
      const bool top = ib >= NBX - 1;
      const bool exactb = (bm(tlo) >= ib) && (ib == 0 || bm(nextafterf(tlo, -CUDART_INF_F)) < ib) &&
                          (top || (bm(thi) > ib && bm(nextafterf(thi, -CUDART_INF_F)) <= ib));
      const unsigned act = __activemask();  // lanes without a valid sample are not here
      if (__all_sync(act, h <= CAND && exactb)) {
        float cand[CAND];
#pragma unroll
        for (int k = 0; k < CAND; ++k) cand[k] = CUDART_INF_F;
        float nextmin = CUDART_INF_F;  // smallest sample above the bin
        int m_in = 0;
        for (int slot = 0; slot < w; ++slot) {
          const int m = nrow[slot];
          const float* ws = win + (size_t)slot * rowcap * 32 + lane;
          for (int j = 0; j < m; ++j) {
            float v = ws[j * 32];
            const bool above = !top && v >= thi;
            if (above) nextmin = fminf(nextmin, v);
            if (v >= tlo && !above) {  // in the bin (false for NaN): sorted insert, the larger value moves on
              ++m_in;
#pragma unroll
              for (int k = 0; k < CAND; ++k) { const float lo = fminf(cand[k], v); v = fmaxf(cand[k], v); cand[k] = lo; }
            }
          }
        }
        const int j0 = r0 - cl;
        a = cand[0];
        float a1 = cand[1];
#pragma unroll
        for (int k = 1; k < CAND; ++k) {
          if (j0 == k) { a = cand[k]; a1 = (k + 1 < CAND) ? cand[k + 1] : CUDART_INF_F; }
        }
        b = (r1 == r0) ? a : ((j0 + 1 < m_in) ? a1 : nextmin);
      } else {
        select_pair(for_each, bm, ib, cl, h, r0, r1, a, b);
      }
      res = f32_lerp(a, b, g);
    }
    if (live) thr[(int64_t)d * N + c] = res;
  }
}

// Queue variant (exact_queue.cuh): per gridpoint a first-in first-out queue of the samples above a pivot instead of
// the window and its histogram -- 8.3 KB of shared memory per warp instead of 51 KB, four warps per CTA.
struct XqDevEnv {
  const char* col;   // the lane's gridpoint in row 0 (bytes)
  uint32_t pitch4;   // row pitch in bytes: one 32 x 32 + 64 bit multiply-add per sample address
  const int32_t* doy_ptr;
  const int32_t* doy_rows;
  float* que_;    // [Q][32] of this warp, already offset by the lane
  uint8_t* cnt_;  // [3][w][32] of this warp, already offset by the lane
  int w;
  static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
  static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
  static __device__ __forceinline__ bool finite(float v) { return is_finite_f(v); }
  static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float level(float lob, float top, int k) {
    return k == 8 ? top : (k == 0 ? lob : lob + (top - lob) * ((float)k * 0.125f));
  }
  float qf;
  __device__ __forceinline__ void rank(int n, int& r0, int& r1, float& g) const { f32_rank(n, qf, r0, r1, g); }
  __device__ __forceinline__ float finish(float a, float b, float g) const { return f32_lerp(a, b, g); }
  __device__ __forceinline__ float load(int j) const {
    return __ldg(reinterpret_cast<const float*>(col + (uint64_t)(uint32_t)__ldg(&doy_rows[j]) * pitch4));
  }
  __device__ __forceinline__ int doy_begin(int dd) const { return __ldg(&doy_ptr[dd]); }
  __device__ __forceinline__ float& que(int pos) { return *reinterpret_cast<float*>(reinterpret_cast<char*>(que_) + (pos << 7)); }
  __device__ __forceinline__ uint8_t& cnt(int s) { return cnt_[s * 32]; }
  __device__ __forceinline__ uint8_t& eqc(int s) { return cnt_[(w + s) * 32]; }
  __device__ __forceinline__ uint8_t& nvc(int s) { return cnt_[(2 * w + s) * 32]; }
  __device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(0xffffffffu, v); }
  __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p); }
  __device__ __forceinline__ bool all(bool p) { return __all_sync(0xffffffffu, p); }
};

constexpr int XQ_WARPS = 4;

template <int Q, int WC>
__global__ void __launch_bounds__(XQ_WARPS * 32, 6) hobday_exact_queue_kernel(const float* __restrict__ anom, int64_t N, int64_t pitch,
                                                                          const int32_t* __restrict__ doy_ptr,
                                                                          const int32_t* __restrict__ doy_rows, int w, float qf,
                                                                          float* __restrict__ thr, int32_t* __restrict__ fail_list,
                                                                          int force_fail) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t group = (int64_t)blockIdx.x * XQ_WARPS + wi;  // 32 adjacent gridpoints; no CTA-wide barrier below
  if (group * 32 >= N) return;
  const int64_t c = group * 32 + lane;
  const bool live = c < N;
  const size_t per_warp = (size_t)Q * 128 + (size_t)3 * w * 32;
  unsigned char* base = smem_raw + wi * per_warp;
  XqDevEnv env{reinterpret_cast<const char*>(anom + (live ? c : N - 1)), (uint32_t)(pitch * 4), doy_ptr, doy_rows, reinterpret_cast<float*>(base) + lane,
               base + (size_t)Q * 128 + lane, w, qf};
  ExactQueue<Q, XqDevEnv, WC> lane_q(env, w);
  bool ok = !(force_fail == 1 || (force_fail == 2 && (group & 1)));
  if (ok) ok = lane_q.run([&](int d, float v) { if (live) thr[(int64_t)d * N + c] = v; });
  if (!ok && lane == 0) fail_list[1 + atomicAdd(&fail_list[0], 1)] = (int32_t)group;
}

// np.nanquantile(a, float64 q) 'linear' (xarray .quantile, detect.py:2899): float64 virtual
// index and lerp, the difference (b - a) still taken in float32.
template <typename CT>
__global__ void __launch_bounds__(32) global_exact_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                          int64_t pitch, double q, double* __restrict__ thr) {
  extern __shared__ unsigned char smem_raw[];
  CT* hist = reinterpret_cast<CT*>(smem_raw);
  const int lane = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const bool live = c < N;
  const int64_t cc = live ? c : N - 1;
  for (int b = 0; b < NBX; ++b) hist[b * 32 + lane] = 0;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll 4
  for (int64_t t = 0; t < T; ++t) {
    const float v = __ldg(&anom[t * pitch + cc]);
    if (is_finite_f(v)) { mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  BinMap bm;
  bm.mn = (mn <= mx) ? mn : 0.f;
  bm.scale = (mx > mn) ? ((float)NBX / (mx - mn)) : 0.f;
  if (!is_finite_f(bm.scale)) bm.scale = 0.f;
  int n = 0;
#pragma unroll 4
  for (int64_t t = 0; t < T; ++t) {
    const float v = __ldg(&anom[t * pitch + cc]);
    if (v == v) { hist[bm(v) * 32 + lane] += 1; ++n; }
  }
  double res = CUDART_NAN;
  if (n > 0) {
    const double vi = (double)(n - 1) * q;
    int r0, r1;
    double g;
    if (vi >= (double)(n - 1)) { r0 = r1 = n - 1; g = 0.0; }
    else if (vi < 0.0) { r0 = r1 = 0; g = 0.0; }
    else { const double lo = floor(vi); r0 = (int)lo; r1 = r0 + 1; g = vi - lo; }
    int ib = 0, cl = 0;
    while (ib < NBX - 1 && cl + (int)hist[ib * 32 + lane] <= r0) { cl += (int)hist[ib * 32 + lane]; ++ib; }
    const int h = (int)hist[ib * 32 + lane];
    auto for_each = [&](auto&& fn) {
      for (int64_t t = 0; t < T; ++t) {
        const float v = __ldg(&anom[t * pitch + cc]);
        if (v == v) fn(v);
      }
    };
    float a, b;
    select_pair(for_each, bm, ib, cl, h, r0, r1, a, b);
    const double diff = (double)__fsub_rn(b, a);
    res = __dadd_rn((double)a, __dmul_rn(diff, g));
    if (g >= 0.5) res = __dsub_rn((double)b, __dmul_rn(diff, __dsub_rn(1.0, g)));
  }
  if (live) thr[c] = res;
}

// ---------------------------------------------------------------------------------------
// Global approximate threshold (_compute_histogram_quantile_1d, detect.py:2737-2865).
// float64 edges, last bin right-closed; pdf = hist / (sum + 1e-10); cdf = sequential cumsum.
// ---------------------------------------------------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(32) global_hist_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                         int64_t pitch, const double* __restrict__ edges,
                                                         const double* __restrict__ centers, int nb, double q,
                                                         double lower_bound, double* __restrict__ thr,
                                                         double* __restrict__ stats,
                                                         const int32_t* __restrict__ cell_list) {
  extern __shared__ unsigned char smem_raw[];
  CT* hist = reinterpret_cast<CT*>(smem_raw);                         // [nb][32]
  double* s_edges = reinterpret_cast<double*>(smem_raw + (((size_t)nb * 32 * sizeof(CT) + 15) & ~(size_t)15));  // [nb+1]
  const int lane = threadIdx.x;
  // all gridpoints, or (list mode) the gridpoints cell_list[1 .. 1 + cell_list[0]) the fast kernel deferred
  const int64_t n_cells = cell_list ? (int64_t)__ldg(&cell_list[0]) : N;
  for (int i = lane; i <= nb; i += 32) s_edges[i] = edges[i];
  for (int64_t g = blockIdx.x; g * 32 < n_cells; g += gridDim.x) {
  const int64_t idx = g * 32 + lane;
  const bool live = idx < n_cells;
  const int64_t c = cell_list ? (int64_t)__ldg(&cell_list[1 + (live ? idx : n_cells - 1)]) : (live ? idx : N - 1);
  const int64_t cc = c;
  for (int b = 0; b < nb; ++b) hist[b * 32 + lane] = 0;
  __syncwarp();
  const double e1 = s_edges[1];
  const double inv_step = (nb > 1) ? 1.0 / (s_edges[2] - s_edges[1]) : 1.0;
  const double e_last = s_edges[nb];
  bool has_nan = false;
  long long total = 0;
#pragma unroll 2
  for (int64_t t = 0; t < T; ++t) {
    const float vf = ld_stream(&anom[t * pitch + cc]);
    if (vf != vf) { has_nan = true; continue; }
    const double v = (double)vf;
    if (v > e_last) continue;  // out of range (xhistogram drops it)
    // largest i in [0, nb-1] with edges[i] <= v  (v == edges[nb] falls in the last bin)
    double g = floor((v - e1) * inv_step) + 1.0;
    g = fmin(fmax(g, 0.0), (double)(nb - 1));
    int i = (int)g;
    while (i > 0 && v < s_edges[i]) --i;
    while (i < nb - 1 && v >= s_edges[i + 1]) ++i;
    hist[i * 32 + lane] += 1;
    ++total;
  }
  const double eps = 1e-10;
  const double hist_sum = (double)total + 1e-10;      // detect.py:2778
  // sweep 1: iu = first bin with cdf >= q - eps (0 if none); cdf_target = cdf[max(iu - 1, 0)]
  int iu = 0;
  {
    double cdf = 0.0;
    bool found = false;
    for (int b = 0; b < nb && !found; ++b) {
      const int h = (int)hist[b * 32 + lane];
      if (h) cdf = __dadd_rn(cdf, __ddiv_rn((double)h, hist_sum));
      if (cdf >= q - eps) { iu = b; found = true; }
    }
  }
  auto cdf_at = [&](int idx) {
    double cdf = 0.0;
    for (int b = 0; b <= idx; ++b) {
      const int h = (int)hist[b * 32 + lane];
      if (h) cdf = __dadd_rn(cdf, __ddiv_rn((double)h, hist_sum));
    }
    return cdf;
  };
  const int ibefore = (iu - 1 > 0) ? iu - 1 : 0;      // detect.py:2793
  const double cdf_target = cdf_at(ibefore);
  int il = 0;                                          // first bin with cdf > cdf_target (0 if none)
  {
    double cdf = 0.0;
    bool found = false;
    for (int b = 0; b < nb && !found; ++b) {
      const int h = (int)hist[b * 32 + lane];
      if (h) cdf = __dadd_rn(cdf, __ddiv_rn((double)h, hist_sum));
      if (cdf > cdf_target) { il = b; found = true; }
    }
  }
  il = il < 0 ? 0 : (il > nb - 2 ? nb - 2 : il);      // detect.py:2804-2805
  iu = iu < 1 ? 1 : (iu > nb - 1 ? nb - 1 : iu);
  const double cdl = cdf_at(il), cdu = cdf_at(iu);
  const double bl = centers[il], bu = centers[iu];
  const double denom = __dsub_rn(cdu, cdl);
  const bool exact = fabs(__dsub_rn(cdl, q)) < eps;
  const bool zero = fabs(denom) <= eps;
  const double frac = __ddiv_rn(__dsub_rn(q, cdl), (fabs(denom) > eps) ? denom : 1.0);
  double res = __dadd_rn(bl, __dmul_rn(frac, __dsub_rn(bu, bl)));
  if (exact) res = bl;
  if (zero && !exact) res = __ddiv_rn(__dadd_rn(bl, bu), 2.0);
  if (has_nan) res = CUDART_NAN;                       // detect.py:2835-2836
  double vmin = CUDART_INF, vmax = -CUDART_INF;
  if (live && res == res) { vmin = res; vmax = res; }
  if (res < lower_bound) res = lower_bound;            // detect.py:2853-2863
  if (live) thr[c] = res;
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if (lane == 0 && stats) {
    if (vmin != CUDART_INF) atomic_min_d(&stats[0], vmin);
    if (vmax != -CUDART_INF) atomic_max_d(&stats[1], vmax);
  }
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------
// Global approximate threshold, fast path.  For a quantile bin iu in [1, nb - 2] the reference's
// rule collapses to thr = centers[iu] (SURVEY.md 3.5 b-global), and iu is decided by
// cdf[b] >= q - 1e-10 with cdf[b] = K_b / (S + 1e-10) up to < 1e-13 of float64 rounding, i.e. by
// the integer test K_b >= R, R = (q - 1e-10) (S + 1e-10), whenever R is not within 1e-3 of an
// integer.  So: thread = gridpoint, pass 1 counts 8-bin blocks (uint16 column in shared memory)
// and finds the block of rank ceil(R), pass 2 counts the 8 bins of that block.  Gridpoints with a
// near-integer R, iu outside [1, nb - 2] or no in-range sample are appended to `slow_list` and
// recomputed by global_hist_kernel in the reference's float64 order.  Binning compares float32
// samples with edges_up[i] = the smallest float32 >= the float64 edge, which is exact.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) global_hist_fast_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                               int64_t pitch, const float* __restrict__ edges_up,
                                                               float e_last_dn, const double* __restrict__ centers,
                                                               int nb, double q, double lower_bound,
                                                               double* __restrict__ thr, double* __restrict__ stats,
                                                               int32_t* __restrict__ slow_list) {
  extern __shared__ unsigned char smem_raw[];
  const int nblk = (nb + 7) >> 3;
  uint16_t* cnt = reinterpret_cast<uint16_t*>(smem_raw);  // [nblk][256]
  float* s_edges = reinterpret_cast<float*>(smem_raw + (size_t)nblk * 256 * 2);  // [nb + 1]
  for (int i = threadIdx.x; i <= nb; i += 256) s_edges[i] = edges_up[i];
  uint16_t* my = cnt + threadIdx.x;
  for (int j = 0; j < nblk; ++j) my[j * 256] = 0;
  __syncthreads();
  const float e1 = s_edges[1];
  const float inv_step = (nb > 1) ? 1.f / (s_edges[2] - s_edges[1]) : 1.f;
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool live = c < N;
  const float* col = anom + (live ? c : N - 1);
  auto bin_of = [&](float v) -> int {  // largest i in [0, nb - 1] with edge[i] <= v, given v <= last edge
    const int i = digitize_f32(v, s_edges, nb + 1, e1, inv_step);
    return i > nb - 1 ? nb - 1 : i;
  };
  bool has_nan = false;
  int total = 0;
  for (int64_t t0 = 0; t0 < T; t0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (t0 + u < T) ? ld_stream(col + (t0 + u) * pitch) : CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (v[u] != v[u]) { has_nan = true; continue; }
      if (v[u] <= e_last_dn) { const int j = bin_of(v[u]) >> 3; my[j * 256] = (uint16_t)(my[j * 256] + 1); ++total; }
    }
  }
  const double R = __dmul_rn(q - 1e-10, (double)total + 1e-10);
  const double Rc = ceil(R);
  bool slow = total <= 0 || fabs(R - rint(R)) < 1e-3;
  const int kneed = (int)Rc;
  int run = 0, jb = -1;
  for (int j = 0; j < nblk && jb < 0; ++j) {
    const int h = my[j * 256];
    if (run + h >= kneed) jb = j; else run += h;
  }
  if (jb < 0) slow = true;
  double res = CUDART_NAN;
  if (!slow) {
    for (int b = 0; b < 8; ++b) my[b * 256] = 0;  // the column is free again: bins of block jb
    // value range of block jb (a superset is enough: bin_of decides): [edge of its first bin, edge after its last]
    const float blk_lo = s_edges[8 * jb];
    const float blk_hi = (8 * jb + 8 <= nb - 1) ? s_edges[8 * jb + 8] : e_last_dn;
    for (int64_t t0 = 0; t0 < T; t0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (t0 + u < T) ? __ldg(col + (t0 + u) * pitch) : CUDART_INF_F;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (v[u] >= blk_lo && v[u] <= blk_hi) {  // two compares reject ~98 % of the samples (false for NaN)
          const int i = bin_of(v[u]);
          if ((i >> 3) == jb) my[(i & 7) * 256] = (uint16_t)(my[(i & 7) * 256] + 1);
        }
      }
    }
    int iu = -1;
    for (int b = 0; b < 8 && iu < 0; ++b) {
      const int h = my[b * 256];
      if (run + h >= kneed) iu = 8 * jb + b; else run += h;
    }
    if (iu < 1 || iu > nb - 2) slow = true;
    else res = centers[iu];
  }
  if (slow && !has_nan) {
    if (live) slow_list[1 + atomicAdd(&slow_list[0], 1)] = (int32_t)c;
    return;  // the exact kernel writes thr and stats of this gridpoint
  }
  if (has_nan) res = CUDART_NAN;  // detect.py:2835-2836
  if (live) {
    if (res == res && stats) { atomic_min_d(&stats[0], res); atomic_max_d(&stats[1], res); }
    if (res < lower_bound) res = lower_bound;
    thr[c] = res;
  }
}

// Finite minimum / maximum of every gridpoint's series: mm[c] = min, mm[N + c] = max (+inf / -inf when no sample is finite).
// Thread = gridpoint, rows split over blockIdx.y and merged with float atomics.
__global__ void __launch_bounds__(256) col_minmax_kernel(const float* __restrict__ a, int64_t T, int64_t N, int64_t pitch,
                                                         float* __restrict__ mm, int rows_per_block) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (int64_t t = t0; t < t1; t += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (t + u < t1) ? ld_stream(a + (t + u) * pitch + c) : CUDART_NAN_F;
#pragma unroll
    for (int u = 0; u < 8; ++u) if (is_finite_f(v[u])) { mn = fminf(mn, v[u]); mx = fmaxf(mx, v[u]); }
  }
  if (mn <= mx) {
    atomic_min_f(&mm[c], mn);
    atomic_max_f(&mm[N + c], mx);
  }
}
__global__ void init_minmax_kernel(float* mm, int64_t N) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < N) { mm[c] = CUDART_INF_F; mm[N + c] = -CUDART_INF_F; }
}

template <typename F>
__global__ void init_stats_kernel(F* stats) {
  stats[0] = (F)CUDART_INF;
  stats[1] = (F)-CUDART_INF;
}

}  // namespace marex

using namespace marex;

template <typename K>
static int set_smem(K kern, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  return MAREX_OK;
}

// Full-range tiled pooled kernel over the whole grid (fail_list == nullptr) or over the band tiles
// listed in fail_list (band_ty x band_tx targets each; pool_band.cu).  MAREX_ERR_UNSUPPORTED when
// no tile fits shared memory.
static int launch_pool_tile(const uint16_t* bins, int64_t ny, int64_t nx, int64_t pitch, const int32_t* doy_ptr,
                            const int32_t* doy_rows, const float* centers, int nb, int w, int ws, double q,
                            const float* anom_row0, float lower_bound, float* thr, float* stats,
                            const int32_t* fail_list, int max_tiles, int band_ty, int band_tx, cudaStream_t st) {
  // pick the own-cell tile OY x OX that fits shared memory
  const int p = ws / 2;
  const int nb1 = (nb + 7) >> 3, nb2 = (nb + 63) >> 6;
  const size_t per_col = (size_t)(nb + nb1 + nb2 + 2) * 2;  // bytes per counter column (own cell or dummy)
  int cs_max = (int)((227 * 1024 - 1024) / per_col);
  if (cs_max > 257) cs_max = 257;
  const int c_max = ((cs_max - 1) | 1) - 1;  // odd column stride spreads a row's cells over all banks
  int best_oy = 0, best_ox = 0, best_t = 0;
  for (int oy = 2 * p + 1; oy <= 64; ++oy)
    for (int ox = 2 * p + 1; ox <= 64; ++ox) {
      if (oy * ox > c_max) continue;
      const int t = (oy - 2 * p) * (ox - 2 * p);
      if (t > best_t || (t == best_t && ox > best_ox)) { best_t = t; best_oy = oy; best_ox = ox; }
    }
  if (best_t <= 0) return MAREX_ERR_UNSUPPORTED;
  const int C = best_oy * best_ox;
  const int CS = (C + 1) | 1;
  const size_t smem_t = per_col * CS;
  const int TY = best_oy - 2 * p, TX = best_ox - 2 * p;
  const int CR = ((C + 31) / 32) * 32;
  int threads = 3 * CR;                       // three event roles
  const int LPT = (8 * TY * TX <= threads) ? 8 : 4;
  if (LPT * TY * TX > threads) threads = ((LPT * TY * TX + 31) / 32) * 32;
  if (threads > 768) return fail(MAREX_ERR_UNSUPPORTED, "pooled tile needs more than 768 threads");
  dim3 grid_t((unsigned)((nx + TX - 1) / TX), (unsigned)((ny + TY - 1) / TY));
  int sub_x = 1, sub_n = 1;
  if (fail_list) {
    sub_x = (band_tx + TX - 1) / TX;
    sub_n = sub_x * ((band_ty + TY - 1) / TY);
    grid_t = dim3((unsigned)((int64_t)max_tiles * sub_n), 1);
  }
#define MAREX_POOL(PP)                                                                                             \
  do {                                                                                                             \
    int rc = set_smem(hobday_pool_tile_kernel<PP>, smem_t);                                                        \
    if (rc) return rc;                                                                                             \
    hobday_pool_tile_kernel<PP><<<grid_t, threads, smem_t, st>>>(bins, ny, nx, pitch, doy_ptr, doy_rows, centers,  \
                                                                 nb, w, q, anom_row0, lower_bound, thr, stats,    \
                                                                 best_oy, best_ox, CS, CR, LPT, fail_list, sub_x, \
                                                                 sub_n, band_ty, band_tx);                        \
  } while (0)
  if (p == 1) MAREX_POOL(1); else if (p == 2) MAREX_POOL(2); else MAREX_POOL(3);
#undef MAREX_POOL
  MAREX_LAUNCH_CHECK("hobday_pool_tile_kernel");
  return MAREX_OK;
}

namespace marex {
int launch_pool_tile_list(const uint16_t* bins, int64_t ny, int64_t nx, int64_t pitch, const int32_t* doy_ptr,
                          const int32_t* doy_rows, const float* centers, int nb, int w, int ws, double q,
                          const float* anom_row0, float lower_bound, float* thr, float* stats,
                          const int32_t* fail_list, int max_tiles, int band_ty, int band_tx, cudaStream_t st) {
  const int rc = launch_pool_tile(bins, ny, nx, pitch, doy_ptr, doy_rows, centers, nb, w, ws, q, anom_row0, lower_bound,
                                  thr, stats, fail_list, max_tiles, band_ty, band_tx, st);
  if (rc == MAREX_ERR_UNSUPPORTED) return fail(rc, "no full-range tile fits shared memory for the band fallback");
  return rc;
}
}  // namespace marex

extern "C" int marex_hobday_thresholds_hist(const uint16_t* bins, int64_t T, int64_t ny, int64_t nx, int64_t pitch,
                                            const int32_t* doy_ptr, const int32_t* doy_rows, int32_t max_window_rows,
                                            const float* centers, int32_t nb, int32_t w, int32_t ws, double q,
                                            const float* anom_row0, float lower_bound, float* thr, float* stats,
                                            void* stream) {
  MAREX_REQUIRE(bins && doy_ptr && doy_rows && centers && anom_row0 && thr, "null pointer");
  MAREX_REQUIRE(T > 0 && ny > 0 && nx > 0 && pitch >= ny * nx, "bad shape");
  MAREX_REQUIRE(nb >= 2 && nb <= 1700, "nb must be in 2..1700 (shared-memory histogram)");
  MAREX_REQUIRE(w >= 3 && w <= 365 && (w & 1), "window_days_hobday must be odd and in 3..365");
  MAREX_REQUIRE(ws >= 1 && (ws & 1), "window_spatial_hobday must be odd");
  MAREX_REQUIRE(ny <= 65535, "ny too large for grid.y");
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) {
    init_stats_kernel<float><<<1, 1, 0, st>>>(stats);
    MAREX_LAUNCH_CHECK("init_stats_kernel");
  }
  if (ws > 1 && ws <= 7 && max_window_rows <= 65535 && !tune_get("pool_v1", 0)) {
    const int rc = launch_pool_tile(bins, ny, nx, pitch, doy_ptr, doy_rows, centers, nb, w, ws, q, anom_row0,
                                    lower_bound, thr, stats, nullptr, 0, 0, 0, st);
    if (rc != MAREX_ERR_UNSUPPORTED) return rc;
  }
  const long long max_count = (long long)max_window_rows * ws * ws;
  const bool wide = max_count > 65535;
  const bool narrow = max_count <= 255 && !tune_get("hist_u16", 0);  // byte counters: half the shared memory, twice the warps per SM
  const size_t smem = (size_t)nb * 32 * (wide ? 4 : (narrow ? 1 : 2));
  MAREX_REQUIRE(smem <= 220 * 1024, "histogram does not fit shared memory");
  dim3 grid((unsigned)((nx + 31) / 32), (unsigned)ny);
#define MAREX_HH(CT, P)                                                                                           \
  do {                                                                                                            \
    int rc = set_smem(hobday_hist_kernel<CT, P>, smem);                                                           \
    if (rc) return rc;                                                                                            \
    hobday_hist_kernel<CT, P><<<grid, 32, smem, st>>>(bins, ny, nx, pitch, doy_ptr, doy_rows, centers, nb, w, ws, \
                                                      q, anom_row0, lower_bound, thr, stats);                     \
  } while (0)
  if (ws > 1) { if (wide) MAREX_HH(uint32_t, true); else if (narrow) MAREX_HH(uint8_t, true); else MAREX_HH(uint16_t, true); }
  else        { if (wide) MAREX_HH(uint32_t, false); else if (narrow) MAREX_HH(uint8_t, false); else MAREX_HH(uint16_t, false); }
#undef MAREX_HH
  MAREX_LAUNCH_CHECK("hobday_hist_kernel");
  return MAREX_OK;
}

extern "C" int marex_hobday_thresholds_exact_f32(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                                 const int32_t* doy_ptr, const int32_t* doy_rows,
                                                 int32_t max_window_rows, int32_t max_doy_rows, int32_t w,
                                                 float percentile, float* thr, float* work, void* stream) {
  MAREX_REQUIRE(anom && doy_ptr && doy_rows && thr, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  MAREX_REQUIRE(w >= 1 && w <= 365 && (w & 1), "window_days_hobday must be odd and in 1..365");
  const float qf = percentile / 100.0f;  // np.true_divide(q, float32(100)) in float32
  const bool wide = max_window_rows > 65535;
  const size_t smem = (size_t)NBX * 32 * (wide ? 4 : 2);
  const unsigned grid = (unsigned)((N + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  // window-in-shared-memory variant (max_doy_rows = most rows any single day of year has; 0 = unknown)
  const int rowcap_day = max_doy_rows > 0 ? max_doy_rows : 1 << 20;
  const size_t smem_win = smem + (size_t)w * (size_t)rowcap_day * 128 + (size_t)w * sizeof(int);
  if (!wide && max_doy_rows > 0 && smem_win <= 200 * 1024 && !tune_get("exact_v1", 0)) {
    int rc = set_smem(hobday_exact_win_kernel<uint16_t>, smem_win);
    if (rc) return rc;
    // Queue kernel (exact_queue.cuh) when the kk largest samples of any window plus the pivot's room fit a queue of 64
    // or 128 floats per gridpoint; the groups of 32 gridpoints it gives up on are listed in `work` and recomputed by
    // the histogram kernel.  marex_tune("exact_queue", 0) selects the histogram kernel for everything;
    // "exact_force_fail" = 1 / 2 sends every / every other group through the list (tests).
    int kk_max;  // largest kk = n - r0 any window can ask for (non-decreasing in n): f32_rank on the host
    {
      const int n = max_window_rows < 1 ? 1 : max_window_rows;
      const float vi = (float)(n - 1) * qf;
      const int r0 = vi >= (float)(n - 1) ? n - 1 : (vi < 0.f ? 0 : (int)floorf(vi));
      kk_max = n - r0;
    }
    const int qcap = kk_max + XQ_ROOM <= 64 - XQ_HEAD ? 64 : (kk_max + XQ_ROOM <= 128 - XQ_HEAD ? 128 : 0);
    const size_t smem_q = (size_t)XQ_WARPS * ((size_t)qcap * 128 + (size_t)3 * w * 32);
    if (work && qcap && max_doy_rows <= 255 && pitch < (1LL << 30) && smem_q <= 200 * 1024 && tune_get("exact_queue", 1)) {
      int32_t* fail_list = reinterpret_cast<int32_t*>(work);  // [0] = count, then group indices (2 * N floats of room)
      cudaError_t e = cudaMemsetAsync(fail_list, 0, sizeof(int32_t), st);
      MAREX_REQUIRE(e == cudaSuccess, "cudaMemsetAsync failed");
      const int ff = (int)tune_get("exact_force_fail", 0);
      const unsigned qgrid = (unsigned)((grid + XQ_WARPS - 1) / XQ_WARPS);
#define MAREX_XQ(QQ, WW)                                                                                              \
  do {                                                                                                                \
    rc = set_smem(hobday_exact_queue_kernel<QQ, WW>, smem_q);                                                         \
    if (rc) return rc;                                                                                                \
    hobday_exact_queue_kernel<QQ, WW><<<qgrid, XQ_WARPS * 32, smem_q, st>>>(anom, N, pitch, doy_ptr, doy_rows, w, qf, \
                                                                            thr, fail_list, ff);                      \
  } while (0)
      if (qcap == 64) { if (w == 11) MAREX_XQ(64, 11); else MAREX_XQ(64, 0); }  // 11 days: the reference's default window
      else MAREX_XQ(128, 0);
#undef MAREX_XQ
      MAREX_LAUNCH_CHECK("hobday_exact_queue_kernel");
      hobday_exact_win_kernel<uint16_t><<<grid, 32, smem_win, st>>>(anom, T, N, pitch, doy_ptr, doy_rows, w, rowcap_day,
                                                                    qf, thr, nullptr, fail_list);
      MAREX_LAUNCH_CHECK("hobday_exact_win_kernel (list)");
      return MAREX_OK;
    }
    if (work) {  // per-gridpoint finite range at full occupancy (optional scratch of 2 * N floats)
      init_minmax_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(work, N);
      MAREX_LAUNCH_CHECK("init_minmax_kernel");
      const int64_t bx = (N + 255) / 256;
      int64_t by = (16LL * sm_count() + bx - 1) / bx;
      by = by < 1 ? 1 : (by > T ? T : by);
      if (by > 65535) by = 65535;
      const int rpb = (int)((T + by - 1) / by);
      by = (T + rpb - 1) / rpb;
      col_minmax_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, st>>>(anom, T, N, pitch, work, rpb);
      MAREX_LAUNCH_CHECK("col_minmax_kernel");
    }
    hobday_exact_win_kernel<uint16_t><<<grid, 32, smem_win, st>>>(anom, T, N, pitch, doy_ptr, doy_rows, w, rowcap_day,
                                                                  qf, thr, work, nullptr);
    MAREX_LAUNCH_CHECK("hobday_exact_win_kernel");
    return MAREX_OK;
  }
  if (wide) {
    int rc = set_smem(hobday_exact_kernel<uint32_t>, smem);
    if (rc) return rc;
    hobday_exact_kernel<uint32_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, doy_ptr, doy_rows, w, qf, thr);
  } else {
    int rc = set_smem(hobday_exact_kernel<uint16_t>, smem);
    if (rc) return rc;
    hobday_exact_kernel<uint16_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, doy_ptr, doy_rows, w, qf, thr);
  }
  MAREX_LAUNCH_CHECK("hobday_exact_kernel");
  return MAREX_OK;
}

extern "C" int marex_global_threshold_exact_f64(const float* anom, int64_t T, int64_t N, int64_t pitch, double q,
                                                double* thr, void* stream) {
  MAREX_REQUIRE(anom && thr, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  const bool wide = T > 65535;
  const size_t smem = (size_t)NBX * 32 * (wide ? 4 : 2);
  const unsigned grid = (unsigned)((N + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (wide) {
    int rc = set_smem(global_exact_kernel<uint32_t>, smem);
    if (rc) return rc;
    global_exact_kernel<uint32_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, q, thr);
  } else {
    int rc = set_smem(global_exact_kernel<uint16_t>, smem);
    if (rc) return rc;
    global_exact_kernel<uint16_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, q, thr);
  }
  MAREX_LAUNCH_CHECK("global_exact_kernel");
  return MAREX_OK;
}

extern "C" int marex_global_threshold_hist_f64(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                               const double* edges, const double* centers, int32_t nb, double q,
                                               double lower_bound, double* thr, double* stats, void* stream) {
  MAREX_REQUIRE(anom && edges && centers && thr, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  MAREX_REQUIRE(nb >= 3 && nb <= 1600, "nb must be in 3..1600 (shared-memory histogram)");
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) {
    init_stats_kernel<double><<<1, 1, 0, st>>>(stats);
    MAREX_LAUNCH_CHECK("init_stats_kernel");
  }
  const bool wide = T > 65535;
  const size_t smem = (((size_t)nb * 32 * (wide ? 4 : 2) + 15) & ~(size_t)15) + (size_t)(nb + 1) * sizeof(double);
  MAREX_REQUIRE(smem <= 220 * 1024, "histogram does not fit shared memory");
  const unsigned grid = (unsigned)((N + 31) / 32);
  if (wide) {
    int rc = set_smem(global_hist_kernel<uint32_t>, smem);
    if (rc) return rc;
    global_hist_kernel<uint32_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, edges, centers, nb, q, lower_bound, thr,
                                                         stats, nullptr);
  } else {
    int rc = set_smem(global_hist_kernel<uint16_t>, smem);
    if (rc) return rc;
    global_hist_kernel<uint16_t><<<grid, 32, smem, st>>>(anom, T, N, pitch, edges, centers, nb, q, lower_bound, thr,
                                                         stats, nullptr);
  }
  MAREX_LAUNCH_CHECK("global_hist_kernel");
  return MAREX_OK;
}

extern "C" int marex_global_threshold_hist_fast_f64(const float* anom, int64_t T, int64_t N, int64_t pitch,
                                                    const double* edges, const float* edges_up, float e_last_dn,
                                                    const double* centers, int32_t nb, double q, double lower_bound,
                                                    double* thr, double* stats, int32_t* work, void* stream) {
  MAREX_REQUIRE(anom && edges && edges_up && centers && thr && work, "null pointer");
  MAREX_REQUIRE(T > 0 && T <= 65535 && N > 0 && pitch >= N, "bad shape (the fast path counts in 16 bits: T <= 65535)");
  MAREX_REQUIRE(nb >= 16 && nb <= 1600, "nb must be in 16..1600");
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) {
    init_stats_kernel<double><<<1, 1, 0, st>>>(stats);
    MAREX_LAUNCH_CHECK("init_stats_kernel");
  }
  MAREX_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t), st));
  const int nblk = (nb + 7) >> 3;
  const size_t smem_f = (size_t)nblk * 256 * 2 + (size_t)(nb + 1) * sizeof(float);
  int rc = set_smem(global_hist_fast_kernel, smem_f);
  if (rc) return rc;
  global_hist_fast_kernel<<<(unsigned)((N + 255) / 256), 256, smem_f, st>>>(anom, T, N, pitch, edges_up, e_last_dn,
                                                                            centers, nb, q, lower_bound, thr, stats, work);
  MAREX_LAUNCH_CHECK("global_hist_fast_kernel");
  // deferred gridpoints, in the reference's float64 order
  const size_t smem = (((size_t)nb * 32 * 2 + 15) & ~(size_t)15) + (size_t)(nb + 1) * sizeof(double);
  rc = set_smem(global_hist_kernel<uint16_t>, smem);
  if (rc) return rc;
  global_hist_kernel<uint16_t><<<2 * sm_count(), 32, smem, st>>>(anom, T, N, pitch, edges, centers, nb, q, lower_bound,
                                                                  thr, stats, work);
  MAREX_LAUNCH_CHECK("global_hist_kernel");
  return MAREX_OK;
}
