// (a) shifting baseline, fast path for a gap-free daily calendar -- reference: detect.py:1691-1816
// (smoothed_rolling_climatology), 1511-1688 (rolling_climatology), 1819-1850, trim 615-641 -- fused with the
// np.digitize of the approximate Hobday thresholds (detect.py:2622-2631).
//
// CTA = 32 * V adjacent gridpoints (one 128 * V-byte row segment) x one strip of D days of year.
// The CTA walks the calendar years in order.  For year i a single TMA box load
// (cp.async.bulk.tensor.2d, D + S - 1 rows x 32 * V cells) stages the strip's rows plus the
// smoothing halo in shared memory, NST years ahead of the arithmetic, so HBM latency is hidden
// by the copy engine and not by resident warps.  Thread = (V adjacent gridpoints, R consecutive days):
// every shared-memory access is one 4 * V-byte vector, so the load / store / addressing instructions
// and the per-year bookkeeping are shared by V * R element-days.
//   * the S-day centred window sum is assembled from sums of R-row blocks (each staged row is touched once
//     per block) and then slides along the R days,
//   * a ring in shared memory keeps the smoothed value of the last W years for every (day, gridpoint) of the
//     strip; the running sum of the ring sits in registers (float32, Kahan-compensated, fed with
//     `new - old` differences; float64 selectable), so clim[year, doy] = sum * (1 / count) costs one multiply.
//     Which ring entries exist is a property of the calendar alone (first / last S/2 days of the series, day 366
//     of non-leap years), so the count and the validity history are per thread, not per gridpoint,
//   * the anomaly row is written straight out (one 128 * V-byte segment per warp and day), and -- when the caller
//     asks for it -- its histogram bin code next to it: uint16, DAY-OF-YEAR-MAJOR rows
//     (row = doy * NY + year index among the output years), the layout the threshold and compare kernels walk.
//     (day, year) slots without a sample (day 366 of non-leap years, days after the end of the series) get the
//     invalid code, so the consumers need no calendar.
// Every input element is read from HBM (D + S - 1) / D times (L2 catches most of the halo) and
// every output element is written once.
//
// A running sum cannot un-add a NaN/inf, so this kernel is only exact for gridpoints whose
// series is all finite or all NaN (land).  It counts the non-finite inputs per gridpoint (the
// same numbers _validate_data_values needs, detect.py:205-279); marex_shift_anomaly_fixup_f32
// then recomputes the few gridpoints with 0 < count < T with the generic kernel (anomaly.cu).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "digitize.cuh"
#include "tma.cuh"

namespace marex {

struct DailyParams {
  int T, N;                 // rows, gridpoints (both < 2^31: TMA coordinates)
  int64_t out_pitch;
  int out_off;              // output row of input row t is t - out_off
  int year0, doy0;          // calendar year and 0-based day of year of row 0
  int n_years, W, S, D, rows_box, n_strips;
  float* out;
  uint8_t* mask0;
  int32_t* nonfinite;
  const float* edges;       // fused digitize (DIG instantiations): float32 edge table, n_edges entries
  int n_edges;
  uint16_t* bins;           // [366 * NY][bins_pitch]
  int64_t bins_pitch;
  int NY;                   // output years = n_years - W
};

__host__ __device__ __forceinline__ bool is_leap(int y) { return (y % 4 == 0 && y % 100 != 0) || (y % 400 == 0); }
// days from 0001-01-01 to Jan 1 of year y (proleptic Gregorian, y >= 1)
__host__ __device__ __forceinline__ int days_before_year(int y) {
  const int m = y - 1;
  return 365 * m + m / 4 - m / 100 + m / 400;
}

template <typename T, int V>
struct alignas(sizeof(T) * V) Pack {
  T a[V];
};
template <typename T, int V>
__device__ __forceinline__ Pack<T, V> ldp(const T* p) {
  return *reinterpret_cast<const Pack<T, V>*>(p);
}
template <typename T, int V>
__device__ __forceinline__ void stp(T* p, const Pack<T, V>& v) {
  *reinterpret_cast<Pack<T, V>*>(p) = v;
}
__device__ __forceinline__ void st_stream_vec(float* p, const Pack<float, 1>& v) { __stcs(p, v.a[0]); }
__device__ __forceinline__ void st_stream_vec(float* p, const Pack<float, 2>& v) {
  __stcs(reinterpret_cast<float2*>(p), make_float2(v.a[0], v.a[1]));
}
__device__ __forceinline__ void st_stream_vec(float* p, const Pack<float, 4>& v) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(v.a[0], v.a[1], v.a[2], v.a[3]));
}
__device__ __forceinline__ void st_codes(uint16_t* p, const uint32_t (&c)[1]) { __stcs(p, (unsigned short)c[0]); }
__device__ __forceinline__ void st_codes(uint16_t* p, const uint32_t (&c)[2]) {
  __stcs(reinterpret_cast<unsigned int*>(p), c[0] | (c[1] << 16));
}
__device__ __forceinline__ void st_codes(uint16_t* p, const uint32_t (&c)[4]) {
  __stcs(reinterpret_cast<uint2*>(p), make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16)));
}

// Running sum of the ring of one (day, gridpoint).  float (default): Kahan-compensated float32, fed with the difference
// `entering - leaving` (one rounding of a small number per year); measured against the float64 oracle: max error
// 2e-7 of the field scale over 41 years, 50 times below the 1e-5 bar (tests/test_f32_accumulation_study.py).
// double (MAREX_SHIFT_F64=1): float64 sums of float32 data are exact and the result is rounded once, what the oracle does.
template <typename Acc>
struct RingSum;
template <>
struct RingSum<double> {
  double s = 0.0;
  __device__ __forceinline__ void add(double v) { s += v; }
  __device__ __forceinline__ double value() const { return s; }
};
template <>
struct RingSum<float> {
  float s = 0.f, c = 0.f;
  __device__ __forceinline__ void add(float v) {
    const float y = v - c;
    const float t = s + y;
    c = (t - s) - y;
    s = t;
  }
  __device__ __forceinline__ float value() const { return s; }
};

constexpr int SD_MAX_YEARS = 1024;
// Register budget: two CTAs of 11 warps per SM need <= 93 registers per thread.  The instantiations without the fused
// digitize are capped through the launch bounds (uncapped they took 115 registers: one CTA per SM, 82 instead of 49 ms);
// the fused-digitize instantiation compiles to 64 registers uncapped and runs 10 % faster that way than capped (measured:
// 65.9 vs 72.5 ms, profiles/r02_shift_shapes.json).
#ifndef MAREX_SD_MAXTHREADS
#define MAREX_SD_MAXTHREADS 384
#endif
#ifndef MAREX_SD_MINBLOCKS
#define MAREX_SD_MINBLOCKS 2
#endif

// SC / WC / DC: smooth_days_baseline, window_year_baseline and the strip length as compile-time constants (0 = taken from
// the parameters).  The instantiation for the reference's defaults (S = 21, W = 15, hence D = 48) lets the compiler unroll
// the window assembly and fold the ring and stage offsets; same arithmetic in the same order, bit-identical results.
template <int V, int R, int NST, int MODE, typename Acc, bool DIG, int SC = 0, int WC = 0, int DC = 0>
__global__ void __launch_bounds__(DIG ? 512 : MAREX_SD_MAXTHREADS, DIG ? 0 : MAREX_SD_MINBLOCKS)
    shift_daily_kernel(const __grid_constant__ CUtensorMap tmap,
                                                          const __grid_constant__ DailyParams p) {
  constexpr int CW = 32 * V;  // gridpoints per CTA
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = DC ? DC : p.D, W = WC ? WC : p.W, S = SC ? SC : p.S, off = S / 2, Tn = p.T;
  const int rows_box = (SC && DC) ? DC + SC - 1 : p.rows_box;
  // shared memory carve-up (all offsets multiples of 128 bytes)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // [NST]
  size_t o = 128;
  Acc* invtab = reinterpret_cast<Acc*>(smem_raw + o);     // [W + 1], invtab[0] = NaN
  o += (((size_t)(W + 1) * sizeof(Acc) + 127) / 128) * 128;
  int* ybase = reinterpret_cast<int*>(smem_raw + o);      // [n_years + 1] input row of Jan 1 of year index i
  o += (((size_t)(p.n_years + 1) * 4 + 127) / 128) * 128;
  float* s_edges = reinterpret_cast<float*>(smem_raw + o);
  if (DIG) o += (((size_t)p.n_edges * 4 + 127) / 128) * 128;
  float* xs = reinterpret_cast<float*>(smem_raw + o);     // [NST][rows_box][CW]
  const int stage_elems = rows_box * CW;
  float* ring = xs + (size_t)NST * stage_elems;           // [W][D][CW]
  Acc* bs = reinterpret_cast<Acc*>(ring + (size_t)W * D * CW);  // [n_blk][CW] sums of R box rows
  const int n_blk = (rows_box + R - 1) / R;

  // strips of one gridpoint group are adjacent CTAs: they run at the same time and at the
  // same pace, so the S - 1 halo rows a strip shares with its neighbour are L2 hits
  const int strip = blockIdx.x % p.n_strips;
  const int c0 = (blockIdx.x / p.n_strips) * CW;
  const int c = c0 + lane * V;
  const bool live = c < p.N;  // N % 4 == 0 and V divides 4: a thread's V gridpoints are all live or all dead
  const int d0 = strip * D;
  const int rbase = warp * R;
  const uint32_t box_bytes = (uint32_t)rows_box * CW * 4u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i <= W; i += blockDim.x) invtab[i] = i ? (Acc)(1.0 / (double)i) : (Acc)CUDART_NAN;
  {
    const int db0 = days_before_year(p.year0) + p.doy0;
    for (int i = threadIdx.x; i <= p.n_years; i += blockDim.x) ybase[i] = days_before_year(p.year0 + i) - db0;
  }
  __syncthreads();
  DigTable dig;
  if (DIG) dig.init_cta(p.edges, p.n_edges, s_edges);

  // producer state (thread 0): the box of the next year to issue
  int issue_year = 0, issue_st = 0;
  auto issue = [&]() {  // a box that lies entirely outside the series is neither loaded nor waited for
    const int tb = ybase[issue_year] + d0 - off;
    if (tb < Tn && tb + rows_box > 0) {
      mbar_expect_tx(&bar[issue_st], box_bytes);
      tma_load_2d(xs + (size_t)issue_st * stage_elems, &tmap, c0, tb, &bar[issue_st]);
    }
    ++issue_year;
    if (++issue_st == NST) issue_st = 0;
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int k = 0; k < NST && k < p.n_years; ++k) issue();
  }

  RingSum<Acc> sum[R][V];
  int cnt[R];
  uint32_t hist[R];  // bit k: (year i - 1 - k, this day) has a smoothed value in the ring (W <= 31)
#pragma unroll
  for (int r = 0; r < R; ++r) { cnt[r] = 0; hist[r] = 0; }
  int bad[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bad[v] = 0;
  const Acc invS = (Acc)(1.0 / (double)S);
  uint32_t phase = 0;  // bit st = parity of the next completion of stage st
  int st = 0, slot = 0;
  const float* Xst = xs + rbase * CW + lane * V;       // this thread's first window row in stage 0
  float* ringp = ring + (size_t)rbase * CW + lane * V;  // this thread's first day in ring slot 0
  float* const outc = p.out + c;                        // column of these gridpoints (dead lanes never store)
  const uint32_t leave_bit = 1u << (W - 1);

  for (int i = 0; i < p.n_years; ++i) {
    const int base = ybase[i], ylen = ybase[i + 1] - base;
    const float* X = Xst + st * stage_elems;            // X[j * CW]: row rbase + j of the box
    float* ringslot = ringp + slot * (D * CW);
    const int nd = min(D, ylen - d0);                   // days of this strip that exist in year i
    const int t0 = base + d0 + rbase;                   // input row of this thread's first day
    const bool target = i >= W;
    {
      const int tb = base + d0 - off;
      if (tb < Tn && tb + rows_box > 0) {
        mbar_wait(&bar[st], (phase >> st) & 1u);
        phase ^= 1u << st;
      }
    }
    float* outp = outc + (int64_t)(t0 - p.out_off) * p.out_pitch;  // only dereferenced for target years
    const int64_t bin_stride = DIG ? (int64_t)p.NY * p.bins_pitch : 0;
    uint16_t* const binp = DIG ? p.bins + ((int64_t)(d0 + rbase) * p.NY + (i - W)) * p.bins_pitch + c : nullptr;

    // ring turnover of one day: year i - W leaves (if it had a value), year i enters (if it has one)
    auto turnover = [&](int r, const Pack<float, V>& s, bool valid_now) {
      const bool leave = hist[r] & leave_bit;
      hist[r] = (hist[r] << 1) | (valid_now ? 1u : 0u);
      if (leave) {
        const Pack<float, V> old = ldp<float, V>(ringslot + r * CW);
        if (valid_now) {
#pragma unroll
          for (int v = 0; v < V; ++v) sum[r][v].add((Acc)s.a[v] - (Acc)old.a[v]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) sum[r][v].add(-(Acc)old.a[v]);
          --cnt[r];
        }
      } else if (valid_now) {
#pragma unroll
        for (int v = 0; v < V; ++v) sum[r][v].add((Acc)s.a[v]);
        ++cnt[r];
      }
      if (valid_now) stp<float, V>(ringslot + r * CW, s);
    };
    auto emit = [&](int r, const Pack<float, V>& xv) {  // anomaly (or climatology) of a target-year day, and its bin
      const Acc inv = invtab[cnt[r]];
      Pack<float, V> ov;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float clim = (float)(sum[r][v].value() * inv);
        ov.a[v] = MODE ? clim : xv.a[v] - clim;
      }
      if (live) {
        st_stream_vec(outp, ov);
        if (DIG) {
          uint32_t code[V];
#pragma unroll
          for (int v = 0; v < V; ++v) code[v] = dig(ov.a[v]);
          st_codes(binp + r * bin_stride, code);
        }
      }
    };
    auto emit_missing = [&](int r) {  // (day, output year) without a sample: the slot of the bin array is invalid
      if (DIG && live && d0 + rbase + r < NDOY) {
        uint32_t code[V];
#pragma unroll
        for (int v = 0; v < V; ++v) code[v] = (uint32_t)BIN_INV;
        st_codes(binp + r * bin_stride, code);
      }
    };

    // Block sums: the sum of every group of R consecutive box rows, each row touched once.
    // A window of S rows is then S / R block sums + S % R single rows instead of S terms.
    Pack<float, V> own[R];
    {
      Pack<Acc, V> b;
#pragma unroll
      for (int v = 0; v < V; ++v) b.a[v] = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        own[r] = ldp<float, V>(X + r * CW);
#pragma unroll
        for (int v = 0; v < V; ++v) b.a[v] += (Acc)own[r].a[v];
      }
      stp<Acc, V>(bs + warp * CW + lane * V, b);
      const int nw = blockDim.x >> 5;
      for (int blk = nw + warp; blk < n_blk; blk += nw) {  // the halo rows behind the last sub-strip
        const float* Xb = xs + st * stage_elems + blk * R * CW + lane * V;
        Pack<Acc, V> e;
#pragma unroll
        for (int v = 0; v < V; ++v) e.a[v] = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (blk * R + r < rows_box) {
            const Pack<float, V> xr = ldp<float, V>(Xb + r * CW);
#pragma unroll
            for (int v = 0; v < V; ++v) e.a[v] += (Acc)xr.a[v];
          }
        stp<Acc, V>(bs + blk * CW + lane * V, e);
      }
    }
    __syncthreads();

    if (rbase + R <= nd && t0 - off >= 0 && t0 + (R - 1) - off + S <= Tn) {
      // ---- whole sub-strip inside the year and the series: no per-day checks ----
      if (t0 == 0 && live) {  // only reachable when S == 1
        const Pack<float, V> x0 = ldp<float, V>(X + off * CW);
#pragma unroll
        for (int v = 0; v < V; ++v) p.mask0[c + v] = is_finite_f(x0.a[v]) ? 1 : 0;
      }
      const float* Xhi = X + (S - 1) * CW;
      const float* Xc = X + off * CW;
      const int nfull = S / R;
      Acc ws[V];
      if (nfull == 0) {  // S < R: the window is a prefix of the own block
#pragma unroll
        for (int v = 0; v < V; ++v) ws[v] = 0;
        for (int k = 0; k < S; ++k) {
          const Pack<float, V> xk = ldp<float, V>(X + k * CW);
#pragma unroll
          for (int v = 0; v < V; ++v) ws[v] += (Acc)xk.a[v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          ws[v] = (Acc)own[0].a[v];
#pragma unroll
          for (int r = 1; r < R; ++r) ws[v] += (Acc)own[r].a[v];
        }
        for (int b = 1; b < nfull; ++b) {
          const Pack<Acc, V> bb = ldp<Acc, V>(bs + (warp + b) * CW + lane * V);
#pragma unroll
          for (int v = 0; v < V; ++v) ws[v] += bb.a[v];
        }
        for (int k = nfull * R; k < S; ++k) {
          const Pack<float, V> xk = ldp<float, V>(X + k * CW);
#pragma unroll
          for (int v = 0; v < V; ++v) ws[v] += (Acc)xk.a[v];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r > 0) {
          const Pack<float, V> xh = ldp<float, V>(Xhi + r * CW);
#pragma unroll
          for (int v = 0; v < V; ++v) ws[v] += (Acc)xh.a[v] - (Acc)own[r - 1].a[v];
        }
        const Pack<float, V> xv = ldp<float, V>(Xc + r * CW);
#pragma unroll
        for (int v = 0; v < V; ++v) bad[v] += is_finite_f(xv.a[v]) ? 0 : 1;
        if (target) {
          emit(r, xv);
          outp += p.out_pitch;
        }
        Pack<float, V> s;
#pragma unroll
        for (int v = 0; v < V; ++v) s.a[v] = (float)(ws[v] * invS);
        turnover(r, s, true);
      }
    } else {
      // ---- series edges / last days of the year: checked path ----
      Acc ws[V];
#pragma unroll
      for (int v = 0; v < V; ++v) ws[v] = 0;
      bool have = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int t = t0 + r;
        if (rbase + r < nd && t >= 0 && t < Tn) {
          const Pack<float, V> xv = ldp<float, V>(X + (r + off) * CW);
          if (t == 0 && live) {
#pragma unroll
            for (int v = 0; v < V; ++v) p.mask0[c + v] = is_finite_f(xv.a[v]) ? 1 : 0;
          }
          Pack<float, V> s;
#pragma unroll
          for (int v = 0; v < V; ++v) s.a[v] = CUDART_NAN_F;
          const bool win = t - off >= 0 && t - off + S <= Tn;  // full window inside the series (min_periods = S)
          if (win) {
            if (have) {
              const Pack<float, V> xh = ldp<float, V>(X + (r + S - 1) * CW), xl = ldp<float, V>(X + (r - 1) * CW);
#pragma unroll
              for (int v = 0; v < V; ++v) ws[v] += (Acc)xh.a[v] - (Acc)xl.a[v];
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v) ws[v] = 0;
              for (int k = 0; k < S; ++k) {
                const Pack<float, V> xk = ldp<float, V>(X + (r + k) * CW);
#pragma unroll
                for (int v = 0; v < V; ++v) ws[v] += (Acc)xk.a[v];
              }
              have = true;
            }
#pragma unroll
            for (int v = 0; v < V; ++v) s.a[v] = (float)(ws[v] * invS);
          } else {
            have = false;
          }
#pragma unroll
          for (int v = 0; v < V; ++v) bad[v] += is_finite_f(xv.a[v]) ? 0 : 1;
          if (target) emit(r, xv);
          turnover(r, s, win);
        } else {
          have = false;  // (year, day) without a sample: year i - W still has to leave the ring
          if (target) emit_missing(r);
          Pack<float, V> s;
#pragma unroll
          for (int v = 0; v < V; ++v) s.a[v] = CUDART_NAN_F;
          turnover(r, s, false);
        }
        outp += p.out_pitch;
      }
    }
    __syncthreads();  // everyone is done with stage `st`
    if (threadIdx.x == 0 && issue_year < p.n_years) {
      fence_proxy_async();
      issue();
    }
    if (++st == NST) st = 0;
    if (++slot == W) slot = 0;
  }
  if (live) {
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (bad[v]) atomicAdd(&p.nonfinite[c + v], bad[v]);
  }
}

// Gridpoints whose series mixes finite and non-finite values: list[1 + k] = cell, list[0] = count.
__global__ void collect_dirty_kernel(const int32_t* __restrict__ nonfinite, int64_t N, int64_t T, int32_t* list) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int32_t b = nonfinite[c];
  if (b > 0 && b < T) list[1 + atomicAdd(&list[0], 1)] = (int32_t)c;
}

// Bin codes of the listed gridpoints, recomputed from the anomalies the fix-up kernel wrote (daily calendar:
// output row j is day (doy_first + j) of the calendar, years counted from year_first).
__global__ void __launch_bounds__(128) redigitize_cells_kernel(const float* __restrict__ anom, int64_t T_out,
                                                               int64_t anom_pitch, const int32_t* __restrict__ list,
                                                               const float* __restrict__ edges, int n_edges,
                                                               uint16_t* __restrict__ bins, int64_t bins_pitch, int NY,
                                                               int year_first) {
  extern __shared__ float s_edges[];
  DigTable dig;
  dig.init_cta(edges, n_edges, s_edges);
  const int n = list[0];
  // output row 0 is Jan 1 of year_first (the trim keeps whole years)
  for (int k = blockIdx.y; k < n; k += gridDim.y) {
    const int64_t c = list[1 + k];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < T_out; j += (int64_t)gridDim.x * blockDim.x) {
      // (year index, day of year) of output row j
      int yi = (int)(j / 366);
      int base = days_before_year(year_first + yi) - days_before_year(year_first);
      while (base > j) { --yi; base = days_before_year(year_first + yi) - days_before_year(year_first); }
      while (days_before_year(year_first + yi + 1) - days_before_year(year_first) <= j) {
        ++yi;
        base = days_before_year(year_first + yi) - days_before_year(year_first);
      }
      const int d = (int)(j - base);
      bins[((int64_t)d * NY + yi) * bins_pitch + c] = (uint16_t)dig(anom[j * anom_pitch + c]);
    }
  }
}

}  // namespace marex

using namespace marex;

namespace {
struct ShiftEnv {
  int v, r, nw, cps;
  bool f64;
};
// tuning knobs (marex_tune / MAREX_SHIFT_V, _R, _NW, _CPS, _F64): not part of the API
ShiftEnv shift_env() {
  ShiftEnv x;
  x.v = (int)tune_get("shift_v", 0); x.r = (int)tune_get("shift_r", 0); x.nw = (int)tune_get("shift_nw", 0);
  x.cps = (int)tune_get("shift_cps", 0);
  x.f64 = tune_get("shift_f64", 0) != 0;
  return x;
}
}  // namespace

extern "C" int marex_shift_anomaly_daily_f32(const float* x, int64_t T, int64_t N, int64_t pitch, int32_t year0,
                                             int32_t doy0, int32_t W, int32_t S, int32_t mode, float* out,
                                             int64_t out_pitch, uint8_t* mask0, int32_t* nonfinite, const float* edges,
                                             int32_t n_edges, uint16_t* bins, int64_t bins_pitch, void* stream) {
  MAREX_REQUIRE(x && out && mask0 && nonfinite, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && out_pitch >= N, "bad shape");
  MAREX_REQUIRE(W >= 1 && S >= 1 && doy0 >= 1 && doy0 <= 366 && year0 >= 1, "bad calendar or window");
  MAREX_REQUIRE((N % 4) == 0 && (pitch % 4) == 0 && (out_pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) % 16) == 0,
                "TMA path needs 16-byte aligned bases, N and the pitches multiples of 4 elements");
  MAREX_REQUIRE(T < (1LL << 31) && N < (1LL << 31), "T and N must fit int32 TMA coordinates");
  const bool digit = bins != nullptr;
  MAREX_REQUIRE(!digit || (edges && n_edges >= 3 && n_edges <= 4096 && mode == 0 && (bins_pitch % 4) == 0 &&
                           bins_pitch >= N && (reinterpret_cast<uintptr_t>(bins) % 8) == 0),
                "fused digitize needs an edge table, mode 0 and an 8-byte aligned bin array with a pitch multiple of 4");
  if (W > 31) return fail(MAREX_ERR_UNSUPPORTED, "window_year_baseline > 31: use the generic kernel");
  cudaStream_t st = (cudaStream_t)stream;
  DailyParams p;
  p.T = (int)T; p.N = (int)N; p.out_pitch = out_pitch;
  p.year0 = year0; p.doy0 = doy0 - 1;
  p.W = W; p.S = S;
  p.out = out; p.mask0 = mask0; p.nonfinite = nonfinite;
  p.edges = edges; p.n_edges = digit ? n_edges : 0; p.bins = bins; p.bins_pitch = bins_pitch;
  // years covered by T daily rows starting at (year0, doy0); row of Jan 1 of year index W
  int64_t base = -(int64_t)(doy0 - 1), base_w = -1;
  int n_years = 0;
  while (base < T) {
    if (n_years == W) base_w = base;
    base += is_leap(year0 + n_years) ? 366 : 365;
    ++n_years;
  }
  if (n_years > SD_MAX_YEARS) return fail(MAREX_ERR_UNSUPPORTED, "too many years for the daily kernel");
  p.n_years = n_years;
  p.NY = n_years - W;
  MAREX_REQUIRE(!digit || p.NY >= 1, "fused digitize needs at least one output year");
  p.out_off = mode ? 0 : (int)(base_w >= 0 ? base_w : T);
  MAREX_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t) * N, st));

  // Strip length D = R * NW days.  Shared memory = ring (W * D rows) + NST staged boxes (D + S - 1
  // rows each) of 128 * V bytes per row; CPS co-resident CTAs split the 227 KB.
  const ShiftEnv env = shift_env();
  const size_t fixed = 128 + (((size_t)(W + 1) * 8 + 127) / 128) * 128 + (((size_t)(n_years + 1) * 4 + 127) / 128) * 128 +
                       (digit ? (((size_t)n_edges * 4 + 127) / 128) * 128 : 0);
  constexpr int MAREX_OK_SPECIALISE = 1;
  bool try_special = !env.v && !env.r && !env.nw && !env.f64 && !tune_get("shift_generic", 0);
  auto launch = [&](auto kern, int V, int R, int nst, int cps, int acc_bytes) -> int {
    const size_t row = (size_t)128 * V;
    auto smem_of = [&](int D) {
      return fixed + (size_t)nst * (D + S - 1) * row + (size_t)W * D * row +
             (size_t)((D + S - 1 + R - 1) / R) * 32 * V * acc_bytes;
    };
    const size_t budget = (size_t)(227 * 1024) / cps - (cps > 1 ? 1024 : 0);
    int NW = env.nw ? std::min(env.nw, MAREX_SD_MAXTHREADS / 32) : MAREX_SD_MAXTHREADS / 32;  // launch bounds of the kernel
    for (; NW >= 1; --NW) {
      const int D = R * NW;
      if (D + S - 1 <= 256 && D <= NDOY + R && smem_of(D) <= budget) break;
    }
    if (NW < 1) return MAREX_ERR_UNSUPPORTED;
    // even out the strips: the fewest strips this R allows, then the smallest NW that still gives that count
    const int n_strips = (NDOY + R * NW - 1) / (R * NW);
    while (!env.nw && NW > 1 && (NDOY + R * (NW - 1) - 1) / (R * (NW - 1)) == n_strips) --NW;
    p.D = R * NW;
    p.rows_box = p.D + S - 1;
    p.n_strips = n_strips;
    if (try_special && V == 1 && R == 4 && nst == 2 && acc_bytes == 4 && !digit && S == 21 && W == 15 && p.D == 48) {
      try_special = false;  // the caller launches the instantiation with these constants instead
      return MAREX_OK_SPECIALISE;
    }
    const size_t smem = smem_of(p.D);
    CUtensorMap tmap;
    int rc = make_tmap_2d(&tmap, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N, pitch, p.rows_box, 32 * V);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(shift_daily)");
    const int64_t n_cta = ((N + 32 * V - 1) / (32 * V)) * n_strips;
    if (n_cta >= (1LL << 31)) return fail(MAREX_ERR_UNSUPPORTED, "grid too large");
    kern<<<(unsigned)n_cta, NW * 32, smem, st>>>(tmap, p);
    MAREX_LAUNCH_CHECK("shift_daily_kernel");
    return MAREX_OK;
  };
  const bool f64 = env.f64;
#define MAREX_SD3(V_, R_, CPS_, ACC_)                                                                  \
  (mode ? launch(shift_daily_kernel<V_, R_, 2, 1, ACC_, false>, V_, R_, 2, CPS_, (int)sizeof(ACC_))    \
        : (digit ? launch(shift_daily_kernel<V_, R_, 2, 0, ACC_, true>, V_, R_, 2, CPS_, (int)sizeof(ACC_)) \
                 : launch(shift_daily_kernel<V_, R_, 2, 0, ACC_, false>, V_, R_, 2, CPS_, (int)sizeof(ACC_))))
#define MAREX_SD(V_, R_, CPS_) (f64 ? MAREX_SD3(V_, R_, CPS_, double) : MAREX_SD3(V_, R_, CPS_, float))
  int rc = MAREX_ERR_UNSUPPORTED;
  if (env.v || env.r) {
    const int v = env.v ? env.v : 1, r = env.r ? env.r : 4, cps = env.cps ? env.cps : 2;
    if (v >= 2) rc = r == 2 ? MAREX_SD(2, 2, cps) : MAREX_SD(2, 4, cps);
    else rc = r == 1 ? MAREX_SD(1, 1, cps) : (r == 2 ? MAREX_SD(1, 2, cps) : MAREX_SD(1, 4, cps));
  } else {
    // measured on B200 (0.25 deg, W = 15, S = 21, fused digitize; profiles/r02_shift_shapes.json): V = 1, R = 4, two CTAs
    // per SM 65.9 ms; V = 2: 68.9 (R = 2) / 91.3 (R = 4); V = 4: 107 - 165 ms (removed) -- the ring's 60 bytes per
    // (day, gridpoint) cap the resident threads, and fewer, fatter threads lose more to latency than they save in
    // instructions.
    rc = MAREX_SD(1, 4, 2);
    if (rc == MAREX_OK_SPECIALISE) {  // the reference's default windows on the default shape: constants folded
      rc = mode ? launch(shift_daily_kernel<1, 4, 2, 1, float, false, 21, 15, 48>, 1, 4, 2, 2, 4)
                : launch(shift_daily_kernel<1, 4, 2, 0, float, false, 21, 15, 48>, 1, 4, 2, 2, 4);
    }
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(2, 2, 2);
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(1, 4, 1);
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(1, 1, 1);
  }
#undef MAREX_SD
#undef MAREX_SD3
  if (rc == MAREX_ERR_UNSUPPORTED)
    return fail(rc, "window_year_baseline / smooth_days_baseline too large for the shared-memory ring");
  return rc;
}

extern "C" int marex_shift_anomaly_fixup_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* tidx,
                                             const int32_t* year_val, int32_t n_years, int32_t W, int32_t S,
                                             const int32_t* out_row, int32_t mode, float* anom, int64_t anom_pitch,
                                             uint8_t* mask0, int32_t* nonfinite, int32_t* work, const float* edges,
                                             int32_t n_edges, uint16_t* bins, int64_t bins_pitch, int64_t T_out,
                                             int32_t year_first, void* stream) {
  MAREX_REQUIRE(x && tidx && year_val && out_row && anom && mask0 && nonfinite && work, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && anom_pitch >= N && n_years > 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  MAREX_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t), st));
  collect_dirty_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(nonfinite, N, T, work);
  MAREX_LAUNCH_CHECK("collect_dirty_kernel");
  int rc = launch_shift_generic(x, T, N, pitch, tidx, year_val, n_years, W, S, out_row, mode, anom, anom_pitch, mask0,
                                nonfinite, work + 1, work, 2 * sm_count(), st);
  if (rc || !bins) return rc;
  MAREX_REQUIRE(edges && n_edges >= 3 && mode == 0 && T_out > 0 && n_years > W, "bad fused-digitize arguments");
  redigitize_cells_kernel<<<dim3(8, (unsigned)(2 * sm_count())), 128, n_edges * sizeof(float), st>>>(
      anom, T_out, anom_pitch, work, edges, n_edges, bins, bins_pitch, n_years - W, year_first);
  MAREX_LAUNCH_CHECK("redigitize_cells_kernel");
  return MAREX_OK;
}
