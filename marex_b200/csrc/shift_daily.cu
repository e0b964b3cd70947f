// (a) shifting baseline, fast path for a gap-free daily calendar -- reference: detect.py:1691-1816
// (smoothed_rolling_climatology), 1511-1688 (rolling_climatology), 1819-1850, trim 615-641.
//
// CTA = 32 adjacent gridpoints (one 128-byte row segment) x one strip of D days of year.
// The CTA walks the calendar years in order.  For year i a single TMA box load
// (cp.async.bulk.tensor.2d, D + S - 1 rows x 32 cells) stages the strip's rows plus the
// smoothing halo in shared memory, NST years ahead of the arithmetic, so HBM latency is hidden
// by the copy engine and not by resident warps.  Thread = (gridpoint, R consecutive days):
//   * the S-day centred window sum is assembled from float64 sums of R-row blocks (each staged row
//     is converted once per block) and then slides along the R days; float64 sums of float32 data
//     are exact, so the grouping does not change the result,
//   * a ring in shared memory keeps the smoothed value of the last W years for every
//     (day, gridpoint) of the strip; the float64 running sum of the ring and its valid count sit
//     in registers, so clim[year, doy] = sum / count costs one multiply,
//   * the anomaly row is written straight out (one 128-byte segment per warp and day).
// Every input element is read from HBM (D + S - 1) / D times (L2 catches most of the halo) and
// every output element is written once.
//
// A running sum cannot un-add a NaN/inf, so this kernel is only exact for gridpoints whose
// series is all finite or all NaN (land).  It counts the non-finite inputs per gridpoint (the
// same numbers _validate_data_values needs, detect.py:205-279); marex_shift_anomaly_fixup_f32
// then recomputes the few gridpoints with 0 < count < T with the generic kernel (anomaly.cu).
#include <cstdlib>
#include <type_traits>

#include "tma.cuh"

namespace marex {

struct DailyParams {
  int64_t T, N, out_pitch, out_off;  // output row of input row t is t - out_off
  int year0, doy0;                   // calendar year and 0-based day of year of row 0
  int n_years, W, S, D, rows_box, n_strips;
  float* out;
  uint8_t* mask0;
  int32_t* nonfinite;
  uint32_t leap_bits[8];  // bit i: year index i is a leap year (first 256 years; only the LEAN instantiations read it)
};

__host__ __device__ __forceinline__ bool is_leap(int y) { return (y % 4 == 0 && y % 100 != 0) || (y % 400 == 0); }

// Running sum of the ring of one (day, gridpoint).  double: float64 sums of float32 data are exact, the result is rounded
// once (what the oracle does).  float (MAREX_SHIFT_ACC=f32, a round-2 experiment, see tests/test_f32_accumulation_study.py):
// Kahan-compensated float32; measured against the oracle on the CPU: max error 2e-7 of the field scale over 41 years,
// 50 times below the 1e-5 bar, and no F2F / DADD on the XU and FP64 pipes.
template <typename Acc>
struct RingSum;
template <>
struct RingSum<double> {
  double s = 0.0;
  __device__ __forceinline__ void add(float v) { s += (double)v; }
  __device__ __forceinline__ void sub(float v) { s -= (double)v; }
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
  __device__ __forceinline__ void sub(float v) { add(-v); }
  __device__ __forceinline__ float value() const { return s; }
};

// LEAN (MAREX_SHIFT_LEAN=1, a round-2 experiment): the two items of the source-level profile that are pure overhead --
// `is_leap` by integer modulo per thread and year (6.4 % of the instructions) becomes a bit test on a host-built mask in
// the parameters, and the row arithmetic is 32-bit (T < 2^31 is required by the TMA coordinates anyway).
template <bool LEAN>
__device__ __forceinline__ bool year_is_leap(const DailyParams& p, int i) {
  if (LEAN && i < 256) return (p.leap_bits[i >> 5] >> (i & 31)) & 1u;
  return is_leap(p.year0 + i);
}

template <int R, int NST, int MODE, typename Acc = double, bool LEAN = false>
__global__ void __launch_bounds__(R >= 12 ? 256 : 512) shift_daily_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                         const DailyParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  using Row = typename std::conditional<LEAN, int, int64_t>::type;  // input row indices
  const Row Tn = (Row)p.T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, W = p.W, S = p.S, off = p.S / 2;
  // shared memory carve-up (all offsets multiples of 128 bytes)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                         // [NST]
  Acc* invtab = reinterpret_cast<Acc*>(smem_raw + 128);                          // [W + 1], invtab[0] = NaN
  const size_t inv_bytes = (((size_t)(W + 1) * 8 + 127) / 128) * 128;
  float* xs = reinterpret_cast<float*>(smem_raw + 128 + inv_bytes);              // [NST][rows_box][32]
  const int stage_elems = p.rows_box * 32;
  float* ring = xs + (size_t)NST * stage_elems;                                  // [W][D][32]
  Acc* bs = reinterpret_cast<Acc*>(ring + (size_t)W * D * 32);                  // [n_blk][32] sums of R box rows
  const int n_blk = (p.rows_box + R - 1) / R;

  // strips of one 32-gridpoint group are adjacent CTAs: they run at the same time and at the
  // same pace, so the S - 1 halo rows a strip shares with its neighbour are L2 hits
  const int strip = blockIdx.x % p.n_strips;
  const int64_t c0 = (int64_t)(blockIdx.x / p.n_strips) * 32;
  const int64_t c = c0 + lane;
  const bool live = c < p.N;
  const int d0 = strip * D;
  const int rbase = warp * R;
  const uint32_t box_bytes = (uint32_t)p.rows_box * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i <= W; i += blockDim.x) invtab[i] = i ? (Acc)(1.0 / (double)i) : (Acc)CUDART_NAN;
  for (int i = threadIdx.x; i < W * D * 32; i += blockDim.x) ring[i] = CUDART_NAN_F;
  __syncthreads();

  // producer state (thread 0): first row of the box of the next year to issue
  int issue_year = 0, issue_st = 0;
  Row issue_base = -(Row)p.doy0;  // row index of day-of-year 0 of year `issue_year`
  auto issue = [&]() {  // a box that lies entirely outside the series is neither loaded nor waited for
    const Row tb = issue_base + d0 - off;
    if (tb < Tn && tb + p.rows_box > 0) {
      mbar_expect_tx(&bar[issue_st], box_bytes);
      tma_load_2d(xs + (size_t)issue_st * stage_elems, &tmap, (int)c0, (int)tb, &bar[issue_st]);
    }
    issue_base += year_is_leap<LEAN>(p, issue_year) ? 366 : 365;
    ++issue_year;
    if (++issue_st == NST) issue_st = 0;
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int k = 0; k < NST && k < p.n_years; ++k) issue();
  }

  RingSum<Acc> sum[R];
  int cnt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) cnt[r] = 0;
  int bad = 0;
  const Acc invS = (Acc)(1.0 / (double)S);
  Row base = -(Row)p.doy0;
  uint32_t phase = 0;  // bit st = parity of the next completion of stage st
  int st = 0, slot = 0;
  const float* Xst = xs + rbase * 32 + lane;          // this thread's first window row in stage 0
  float* ringp = ring + (size_t)rbase * 32 + lane;    // this thread's first day in ring slot 0
  float* const outc = p.out + c;                      // column of this gridpoint (dead lanes never store)

  for (int i = 0; i < p.n_years; ++i) {
    const int ylen = year_is_leap<LEAN>(p, i) ? 366 : 365;
    const float* X = Xst + st * stage_elems;          // X[j * 32]: row rbase + j of the box
    float* ringslot = ringp + slot * (D * 32);
    const int nd = min(D, ylen - d0);                 // days of this strip that exist in year i
    const Row t0 = base + d0 + rbase;                 // input row of this thread's first day
    const bool target = i >= W;
    {
      const Row tb = base + d0 - off;
      if (tb < Tn && tb + p.rows_box > 0) {
        mbar_wait(&bar[st], (phase >> st) & 1u);
        phase ^= 1u << st;
      }
    }
    float* outp = outc + (int64_t)(t0 - (Row)p.out_off) * p.out_pitch;  // only dereferenced for target years

    // ring turnover of one (day, gridpoint): year i - W leaves, year i enters
    auto turnover = [&](int r, float s) {
      const float old = ringslot[r * 32];
      if (old == old) { sum[r].sub(old); --cnt[r]; }
      ringslot[r * 32] = s;
      if (s == s) { sum[r].add(s); ++cnt[r]; }
    };
    auto emit = [&](int r, float xv) {  // anomaly (or climatology) of a target-year day
      const float clim = (float)(sum[r].value() * invtab[cnt[r]]);
      if (live) st_stream(outp, MODE ? clim : xv - clim);
    };

    // Block sums: the float64 sum of every group of R consecutive box rows, each row converted once.
    // A window of S rows is then S / R block sums + S % R single rows instead of S conversions
    // (the sums are exact for float32 data, so the grouping does not change the result).
    Acc own[R];
    {
      Acc b = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) { own[r] = (Acc)X[r * 32]; b += own[r]; }
      bs[warp * 32 + lane] = b;
      const int nw = blockDim.x >> 5;
      for (int blk = nw + warp; blk < n_blk; blk += nw) {  // the halo rows behind the last sub-strip
        const float* Xb = xs + st * stage_elems + blk * R * 32 + lane;
        Acc e = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) if (blk * R + r < p.rows_box) e += (Acc)Xb[r * 32];
        bs[blk * 32 + lane] = e;
      }
    }
    __syncthreads();

    if (rbase + R <= nd && t0 - off >= 0 && t0 + (R - 1) - off + S <= Tn) {
      // ---- whole sub-strip inside the year and the series: no per-day checks ----
      if (t0 == 0 && live) p.mask0[c] = is_finite_f(X[off * 32]) ? 1 : 0;  // only reachable when S == 1
      const float* Xhi = X + (S - 1) * 32;
      const float* Xc = X + off * 32;
      const int nfull = S / R;
      Acc ws = own[0];
#pragma unroll
      for (int r = 1; r < R; ++r) ws += own[r];
      if (nfull == 0) {  // S < R: the window is a prefix of the own block
        ws = 0;
        for (int k = 0; k < S; ++k) ws += (Acc)X[k * 32];
      } else {
        for (int b = 1; b < nfull; ++b) ws += bs[(warp + b) * 32 + lane];
        for (int k = nfull * R; k < S; ++k) ws += (Acc)X[k * 32];
      }
      if (target) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r > 0) ws += (Acc)Xhi[r * 32] - own[r - 1];
          const float xv = Xc[r * 32];
          bad += is_finite_f(xv) ? 0 : 1;
          emit(r, xv);
          outp += p.out_pitch;
          turnover(r, (float)(ws * invS));
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r > 0) ws += (Acc)Xhi[r * 32] - own[r - 1];
          bad += is_finite_f(Xc[r * 32]) ? 0 : 1;
          turnover(r, (float)(ws * invS));
        }
      }
    } else {
      // ---- series edges / last days of the year: checked path ----
      Acc ws = 0;
      bool have = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const Row t = t0 + r;
        if (rbase + r < nd && t >= 0 && t < Tn) {
          const float xv = X[(r + off) * 32];
          if (t == 0 && live) p.mask0[c] = is_finite_f(xv) ? 1 : 0;
          float s = CUDART_NAN_F;
          if (t - off >= 0 && t - off + S <= Tn) {  // full window inside the series (min_periods = S)
            if (have) {
              ws += (Acc)X[(r + S - 1) * 32] - (Acc)X[(r - 1) * 32];
            } else {
              ws = 0;
              for (int k = 0; k < S; ++k) ws += (Acc)X[(r + k) * 32];
              have = true;
            }
            s = (float)(ws * invS);
          } else {
            have = false;
          }
          bad += is_finite_f(xv) ? 0 : 1;
          if (target) emit(r, xv);
          turnover(r, s);
        } else {
          have = false;  // (year, day) without a sample: year i - W still has to leave the ring
          const float old = ringslot[r * 32];
          if (old == old) { sum[r].sub(old); --cnt[r]; ringslot[r * 32] = CUDART_NAN_F; }
        }
        outp += p.out_pitch;
      }
    }
    __syncthreads();  // everyone is done with stage `st`
    if (threadIdx.x == 0 && issue_year < p.n_years) {
      fence_proxy_async();
      issue();
    }
    base += ylen;
    if (++st == NST) st = 0;
    if (++slot == W) slot = 0;
  }
  if (live && bad) atomicAdd(&p.nonfinite[c], bad);
}

// Gridpoints whose series mixes finite and non-finite values: list[1 + k] = cell, list[0] = count.
__global__ void collect_dirty_kernel(const int32_t* __restrict__ nonfinite, int64_t N, int64_t T, int32_t* list) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int32_t b = nonfinite[c];
  if (b > 0 && b < T) list[1 + atomicAdd(&list[0], 1)] = (int32_t)c;
}

}  // namespace marex

using namespace marex;

extern "C" int marex_shift_anomaly_daily_f32(const float* x, int64_t T, int64_t N, int64_t pitch, int32_t year0,
                                             int32_t doy0, int32_t W, int32_t S, int32_t mode, float* out,
                                             int64_t out_pitch, uint8_t* mask0, int32_t* nonfinite, void* stream) {
  MAREX_REQUIRE(x && out && mask0 && nonfinite, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && out_pitch >= N, "bad shape");
  MAREX_REQUIRE(W >= 1 && S >= 1 && doy0 >= 1 && doy0 <= 366, "bad calendar or window");
  MAREX_REQUIRE((pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0,
                "TMA path needs a 16-byte aligned base and a pitch that is a multiple of 4 elements");
  MAREX_REQUIRE(T < (1LL << 31) && N < (1LL << 31), "T and N must fit int32 TMA coordinates");
  cudaStream_t st = (cudaStream_t)stream;
  DailyParams p;
  p.T = T; p.N = N; p.out_pitch = out_pitch;
  p.year0 = year0; p.doy0 = doy0 - 1;
  p.W = W; p.S = S;
  p.out = out; p.mask0 = mask0; p.nonfinite = nonfinite;
  // years covered by T daily rows starting at (year0, doy0); row of Jan 1 of year index W
  int64_t base = -(int64_t)(doy0 - 1), base_w = -1;
  int n_years = 0;
  for (int k = 0; k < 8; ++k) p.leap_bits[k] = 0;
  while (base < T) {
    if (n_years == W) base_w = base;
    const bool leap = is_leap(year0 + n_years);
    if (leap && n_years < 256) p.leap_bits[n_years >> 5] |= 1u << (n_years & 31);
    base += leap ? 366 : 365;
    ++n_years;
  }
  p.n_years = n_years;
  p.out_off = mode ? 0 : (base_w >= 0 ? base_w : T);
  MAREX_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t) * N, st));

  // Strip length D = R * NW days.  Shared memory = ring (W * D rows) + NST staged boxes (D + S - 1
  // rows each) of 128 bytes per row; CTAS_PER_SM co-resident CTAs split the 227 KB.
  const int env_r = getenv("MAREX_SHIFT_R") ? atoi(getenv("MAREX_SHIFT_R")) : 0;
  const int env_nw = getenv("MAREX_SHIFT_NW") ? atoi(getenv("MAREX_SHIFT_NW")) : 0;
  const int env_nst = getenv("MAREX_SHIFT_NST") ? atoi(getenv("MAREX_SHIFT_NST")) : 0;
  const int env_cps = getenv("MAREX_SHIFT_CPS") ? atoi(getenv("MAREX_SHIFT_CPS")) : 0;
  const char* env_acc = getenv("MAREX_SHIFT_ACC");
  const bool acc_f32 = env_acc && std::string(env_acc) == "f32";  // experiment: float32 sums (Kahan ring), within 1e-5
  const bool lean = getenv("MAREX_SHIFT_LEAN") && atoi(getenv("MAREX_SHIFT_LEAN")) == 1;  // experiment, default shape only
  const size_t fixed = 128 + (((size_t)(W + 1) * 8 + 127) / 128) * 128;
  auto launch = [&](auto kern, int R, int nst, int cps) -> int {
    auto smem_of = [&](int D, int ns) {
      return fixed + (size_t)ns * (D + S - 1) * 128 + (size_t)W * D * 128 + (size_t)((D + S - 1 + R - 1) / R) * 256;
    };
    const size_t budget = (size_t)(227 * 1024) / cps - (cps > 1 ? 1024 : 0);
    const int nw_max = R >= 12 ? 8 : 16;
    int NW = env_nw ? env_nw : nw_max;
    for (; NW >= 1; --NW) {
      const int D = R * NW;
      if (D + S - 1 <= 256 && D <= NDOY + R && smem_of(D, nst) <= budget) break;
    }
    if (NW < 1) return MAREX_ERR_UNSUPPORTED;
    // even out the strips: the fewest strips this R allows, then the smallest NW that still gives that count
    const int n_strips = (NDOY + R * NW - 1) / (R * NW);
    while (!env_nw && NW > 1 && (NDOY + R * (NW - 1) - 1) / (R * (NW - 1)) == n_strips) --NW;
    p.D = R * NW;
    p.rows_box = p.D + S - 1;
    p.n_strips = n_strips;
    const size_t smem = smem_of(p.D, nst);
    CUtensorMap tmap;
    int rc = make_tmap_2d(&tmap, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N, pitch, p.rows_box, 32);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(shift_daily)");
    const int64_t n_cta = ((N + 31) / 32) * n_strips;
    if (n_cta >= (1LL << 31)) return fail(MAREX_ERR_UNSUPPORTED, "grid too large");
    kern<<<(unsigned)n_cta, NW * 32, smem, st>>>(tmap, p);
    MAREX_LAUNCH_CHECK("shift_daily_kernel");
    return MAREX_OK;
  };
#define MAREX_SD(R_, NST_, CPS_)                                                                                  \
  (acc_f32 ? (mode ? launch(shift_daily_kernel<R_, NST_, 1, float>, R_, NST_, CPS_)                              \
                   : launch(shift_daily_kernel<R_, NST_, 0, float>, R_, NST_, CPS_))                             \
           : (mode ? launch(shift_daily_kernel<R_, NST_, 1, double>, R_, NST_, CPS_)                             \
                   : launch(shift_daily_kernel<R_, NST_, 0, double>, R_, NST_, CPS_)))
  int rc = MAREX_ERR_UNSUPPORTED;
  if (env_r) {  // tuning knobs (MAREX_SHIFT_R / _NW / _NST / _CPS), not part of the API
    const int nst = env_nst ? env_nst : 2, cps = env_cps ? env_cps : 1;
    if (env_r == 12) rc = nst == 3 ? MAREX_SD(12, 3, cps) : MAREX_SD(12, 2, cps);
    else if (env_r == 6) rc = nst == 3 ? MAREX_SD(6, 3, cps) : MAREX_SD(6, 2, cps);
    else if (env_r == 4) rc = MAREX_SD(4, 2, cps);
    else if (env_r == 3) rc = MAREX_SD(3, 2, cps);
    else rc = MAREX_SD(1, 2, cps);
  } else {
    // measured on B200 (0.25 deg, W=15, S=21): two co-resident CTAs of 11 warps x 4 days: 56-57 ms; one CTA of
    // 8 x 12: 78-89 ms.  The arithmetic is latency-bound (8-22 warps per SM: the ring limits residency), the TMA
    // staging alone takes 21 ms and the output stores 7 ms.
    if (lean) {
      rc = acc_f32 ? (mode ? launch(shift_daily_kernel<4, 2, 1, float, true>, 4, 2, 2) : launch(shift_daily_kernel<4, 2, 0, float, true>, 4, 2, 2))
                   : (mode ? launch(shift_daily_kernel<4, 2, 1, double, true>, 4, 2, 2) : launch(shift_daily_kernel<4, 2, 0, double, true>, 4, 2, 2));
    } else {
      rc = MAREX_SD(4, 2, 2);
    }
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(6, 2, 2);
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(12, 2, 1);
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(4, 2, 1);
    if (rc == MAREX_ERR_UNSUPPORTED) rc = MAREX_SD(1, 2, 1);
  }
#undef MAREX_SD
  if (rc == MAREX_ERR_UNSUPPORTED)
    return fail(rc, "window_year_baseline / smooth_days_baseline too large for the shared-memory ring");
  return rc;
}

extern "C" int marex_shift_anomaly_fixup_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* tidx,
                                             const int32_t* year_val, int32_t n_years, int32_t W, int32_t S,
                                             const int32_t* out_row, int32_t mode, float* anom, int64_t anom_pitch,
                                             uint8_t* mask0, int32_t* nonfinite, int32_t* work, void* stream) {
  MAREX_REQUIRE(x && tidx && year_val && out_row && anom && mask0 && nonfinite && work, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && anom_pitch >= N && n_years > 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  MAREX_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t), st));
  collect_dirty_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(nonfinite, N, T, work);
  MAREX_LAUNCH_CHECK("collect_dirty_kernel");
  return launch_shift_generic(x, T, N, pitch, tidx, year_val, n_years, W, S, out_row, mode, anom, anom_pitch, mask0,
                              nonfinite, work + 1, work, 2 * sm_count(), st);
}
