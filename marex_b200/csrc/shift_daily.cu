// (a) shifting baseline, fast path for a gap-free daily calendar -- reference: detect.py:1691-1816
// (smoothed_rolling_climatology), 1511-1688 (rolling_climatology), 1819-1850, trim 615-641.
//
// CTA = 32 adjacent gridpoints (one 128-byte row segment) x one strip of D days of year.
// The CTA walks the calendar years in order.  For year i a single TMA box load
// (cp.async.bulk.tensor.2d, D + S - 1 rows x 32 cells) stages the strip's rows plus the
// smoothing halo in shared memory, NST years ahead of the arithmetic, so HBM latency is hidden
// by the copy engine and not by resident warps.  Thread = (gridpoint, R consecutive days):
//   * the S-day centred window sum slides along the R days in float64 (exact for float32 data),
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
#include "tma.cuh"

namespace marex {

struct DailyParams {
  int64_t T, N, out_pitch, out_off;  // output row of input row t is t - out_off
  int year0, doy0;                   // calendar year and 0-based day of year of row 0
  int n_years, W, S, D, rows_box, mode;
  float* out;
  uint8_t* mask0;
  int32_t* nonfinite;
};

__host__ __device__ __forceinline__ bool is_leap(int y) { return (y % 4 == 0 && y % 100 != 0) || (y % 400 == 0); }

template <int R, int NST>
__global__ void __launch_bounds__(512) shift_daily_kernel(const __grid_constant__ CUtensorMap tmap, const DailyParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, W = p.W, S = p.S, off = p.S / 2;
  // shared memory carve-up (all offsets multiples of 128 bytes)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                         // [NST]
  double* invtab = reinterpret_cast<double*>(smem_raw + 128);                    // [W + 1]
  const size_t inv_bytes = (((size_t)(W + 1) * 8 + 127) / 128) * 128;
  float* xs = reinterpret_cast<float*>(smem_raw + 128 + inv_bytes);              // [NST][rows_box][32]
  float* ring = xs + (size_t)NST * p.rows_box * 32;                              // [W][D][32]

  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int64_t c = c0 + lane;
  const bool live = c < p.N;
  const int d0 = blockIdx.y * D;
  const int rbase = warp * R;
  const uint32_t box_bytes = (uint32_t)p.rows_box * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i <= W; i += blockDim.x) invtab[i] = i ? 1.0 / (double)i : 0.0;
  for (int i = threadIdx.x; i < W * D * 32; i += blockDim.x) ring[i] = CUDART_NAN_F;
  __syncthreads();

  // producer state (thread 0): first row of the box of the next year to issue
  int issue_year = 0;
  int64_t issue_base = -(int64_t)p.doy0;  // row index of day-of-year 0 of year `issue_year`
  auto issue = [&]() {  // a box that lies entirely outside the series is neither loaded nor waited for
    const int st = issue_year % NST;
    const int64_t tb = issue_base + d0 - off;
    if (tb < p.T && tb + p.rows_box > 0) {
      mbar_expect_tx(&bar[st], box_bytes);
      tma_load_2d(xs + (size_t)st * p.rows_box * 32, &tmap, (int)c0, (int)tb, &bar[st]);
    }
    issue_base += is_leap(p.year0 + issue_year) ? 366 : 365;
    ++issue_year;
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int k = 0; k < NST && k < p.n_years; ++k) issue();
  }

  double sum[R];
  int cnt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { sum[r] = 0.0; cnt[r] = 0; }
  int bad = 0;
  const double invS = 1.0 / (double)S;
  int64_t base = -(int64_t)p.doy0;
  uint32_t phase = 0;  // bit st = parity of the next completion of stage st

  for (int i = 0; i < p.n_years; ++i) {
    const int ylen = is_leap(p.year0 + i) ? 366 : 365;
    const int st = i % NST;
    const float* X = xs + (size_t)st * p.rows_box * 32 + lane;  // X[j * 32]: row j of the box, this lane's column
    float* ringslot = ring + ((size_t)(i % W) * D + rbase) * 32 + lane;
    const int nd = min(D, ylen - d0);                 // days of this strip that exist in year i
    const int64_t t0 = base + d0 + rbase;             // input row of this thread's first day
    const bool target = i >= W;
    {
      const int64_t tb = base + d0 - off;
      if (tb < p.T && tb + p.rows_box > 0) {
        mbar_wait(&bar[st], (phase >> st) & 1u);
        phase ^= 1u << st;
      }
    }

    auto finish_day = [&](int r, int64_t t, float xv, float s) {
      // anomaly of a target-year day against the W previous years, then ring turnover
      if (!is_finite_f(xv)) ++bad;
      if (target) {
        const float clim = cnt[r] ? (float)(sum[r] * invtab[cnt[r]]) : CUDART_NAN_F;
        if (live) st_stream(&p.out[(t - p.out_off) * p.out_pitch + c], p.mode ? clim : xv - clim);
      }
      const float old = ringslot[r * 32];
      if (old == old) { sum[r] -= (double)old; --cnt[r]; }
      ringslot[r * 32] = s;
      if (s == s) { sum[r] += (double)s; ++cnt[r]; }
    };
    auto expire_only = [&](int r) {  // (year, day) without a sample: year i - W still has to leave the ring
      const float old = ringslot[r * 32];
      if (old == old) { sum[r] -= (double)old; --cnt[r]; ringslot[r * 32] = CUDART_NAN_F; }
    };

    if (rbase + R <= nd && t0 - off >= 0 && t0 + (R - 1) - off + S <= p.T) {
      // ---- whole sub-strip inside the year and the series: no per-day checks ----
      if (t0 <= 0 && live) {  // only reachable when S == 1
        if (t0 == 0) p.mask0[c] = is_finite_f(X[(rbase + off) * 32]) ? 1 : 0;
      }
      double ws = 0.0;
#pragma unroll 7
      for (int k = 0; k < S; ++k) ws += (double)X[(rbase + k) * 32];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r > 0) ws += (double)X[(rbase + r + S - 1) * 32] - (double)X[(rbase + r - 1) * 32];
        finish_day(r, t0 + r, X[(rbase + r + off) * 32], (float)(ws * invS));
      }
    } else {
      // ---- series edges / last days of the year: checked path ----
      double ws = 0.0;
      bool have = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t t = t0 + r;
        if (rbase + r < nd && t >= 0 && t < p.T) {
          const float xv = X[(rbase + r + off) * 32];
          if (t == 0 && live) p.mask0[c] = is_finite_f(xv) ? 1 : 0;
          float s = CUDART_NAN_F;
          if (t - off >= 0 && t - off + S <= p.T) {  // full window inside the series (min_periods = S)
            if (have) {
              ws += (double)X[(rbase + r + S - 1) * 32] - (double)X[(rbase + r - 1) * 32];
            } else {
              ws = 0.0;
              for (int k = 0; k < S; ++k) ws += (double)X[(rbase + r + k) * 32];
              have = true;
            }
            s = (float)(ws * invS);
          } else {
            have = false;
          }
          finish_day(r, t, xv, s);
        } else {
          have = false;
          expire_only(r);
        }
      }
    }
    __syncthreads();  // everyone is done with stage `st`
    if (threadIdx.x == 0 && issue_year < p.n_years) {
      fence_proxy_async();
      issue();
    }
    base += ylen;
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
  p.W = W; p.S = S; p.mode = mode;
  p.out = out; p.mask0 = mask0; p.nonfinite = nonfinite;
  // years covered by T daily rows starting at (year0, doy0); row of Jan 1 of year index W
  int64_t base = -(int64_t)(doy0 - 1), base_w = -1;
  int n_years = 0;
  while (base < T) {
    if (n_years == W) base_w = base;
    base += is_leap(year0 + n_years) ? 366 : 365;
    ++n_years;
  }
  p.n_years = n_years;
  p.out_off = mode ? 0 : (base_w >= 0 ? base_w : T);
  MAREX_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t) * N, st));

  constexpr int NST = 2;
  const size_t budget = 225 * 1024;
  auto plan = [&](int R, int& NW, size_t& smem) -> bool {  // largest NW (<= 16) whose strip fits shared memory
    const size_t fixed = 128 + (((size_t)(W + 1) * 8 + 127) / 128) * 128;
    for (NW = 16; NW >= 1; --NW) {
      const int D = R * NW;
      const int rows = D + S - 1;
      if (rows > 256) continue;
      if (D > NDOY + R) continue;
      smem = fixed + (size_t)NST * rows * 128 + (size_t)W * D * 128;
      if (smem <= budget) return true;
    }
    return false;
  };
  auto launch = [&](auto kern, int R) -> int {
    int NW;
    size_t smem;
    if (!plan(R, NW, smem)) return MAREX_ERR_UNSUPPORTED;
    // even out the strips: the fewest strips this R allows, then the smallest NW that still gives that count
    const int n_strips = (NDOY + R * NW - 1) / (R * NW);
    while (NW > 1 && (NDOY + R * (NW - 1) - 1) / (R * (NW - 1)) == n_strips) --NW;
    p.D = R * NW;
    p.rows_box = p.D + S - 1;
    smem = 128 + (((size_t)(W + 1) * 8 + 127) / 128) * 128 + (size_t)NST * p.rows_box * 128 + (size_t)W * p.D * 128;
    CUtensorMap tmap;
    int rc = make_tmap_2d(&tmap, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N, pitch, p.rows_box, 32);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(shift_daily)");
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)n_strips);
    kern<<<grid, NW * 32, smem, st>>>(tmap, p);
    MAREX_LAUNCH_CHECK("shift_daily_kernel");
    return MAREX_OK;
  };
  int rc = launch(shift_daily_kernel<12, NST>, 12);
  if (rc == MAREX_ERR_UNSUPPORTED) rc = launch(shift_daily_kernel<4, NST>, 4);
  if (rc == MAREX_ERR_UNSUPPORTED) rc = launch(shift_daily_kernel<1, NST>, 1);
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
