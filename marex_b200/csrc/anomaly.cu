// Anomaly kernels: shifting baseline (a), fixed baseline (a'), polynomial detrend (a'').
// Layout: time-major fields, a warp's 32 lanes own 32 adjacent gridpoints, so every row
// access is one coalesced 128-byte segment.
#include "common.cuh"

namespace marex {

// =======================================================================================
// (a) shifting baseline  -- reference: detect.py:1511-1688, 1691-1816, 1819-1850
//
// Thread = (gridpoint c, strip of R consecutive days of year).  The thread sweeps the
// calendar years in order.  For its R days it keeps, per day,
//   * a ring of the last W smoothed values s[year, doy] in shared memory (slot = year % W),
//   * the running float64 sum of the finite ring entries and a packed valid/inf counter,
// so clim[year, doy] = sum / count is available without re-reading any earlier year.
// The S-day centred window sum slides along the strip (2 loads per day after the first
// window); neighbouring strips share their halo rows through L1/L2.
// =======================================================================================
template <int R>
__global__ void __launch_bounds__(256) shift_anomaly_kernel(
    const float* __restrict__ x, int64_t T, int64_t N, int64_t pitch,
    const int32_t* __restrict__ tidx, const int32_t* __restrict__ year_val, int n_years, int W, int S,
    const int32_t* __restrict__ out_row, float* __restrict__ anom, int64_t anom_pitch,
    uint8_t* __restrict__ mask0, int32_t* __restrict__ nonfinite, int n_strips, int spb, int mode,
    const int32_t* __restrict__ cell_list, const int32_t* __restrict__ n_list) {
  extern __shared__ float ring[];  // [W][R][spb][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = blockIdx.y * spb + warp;
  if (strip >= n_strips) return;  // whole warp leaves together
  // Cell groups: all of 0..N-1, or (fix-up pass of the daily kernel) the cells of `cell_list`,
  // walked grid-stride so the launch does not need the list length on the host.
  const int64_t n_cells = cell_list ? (int64_t)__ldg(n_list) : N;
  const int64_t n_groups = (n_cells + 31) / 32;
  const int d0 = strip * R;
  const int off = S / 2;
  const int first_target = year_val[0] + W;
  auto ring_at = [&](int slot, int r) -> float& { return ring[((slot * R + r) * spb + warp) * 32 + lane]; };

  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
  const int64_t idx = g * 32 + lane;
  const bool live = idx < n_cells;
  const int64_t c = cell_list ? (int64_t)__ldg(&cell_list[live ? idx : n_cells - 1]) : (live ? idx : N - 1);
  const int64_t cc = c;  // dead lanes read a valid column, never write

  double sum[R];
  uint32_t cnt[R];  // bits 0..9 valid (non-NaN) count, 10..19 +inf, 20..29 -inf
#pragma unroll
  for (int r = 0; r < R; ++r) { sum[r] = 0.0; cnt[r] = 0u; }
  int bad = 0;

  auto ring_update = [&](float s, int r, int sign) {
    if (s == s) {  // NaN contributes nothing to a nanmean
      const bool fin = is_finite_f(s);
      if (fin) sum[r] += sign * (double)s;
      const uint32_t code = fin ? 1u : (1u + ((s > 0.f) ? (1u << 10) : (1u << 20)));
      cnt[r] += sign > 0 ? code : (0u - code);
    }
  };

  if (strip == 0 && live && !cell_list) mask0[c] = is_finite_f(x[c]) ? 1 : 0;

  int rm = 0;  // oldest year index still in the ring
  for (int i = 0; i < n_years; ++i) {
    const int Ty = year_val[i];
    // (1) expire years older than Ty - W: they do not contribute to target year Ty
    while (rm < i && year_val[rm] < Ty - W) {
#pragma unroll
      for (int r = 0; r < R; ++r) ring_update(ring_at(rm % W, r), r, -1);
      ++rm;
    }
    const bool is_target = Ty >= first_target;
    // sliding S-day window state (float64 sum of finite samples + packed non-finite counts)
    double ws = 0.0;
    uint32_t wnf = 0u;
    int64_t wlo = -(1LL << 40);
    bool wvalid = false;
    float snew[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int d = d0 + r;
      const int u = (d < NDOY) ? __ldg(&tidx[i * NDOY + d]) : -1;
      float s = CUDART_NAN_F;
      if (u >= 0) {
        const float xu = __ldg(&x[(int64_t)u * pitch + cc]);
        if (!is_finite_f(xu)) ++bad;
        // (2) anomaly of a target-year sample against the W previous years
        if (is_target) {
          const int orow = __ldg(&out_row[u]);
          if (orow >= 0) {
            const uint32_t n = cnt[r] & 1023u;
            float clim;
            if (n == 0u) clim = CUDART_NAN_F;
            else if (cnt[r] >> 10) clim = nonfinite_result(cnt[r] & ~1023u);
            else clim = (float)(sum[r] / (double)n);
            if (live) st_stream(&anom[(int64_t)orow * anom_pitch + c], mode ? clim : xu - clim);
          }
        }
        // smoothed value of this sample: window rows [u - off, u - off + S - 1]
        const int64_t lo = (int64_t)u - off;
        if (lo >= 0 && lo + S <= T) {
          if (wvalid && lo == wlo + 1) {
            const float vo = __ldg(&x[wlo * pitch + cc]);
            const float vn = __ldg(&x[(wlo + S) * pitch + cc]);
            if (is_finite_f(vo)) ws -= (double)vo; else wnf -= nonfinite_code(vo);
            if (is_finite_f(vn)) ws += (double)vn; else wnf += nonfinite_code(vn);
          } else if (!(wvalid && lo == wlo)) {
            ws = 0.0;
            wnf = 0u;
#pragma unroll 7
            for (int k = 0; k < S; ++k) {
              const float v = __ldg(&x[(lo + k) * pitch + cc]);
              if (is_finite_f(v)) ws += (double)v; else wnf += nonfinite_code(v);
            }
          }
          wlo = lo;
          wvalid = true;
          s = wnf ? nonfinite_result(wnf) : (float)(ws / (double)S);
        }
      }
      snew[r] = s;
    }
    // (3) expire year Ty - W (it contributed to Ty but not to any later year) ...
    while (rm < i && year_val[rm] < Ty + 1 - W) {
#pragma unroll
      for (int r = 0; r < R; ++r) ring_update(ring_at(rm % W, r), r, -1);
      ++rm;
    }
    // (4) ... and enter year i for the targets to come
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ring_at(i % W, r) = snew[r];
      ring_update(snew[r], r, +1);
    }
  }
  if (live && bad && !cell_list) atomicAdd(&nonfinite[c], bad);  // the fix-up pass keeps the first count
  }
}

// =======================================================================================
// (a') fixed baseline -- reference: detect.py:2299-2397
// =======================================================================================
// clim[d, c] = nanmean over the rows of day-of-year d (CSR list).  Thread = gridpoint,
// blockIdx.y strides over days of year.
__global__ void __launch_bounds__(256) doy_climatology_kernel(
    const float* __restrict__ x, int64_t N, int64_t pitch, const int32_t* __restrict__ doy_ptr,
    const int32_t* __restrict__ doy_rows, const float* __restrict__ shift, float* __restrict__ clim) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const float sh = shift ? shift[c] : 0.f;
  for (int d = blockIdx.y; d < NDOY; d += gridDim.y) {
    const int b = __ldg(&doy_ptr[d]), e = __ldg(&doy_ptr[d + 1]);
    double sum = 0.0;
    uint32_t n = 0u, nf = 0u;
#pragma unroll 4
    for (int j = b; j < e; ++j) {
      float v = ld_stream(&x[(int64_t)__ldg(&doy_rows[j]) * pitch + c]);
      if (shift) v = v - sh;
      if (v == v) {
        ++n;
        if (is_finite_f(v)) sum += (double)v; else nf |= nonfinite_code(v);
      }
    }
    float m;
    if (n == 0u) m = CUDART_NAN_F;
    else if (nf) m = nonfinite_result(nf);
    else m = (float)(sum / (double)n);
    clim[(int64_t)d * N + c] = m;
  }
}

// anom[t, c] = f32(f32(x - shift) - clim[doy[t], c]); thread = gridpoint, blockIdx.y = row chunk.
__global__ void __launch_bounds__(256) sub_doy_climatology_kernel(
    const float* x, int64_t T, int64_t N, int64_t pitch, const int16_t* __restrict__ doy,
    const float* __restrict__ shift, const float* __restrict__ clim, float* anom, int64_t anom_pitch,
    uint8_t* __restrict__ mask0, int32_t* __restrict__ nonfinite, int rows_per_block) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const float sh = shift ? shift[c] : 0.f;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t t1 = min(T, t0 + rows_per_block);
  int bad = 0;
#pragma unroll 4
  for (int64_t t = t0; t < t1; ++t) {
    float v = ld_stream(&x[t * pitch + c]);
    if (!is_finite_f(v)) ++bad;
    if (shift) v = v - sh;
    if (t == 0 && mask0) mask0[c] = is_finite_f(v) ? 1 : 0;
    const int d = __ldg(&doy[t]) - 1;
    st_stream(&anom[t * anom_pitch + c], clim ? v - __ldg(&clim[(int64_t)d * N + c]) : v);
  }
  if (nonfinite && bad) atomicAdd(&nonfinite[c], bad);
}

// Same, four adjacent gridpoints per thread (16-byte loads and stores, four rows in flight).
__global__ void __launch_bounds__(128) sub_doy_climatology4_kernel(
    const float* x, int64_t T, int64_t N, int64_t pitch, const int16_t* __restrict__ doy,
    const float* __restrict__ shift, const float* __restrict__ clim, float* anom, int64_t anom_pitch,
    uint8_t* __restrict__ mask0, int32_t* __restrict__ nonfinite, int rows_per_block) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (shift) sh = *reinterpret_cast<const float4*>(shift + c);
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t t1 = min(T, t0 + rows_per_block);
  int bad[4] = {0, 0, 0, 0};
  for (int64_t tb = t0; tb < t1; tb += 4) {
    float4 v[4], cl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (tb + u < t1) {
        v[u] = __ldcs(reinterpret_cast<const float4*>(x + (tb + u) * pitch + c));
        cl[u] = clim ? __ldg(reinterpret_cast<const float4*>(clim + (int64_t)(__ldg(&doy[tb + u]) - 1) * N + c))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (tb + u >= t1) continue;
      float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      const float s4[4] = {sh.x, sh.y, sh.z, sh.w};
      const float c4[4] = {cl[u].x, cl[u].y, cl[u].z, cl[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!is_finite_f(f[j])) ++bad[j];
        if (shift) f[j] = f[j] - s4[j];
      }
      if (tb + u == 0 && mask0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mask0[c + j] = is_finite_f(f[j]) ? 1 : 0;
      }
      __stcs(reinterpret_cast<float4*>(anom + (tb + u) * anom_pitch + c),
             clim ? make_float4(f[0] - c4[0], f[1] - c4[1], f[2] - c4[2], f[3] - c4[3]) : make_float4(f[0], f[1], f[2], f[3]));
    }
  }
  if (nonfinite) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (bad[j]) atomicAdd(&nonfinite[c + j], bad[j]);
  }
}

// =======================================================================================
// (a'') polynomial detrend -- reference: detect.py:2061-2296 (remove_harmonics=False)
// =======================================================================================

// coef[k, c] = sum_t P[t, k] * x[t, c], float64 accumulate, sequential in t (deterministic).
template <int K>
__global__ void __launch_bounds__(256) detrend_coef_kernel(
    const float* __restrict__ x, int64_t T, int64_t N, int64_t pitch, const double* __restrict__ P,
    double* __restrict__ coef, uint8_t* __restrict__ mask0, int32_t* __restrict__ nonfinite) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  int bad = 0;
#pragma unroll 4
  for (int64_t t = 0; t < T; ++t) {
    const float v = ld_stream(&x[t * pitch + c]);
    if (!is_finite_f(v)) ++bad;
    if (t == 0 && mask0) mask0[c] = is_finite_f(v) ? 1 : 0;
    const double dv = (double)v;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = fma(__ldg(&P[t * K + k]), dv, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) coef[(int64_t)k * N + c] = acc[k];
  if (nonfinite) nonfinite[c] = bad;
}

// Same, four adjacent gridpoints per thread (16-byte loads, four rows in flight); N % 4 == 0.
template <int K>
__global__ void __launch_bounds__(128) detrend_coef4_kernel(
    const float* __restrict__ x, int64_t T, int64_t N, int64_t pitch, const double* __restrict__ P,
    double* __restrict__ coef, uint8_t* __restrict__ mask0, int32_t* __restrict__ nonfinite) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  double acc[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.0;
  int bad[4] = {0, 0, 0, 0};
  for (int64_t t0 = 0; t0 < T; t0 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (t0 + u < T) v[u] = __ldcs(reinterpret_cast<const float4*>(x + (t0 + u) * pitch + c));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u >= T) continue;
      const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!is_finite_f(f[j])) ++bad[j];
        const double dv = (double)f[j];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k][j] = fma(__ldg(&P[(t0 + u) * K + k]), dv, acc[k][j]);
      }
      if (t0 + u == 0 && mask0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mask0[c + j] = is_finite_f(f[j]) ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) coef[(int64_t)k * N + c + j] = acc[k][j];
  if (nonfinite) {
#pragma unroll
    for (int j = 0; j < 4; ++j) nonfinite[c + j] = bad[j];
  }
}

// xd[t, c] = x[t, c] - f32(sum_k M[k, t] * coef[k, c]);  mean[c] = f32(nanmean_t xd).
template <int K>
__global__ void __launch_bounds__(256) detrend_apply_kernel(
    const float* x, int64_t T, int64_t N, int64_t pitch, const double* __restrict__ M,
    const double* __restrict__ coef, float* xd, int64_t xd_pitch, float* __restrict__ mean) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double cf[K];
#pragma unroll
  for (int k = 0; k < K; ++k) cf[k] = coef[(int64_t)k * N + c];
  double sum = 0.0;
  int64_t cnt = 0;
  uint32_t nf = 0u;
#pragma unroll 4
  for (int64_t t = 0; t < T; ++t) {
    const float v = ld_stream(&x[t * pitch + c]);
    double fit = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) fit = fma(__ldg(&M[(int64_t)k * T + t]), cf[k], fit);
    const float r = v - (float)fit;
    st_stream(&xd[t * xd_pitch + c], r);
    if (r == r) {
      ++cnt;
      if (is_finite_f(r)) sum += (double)r; else nf |= nonfinite_code(r);
    }
  }
  if (mean) {
    float m;
    if (cnt == 0) m = CUDART_NAN_F;
    else if (nf) m = nonfinite_result(nf);
    else m = (float)(sum / (double)cnt);
    mean[c] = m;
  }
}

// Same, four adjacent gridpoints per thread; N % 4 == 0.
template <int K>
__global__ void __launch_bounds__(128) detrend_apply4_kernel(
    const float* x, int64_t T, int64_t N, int64_t pitch, const double* __restrict__ M,
    const double* __restrict__ coef, float* xd, int64_t xd_pitch, float* __restrict__ mean) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  double cf[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) cf[k][j] = coef[(int64_t)k * N + c + j];
  double sum[4] = {0.0, 0.0, 0.0, 0.0};
  int cnt[4] = {0, 0, 0, 0};
  uint32_t nf[4] = {0u, 0u, 0u, 0u};
  for (int64_t t0 = 0; t0 < T; t0 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (t0 + u < T) v[u] = __ldcs(reinterpret_cast<const float4*>(x + (t0 + u) * pitch + c));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u >= T) continue;
      const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) fit = fma(__ldg(&M[(int64_t)k * T + t0 + u]), cf[k][j], fit);
        r[j] = f[j] - (float)fit;
        if (r[j] == r[j]) {
          ++cnt[j];
          if (is_finite_f(r[j])) sum[j] += (double)r[j]; else nf[j] |= nonfinite_code(r[j]);
        }
      }
      __stcs(reinterpret_cast<float4*>(xd + (t0 + u) * xd_pitch + c), make_float4(r[0], r[1], r[2], r[3]));
    }
  }
  if (mean) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float m;
      if (cnt[j] == 0) m = CUDART_NAN_F;
      else if (nf[j]) m = nonfinite_result(nf[j]);
      else m = (float)(sum[j] / (double)cnt[j]);
      mean[c + j] = m;
    }
  }
}

// =======================================================================================
// std_normalise of detrend_harmonic -- reference: detect.py:2257-2293
// =======================================================================================
// std[d, c] = population standard deviation (ddof = 0) of the rows of day-of-year d (flox "std", detect.py:2260-2268):
// float64 two-pass; a non-finite sample makes the group NaN, an empty group is NaN.
__global__ void __launch_bounds__(128) doy_std_kernel(const float* __restrict__ x, int64_t N, int64_t pitch,
                                                      const int32_t* __restrict__ doy_ptr,
                                                      const int32_t* __restrict__ doy_rows, float* __restrict__ sd) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  for (int d = blockIdx.y; d < NDOY; d += gridDim.y) {
    const int b = __ldg(&doy_ptr[d]), e = __ldg(&doy_ptr[d + 1]);
    double sum = 0.0;
    for (int j = b; j < e; ++j) sum += (double)__ldg(&x[(int64_t)__ldg(&doy_rows[j]) * pitch + c]);
    const double n = (double)(e - b);
    const double mean = sum / n;
    double m2 = 0.0;
    for (int j = b; j < e; ++j) {
      const double dv = (double)__ldg(&x[(int64_t)__ldg(&doy_rows[j]) * pitch + c]) - mean;
      m2 += dv * dv;
    }
    sd[(int64_t)d * N + c] = (e > b) ? (float)sqrt(m2 / n) : CUDART_NAN_F;
  }
}

// out[d, c] = sqrt(mean over the `win` days of year [d - win/2, d - win/2 + win - 1] (wrapping at 366) of sd^2): the
// wrap-padded centred rolling mean of detect.py:2271-2272 (xarray centres an even window on [i - win/2, i + win/2 - 1];
// a NaN anywhere in the window gives NaN: min_periods = win); values <= 1e-10 become NaN (detect.py:2276).
__global__ void __launch_bounds__(128) doy_rolling_rms_kernel(const float* __restrict__ sd, int64_t N, int win,
                                                              float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int half = win / 2;
  for (int d = blockIdx.y; d < NDOY; d += gridDim.y) {
    double acc = 0.0;
    for (int k = 0; k < win; ++k) {
      int dd = (d - half + k) % NDOY;
      if (dd < 0) dd += NDOY;
      const float v = __ldg(&sd[(int64_t)dd * N + c]);
      acc += (double)__fmul_rn(v, v);  // (std_day_wrap**2) is a float32 array
    }
    const float r = sqrtf((float)(acc / (double)win));
    out[(int64_t)d * N + c] = (r > 1e-10f) ? r : CUDART_NAN_F;
  }
}

// out[t, c] = x[t, c] / sd[doy[t], c]  (groupby division, detect.py:2278)
__global__ void __launch_bounds__(256) div_doy_kernel(const float* __restrict__ x, int64_t T, int64_t N, int64_t pitch,
                                                      const int16_t* __restrict__ doy, const float* __restrict__ sd,
                                                      float* __restrict__ out, int64_t out_pitch, int rows_per_block) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
#pragma unroll 4
  for (int64_t t = t0; t < t1; ++t) {
    const float v = ld_stream(&x[t * pitch + c]);
    st_stream(&out[t * out_pitch + c], __fdiv_rn(v, __ldg(&sd[(int64_t)(__ldg(&doy[t]) - 1) * N + c])));
  }
}

}  // namespace marex

using namespace marex;

// Launches the generic kernel over all N cells (cell_list == nullptr) or, grid-stride, over the
// cells listed in cell_list[0 .. *n_list) (device memory), with `list_ctas` CTAs per strip row.
namespace marex {
int launch_shift_generic(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* tidx,
                         const int32_t* year_val, int32_t n_years, int32_t W, int32_t S, const int32_t* out_row,
                         int32_t mode, float* anom, int64_t anom_pitch, uint8_t* mask0, int32_t* nonfinite,
                         const int32_t* cell_list, const int32_t* n_list, int list_ctas, cudaStream_t st) {
  const size_t smem_budget = 96 * 1024;  // two CTAs per SM
  const int64_t nblk_x = cell_list ? list_ctas : (N + 31) / 32;
  auto launch = [&](auto kern, int R) -> int {
    const size_t per_warp = (size_t)W * R * 32 * sizeof(float);
    if (per_warp > 200 * 1024) return MAREX_ERR_UNSUPPORTED;
    int spb = (int)(smem_budget / per_warp);
    if (spb < 1) spb = 1;
    if (spb > 8) spb = 8;
    const int n_strips = (NDOY + R - 1) / R;
    if (spb > n_strips) spb = n_strips;
    const size_t smem = per_warp * spb;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(shift_anomaly)");
    dim3 grid((unsigned)nblk_x, (unsigned)((n_strips + spb - 1) / spb));
    kern<<<grid, spb * 32, smem, st>>>(x, T, N, pitch, tidx, year_val, n_years, W, S, out_row, anom, anom_pitch,
                                       mask0, nonfinite, n_strips, spb, mode, cell_list, n_list);
    MAREX_LAUNCH_CHECK("shift_anomaly_kernel");
    return MAREX_OK;
  };
  int rc = launch(shift_anomaly_kernel<8>, 8);
  if (rc == MAREX_ERR_UNSUPPORTED) rc = launch(shift_anomaly_kernel<2>, 2);
  if (rc == MAREX_ERR_UNSUPPORTED) rc = launch(shift_anomaly_kernel<1>, 1);
  if (rc == MAREX_ERR_UNSUPPORTED) return fail(rc, "window_year_baseline too large for the shared-memory ring");
  return rc;
}
}  // namespace marex

extern "C" int marex_shift_anomaly_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* tidx,
                                       const int32_t* year_val, int32_t n_years, int32_t W, int32_t S,
                                       const int32_t* out_row, int32_t mode, float* anom, int64_t anom_pitch,
                                       uint8_t* mask0, int32_t* nonfinite, void* stream) {
  MAREX_REQUIRE(x && tidx && year_val && out_row && anom && mask0 && nonfinite, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && anom_pitch >= N && n_years > 0, "bad shape");
  MAREX_REQUIRE(W >= 1 && W <= 1023 && S >= 1 && S <= 1023, "W and S must be in 1..1023");
  cudaStream_t st = (cudaStream_t)stream;
  MAREX_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t) * N, st));
  return launch_shift_generic(x, T, N, pitch, tidx, year_val, n_years, W, S, out_row, mode, anom, anom_pitch, mask0,
                              nonfinite, nullptr, nullptr, 0, st);
}

extern "C" int marex_doy_climatology_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* doy_ptr,
                                         const int32_t* doy_rows, const float* shift, float* clim, void* stream) {
  MAREX_REQUIRE(x && doy_ptr && doy_rows && clim, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  const int threads = 128;
  const int64_t bx = (N + threads - 1) / threads;
  int by = (int)((4LL * sm_count() * 4 + bx - 1) / bx);  // enough CTAs to fill the chip for small N
  by = by < 1 ? 1 : (by > NDOY ? NDOY : by);
  doy_climatology_kernel<<<dim3((unsigned)bx, by), threads, 0, (cudaStream_t)stream>>>(x, N, pitch, doy_ptr, doy_rows,
                                                                                        shift, clim);
  MAREX_LAUNCH_CHECK("doy_climatology_kernel");
  return MAREX_OK;
}

extern "C" int marex_sub_doy_climatology_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int16_t* doy,
                                             const float* shift, const float* clim, float* anom, int64_t anom_pitch,
                                             uint8_t* mask0, int32_t* nonfinite, void* stream) {
  MAREX_REQUIRE(x && doy && anom, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && anom_pitch >= N, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (nonfinite) MAREX_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t) * N, st));
  const int threads = 128;
  if ((N % 4) == 0 && (pitch % 4) == 0 && (anom_pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(anom) % 16) == 0 && (!clim || (reinterpret_cast<uintptr_t>(clim) % 16) == 0) &&
      (!shift || (reinterpret_cast<uintptr_t>(shift) % 16) == 0)) {
    const int64_t bx4 = (N / 4 + threads - 1) / threads;
    int64_t by4 = (16LL * sm_count() + bx4 - 1) / bx4;
    by4 = by4 < 1 ? 1 : (by4 > T ? T : by4);
    if (by4 > 65535) by4 = 65535;
    int rpb = (int)((T + by4 - 1) / by4);
    rpb = (rpb + 3) & ~3;
    by4 = (T + rpb - 1) / rpb;
    sub_doy_climatology4_kernel<<<dim3((unsigned)bx4, (unsigned)by4), threads, 0, st>>>(
        x, T, N, pitch, doy, shift, clim, anom, anom_pitch, mask0, nonfinite, rpb);
    MAREX_LAUNCH_CHECK("sub_doy_climatology4_kernel");
    return MAREX_OK;
  }
  const int64_t bx = (N + threads - 1) / threads;
  int64_t by = (8LL * sm_count() + bx - 1) / bx;
  if (by < 1) by = 1;
  if (by > T) by = T;
  if (by > 65535) by = 65535;
  const int rows_per_block = (int)((T + by - 1) / by);
  by = (T + rows_per_block - 1) / rows_per_block;
  sub_doy_climatology_kernel<<<dim3((unsigned)bx, (unsigned)by), threads, 0, st>>>(
      x, T, N, pitch, doy, shift, clim, anom, anom_pitch, mask0, nonfinite, rows_per_block);
  MAREX_LAUNCH_CHECK("sub_doy_climatology_kernel");
  return MAREX_OK;
}

template <int K>
static int launch_detrend_coef(const float* x, int64_t T, int64_t N, int64_t pitch, const double* P, double* coef,
                               uint8_t* mask0, int32_t* nonfinite, cudaStream_t st) {
  const int threads = 128;
  if (K <= 8 && (N % 4) == 0 && (pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) {
    detrend_coef4_kernel<K><<<(unsigned)((N / 4 + threads - 1) / threads), threads, 0, st>>>(x, T, N, pitch, P, coef,
                                                                                            mask0, nonfinite);
    MAREX_LAUNCH_CHECK("detrend_coef4_kernel");
    return MAREX_OK;
  }
  detrend_coef_kernel<K><<<(unsigned)((N + threads - 1) / threads), threads, 0, st>>>(x, T, N, pitch, P, coef, mask0,
                                                                                      nonfinite);
  MAREX_LAUNCH_CHECK("detrend_coef_kernel");
  return MAREX_OK;
}
template <int K>
static int launch_detrend_apply(const float* x, int64_t T, int64_t N, int64_t pitch, const double* M,
                                const double* coef, float* xd, int64_t xd_pitch, float* mean, cudaStream_t st) {
  const int threads = 128;
  if (K <= 8 && (N % 4) == 0 && (pitch % 4) == 0 && (xd_pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(xd) % 16) == 0) {
    detrend_apply4_kernel<K><<<(unsigned)((N / 4 + threads - 1) / threads), threads, 0, st>>>(x, T, N, pitch, M, coef,
                                                                                             xd, xd_pitch, mean);
    MAREX_LAUNCH_CHECK("detrend_apply4_kernel");
    return MAREX_OK;
  }
  detrend_apply_kernel<K><<<(unsigned)((N + threads - 1) / threads), threads, 0, st>>>(x, T, N, pitch, M, coef, xd,
                                                                                       xd_pitch, mean);
  MAREX_LAUNCH_CHECK("detrend_apply_kernel");
  return MAREX_OK;
}

#define MAREX_K_DISPATCH(fn, ...)                                                 \
  switch (K) {                                                                    \
    case 1: return fn<1>(__VA_ARGS__);                                            \
    case 2: return fn<2>(__VA_ARGS__);                                            \
    case 3: return fn<3>(__VA_ARGS__);                                            \
    case 4: return fn<4>(__VA_ARGS__);                                            \
    case 5: return fn<5>(__VA_ARGS__);                                            \
    case 6: return fn<6>(__VA_ARGS__);                                            \
    case 7: return fn<7>(__VA_ARGS__);                                            \
    case 8: return fn<8>(__VA_ARGS__);                                            \
    case 9: return fn<9>(__VA_ARGS__);                                            \
    case 10: return fn<10>(__VA_ARGS__);                                          \
    case 11: return fn<11>(__VA_ARGS__);                                          \
    case 12: return fn<12>(__VA_ARGS__);                                          \
    default: return fail(MAREX_ERR_UNSUPPORTED, "K (model columns) must be 1..12"); \
  }

extern "C" int marex_detrend_coef_f64(const float* x, int64_t T, int64_t N, int64_t pitch, const double* P, int32_t K,
                                      double* coef, uint8_t* mask0, int32_t* nonfinite, void* stream) {
  MAREX_REQUIRE(x && P && coef, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  MAREX_K_DISPATCH(launch_detrend_coef, x, T, N, pitch, P, coef, mask0, nonfinite, (cudaStream_t)stream);
}

extern "C" int marex_detrend_apply_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const double* M, int32_t K,
                                       const double* coef, float* xd, int64_t xd_pitch, float* mean, void* stream) {
  MAREX_REQUIRE(x && M && coef && xd, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && xd_pitch >= N, "bad shape");
  MAREX_K_DISPATCH(launch_detrend_apply, x, T, N, pitch, M, coef, xd, xd_pitch, mean, (cudaStream_t)stream);
}

extern "C" int marex_doy_std_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* doy_ptr,
                                 const int32_t* doy_rows, float* sd, void* stream) {
  MAREX_REQUIRE(x && doy_ptr && doy_rows && sd, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  const int threads = 128;
  const int64_t bx = (N + threads - 1) / threads;
  int by = (int)((16LL * sm_count() + bx - 1) / bx);
  by = by < 1 ? 1 : (by > NDOY ? NDOY : by);
  doy_std_kernel<<<dim3((unsigned)bx, by), threads, 0, (cudaStream_t)stream>>>(x, N, pitch, doy_ptr, doy_rows, sd);
  MAREX_LAUNCH_CHECK("doy_std_kernel");
  return MAREX_OK;
}

extern "C" int marex_doy_rolling_rms_f32(const float* sd, int64_t N, int32_t win, float* out, void* stream) {
  MAREX_REQUIRE(sd && out && sd != out, "null or aliased pointer");
  MAREX_REQUIRE(N > 0 && win >= 1 && win <= NDOY, "bad shape");
  const int threads = 128;
  const int64_t bx = (N + threads - 1) / threads;
  int by = (int)((16LL * sm_count() + bx - 1) / bx);
  by = by < 1 ? 1 : (by > NDOY ? NDOY : by);
  doy_rolling_rms_kernel<<<dim3((unsigned)bx, by), threads, 0, (cudaStream_t)stream>>>(sd, N, win, out);
  MAREX_LAUNCH_CHECK("doy_rolling_rms_kernel");
  return MAREX_OK;
}

extern "C" int marex_div_doy_f32(const float* x, int64_t T, int64_t N, int64_t pitch, const int16_t* doy,
                                 const float* sd, float* out, int64_t out_pitch, void* stream) {
  MAREX_REQUIRE(x && doy && sd && out, "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && out_pitch >= N, "bad shape");
  const int threads = 256;
  const int64_t bx = (N + threads - 1) / threads;
  int64_t by = (8LL * sm_count() + bx - 1) / bx;
  by = by < 1 ? 1 : (by > T ? T : by);
  if (by > 65535) by = 65535;
  const int rows_per_block = (int)((T + by - 1) / by);
  by = (T + rows_per_block - 1) / rows_per_block;
  div_doy_kernel<<<dim3((unsigned)bx, (unsigned)by), threads, 0, (cudaStream_t)stream>>>(x, T, N, pitch, doy, sd, out,
                                                                                       out_pitch, rows_per_block);
  MAREX_LAUNCH_CHECK("div_doy_kernel");
  return MAREX_OK;
}
