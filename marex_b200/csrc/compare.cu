// (c) fused compare -> bool bytes and/or bit-packed extreme mask; utilities (transpose,
// synthetic SST generator, library globals).
#include <climits>
#include <cstdlib>
#include <map>
#include <mutex>

#include "digitize.cuh"

namespace marex {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

namespace {
std::mutex g_tune_mutex;
std::map<std::string, long long> g_tune;      // pinned by marex_tune
std::map<std::string, long long> g_tune_env;  // cached environment look-ups (LLONG_MIN = not set)
}  // namespace

long long tune_get(const char* key, long long dflt) {
  std::lock_guard<std::mutex> lock(g_tune_mutex);
  auto it = g_tune.find(key);
  if (it != g_tune.end()) return it->second;
  auto ie = g_tune_env.find(key);
  if (ie == g_tune_env.end()) {
    std::string name = "MAREX_";
    for (const char* c = key; *c; ++c) name += (char)toupper((unsigned char)*c);
    const char* v = getenv(name.c_str());
    ie = g_tune_env.emplace(key, v ? atoll(v) : LLONG_MIN).first;
  }
  return ie->second == LLONG_MIN ? dflt : ie->second;
}

// events[t, c] = anom[t, c] >= thr.  A warp covers 32 consecutive gridpoints of one row, so the
// packed word is one __ballot_sync (bit = lane = c & 31).  THR_GLOBAL: thr is float64[N] and the
// comparison is done in float64 (detect.py:2915); otherwise thr is float32[366, N] selected by
// doy[t] (detect.py:2001-2004).  NaN on either side compares false.
template <bool THR_GLOBAL>
__global__ void __launch_bounds__(256) compare_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                      int64_t pitch, const int16_t* __restrict__ doy,
                                                      const void* __restrict__ thr_v, int64_t thr_pitch,
                                                      uint8_t* __restrict__ events, int64_t events_pitch,
                                                      uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                      unsigned long long* __restrict__ count, int rows_per_block) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x multiple of 32
  const bool live = c < N;
  const int64_t cc = live ? c : N - 1;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  double thr_g = 0.0;
  if (THR_GLOBAL) thr_g = reinterpret_cast<const double*>(thr_v)[cc];
  const float* thr_f = reinterpret_cast<const float*>(thr_v);
  unsigned int local = 0;
#pragma unroll 4
  for (int64_t t = t0; t < t1; ++t) {
    const float a = ld_stream(&anom[t * pitch + cc]);
    bool e;
    if (THR_GLOBAL) e = (double)a >= thr_g;
    else e = a >= __ldg(&thr_f[(int64_t)(__ldg(&doy[t]) - 1) * thr_pitch + cc]);
    e = e && live;
    if (events && live) events[t * events_pitch + c] = e ? 1 : 0;
    const unsigned int word = __ballot_sync(0xffffffffu, e);
    if (bits && (threadIdx.x & 31) == 0 && (c >> 5) < bits_pitch) bits[t * bits_pitch + (c >> 5)] = word;
    local += e ? 1u : 0u;
  }
  if (count) {
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
  }
}

// Hobday compare organised by day of year: a CTA owns (512 gridpoints, one day of year), keeps that
// day's thresholds in registers and walks the rows (one per year) of the day.  The threshold field
// is read once instead of once per row, loads are 16 bytes and stores 4 bytes per thread.
__global__ void __launch_bounds__(128) compare_doy_kernel(const float* __restrict__ anom, int64_t N, int64_t pitch,
                                                          const int32_t* __restrict__ doy_ptr,
                                                          const int32_t* __restrict__ doy_rows,
                                                          const float* __restrict__ thr, int64_t thr_pitch,
                                                          uint8_t* __restrict__ events, int64_t events_pitch,
                                                          uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                          unsigned long long* __restrict__ count) {
  const int d = blockIdx.y;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // N % 4 == 0: all four live or none
  const bool live = c < N;
  const int64_t cc = live ? c : 0;
  const int b0 = __ldg(&doy_ptr[d]), b1 = __ldg(&doy_ptr[d + 1]);
  const float4 th = __ldg(reinterpret_cast<const float4*>(thr + (int64_t)d * thr_pitch + cc));
  const int lane = threadIdx.x & 31;
  unsigned int local = 0;
  for (int j0 = b0; j0 < b1; j0 += 5) {
    float4 a[5];
    int64_t row[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      row[u] = (j0 + u < b1) ? (int64_t)__ldg(&doy_rows[j0 + u]) : -1;
      if (row[u] >= 0) a[u] = __ldcs(reinterpret_cast<const float4*>(anom + row[u] * pitch + cc));
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      if (row[u] < 0) continue;
      const unsigned e0 = (a[u].x >= th.x) && live, e1 = (a[u].y >= th.y) && live, e2 = (a[u].z >= th.z) && live,
                     e3 = (a[u].w >= th.w) && live;
      const unsigned nib = e0 | (e1 << 1) | (e2 << 2) | (e3 << 3);
      local += __popc(nib);
      if (events && live)
        __stcs(reinterpret_cast<unsigned int*>(events + row[u] * events_pitch + c), e0 | (e1 << 8) | (e2 << 16) | (e3 << 24));
      if (bits) {  // 8 lanes x 4 gridpoints = one 32-bit word
        unsigned w = nib << ((lane & 7) * 4);
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0 && live) bits[row[u] * bits_pitch + (c >> 5)] = w;
      }
    }
  }
  if (count) {
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(count, (unsigned long long)local);
  }
}

// Hobday compare from the 2-byte bin codes (the day-of-year-major array the anomaly / digitize kernels wrote for the
// approximate thresholds) instead of the 4-byte anomalies: with bt = bin of the threshold under the same edge table,
//   bin(a) > bt  =>  a >= edges[bin(a)] >= edges[bt + 1] > thr   : extreme
//   bin(a) < bt  =>  a <  edges[bin(a) + 1] <= edges[bt] <= thr  : not extreme
// so the float anomaly is only read where bin(a) == bt or where the sample carries the invalid code (NaN, or a beyond
// the last edge) -- a fraction of a percent of the samples.  A CTA owns (1024 gridpoints, one day of year): the
// threshold bins sit in registers, the rows of the day are consecutive slots of the bin array; 16-byte loads of 8
// codes, 8-byte stores of 8 bool bytes.  slot_row[s] = output row of slot s, or -1.
__global__ void __launch_bounds__(128) compare_bins_kernel(const uint16_t* __restrict__ bins, int NY, int64_t bins_pitch,
                                                           const int32_t* __restrict__ slot_row,
                                                           const float* __restrict__ anom, int64_t pitch, int64_t N,
                                                           const float* __restrict__ thr, int64_t thr_pitch,
                                                           const float* __restrict__ edges, int n_edges,
                                                           uint8_t* __restrict__ events, int64_t events_pitch,
                                                           uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                           unsigned long long* __restrict__ count) {
  extern __shared__ float s_edges[];
  DigTable dig;
  dig.init_cta(edges, n_edges, s_edges);
  const int d = blockIdx.y;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;  // N % 8 == 0: all eight live or none
  const bool live = c < N;
  const int64_t cc = live ? c : 0;
  const int lane = threadIdx.x & 31;
  const float* const thr_row = thr + (int64_t)d * thr_pitch + cc;
  // Two codes per 32-bit word, decided with 16-bit-lane arithmetic (codes are < 0x8000):
  //   extreme without looking at the anomaly:  bit 15 of (code | 0x8000) - (bt + 1)            [code > bt]
  //   needs the float compare:                 code == bt, or code == invalid (zero-lane test of the XOR)
  // NaN threshold (land) or a dead thread: bt + 1 := 0x8000 (never above), bt := 0xFFFF (never equal), invalid ignored.
  unsigned p1[4], btp[4], mk[4];
  {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(thr_row));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(thr_row + 4));
    const float th[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned a1 = 0, ab = 0, am = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float t = th[2 * j + h];
        unsigned q1 = 0x8000u, qb = 0xFFFFu, qm = 0u;
        if (t == t && live) {
          const unsigned bt = dig(t);  // BIN_INV when the threshold lies beyond the last edge
          q1 = bt < (unsigned)BIN_INV ? bt + 1u : 0x8000u;
          qb = bt;
          qm = 0x8000u;
        }
        a1 |= q1 << (16 * h); ab |= qb << (16 * h); am |= qm << (16 * h);
      }
      p1[j] = a1; btp[j] = ab; mk[j] = am;
    }
  }
  unsigned int local = 0;
  const int64_t s0 = (int64_t)d * NY;
  for (int j0 = 0; j0 < NY; j0 += 5) {
    uint4 b[5];
    int64_t row[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      row[u] = (j0 + u < NY) ? (int64_t)__ldg(&slot_row[s0 + j0 + u]) : -1;
      if (row[u] >= 0) b[u] = __ldcs(reinterpret_cast<const uint4*>(bins + (s0 + j0 + u) * bins_pitch + cc));
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      if (row[u] < 0) continue;
      const unsigned wv[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
      unsigned g[4], am[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        g[j] = (((wv[j] | 0x80008000u) - p1[j]) >> 15) & 0x00010001u;          // bytes 0 and 2: code > bt
        const unsigned x = wv[j] ^ btp[j], y = wv[j] ^ 0x7FFF7FFFu;
        am[j] = ((x - 0x00010001u) & ~x & 0x80008000u) | ((y - 0x00010001u) & ~y & mk[j]);  // bits 15 / 31: ambiguous
      }
      unsigned ex = __byte_perm(g[0], g[1], 0x6420), ey = __byte_perm(g[2], g[3], 0x6420);  // one bool byte per gridpoint
      if ((am[0] | am[1]) | (am[2] | am[3])) {
        // rare: decide those samples on the float anomaly (detect.py:2001-2004).  A zero-lane test may also flag
        // the upper code of a word whose lower code matched: harmless, the float compare is the definition.
        unsigned mx = __byte_perm((am[0] >> 15) & 0x00010001u, (am[1] >> 15) & 0x00010001u, 0x6420);
        unsigned my = __byte_perm((am[2] >> 15) & 0x00010001u, (am[3] >> 15) & 0x00010001u, 0x6420);
        const float* arow = anom + row[u] * pitch + cc;
        while (mx) {
          const int pos = __ffs(mx) - 1;  // 8 * k
          mx &= mx - 1;
          const int k = pos >> 3;
          const unsigned e = __ldg(arow + k) >= __ldg(thr_row + k) ? 1u : 0u;
          ex = (ex & ~(1u << pos)) | (e << pos);
        }
        while (my) {
          const int pos = __ffs(my) - 1;
          my &= my - 1;
          const int k = pos >> 3;
          const unsigned e = __ldg(arow + 4 + k) >= __ldg(thr_row + 4 + k) ? 1u : 0u;
          ey = (ey & ~(1u << pos)) | (e << pos);
        }
      }
      local += __popc(ex) + __popc(ey);
      if (events && live) __stcs(reinterpret_cast<uint2*>(events + row[u] * events_pitch + c), make_uint2(ex, ey));
      if (bits) {  // 4 lanes x 8 gridpoints = one 32-bit word
        // bool bytes -> 8 bits: byte k -> bit k
        const unsigned lo = (ex * 0x01020408u) >> 24, hi = (ey * 0x01020408u) >> 24;  // four 0/1 bytes gathered into a nibble
        unsigned w = ((lo & 0xFu) | ((hi & 0xFu) << 4)) << ((lane & 3) * 8);
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((lane & 3) == 0 && live) bits[row[u] * bits_pitch + (c >> 5)] = w;
      }
    }
  }
  if (count) {
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(count, (unsigned long long)local);
  }
}

// Global compare, four gridpoints per thread.  a >= thr with a float32 and thr float64 is decided
// exactly by a >= (smallest float32 >= thr), so the row loop stays in float32 (detect.py:2915).
__global__ void __launch_bounds__(128) compare_global4_kernel(const float* __restrict__ anom, int64_t T, int64_t N,
                                                              int64_t pitch, const double* __restrict__ thr,
                                                              uint8_t* __restrict__ events, int64_t events_pitch,
                                                              uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                              unsigned long long* __restrict__ count,
                                                              int rows_per_block) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const bool live = c < N;
  const int64_t cc = live ? c : 0;
  const int lane = threadIdx.x & 31;
  float th[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double t64 = thr[cc + k];
    float f = (float)t64;  // round to nearest; NaN stays NaN (compares false)
    if ((double)f < t64) f = nextafterf(f, CUDART_INF_F);
    th[k] = f;
  }
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  unsigned int local = 0;
  for (int64_t tb = t0; tb < t1; tb += 4) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tb + u < t1) a[u] = __ldcs(reinterpret_cast<const float4*>(anom + (tb + u) * pitch + cc));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (tb + u >= t1) continue;
      const unsigned e0 = (a[u].x >= th[0]) && live, e1 = (a[u].y >= th[1]) && live, e2 = (a[u].z >= th[2]) && live,
                     e3 = (a[u].w >= th[3]) && live;
      const unsigned nib = e0 | (e1 << 1) | (e2 << 2) | (e3 << 3);
      local += __popc(nib);
      if (events && live)
        __stcs(reinterpret_cast<unsigned int*>(events + (tb + u) * events_pitch + c), e0 | (e1 << 8) | (e2 << 16) | (e3 << 24));
      if (bits) {
        unsigned w = nib << ((lane & 7) * 4);
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0 && live) bits[(tb + u) * bits_pitch + (c >> 5)] = w;
      }
    }
  }
  if (count) {
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(count, (unsigned long long)local);
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int64_t rows, int64_t cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
  }
}

// ---- synthetic SST ------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)((h >> 40) + 1) * (1.0f / 16777217.0f); }  // (0, 1)

__global__ void __launch_bounds__(128) synth_sst_kernel(float* __restrict__ x, int64_t T, int64_t N, int64_t pitch,
                                                        int64_t c0, int64_t ny_g, int64_t nx_g,
                                                        const float* __restrict__ dec_year, uint64_t seed,
                                                        float land_fraction) {
  const int64_t cl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cl >= N) return;
  const int64_t cg = c0 + cl;
  const int64_t iy = cg / nx_g, ix = cg % nx_g;
  // land: 8 x 8-cell blobs chosen by a hash of the coarse block
  const uint64_t hb = splitmix64(seed ^ (uint64_t)((iy >> 3) * 1315423911ull + (ix >> 3)));
  const bool land = u01(hb) < land_fraction;
  const uint64_t hc = splitmix64(seed * 0x5851F42D4C957F2Dull + (uint64_t)cg);
  const float latf = ny_g > 1 ? (float)iy / (float)(ny_g - 1) : 0.5f;  // 0 south .. 1 north
  const float mu = -1.8f + 31.8f * (1.f - fabsf(2.f * latf - 1.f));     // [-1.8, 30]
  const float amp = 0.5f + 5.5f * u01(hc);
  const float phase = u01(splitmix64(hc));
  const float rho = 0.9f, sig = 0.6f * sqrtf(1.f - rho * rho);
  float ar = 0.f;
  uint64_t state = splitmix64(hc ^ 0xD1B54A32D192ED03ull);
  const float y0 = dec_year[0];
  for (int64_t t = 0; t < T; ++t) {
    state = splitmix64(state);
    const float u1 = u01(state), u2 = u01(state * 0x9E3779B97F4A7C15ull + 1);
    const float z = sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
    ar = rho * ar + sig * z;
    const float dy = dec_year[t];
    const float v = mu + amp * __cosf(6.2831853f * (dy - floorf(dy) - phase)) + 0.02f * (dy - y0) + ar;
    x[t * pitch + cl] = land ? CUDART_NAN_F : v;
  }
}

}  // namespace marex

using namespace marex;

extern "C" int marex_version(void) { return 100; }
extern "C" const char* marex_last_error(void) { return g_last_error.c_str(); }
extern "C" long long marex_launch_count(void) { return g_launches.load(); }
extern "C" int marex_tune(const char* key, long long value, int32_t set) {
  if (!key) return fail(MAREX_ERR_INVALID_ARG, "null key");
  std::lock_guard<std::mutex> lock(g_tune_mutex);
  if (set) g_tune[key] = value;
  else g_tune.erase(key);
  return MAREX_OK;
}

template <bool G>
static int launch_compare(const float* anom, int64_t T, int64_t N, int64_t pitch, const int16_t* doy, const void* thr,
                          int64_t thr_pitch, uint8_t* events, int64_t events_pitch, uint32_t* bits, int64_t bits_pitch,
                          unsigned long long* count, cudaStream_t st) {
  const int threads = 128;
  const int64_t bx = (N + threads - 1) / threads;
  int64_t by = (8LL * sm_count() + bx - 1) / bx;
  by = by < 1 ? 1 : (by > T ? T : by);
  if (by > 65535) by = 65535;
  const int rows_per_block = (int)((T + by - 1) / by);
  by = (T + rows_per_block - 1) / rows_per_block;
  compare_kernel<G><<<dim3((unsigned)bx, (unsigned)by), threads, 0, st>>>(anom, T, N, pitch, doy, thr, thr_pitch, events,
                                                                          events_pitch, bits, bits_pitch, count,
                                                                          rows_per_block);
  MAREX_LAUNCH_CHECK("compare_kernel");
  return MAREX_OK;
}

extern "C" int marex_compare_hobday(const float* anom, int64_t T, int64_t N, int64_t pitch, const int16_t* doy,
                                    const int32_t* doy_ptr, const int32_t* doy_rows, const float* thr,
                                    int64_t thr_pitch, uint8_t* events, int64_t events_pitch, uint32_t* bits,
                                    int64_t bits_pitch, unsigned long long* count, void* stream) {
  MAREX_REQUIRE(anom && doy && thr && (events || bits || count), "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N && thr_pitch >= N, "bad shape");
  MAREX_REQUIRE(!events || events_pitch >= N, "events_pitch < N");
  MAREX_REQUIRE(!bits || bits_pitch >= (N + 31) / 32, "bits_pitch < ceil(N/32)");
  const bool aligned = (N % 4) == 0 && (pitch % 4) == 0 && (thr_pitch % 4) == 0 && (!events || (events_pitch % 4) == 0) &&
                       (reinterpret_cast<uintptr_t>(anom) % 16) == 0 && (reinterpret_cast<uintptr_t>(thr) % 16) == 0 &&
                       (!events || (reinterpret_cast<uintptr_t>(events) % 4) == 0);
  if (doy_ptr && doy_rows && aligned && (N % 32 == 0 || !bits)) {
    const int threads = 128;
    dim3 grid((unsigned)((N / 4 + threads - 1) / threads), NDOY);
    compare_doy_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(anom, N, pitch, doy_ptr, doy_rows, thr, thr_pitch,
                                                                   events, events_pitch, bits, bits_pitch, count);
    MAREX_LAUNCH_CHECK("compare_doy_kernel");
    return MAREX_OK;
  }
  return launch_compare<false>(anom, T, N, pitch, doy, thr, thr_pitch, events, events_pitch, bits, bits_pitch, count,
                               (cudaStream_t)stream);
}

extern "C" int marex_compare_hobday_bins(const uint16_t* bins, int64_t NY, int64_t bins_pitch, const int32_t* slot_row,
                                         const float* anom, int64_t pitch, int64_t N, const float* thr,
                                         int64_t thr_pitch, const float* edges, int32_t n_edges, uint8_t* events,
                                         int64_t events_pitch, uint32_t* bits, int64_t bits_pitch,
                                         unsigned long long* count, void* stream) {
  MAREX_REQUIRE(bins && slot_row && anom && thr && edges && (events || bits || count), "null pointer");
  MAREX_REQUIRE(NY > 0 && N > 0 && pitch >= N && thr_pitch >= N && bins_pitch >= N, "bad shape");
  MAREX_REQUIRE(n_edges >= 3 && n_edges <= 4096, "n_edges must be in 3..4096");
  MAREX_REQUIRE(!events || events_pitch >= N, "events_pitch < N");
  MAREX_REQUIRE(!bits || bits_pitch >= (N + 31) / 32, "bits_pitch < ceil(N/32)");
  const bool aligned = (N % 8) == 0 && (bins_pitch % 8) == 0 && (thr_pitch % 4) == 0 &&
                       (reinterpret_cast<uintptr_t>(bins) % 16) == 0 && (reinterpret_cast<uintptr_t>(thr) % 16) == 0 &&
                       (!events || ((events_pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(events) % 8) == 0)) &&
                       (N % 32 == 0 || !bits);
  if (!aligned) return fail(MAREX_ERR_UNSUPPORTED, "compare from bins needs N % 8 == 0 and 16-byte aligned rows");
  const int threads = 128;
  dim3 grid((unsigned)((N / 8 + threads - 1) / threads), NDOY);
  compare_bins_kernel<<<grid, threads, n_edges * sizeof(float), (cudaStream_t)stream>>>(
      bins, (int)NY, bins_pitch, slot_row, anom, pitch, N, thr, thr_pitch, edges, n_edges, events, events_pitch, bits,
      bits_pitch, count);
  MAREX_LAUNCH_CHECK("compare_bins_kernel");
  return MAREX_OK;
}

extern "C" int marex_compare_global(const float* anom, int64_t T, int64_t N, int64_t pitch, const double* thr,
                                    uint8_t* events, int64_t events_pitch, uint32_t* bits, int64_t bits_pitch,
                                    unsigned long long* count, void* stream) {
  MAREX_REQUIRE(anom && thr && (events || bits || count), "null pointer");
  MAREX_REQUIRE(T > 0 && N > 0 && pitch >= N, "bad shape");
  MAREX_REQUIRE(!events || events_pitch >= N, "events_pitch < N");
  MAREX_REQUIRE(!bits || bits_pitch >= (N + 31) / 32, "bits_pitch < ceil(N/32)");
  const bool aligned = (N % 4) == 0 && (pitch % 4) == 0 && (!events || (events_pitch % 4) == 0) &&
                       (reinterpret_cast<uintptr_t>(anom) % 16) == 0 &&
                       (!events || (reinterpret_cast<uintptr_t>(events) % 4) == 0) && (N % 32 == 0 || !bits);
  if (aligned) {
    const int threads = 128;
    const int64_t bx = (N / 4 + threads - 1) / threads;
    int64_t by = (16LL * sm_count() + bx - 1) / bx;
    by = by < 1 ? 1 : (by > T ? T : by);
    if (by > 65535) by = 65535;
    int rows_per_block = (int)((T + by - 1) / by);
    rows_per_block = (rows_per_block + 3) & ~3;
    by = (T + rows_per_block - 1) / rows_per_block;
    compare_global4_kernel<<<dim3((unsigned)bx, (unsigned)by), threads, 0, (cudaStream_t)stream>>>(
        anom, T, N, pitch, thr, events, events_pitch, bits, bits_pitch, count, rows_per_block);
    MAREX_LAUNCH_CHECK("compare_global4_kernel");
    return MAREX_OK;
  }
  return launch_compare<true>(anom, T, N, pitch, nullptr, thr, 0, events, events_pitch, bits, bits_pitch, count,
                              (cudaStream_t)stream);
}

extern "C" int marex_memcpy2d_async(void* dst, int64_t dpitch_bytes, const void* src, int64_t spitch_bytes,
                                    int64_t width_bytes, int64_t height, int32_t to_device, void* stream) {
  MAREX_REQUIRE(dst && src && width_bytes > 0 && height > 0 && dpitch_bytes >= width_bytes && spitch_bytes >= width_bytes,
                "bad argument");
  MAREX_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch_bytes, src, (size_t)spitch_bytes, (size_t)width_bytes, (size_t)height,
                               to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return MAREX_OK;
}

extern "C" int marex_transpose_f32(const float* in, int64_t rows, int64_t cols, float* out, void* stream) {
  MAREX_REQUIRE(in && out && rows > 0 && cols > 0, "bad argument");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  MAREX_REQUIRE(grid.y <= 65535, "too many rows for transpose grid");
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, rows, cols, out);
  MAREX_LAUNCH_CHECK("transpose_kernel");
  return MAREX_OK;
}

extern "C" int marex_synth_sst_f32(float* x, int64_t T, int64_t N, int64_t pitch, int64_t c0, int64_t ny_global,
                                   int64_t nx_global, const float* dec_year, uint64_t seed, float land_fraction,
                                   void* stream) {
  MAREX_REQUIRE(x && dec_year && T > 0 && N > 0 && pitch >= N && ny_global > 0 && nx_global > 0, "bad argument");
  synth_sst_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, T, N, pitch, c0, ny_global,
                                                                                  nx_global, dec_year, seed,
                                                                                  land_fraction);
  MAREX_LAUNCH_CHECK("synth_sst_kernel");
  return MAREX_OK;
}
