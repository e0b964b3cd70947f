// Shared helpers for libmarex_b200 (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/marex_b200.h"

namespace marex {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

// Tuning / test knobs (not part of the computation's API): marex_tune("shift_v", 4) pins a value for this process;
// unset keys fall back to the environment variable MAREX_<KEY IN CAPITALS>, read once, then to `dflt`.
long long tune_get(const char* key, long long dflt);

inline int fail(int code, const char* what, const char* detail = "") {
  g_last_error = std::string(what) + (detail[0] ? ": " : "") + detail;
  return code;
}

inline int cuda_fail(cudaError_t e, const char* where) {
  g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
  return MAREX_ERR_CUDA;
}

#define MAREX_CUDA(call)                                   \
  do {                                                     \
    cudaError_t _e = (call);                               \
    if (_e != cudaSuccess) return cuda_fail(_e, #call);    \
  } while (0)

#define MAREX_LAUNCH_CHECK(name)                           \
  do {                                                     \
    marex::g_launches.fetch_add(1);                        \
    cudaError_t _e = cudaGetLastError();                   \
    if (_e != cudaSuccess) return cuda_fail(_e, name);     \
  } while (0)

#define MAREX_REQUIRE(cond, msg)                                          \
  do {                                                                    \
    if (!(cond)) return marex::fail(MAREX_ERR_INVALID_ARG, msg, #cond);   \
  } while (0)

constexpr int NDOY = MAREX_NDOY;

// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ bool is_finite_f(float v) { return fabsf(v) < CUDART_INF_F; }  // false for NaN/inf

// Streaming loads/stores: inputs and outputs are touched once (or re-read out of L2 soon),
// so keep them out of L1 where they would only evict the tables.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// Non-finite bookkeeping for running sums: a running sum cannot "un-add" NaN/inf, so
// non-finite samples are kept out of the sum and counted in three 10-bit fields:
//   bits 0..9  NaN count, bits 10..19 +inf count, bits 20..29 -inf count.
__device__ __forceinline__ uint32_t nonfinite_code(float v) {
  return (v != v) ? 1u : ((v > 0.f) ? (1u << 10) : (1u << 20));
}
// Value the IEEE sum would have given the counts in `nf` (only called when nf != 0).
__device__ __forceinline__ float nonfinite_result(uint32_t nf) {
  const uint32_t n_nan = nf & 1023u, n_pi = (nf >> 10) & 1023u, n_ni = (nf >> 20) & 1023u;
  if (n_nan || (n_pi && n_ni)) return CUDART_NAN_F;
  return n_pi ? CUDART_INF_F : -CUDART_INF_F;
}

__device__ __forceinline__ float warp_min(float v) {
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Float atomic min/max through the sign-split integer trick (addr initialised to +/-inf).
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_d(double* addr, double v) {
  if (v >= 0.0) atomicMin(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
  else atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_max_d(double* addr, double v) {
  if (v >= 0.0) atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
  else atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// anomaly.cu: generic shifting-baseline kernel over all cells or over a device-side cell list.
int launch_shift_generic(const float* x, int64_t T, int64_t N, int64_t pitch, const int32_t* tidx,
                         const int32_t* year_val, int32_t n_years, int32_t W, int32_t S, const int32_t* out_row,
                         int32_t mode, float* anom, int64_t anom_pitch, uint8_t* mask0, int32_t* nonfinite,
                         const int32_t* cell_list, const int32_t* n_list, int list_ctas, cudaStream_t st);

inline int sm_count() {  // of the CURRENT device (cached per device: a process may drive several GPUs)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

}  // namespace marex
