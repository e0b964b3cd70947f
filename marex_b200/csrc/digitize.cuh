// np.digitize(a, edges) - 1 against the reference's float32 edge table (detect.py:2601-2605, 2622-2631).
// The table is near-uniform (np.arange in float32 accumulates rounding), so the bin is guessed
// from the spacing and then corrected against the REAL edges held in shared memory: bit-exact for any
// monotone table.  A near-uniform table (checked when it is loaded) needs exactly two look-ups per sample; any
// other table is walked.
#pragma once
#include "common.cuh"

namespace marex {

constexpr int BIN_INV = 0x7FFF;  // code of a sample that is not counted (NaN or a >= last edge); sorts above every bin

struct DigTable {
  const float* s_edges;  // shared-memory copy of edges[0 .. n_edges)
  int n_edges;
  float e1, inv_step;
  bool near_uniform;     // every edge within a quarter step of e1 + (i - 1) * step: the guess is at most one bin off
  // Cooperative set-up by every thread of the CTA (contains __syncthreads): copies the table to shared memory and
  // checks how uniform it is.
  __device__ __forceinline__ void init_cta(const float* __restrict__ g_edges, int n, float* s) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = g_edges[i];
    __syncthreads();
    s_edges = s;
    n_edges = n;
    e1 = s[1];
    const float step = (n > 2) ? s[2] - s[1] : 1.f;
    inv_step = 1.f / step;
    bool ok = step > 0.f && n > 2;
    for (int i = 1 + threadIdx.x; i < n; i += blockDim.x) ok = ok && fabsf(s[i] - (e1 + (float)(i - 1) * step)) < 0.25f * step;
    near_uniform = __syncthreads_and(ok);
  }
  __device__ __forceinline__ uint32_t operator()(float v) const {
    float g = floorf((v - e1) * inv_step) + 1.f;
    g = fminf(fmaxf(g, 0.f), (float)(n_edges - 2));  // NaN -> 0 (fmaxf drops it); the NaN test comes last
    int i = (int)g;
    i -= (v < s_edges[i]) ? 1 : 0;  // edges[0] = -inf: never below 0
    i += (v >= s_edges[i + 1]) ? 1 : 0;
    if (!near_uniform) {  // arbitrary monotone table: walk to the bin
      while (i > 0 && v < s_edges[i]) --i;
      while (i < n_edges - 1 && v >= s_edges[i + 1]) ++i;
    }
    return (i >= n_edges - 1 || v != v) ? (uint32_t)BIN_INV : (uint32_t)i;
  }
};

}  // namespace marex
