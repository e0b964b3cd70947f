// np.digitize(a, edges) - 1 against the reference's float32 edge table (detect.py:2601-2605, 2622-2631).
// The table is near-uniform (np.arange in float32 accumulates rounding), so the bin is guessed
// from the spacing and then corrected against the REAL edges held in shared memory: bit-exact for any
// monotone table, two table look-ups per sample for a near-uniform one.
#pragma once
#include "common.cuh"

namespace marex {

constexpr int BIN_INV = 0x7FFF;  // code of a sample that is not counted (NaN or a >= last edge); sorts above every bin

struct DigTable {
  const float* s_edges;  // shared-memory copy of edges[0 .. n_edges)
  int n_edges;
  float e1, inv_step;
  __device__ __forceinline__ void init(const float* s, int n) {
    s_edges = s;
    n_edges = n;
    e1 = s[1];
    inv_step = (n > 2) ? 1.f / (s[2] - s[1]) : 1.f;
  }
  __device__ __forceinline__ uint32_t operator()(float v) const {
    float g = floorf((v - e1) * inv_step) + 1.f;
    g = fminf(fmaxf(g, 0.f), (float)(n_edges - 2));  // NaN -> 0 (fmaxf drops it); the NaN test comes last
    int i = (int)g;
    i -= (v < s_edges[i]) ? 1 : 0;  // edges[0] = -inf: never below 0
    i += (v >= s_edges[i + 1]) ? 1 : 0;
    if (i < n_edges - 1 && (v < s_edges[i] || v >= s_edges[i + 1])) {  // table not near-uniform: walk
      while (i > 0 && v < s_edges[i]) --i;
      while (i < n_edges - 1 && v >= s_edges[i + 1]) ++i;
    }
    return (i >= n_edges - 1 || v != v) ? (uint32_t)BIN_INV : (uint32_t)i;
  }
};

}  // namespace marex
