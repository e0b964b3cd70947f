// Word-level core of the binary morphology on bit-packed extreme-event masks (tracker stage 1:
// marEx/track.py:1520-1726 fill_holes / fill_time_gaps).  Everything here is __host__ __device__ so that
// tests/morph_host.cu can run the very same per-word code on the build box (no GPU) against scipy.ndimage;
// the product only ever calls it from the kernels in morph.cu.
//
// Slab layout: a padded time step is Hp = ny + 2*pad rows of Wpw = ceil((nx + 2*pad) / 32) uint32 words,
// every row word-aligned; bit i of word w of row y is padded column 32*w + i.  Bits at columns >= Wp of the
// last word of a row are always stored as 0 (every kernel re-establishes this).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MAREX_HD __host__ __device__ __forceinline__
#else
#define MAREX_HD inline
#endif

namespace marex {

constexpr int MORPH_MAX_R = 32;  // the 96-bit (previous, current, next word) window covers half-widths <= 32

// Where the cells of one time step come from: bool bytes or bits; `origin` + y * row_stride + x is the
// element (byte) or bit offset of cell (y, x) inside a time step, t_pitch the stride between time steps
// in bytes (bytes source) or uint32 words (bits source).  `mask` (optional, [ny * nx] bytes) zeroes
// cells outside the ocean mask at read time (`data_bin.where(self.mask, other=False)`, track.py:1667).
struct MorphSrc {
  const uint8_t* bytes;
  const uint32_t* bits;
  int64_t t_pitch;
  int64_t row_stride;
  int64_t origin;
  const uint8_t* mask;
};

MAREX_HD uint32_t morph_src_bit(const MorphSrc& s, int64_t t, int y, int x, int nx) {
  if (s.mask && !s.mask[(int64_t)y * nx + x]) return 0u;
  const int64_t o = s.origin + (int64_t)y * s.row_stride + x;
  if (s.bytes) return s.bytes[t * s.t_pitch + o] != 0 ? 1u : 0u;
  return (s.bits[t * s.t_pitch + (o >> 5)] >> (o & 31)) & 1u;
}

// Padded index -> source index: np.pad mode "wrap" (periodic, track.py:1617) or "edge" (regional_mode).
MAREX_HD int morph_pad_index(int ip, int pad, int n, int wrap) {
  int i = ip - pad;
  if (wrap) {
    i %= n;
    if (i < 0) i += n;
  } else {
    i = i < 0 ? 0 : (i >= n ? n - 1 : i);
  }
  return i;
}

// Disk structuring element `x^2 + y^2 < R^2 + 1` (track.py:1613-1616) as the half-width of every row:
// hw[|dy|] = largest dx with dx^2 + dy^2 <= R^2.  Non-increasing in |dy|.
struct MorphDisk {
  int32_t R;
  int8_t hw[MORPH_MAX_R + 1];
};

// One step of 1-D dilation (radius 1) of the 96-bit window (p, c, n) = words w-1, w, w+1.  Nothing is known
// left of p or right of n, so p's low bits / n's high bits go stale by one bit per step; after <= 32 steps the
// stale bits have not reached c.
MAREX_HD void morph_h1(uint32_t& p, uint32_t& c, uint32_t& n) {
  const uint32_t np = p | (p << 1) | (p >> 1) | (c << 31);
  const uint32_t nc = c | (c << 1) | (c >> 1) | (p >> 31) | (n << 31);
  const uint32_t nn = n | (n << 1) | (n >> 1) | (c >> 31);
  p = np;
  c = nc;
  n = nn;
}

// Output word (y, w) of the dilation (erode = 0) or erosion (erode = 1) of one padded time step by the disk,
// with scipy's border_value = 0 (cells outside the padded array are False for both operations).
//   dilation = OR over rows dy of the row dilated horizontally by hw[|dy|]; since hw is non-increasing the
//   rows are OR-ed in from the centre outwards and the accumulated window is widened by the DIFFERENCE of
//   consecutive half-widths, R single-bit steps in total (H_a o H_b = H_{a+b}, H distributes over OR).
//   erosion by a symmetric element = complement of the dilation of the complement (outside -> True).
MAREX_HD uint32_t morph_disk_word(const uint32_t* in, int Hp, int Wpw, uint32_t tailmask, int y, int w,
                                  const MorphDisk& d, int erode) {
  const uint32_t flip = erode ? 0xffffffffu : 0u;
  const bool wl = w - 1 >= 0, wr = w + 1 < Wpw;
  uint32_t p, c, n;
  {
    const uint32_t* row = in + (int64_t)y * Wpw;
    p = (wl ? row[w - 1] : 0u) ^ flip;
    c = row[w] ^ flip;
    n = (wr ? row[w + 1] : 0u) ^ flip;
  }
  int hcur = d.hw[0];
  for (int a = 1; a <= d.R; ++a) {
    const int h = d.hw[a];
    for (int k = hcur - h; k > 0; --k) morph_h1(p, c, n);
    hcur = h;
#pragma unroll
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const int yy = y + sgn * a;
      if (yy >= 0 && yy < Hp) {
        const uint32_t* row = in + (int64_t)yy * Wpw;
        p |= (wl ? row[w - 1] : 0u) ^ flip;
        c |= row[w] ^ flip;
        n |= (wr ? row[w + 1] : 0u) ^ flip;
      } else {
        p |= flip;
        c |= flip;
        n |= flip;
      }
    }
  }
  for (int k = hcur; k > 0; --k) morph_h1(p, c, n);
  uint32_t res = c ^ flip;
  if (w == Wpw - 1) res &= tailmask;
  return res;
}

// Temporal operator on whole slabs (per word, bit-parallel over 32 cells): out[t] = OP_{k=0..K-1} in[t + off + k],
// time steps outside [0, T_in) read as False (scipy border_value = 0 / the constant False padding of track.py:1706).
MAREX_HD uint32_t morph_time_word(const uint32_t* in, int64_t T_in, int64_t words, int64_t t, int64_t i, int off,
                                  int K, int erode) {
  uint32_t acc = erode ? 0xffffffffu : 0u;
  for (int k = 0; k < K; ++k) {
    const int64_t tt = t + off + k;
    const uint32_t v = (tt >= 0 && tt < T_in) ? in[tt * words + i] : 0u;
    acc = erode ? (acc & v) : (acc | v);
  }
  return acc;
}

// ---- unstructured: cell-major, time-packed words (word k of a cell = time steps 32*(k-1) .. 32*(k-1)+31) ----------

MAREX_HD uint32_t morph_tpack_word(const MorphSrc& src, int64_t T, int64_t c, int k, int Tw) {
  uint32_t word = 0;
  if (k >= 1 && k < Tw - 1) {
    const int64_t t0 = (int64_t)(k - 1) * 32;
    const int nt = (int)((T - t0) < 32 ? (T - t0) : 32);
    const int64_t o = src.origin + c;
    for (int j = 0; j < nt; ++j) {
      uint32_t b;
      if (src.bytes) b = src.bytes[(t0 + j) * src.t_pitch + o] != 0 ? 1u : 0u;
      else b = (src.bits[(t0 + j) * src.t_pitch + (o >> 5)] >> (o & 31)) & 1u;
      word |= b << j;
    }
  }
  return word;
}

// One application of "neighbours + identity" to word k of cell c (see morph_nbr_kernel).
MAREX_HD uint32_t morph_nbr_word(const uint32_t* in, int64_t N, int Tw, const int32_t* nbr, int nv, const uint8_t* mask,
                                 int flip, int set_land, int64_t c, int k) {
  const uint32_t f = flip ? 0xffffffffu : 0u;
  uint32_t v = in[c * Tw + k];
  if (set_land && !mask[c]) v = 0xffffffffu;
  uint32_t acc = v ^ f;
  for (int j = 0; j < nv; ++j) {
    const int32_t m = nbr[(int64_t)j * N + c];
    if (m >= 0) {
      uint32_t u = in[(int64_t)m * Tw + k];
      if (set_land && !mask[m]) u = 0xffffffffu;
      acc |= u ^ f;
    }
  }
  return acc ^ f;
}

// Dilation / erosion by +-half time steps along the bit axis of cell c, output word k (see morph_tshift_kernel).
MAREX_HD uint32_t morph_tshift_word(const uint32_t* in, int Tw, int64_t T, int half, int flip, int clip, int64_t c, int k) {
  const uint32_t f = flip ? 0xffffffffu : 0u;
  uint32_t win[3];
  for (int d = -1; d <= 1; ++d) {
    const int kk = k + d;
    uint32_t v = f;  // beyond the margins: False
    if (kk >= 0 && kk < Tw) {
      v = in[c * Tw + kk];
      if (clip) {
        const int64_t lo = (int64_t)(kk - 1) * 32;  // time step of bit 0
        uint32_t valid = 0;
        if (kk >= 1 && lo < T) valid = (T - lo >= 32) ? 0xffffffffu : ((1u << (int)(T - lo)) - 1u);
        v &= valid;
      }
      v ^= f;
    }
    win[d + 1] = v;
  }
  for (int s = 0; s < half; ++s) morph_h1(win[0], win[1], win[2]);
  return win[1] ^ f;
}

MAREX_HD uint32_t morph_tailmask(int Wp) { return (Wp & 31) ? ((1u << (Wp & 31)) - 1u) : 0xffffffffu; }

}  // namespace marex
