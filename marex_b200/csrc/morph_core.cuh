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
// in bytes (bytes source) or uint32 words (bits source).  `mask_bits` (optional, the ocean mask as flattened bits,
// bit c & 31 of word c >> 5, c = y * nx + x) zeroes cells outside the mask at read time
// (`data_bin.where(self.mask, other=False)`, track.py:1667).
struct MorphSrc {
  const uint8_t* bytes;
  const uint32_t* bits;
  int64_t t_pitch;
  int64_t row_stride;
  int64_t origin;
  const uint32_t* mask_bits;
};

MAREX_HD uint32_t morph_src_bit(const MorphSrc& s, int64_t t, int y, int x, int nx) {
  if (s.mask_bits) {
    const int64_t c = (int64_t)y * nx + x;
    if (!((s.mask_bits[c >> 5] >> (c & 31)) & 1u)) return 0u;
  }
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

// The same step written through 64-bit shifts of (hi:lo) pairs, which compile to one funnel shift (SHF) per neighbour
// term: 13 instead of 16 instructions.  Used by the third disk variant only (not yet measured on a GPU).
MAREX_HD uint32_t morph_shl1(uint32_t lo, uint32_t hi) { return (uint32_t)((((uint64_t)hi << 32) | lo) << 1 >> 32); }  // hi<<1 | lo>>31
MAREX_HD uint32_t morph_shr1(uint32_t lo, uint32_t hi) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> 1); }        // lo>>1 | hi<<31
MAREX_HD void morph_h1f(uint32_t& p, uint32_t& c, uint32_t& n) {
  const uint32_t np = p | (p << 1) | morph_shr1(p, c);
  const uint32_t nc = c | morph_shl1(p, c) | morph_shr1(c, n);
  const uint32_t nn = n | morph_shl1(c, n) | (n >> 1);
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
template <bool ERODE>
MAREX_HD uint32_t morph_disk_word_t(const uint32_t* in, int Hp, int Wpw, uint32_t tailmask, int y, int w,
                                    const MorphDisk& d) {
  constexpr uint32_t flip = ERODE ? 0xffffffffu : 0u;
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

// Third variant (variant = 3, MAREX_MORPH_DISK=3; its bits are checked on the host, its speed is NOT yet measured on a GPU):
// the SASS of the version above spends ~100 of ~120 instructions per row pair on bounds branches and 64-bit addressing.
template <bool ERODE>
MAREX_HD uint32_t morph_disk_word3_t(const uint32_t* in, int Hp, int Wpw, uint32_t tailmask, int y, int w,
                                    const MorphDisk& d) {
  constexpr uint32_t flip = ERODE ? 0xffffffffu : 0u;
  // Branch-free borders: rows are CLAMPED into the slab (every load is in bounds) and a row outside is masked to 0 before
  // the complement, so `acc |= (v & m) ^ flip` is one LOP3 per loaded word; one pointer per row, the neighbour columns
  // at immediate offsets -1 / +1.  (The first version branched on every row and did 64-bit address arithmetic per
  // load: ~100 of its ~120 instructions per row pair were checks and addressing.)
  const bool wl = w > 0, wr = w + 1 < Wpw;  // per thread constants: the column loads are predicated, not branched
  const uint32_t* ctr = in + (y * Wpw + w);  // 32-bit word offsets inside the time step (Hp * Wpw < 2^31)
  uint32_t p = (wl ? ctr[-1] : 0u) ^ flip;
  uint32_t c = ctr[0] ^ flip;
  uint32_t n = (wr ? ctr[1] : 0u) ^ flip;
  int hcur = d.hw[0];
  for (int a = 1; a <= d.R; ++a) {
    const int h = d.hw[a];
    for (int k = hcur - h; k > 0; --k) morph_h1f(p, c, n);
    hcur = h;
    const int yu = y - a, yd = y + a;
    const uint32_t mu = yu >= 0 ? 0xffffffffu : 0u, md = yd < Hp ? 0xffffffffu : 0u;
    const uint32_t* pu = in + ((yu >= 0 ? yu : 0) * Wpw + w);
    const uint32_t* pd = in + ((yd < Hp ? yd : Hp - 1) * Wpw + w);
    p |= (((wl ? pu[-1] : 0u) & mu) ^ flip) | (((wl ? pd[-1] : 0u) & md) ^ flip);
    c |= ((pu[0] & mu) ^ flip) | ((pd[0] & md) ^ flip);
    n |= (((wr ? pu[1] : 0u) & mu) ^ flip) | (((wr ? pd[1] : 0u) & md) ^ flip);
  }
  for (int k = hcur; k > 0; --k) morph_h1f(p, c, n);
  uint32_t res = c ^ flip;
  if (w == Wpw - 1) res &= tailmask;
  return res;
}

// `erode` and `variant` are uniform over a launch, so the dispatch does not diverge.  (Measured on B200: a dispatch on
// "all inputs inside the slab" made the kernel SLOWER -- rows are 46 words at 0.25 degree, so almost every warp
// straddles a row end and ran both variants.)
MAREX_HD uint32_t morph_disk_word(const uint32_t* in, int Hp, int Wpw, uint32_t tailmask, int y, int w,
                                  const MorphDisk& d, int erode, int variant = 2) {
  if (variant == 3)
    return erode ? morph_disk_word3_t<true>(in, Hp, Wpw, tailmask, y, w, d)
                 : morph_disk_word3_t<false>(in, Hp, Wpw, tailmask, y, w, d);
  return erode ? morph_disk_word_t<true>(in, Hp, Wpw, tailmask, y, w, d)
               : morph_disk_word_t<false>(in, Hp, Wpw, tailmask, y, w, d);
}

inline MorphDisk morph_make_disk(int R) {
  MorphDisk d;
  d.R = R;
  for (int a = 0; a <= MORPH_MAX_R; ++a) d.hw[a] = 0;
  for (int a = 0; a <= R; ++a) {
    int h = 0;
    while ((h + 1) * (h + 1) + a * a < R * R + 1) ++h;  // x^2 + y^2 < R^2 + 1 (track.py:1614-1616)
    d.hw[a] = (int8_t)h;
  }
  return d;
}

// ---- separable form of the same pass ---------------------------------------------------------------------------
// The disk is a stack of horizontal runs, so  dilate(x)[y] = OR_dy  H_{hw[|dy|]}(x[y + dy])  with H_h the 1-D dilation
// of a row by +-h.  Pass H widens every input row ONCE step by step and stores it at the distinct half-widths the
// disk uses ("levels", 5 for R = 8); pass V then ORs 2R+1 single words per output word.  The level buffers live in a
// caller-provided scratch that is sized for a chunk of time steps small enough to stay in L2 between the two passes.
struct MorphPlan {
  int32_t R, hmax, nlev;
  int8_t row_lvl[MORPH_MAX_R + 1];   // level buffer read for rows at distance a, -1 = the input itself (half-width 0)
  int8_t store_at[MORPH_MAX_R + 1];  // level buffer written after s widening steps, -1 = none
};

inline MorphPlan morph_make_plan(int R) {
  const MorphDisk d = morph_make_disk(R);
  MorphPlan pl;
  pl.R = R;
  pl.hmax = d.hw[0];
  pl.nlev = 0;
  for (int i = 0; i <= MORPH_MAX_R; ++i) pl.row_lvl[i] = pl.store_at[i] = -1;
  for (int a = 0; a <= R; ++a) {
    const int h = d.hw[a];
    if (h > 0 && pl.store_at[h] < 0) pl.store_at[h] = (int8_t)pl.nlev++;
    pl.row_lvl[a] = h > 0 ? pl.store_at[h] : (int8_t)-1;
  }
  return pl;
}

// Pass H for input word (y, w) of one time step: `hb` points at this word in level buffer 0, level j is lvl_stride
// words further.  Bits beyond the row end may be set in the stored words; pass V clears them.
template <bool ERODE>
MAREX_HD void morph_disk_h_word_t(const uint32_t* in, int Wpw, int y, int w, const MorphPlan& pl, uint32_t* hb,
                                  int64_t lvl_stride) {
  constexpr uint32_t flip = ERODE ? 0xffffffffu : 0u;
  const uint32_t* row = in + (int64_t)y * Wpw;
  uint32_t p = (w - 1 >= 0 ? row[w - 1] : 0u) ^ flip;
  uint32_t c = row[w] ^ flip;
  uint32_t n = (w + 1 < Wpw ? row[w + 1] : 0u) ^ flip;
  for (int s = 1; s <= pl.hmax; ++s) {
    morph_h1(p, c, n);
    const int j = pl.store_at[s];
    if (j >= 0) hb[j * lvl_stride] = c;
  }
}
MAREX_HD void morph_disk_h_word(const uint32_t* in, int Wpw, int y, int w, const MorphPlan& pl, uint32_t* hb,
                                int64_t lvl_stride, int erode) {
  if (erode) morph_disk_h_word_t<true>(in, Wpw, y, w, pl, hb, lvl_stride);
  else morph_disk_h_word_t<false>(in, Wpw, y, w, pl, hb, lvl_stride);
}

// Pass V for output word (y, w): `in` and `hb0` are the time step's input and its level-0 buffer.
template <bool ERODE>
MAREX_HD uint32_t morph_disk_v_word_t(const uint32_t* in, const uint32_t* hb0, int64_t lvl_stride, int Hp, int Wpw,
                                      uint32_t tailmask, int y, int w, const MorphPlan& pl) {
  constexpr uint32_t flip = ERODE ? 0xffffffffu : 0u;
  uint32_t acc = 0;
  for (int a = 0; a <= pl.R; ++a) {
    const int j = pl.row_lvl[a];
    const uint32_t* src = j < 0 ? in : hb0 + j * lvl_stride;
    const uint32_t f = j < 0 ? flip : 0u;  // the level buffers already hold the complemented field
    for (int sgn = (a == 0 ? 1 : -1); sgn <= 1; sgn += 2) {
      const int yy = y + sgn * a;
      acc |= (yy >= 0 && yy < Hp) ? (src[(int64_t)yy * Wpw + w] ^ f) : flip;
    }
  }
  uint32_t res = acc ^ flip;
  if (w == Wpw - 1) res &= tailmask;
  return res;
}
MAREX_HD uint32_t morph_disk_v_word(const uint32_t* in, const uint32_t* hb0, int64_t lvl_stride, int Hp, int Wpw,
                                    uint32_t tailmask, int y, int w, const MorphPlan& pl, int erode) {
  return erode ? morph_disk_v_word_t<true>(in, hb0, lvl_stride, Hp, Wpw, tailmask, y, w, pl)
               : morph_disk_v_word_t<false>(in, hb0, lvl_stride, Hp, Wpw, tailmask, y, w, pl);
}

// ---- the separable pass inside ONE shared-memory tile (variant 4, MAREX_MORPH_DISK=4; host-checked, not yet measured) --------
// A CTA owns TH output rows of one time step: it stages rows y0-R .. y0+TH+R-1 (complemented for erosion, rows outside the
// slab as "False"), widens every staged row ONCE into one shared buffer per level (phase H), then ORs 2R+1 single words per
// output word (phase V).  Each input row is widened once per tile instead of once per output row it feeds, and nothing
// but the input and the output touches global memory.  The three per-item functions below are what both the kernel
// (threads striding over items, __syncthreads between phases) and the host harness (plain loops) execute.
MAREX_HD uint32_t morph_tile_load_item(const uint32_t* in, int Hp, int Wpw, int y0, int R, int item, uint32_t flip) {
  const int row = item / Wpw, w = item - row * Wpw;
  const int yy = y0 - R + row;
  const uint32_t v = (yy >= 0 && yy < Hp) ? in[yy * Wpw + w] : 0u;
  return v ^ flip;
}

MAREX_HD void morph_tile_h_item(const uint32_t* in_s, uint32_t* lvl_s, int lvl_stride, int Wpw, const MorphPlan& pl, int item,
                                uint32_t flip) {
  const int row = item / Wpw, w = item - row * Wpw;
  uint32_t p = w > 0 ? in_s[item - 1] : flip;
  uint32_t c = in_s[item];
  uint32_t n = w + 1 < Wpw ? in_s[item + 1] : flip;
  for (int s = 1; s <= pl.hmax; ++s) {
    morph_h1(p, c, n);
    const int j = pl.store_at[s];
    if (j >= 0) lvl_s[j * lvl_stride + item] = c;
  }
}

MAREX_HD uint32_t morph_tile_v_item(const uint32_t* in_s, const uint32_t* lvl_s, int lvl_stride, int Wpw, const MorphPlan& pl,
                                    int orow, int w, uint32_t flip) {
  uint32_t acc = 0;
  const int ctr = (orow + pl.R) * Wpw + w;  // staged row of the output row
  for (int a = 0; a <= pl.R; ++a) {
    const int j = pl.row_lvl[a];
    const uint32_t* src = j < 0 ? in_s : lvl_s + j * lvl_stride;
    acc |= src[ctr - a * Wpw] | src[ctr + a * Wpw];
  }
  return acc ^ flip;
}

// Rows per tile that fit `smem_budget` bytes: (1 + nlev) buffers of (TH + 2R) x Wpw words; 0 = does not fit.
inline int morph_tile_rows(int Wpw, int R, int nlev, int64_t smem_budget) {
  for (int th = 32; th >= 4; th >>= 1)
    if ((int64_t)(1 + nlev) * (th + 2 * R) * Wpw * 4 <= smem_budget) return th;
  return 0;
}

// Chunk loop shared by the library (kernel launches) and the host test harness (plain loops): pass_h(t0, n, lvl_stride)
// then pass_v(t0, n, lvl_stride) for consecutive chunks of n <= chunk time steps.
template <class FH, class FV>
inline int morph_disk_separable_chunks(int64_t T, int64_t per_t, int nlev, int64_t scratch_words, FH&& pass_h, FV&& pass_v) {
  int64_t chunk = scratch_words / ((int64_t)nlev * per_t);
  if (chunk < 1) return -1;
  if (chunk > 65535) chunk = 65535;  // gridDim.y
  if (chunk > T) chunk = T;
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t n = chunk < T - t0 ? chunk : T - t0;
    int rc = pass_h(t0, n, chunk * per_t);
    if (rc) return rc;
    rc = pass_v(t0, n, chunk * per_t);
    if (rc) return rc;
  }
  return 0;
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

// ---- whole-word gathers for BITS sources (no per-bit work) ---------------------------------------------------------

MAREX_HD uint32_t morph_lowmask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// 32 consecutive bits of a bit array starting at bit offset o >= 0; words at index >= nwords read as 0.
MAREX_HD uint32_t morph_get32(const uint32_t* base, int64_t nwords, int64_t o) {
  const int64_t i = o >> 5;
  const int sh = (int)(o & 31);
  const uint32_t lo = i < nwords ? base[i] : 0u;
  if (sh == 0) return lo;
  const uint32_t hi = (i + 1 < nwords) ? base[i + 1] : 0u;
  return (lo >> sh) | (hi << (32 - sh));
}

// Word wp of padded row (source row ys) of time step t: runs of consecutive source cells are fetched 32 bits at a time.
// wrap: a run ends where the source row ends (then continues at x = 0); edge: the pad columns replicate one cell.
MAREX_HD uint32_t morph_pad_word(const MorphSrc& s, int64_t T, int64_t t, int ys, int wp, int Wp, int pad, int ny, int nx,
                                 int wrap) {
  const uint32_t* base = s.bits + t * s.t_pitch;
  const int64_t left = (T - t) * s.t_pitch, mwords = ((int64_t)ny * nx + 31) >> 5;
  uint32_t out = 0;
  int filled = 0, xp = wp * 32;
  while (filled < 32 && xp < Wp) {
    const int want = (32 - filled) < (Wp - xp) ? (32 - filled) : (Wp - xp);
    int xs, n;
    bool rep = false;
    if (wrap) {
      xs = morph_pad_index(xp, pad, nx, 1);
      n = want < nx - xs ? want : nx - xs;
    } else if (xp < pad) {
      xs = 0;
      n = want < pad - xp ? want : pad - xp;
      rep = true;
    } else if (xp >= pad + nx) {
      xs = nx - 1;
      n = want;
      rep = true;
    } else {
      xs = xp - pad;
      n = want < nx - xs ? want : nx - xs;
    }
    uint32_t v = morph_get32(base, left, s.origin + (int64_t)ys * s.row_stride + xs);
    if (s.mask_bits) v &= morph_get32(s.mask_bits, mwords, (int64_t)ys * nx + xs);
    if (rep) v = (v & 1u) ? 0xffffffffu : 0u;
    out |= (v & morph_lowmask(n)) << filled;
    filled += n;
    xp += n;
  }
  return out;
}

// Word w of the flattened bit mask of time step t (cells 32*w .. 32*w+31, row-major over (y, x)).
MAREX_HD uint32_t morph_extract_word(const MorphSrc& s, int64_t T, int64_t t, int64_t w, int nx, int64_t N) {
  const uint32_t* base = s.bits + t * s.t_pitch;
  const int64_t left = (T - t) * s.t_pitch;
  const int64_t c = w * 32;
  int y = (int)(c / nx), x = (int)(c - (int64_t)y * nx);
  uint32_t out = 0;
  int filled = 0;
  while (filled < 32 && c + filled < N) {
    const int n = (32 - filled) < (nx - x) ? (32 - filled) : (nx - x);
    const uint32_t v = morph_get32(base, left, s.origin + (int64_t)y * s.row_stride + x);
    out |= (v & morph_lowmask(n)) << filled;
    filled += n;
    x = 0;
    ++y;
  }
  if (s.mask_bits) out &= s.mask_bits[w];
  return out;
}

// 4 bits -> 4 bool bytes (bit k -> byte k): the products b_j * 2^(7i) land on distinct bit positions, no carries.
MAREX_HD uint32_t morph_expand4(uint32_t b) { return ((b & 0xfu) * 0x00204081u) & 0x01010101u; }

// 32 bool bytes (each 0 or 1 -- the host mirror normalises other dtypes with `!= 0`) -> one word of the flattened bit mask.
// 4 bytes at a time: (w * 0x01020408) >> 24 gathers b0 + 2 b1 + 4 b2 + 8 b3 (the partial products land on distinct bits).
MAREX_HD uint32_t morph_nibble(uint32_t w) { return ((w & 0x01010101u) * 0x01020408u) >> 24; }
MAREX_HD uint32_t morph_pack_word(const uint8_t* p, int n) {
  uint32_t word = 0;
  if (n == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
    word = morph_nibble(a.x) | (morph_nibble(a.y) << 4) | (morph_nibble(a.z) << 8) | (morph_nibble(a.w) << 12) |
           (morph_nibble(b.x) << 16) | (morph_nibble(b.y) << 20) | (morph_nibble(b.z) << 24) | (morph_nibble(b.w) << 28);
  } else {
    for (int j = 0; j < n; ++j) word |= (p[j] != 0 ? 1u : 0u) << j;
  }
  return word;
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
