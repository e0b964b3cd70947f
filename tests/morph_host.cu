// TEST HARNESS (never part of libmarex_b200.so, never loaded by the product): the marex_morph_* entry points of
// include/marex_b200.h re-implemented as plain host loops around the __host__ __device__ per-word code of
// marex_b200/csrc/morph_core.cuh -- the very functions the CUDA kernels of morph.cu call.  The build box (no GPU) can
// so check the bit arithmetic AND the host-side composition in marex_b200/track.py against scipy.ndimage; the GPU
// tests then only have to establish that the kernels index and launch correctly.  Built by tests/test_track_cpu.py
// with `nvcc -shared` (host code only, no CUDA runtime call).  Pointers are HOST pointers here.
#include <stdint.h>
#include <stdlib.h>

#include "../marex_b200/csrc/morph_core.cuh"

using namespace marex;

static int popc(uint32_t v) { return __builtin_popcount(v); }

extern "C" {

int64_t marex_morph_slab_words(int64_t ny, int64_t nx, int32_t pad) { return (ny + 2 * pad) * ((nx + 2 * pad + 31) / 32); }
int64_t marex_morph_tpack_words(int64_t T) { return (T + 31) / 32 + 2; }

int marex_morph_pad_bits(const uint8_t* bytes, const uint32_t* bits, int64_t t_pitch, int64_t row_stride, int64_t origin,
                         const uint32_t* mask_bits, int64_t T, int64_t ny, int64_t nx, int32_t pad, int32_t wrap, uint32_t* dst,
                         void*) {
  MorphSrc s{bytes, bits, t_pitch, row_stride, origin, mask_bits};
  const int Hp = (int)ny + 2 * pad, Wp = (int)nx + 2 * pad, Wpw = (Wp + 31) >> 5;
  for (int64_t t = 0; t < T; ++t)
    for (int yp = 0; yp < Hp; ++yp)
      for (int w = 0; w < Wpw; ++w) {
        const int ys = morph_pad_index(yp, pad, (int)ny, wrap);
        uint32_t word = 0;
        if (bits) {  // the word-gather path of morph_pad_words_kernel
          word = morph_pad_word(s, T, t, ys, w, Wp, pad, (int)ny, (int)nx, wrap);
        } else {  // the lane-per-cell path of morph_pad_kernel
          for (int lane = 0; lane < 32; ++lane) {
            const int xp = w * 32 + lane;
            if (xp < Wp) word |= morph_src_bit(s, t, ys, morph_pad_index(xp, pad, (int)nx, wrap), (int)nx) << lane;
          }
        }
        dst[(t * Hp + yp) * Wpw + w] = word;
      }
  return 0;
}

int marex_morph_disk(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R, int32_t erode, void*) {
  if (R < 0 || R > MORPH_MAX_R) return -3;
  const MorphDisk d = morph_make_disk(R);
  const int env_variant = getenv("MAREX_MORPH_DISK") ? atoi(getenv("MAREX_MORPH_DISK")) : 2;  // as morph.cu
  const int variant = env_variant == 3 ? 3 : 2;
  const int Wpw = (int)((Wp + 31) >> 5);
  const int64_t per_t = Hp * Wpw;
  if (env_variant == 4 && R >= 1) {  // morph_disk_tile_kernel: phases separated by __syncthreads() = three loops per tile
    const MorphPlan pl = morph_make_plan(R);
    const char* tb = getenv("MORPH_HOST_TILE_BUDGET");  // tests shrink the budget to force small tiles
    const int TH = morph_tile_rows(Wpw, R, pl.nlev, tb ? atoll(tb) : 200 * 1024);
    if (TH > 0) {
      const int n_stage = (TH + 2 * R) * Wpw;
      uint32_t* smem = new uint32_t[(size_t)(1 + pl.nlev) * n_stage];
      uint32_t *in_s = smem, *lvl_s = smem + n_stage;
      const uint32_t flip = erode ? 0xffffffffu : 0u;
      for (int64_t t = 0; t < T; ++t)
        for (int y0 = 0; y0 < Hp; y0 += TH) {  // blockIdx.x
          for (int i = 0; i < (1 + pl.nlev) * n_stage; ++i) smem[i] = 0xdeadbeefu;  // shared memory starts undefined
          for (int item = 0; item < n_stage; ++item) in_s[item] = morph_tile_load_item(in + t * per_t, (int)Hp, Wpw, y0, R, item, flip);
          for (int item = 0; item < n_stage; ++item) morph_tile_h_item(in_s, lvl_s, n_stage, Wpw, pl, item, flip);
          const int n_out = ((int)Hp - y0 < TH ? (int)Hp - y0 : TH) * Wpw;
          for (int item = 0; item < n_out; ++item) {
            const int orow = item / Wpw, w = item - orow * Wpw;
            uint32_t res = morph_tile_v_item(in_s, lvl_s, n_stage, Wpw, pl, orow, w, flip);
            if (w == Wpw - 1) res &= morph_tailmask((int)Wp);
            out[t * per_t + (int64_t)(y0 + orow) * Wpw + w] = res;
          }
        }
      delete[] smem;
      return 0;
    }
  }
  for (int64_t t = 0; t < T; ++t)
    for (int y = 0; y < Hp; ++y)
      for (int w = 0; w < Wpw; ++w)
        out[t * per_t + (int64_t)y * Wpw + w] =
            morph_disk_word(in + t * per_t, (int)Hp, Wpw, morph_tailmask((int)Wp), y, w, d, erode ? 1 : 0, variant);
  return 0;
}

int32_t marex_morph_disk_levels(int32_t R) { return (R < 1 || R > MORPH_MAX_R) ? 0 : morph_make_plan(R).nlev; }

int marex_morph_disk_sep(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R, int32_t erode,
                         uint32_t* scratch, int64_t scratch_words, void*) {
  if (R < 1 || R > MORPH_MAX_R) return -3;
  const MorphPlan pl = morph_make_plan(R);
  const int Wpw = (int)((Wp + 31) >> 5);
  const int64_t per_t = Hp * Wpw;
  const int e = erode ? 1 : 0;
  return morph_disk_separable_chunks(
      T, per_t, pl.nlev, scratch_words,
      [&](int64_t t0, int64_t n, int64_t lvl_stride) -> int {
        for (int64_t tc = 0; tc < n; ++tc)  // blockIdx.y
          for (int r = 0; r < per_t; ++r)
            morph_disk_h_word(in + (t0 + tc) * per_t, Wpw, r / Wpw, r % Wpw, pl, scratch + tc * per_t + r, lvl_stride, e);
        return 0;
      },
      [&](int64_t t0, int64_t n, int64_t lvl_stride) -> int {
        for (int64_t tc = 0; tc < n; ++tc)
          for (int r = 0; r < per_t; ++r)
            out[(t0 + tc) * per_t + r] = morph_disk_v_word(in + (t0 + tc) * per_t, scratch + tc * per_t, lvl_stride, (int)Hp, Wpw,
                                                           morph_tailmask((int)Wp), r / Wpw, r % Wpw, pl, e);
        return 0;
      });
}

int marex_morph_time(const uint32_t* in, int64_t T_in, int64_t words, uint32_t* out, int64_t T_out, int32_t off, int32_t K,
                     int32_t erode, void*) {
  for (int64_t t = 0; t < T_out; ++t)
    for (int64_t i = 0; i < words; ++i) out[t * words + i] = morph_time_word(in, T_in, words, t, i, off, K, erode ? 1 : 0);
  return 0;
}

int marex_morph_extract(const uint8_t* bytes, const uint32_t* bits_in, int64_t t_pitch, int64_t row_stride, int64_t origin,
                        const uint32_t* mask_bits, int64_t T, int64_t ny, int64_t nx, uint8_t* events, int64_t events_pitch,
                        uint32_t* bits, int64_t bits_pitch, unsigned long long* count, void*) {
  MorphSrc s{bytes, bits_in, t_pitch, row_stride, origin, mask_bits};
  const int64_t N = ny * nx, nwords = (N + 31) / 32;
  for (int64_t t = 0; t < T; ++t)
    for (int64_t w = 0; w < nwords; ++w) {
      uint32_t word = 0;
      if (bits_in) {  // morph_extract_words_kernel
        word = morph_extract_word(s, T, t, w, (int)nx, N);
        if (events) {
          uint8_t* dst = events + t * events_pitch + w * 32;
          const int64_t ncell = N - w * 32 < 32 ? N - w * 32 : 32;
          if (ncell == 32) {
            for (int q = 0; q < 8; ++q) {
              const uint32_t four = morph_expand4(word >> (4 * q));
              for (int j = 0; j < 4; ++j) dst[4 * q + j] = (uint8_t)(four >> (8 * j));
            }
          } else {
            for (int j = 0; j < (int)ncell; ++j) dst[j] = (uint8_t)((word >> j) & 1u);
          }
        }
      } else {  // morph_extract_kernel
        for (int lane = 0; lane < 32; ++lane) {
          const int64_t c = w * 32 + lane;
          if (c >= N) break;
          const uint32_t b = morph_src_bit(s, t, (int)(c / nx), (int)(c % nx), (int)nx);
          if (events) events[t * events_pitch + c] = (uint8_t)b;
          word |= b << lane;
        }
      }
      if (bits) bits[t * bits_pitch + w] = word;
      if (count) *count += popc(word);
    }
  return 0;
}

int marex_morph_pack_u8(const uint8_t* events, int64_t T, int64_t N, int64_t pitch, uint32_t* bits, int64_t bits_pitch, void*) {
  const int64_t nwords = (N + 31) / 32;
  for (int64_t t = 0; t < T; ++t)
    for (int64_t w = 0; w < nwords; ++w) {
      const int64_t c0 = w * 32;
      bits[t * bits_pitch + w] = morph_pack_word(events + t * pitch + c0, (int)(N - c0 < 32 ? N - c0 : 32));
    }
  return 0;
}

int marex_morph_tpack(const uint8_t* bytes, const uint32_t* bits, int64_t t_pitch, int64_t T, int64_t N, uint32_t* dst, void*) {
  MorphSrc s{bytes, bits, t_pitch, 0, 0, nullptr};
  const int Tw = (int)marex_morph_tpack_words(T);
  for (int64_t c = 0; c < N; ++c)
    for (int k = 0; k < Tw; ++k) dst[c * Tw + k] = morph_tpack_word(s, T, c, k, Tw);
  return 0;
}

int marex_morph_nbr(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, const int32_t* nbr, int32_t nv,
                    const uint8_t* mask, int32_t erode, int32_t set_land, void*) {
  const int Tw = (int)marex_morph_tpack_words(T);
  for (int64_t c = 0; c < N; ++c)
    for (int k = 0; k < Tw; ++k) out[c * Tw + k] = morph_nbr_word(in, N, Tw, nbr, nv, mask, erode ? 1 : 0, set_land ? 1 : 0, c, k);
  return 0;
}

int marex_morph_tshift(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, int32_t half, int32_t erode, int32_t clip, void*) {
  if (half < 0 || half > 16) return -3;
  const int Tw = (int)marex_morph_tpack_words(T);
  for (int64_t c = 0; c < N; ++c)
    for (int k = 0; k < Tw; ++k) out[c * Tw + k] = morph_tshift_word(in, Tw, T, half, erode ? 1 : 0, clip ? 1 : 0, c, k);
  return 0;
}

int marex_morph_tunpack(const uint32_t* packed, int64_t T, int64_t N, const uint8_t* mask, uint8_t* events, int64_t events_pitch,
                        uint32_t* bits, int64_t bits_pitch, unsigned long long* count, void*) {
  const int Tw = (int)marex_morph_tpack_words(T);
  if (bits)
    for (int64_t t = 0; t < T; ++t)
      for (int64_t w = 0; w < (N + 31) / 32; ++w) bits[t * bits_pitch + w] = 0;
  for (int64_t c = 0; c < N; ++c)
    for (int64_t t = 0; t < T; ++t) {
      uint32_t b = (packed[c * Tw + 1 + t / 32] >> (t & 31)) & 1u;
      if (mask && !mask[c]) b = 0;
      if (events) events[t * events_pitch + c] = (uint8_t)b;
      if (bits) bits[t * bits_pitch + (c >> 5)] |= b << (c & 31);
      if (count) *count += b;
    }
  return 0;
}

void host_h1_steps(uint32_t* pcn, int steps) {
  for (int s = 0; s < steps; ++s) morph_h1(pcn[0], pcn[1], pcn[2]);
}

}  // extern "C"
