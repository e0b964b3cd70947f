// Tracker stage 1 on BIT-PACKED extreme-event masks: fill_holes (binary closing then opening
// with a disk, marEx/track.py:1520-1669) and fill_time_gaps (temporal closing, track.py:1671-1726), for gridded
// fields (row-aligned padded bit slabs, 32 cells per word) and unstructured meshes (cell-major words holding
// 32 TIME steps, so that one neighbour gather serves 32 days).  Integer / bit work, HBM- and L1-bound: no tensor cores.
// The per-word arithmetic lives in morph_core.cuh (shared with the host test harness).
#include <cstdlib>

#include "common.cuh"
#include "morph_core.cuh"

namespace marex {

// ---------------------------------------------------------------------------------------------------------
// gridded
// ---------------------------------------------------------------------------------------------------------

// Source cells (bool bytes, flattened bits, or the interior of another slab; optional ocean mask) -> padded slab.
// A warp produces 32 consecutive words of one padded row: lane = bit (coalesced byte / bit reads), the words are
// assembled with __ballot_sync and lane j keeps word j, so the store is one coalesced 128-byte line.
__global__ void __launch_bounds__(256) morph_pad_kernel(MorphSrc src, int64_t T, int ny, int nx, int pad, int wrap,
                                                        uint32_t* __restrict__ dst) {
  const int Hp = ny + 2 * pad, Wp = nx + 2 * pad, Wpw = (Wp + 31) >> 5, groups = (Wpw + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_items = T * Hp * groups;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    const int g = (int)(item % groups);
    const int yp = (int)((item / groups) % Hp);
    const int64_t t = item / ((int64_t)groups * Hp);
    const int ys = morph_pad_index(yp, pad, ny, wrap);
    uint32_t mine = 0;
    const int nw = min(32, Wpw - g * 32);
    for (int j = 0; j < nw; ++j) {
      const int xp = (g * 32 + j) * 32 + lane;
      uint32_t b = 0;
      if (xp < Wp) b = morph_src_bit(src, t, ys, morph_pad_index(xp, pad, nx, wrap), nx);
      const uint32_t word = __ballot_sync(0xffffffffu, b != 0);
      if (lane == j) mine = word;
    }
    if (lane < nw) dst[(t * Hp + yp) * Wpw + g * 32 + lane] = mine;
  }
}

// The same for a BITS source (flattened bits or the interior of another slab): thread = output word, runs of source
// cells are fetched 32 bits at a time with two loads and a funnel shift (morph_pad_word), no per-bit work.
__global__ void __launch_bounds__(256) morph_pad_words_kernel(MorphSrc src, int64_t T, int ny, int nx, int pad, int wrap,
                                                              uint32_t* __restrict__ dst) {
  const int Hp = ny + 2 * pad, Wp = nx + 2 * pad, Wpw = (Wp + 31) >> 5;
  const int64_t per_t = (int64_t)Hp * Wpw, total = T * per_t;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / per_t;
    const int r = (int)(idx - t * per_t);
    const int yp = r / Wpw, wp = r - yp * Wpw;
    dst[idx] = morph_pad_word(src, T, t, morph_pad_index(yp, pad, ny, wrap), wp, Wp, pad, ny, nx, wrap);
  }
}

// One morphological pass over all padded time steps; thread = output word, w fastest (coalesced; the 3 x (2R+1)
// input words of neighbouring threads overlap and are served by L1).
__global__ void __launch_bounds__(256) morph_disk_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         int Hp, int Wpw, uint32_t tailmask, const __grid_constant__ MorphDisk disk, int erode,
                                                         int variant) {
  const int per_t = Hp * Wpw;  // < 2^31 (checked by the caller); blockIdx.y = time step: no 64-bit division
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= per_t) return;
  const int64_t base = (int64_t)blockIdx.y * per_t;
  const int y = r / Wpw, w = r - y * Wpw;
  out[base + r] = morph_disk_word(in + base, Hp, Wpw, tailmask, y, w, disk, erode, variant);
}

// Separable form (morph_core.cuh): pass H widens every input word once and stores it at the disk's distinct half-widths,
// pass V ORs 2R+1 single words.  blockIdx.y = time step inside the chunk whose level buffers fit the scratch (and L2).
__global__ void __launch_bounds__(256) morph_disk_h_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ hbuf,
                                                           int64_t lvl_stride, int Hp, int Wpw,
                                                           const __grid_constant__ MorphPlan plan, int erode) {
  const int per_t = Hp * Wpw;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= per_t) return;
  const int64_t base = (int64_t)blockIdx.y * per_t;
  const int y = r / Wpw, w = r - y * Wpw;
  morph_disk_h_word(in + base, Wpw, y, w, plan, hbuf + base + r, lvl_stride, erode);
}

__global__ void __launch_bounds__(256) morph_disk_v_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ hbuf,
                                                           int64_t lvl_stride, uint32_t* __restrict__ out, int Hp, int Wpw,
                                                           uint32_t tailmask, const __grid_constant__ MorphPlan plan,
                                                           int erode) {
  const int per_t = Hp * Wpw;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= per_t) return;
  const int64_t base = (int64_t)blockIdx.y * per_t;
  const int y = r / Wpw, w = r - y * Wpw;
  out[base + r] = morph_disk_v_word(in + base, hbuf + base, lvl_stride, Hp, Wpw, tailmask, y, w, plan, erode);
}

// Variant 4: the separable pass inside one shared-memory tile (morph_core.cuh).  grid = (tiles of TH rows, time steps).
__global__ void __launch_bounds__(256) morph_disk_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int Hp,
                                                              int Wpw, uint32_t tailmask, const __grid_constant__ MorphPlan plan,
                                                              int erode, int TH) {
  extern __shared__ uint32_t tile_smem[];
  const int R = plan.R, rows = TH + 2 * R, n_stage = rows * Wpw;
  uint32_t* in_s = tile_smem;
  uint32_t* lvl_s = tile_smem + n_stage;
  const uint32_t flip = erode ? 0xffffffffu : 0u;
  const int64_t base = (int64_t)blockIdx.y * Hp * Wpw;
  const int y0 = blockIdx.x * TH;
  for (int item = threadIdx.x; item < n_stage; item += blockDim.x)
    in_s[item] = morph_tile_load_item(in + base, Hp, Wpw, y0, R, item, flip);
  __syncthreads();
  for (int item = threadIdx.x; item < n_stage; item += blockDim.x) morph_tile_h_item(in_s, lvl_s, n_stage, Wpw, plan, item, flip);
  __syncthreads();
  const int n_out = min(TH, Hp - y0) * Wpw;
  for (int item = threadIdx.x; item < n_out; item += blockDim.x) {
    const int orow = item / Wpw, w = item - orow * Wpw;
    uint32_t res = morph_tile_v_item(in_s, lvl_s, n_stage, Wpw, plan, orow, w, flip);
    if (w == Wpw - 1) res &= tailmask;
    out[base + (int64_t)(y0 + orow) * Wpw + w] = res;
  }
}

// Temporal dilation / erosion of whole slabs, bit-parallel over the 32 cells of a word.
__global__ void __launch_bounds__(256) morph_time_kernel(const uint32_t* __restrict__ in, int64_t T_in, int64_t words,
                                                         uint32_t* __restrict__ out, int64_t T_out, int off, int K,
                                                         int erode) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T_out * words) return;
  const int64_t t = idx / words, i = idx - t * words;
  out[idx] = morph_time_word(in, T_in, words, t, i, off, K, erode);
}

// Interior of a slab (or any MorphSrc) -> the layouts the rest of the pipeline uses: bool bytes [T, N] and / or
// flattened bits (bit c & 31 of word c >> 5, the layout of marex_compare_*), ocean mask applied, cells counted.
__global__ void __launch_bounds__(256) morph_extract_kernel(MorphSrc src, int64_t T, int ny, int nx,
                                                            uint8_t* __restrict__ events, int64_t events_pitch,
                                                            uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                            unsigned long long* __restrict__ count) {
  const int64_t N = (int64_t)ny * nx, nwords = (N + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_items = T * nwords;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned int local = 0;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    const int64_t t = item / nwords, w = item - t * nwords;
    const int64_t c = w * 32 + lane;
    uint32_t b = 0;
    if (c < N) {
      const int y = (int)(c / nx), x = (int)(c - (int64_t)y * nx);
      b = morph_src_bit(src, t, y, x, nx);
      if (events) events[t * events_pitch + c] = (uint8_t)b;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, b != 0);
    if (lane == 0) {
      if (bits) bits[t * bits_pitch + w] = word;
      local += __popc(word);
    }
  }
  if (count && lane == 0 && local) atomicAdd(count, (unsigned long long)local);
}

// The same for a BITS source: thread = output word (morph_extract_word), bool bytes written as two 16-byte stores,
// True cells counted per thread and reduced once per block.
__global__ void __launch_bounds__(256) morph_extract_words_kernel(MorphSrc src, int64_t T, int ny, int nx,
                                                                  uint8_t* __restrict__ events, int64_t events_pitch,
                                                                  uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                                  unsigned long long* __restrict__ count) {
  const int64_t N = (int64_t)ny * nx, nwords = (N + 31) >> 5, total = T * nwords;
  unsigned int local = 0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / nwords, w = idx - t * nwords;
    const uint32_t word = morph_extract_word(src, T, t, w, nx, N);
    local += __popc(word);
    if (bits) bits[t * bits_pitch + w] = word;
    if (events) {
      uint8_t* dst = events + t * events_pitch + w * 32;
      const int64_t ncell = min((int64_t)32, N - w * 32);
      if (ncell == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        uint4 a, b;
        a.x = morph_expand4(word), a.y = morph_expand4(word >> 4), a.z = morph_expand4(word >> 8), a.w = morph_expand4(word >> 12);
        b.x = morph_expand4(word >> 16), b.y = morph_expand4(word >> 20), b.z = morph_expand4(word >> 24), b.w = morph_expand4(word >> 28);
        reinterpret_cast<uint4*>(dst)[0] = a;
        reinterpret_cast<uint4*>(dst)[1] = b;
      } else {
        for (int j = 0; j < (int)ncell; ++j) dst[j] = (uint8_t)((word >> j) & 1u);
      }
    }
  }
  if (count) {
    __shared__ unsigned int block_sum;
    if (threadIdx.x == 0) block_sum = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&block_sum, local);
    __syncthreads();
    if (threadIdx.x == 0 && block_sum) atomicAdd(count, (unsigned long long)block_sum);
  }
}

// Bool bytes [T, N] -> flattened bits: thread = output word, two 16-byte loads, eight multiply-gathers.
__global__ void __launch_bounds__(256) morph_pack_kernel(const uint8_t* __restrict__ events, int64_t T, int64_t N, int64_t pitch,
                                                         uint32_t* __restrict__ bits, int64_t bits_pitch) {
  const int64_t nwords = (N + 31) >> 5, total = T * nwords;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / nwords, w = idx - t * nwords;
    const int64_t c0 = w * 32;
    bits[t * bits_pitch + w] = morph_pack_word(events + t * pitch + c0, (int)min((int64_t)32, N - c0));
  }
}

// ---------------------------------------------------------------------------------------------------------
// unstructured: cell-major, time-packed.  Word (c, k) holds time steps 32*(k-1) .. 32*(k-1)+31 of cell c: word 0 and
// the last word of every cell are margins, so that the temporal closing can look 32 steps past either end.
// ---------------------------------------------------------------------------------------------------------

// [T, N] bytes or flattened bits -> cell-major time-packed.  lane = cell (coalesced reads of one time step), every
// lane assembles its own word over 32 time steps.
__global__ void __launch_bounds__(256) morph_tpack_kernel(MorphSrc src, int64_t T, int64_t N, int Tw,
                                                          uint32_t* __restrict__ dst) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;  // 0 .. Tw-1
  if (c >= N) return;
  dst[c * Tw + k] = morph_tpack_word(src, T, c, k, Tw);
}

// One application of the sparse dilation matrix (neighbours + identity, track.py:1093-1115, 5423-5470) to 32 time
// steps at once: out[c] = in[c] | OR_j in[nb[j, c]] (negative neighbour = none).  flip = 1 computes the erosion
// `~dilate(~x)`.  set_land = 1 first forces cells outside the mask to True (`bitmap[:, ~mask] = True`,
// track.py:1566, 1574), which is applied to the INPUT of this pass (own cell and gathered neighbours alike).
__global__ void __launch_bounds__(256) morph_nbr_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        int64_t N, int Tw, const int32_t* __restrict__ nbr, int nv,
                                                        const uint8_t* __restrict__ mask, int flip, int set_land) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Tw) return;
  const int64_t c = idx / Tw;
  const int k = (int)(idx - c * Tw);
  out[idx] = morph_nbr_word(in, N, Tw, nbr, nv, mask, flip, set_land, c, k);
}

// Dilation (flip = 0) or erosion (flip = 1) by +-half steps ALONG TIME, i.e. along the bit axis of a cell's words.
// clip = 1 reads only the bits of real time steps [0, T) (everything else False: the constant padding of
// track.py:1706); the erosion pass reads the dilated margins as they are, exactly like scipy on the padded axis.
__global__ void __launch_bounds__(256) morph_tshift_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                           int64_t N, int Tw, int64_t T, int half, int flip, int clip) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Tw) return;
  const int64_t c = idx / Tw;
  const int k = (int)(idx - c * Tw);
  out[idx] = morph_tshift_word(in, Tw, T, half, flip, clip, c, k);
}

// cell-major time-packed -> [T, N] bool bytes and / or flattened bits (+ optional ocean mask and count).
__global__ void __launch_bounds__(256) morph_tunpack_kernel(const uint32_t* __restrict__ src, int64_t T, int64_t N, int Tw,
                                                            const uint8_t* __restrict__ mask,
                                                            uint8_t* __restrict__ events, int64_t events_pitch,
                                                            uint32_t* __restrict__ bits, int64_t bits_pitch,
                                                            unsigned long long* __restrict__ count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim multiple of 32: a warp = one bits word
  const int k = blockIdx.y + 1;                                      // 1 .. Tw-2
  const bool live = c < N;
  uint32_t word = live ? src[c * Tw + k] : 0u;
  if (live && mask && !mask[c]) word = 0u;
  const int64_t t0 = (int64_t)(k - 1) * 32;
  const int nt = (int)min((int64_t)32, T - t0);
  unsigned int local = 0;
  for (int j = 0; j < nt; ++j) {
    const uint32_t b = (word >> j) & 1u;
    if (events && live) events[(t0 + j) * events_pitch + c] = (uint8_t)b;
    const uint32_t packed = __ballot_sync(0xffffffffu, b != 0);
    if ((threadIdx.x & 31) == 0) {
      if (bits && (c >> 5) < bits_pitch) bits[(t0 + j) * bits_pitch + (c >> 5)] = packed;
      local += __popc(packed);
    }
  }
  if (count && (threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
}

static int grid_for(int64_t n_threads, int block, int64_t* out_blocks) {
  const int64_t b = (n_threads + block - 1) / block;
  if (b <= 0 || b > 2147483647LL) return -1;
  *out_blocks = b;
  return 0;
}

static MorphSrc make_src(const uint8_t* bytes, const uint32_t* bits, int64_t t_pitch, int64_t row_stride, int64_t origin,
                         const uint32_t* mask_bits) {
  MorphSrc s;
  s.bytes = bytes;
  s.bits = bits;
  s.t_pitch = t_pitch;
  s.row_stride = row_stride;
  s.origin = origin;
  s.mask_bits = mask_bits;
  return s;
}

}  // namespace marex

using namespace marex;

extern "C" int64_t marex_morph_slab_words(int64_t ny, int64_t nx, int32_t pad) {
  if (ny <= 0 || nx <= 0 || pad < 0) return -1;
  return (ny + 2 * pad) * ((nx + 2 * pad + 31) / 32);
}

extern "C" int marex_morph_pad_bits(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch,
                                    int64_t src_row_stride, int64_t src_origin, const uint32_t* mask_bits, int64_t T,
                                    int64_t ny, int64_t nx, int32_t pad, int32_t wrap, uint32_t* slab, void* stream) {
  MAREX_REQUIRE((src_bytes != nullptr) != (src_bits != nullptr), "exactly one of src_bytes / src_bits");
  MAREX_REQUIRE(slab && T > 0 && ny > 0 && nx > 0 && pad >= 0, "bad shape");
  MAREX_REQUIRE(ny + 2 * (int64_t)pad < (1 << 30) && nx + 2 * (int64_t)pad < (1 << 30), "grid too large");
  const int Hp = (int)ny + 2 * pad, Wpw = ((int)nx + 2 * pad + 31) >> 5, groups = (Wpw + 31) >> 5;
  const MorphSrc src = make_src(src_bytes, src_bits, src_t_pitch, src_row_stride, src_origin, mask_bits);
  if (src_bits) {
    const int64_t blocks = std::min<int64_t>((T * Hp * Wpw + 255) / 256, (int64_t)sm_count() * 64);
    morph_pad_words_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, T, (int)ny, (int)nx, pad, wrap, slab);
    MAREX_LAUNCH_CHECK("morph_pad_words_kernel");
    return MAREX_OK;
  }
  const int64_t items = T * Hp * groups;
  const int64_t blocks = std::min<int64_t>((items + 7) / 8, (int64_t)sm_count() * 64);
  morph_pad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, T, (int)ny, (int)nx, pad, wrap, slab);
  MAREX_LAUNCH_CHECK("morph_pad_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_disk(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R,
                                int32_t erode, void* stream) {
  MAREX_REQUIRE(in && out && in != out && T > 0 && Hp > 0 && Wp > 0, "bad arguments");
  if (R < 0 || R > MORPH_MAX_R) return fail(MAREX_ERR_UNSUPPORTED, "R_fill must be in 0..32");
  const MorphDisk d = morph_make_disk(R);
  // measured on B200 (0.25 deg, R = 8, 2048 days, profiles/r02_first_call_parked_variants.json): shared-memory tile
  // (4) 1.41 ms per pass, branch-free direct kernel (3) 1.45 / 1.62 ms, first direct kernel (2) 1.77 / 1.96 ms
  const int env_variant = (int)tune_get("morph_disk", 4);
  const int variant = env_variant == 2 ? 2 : 3;
  const int Wpw = (int)((Wp + 31) >> 5);
  MAREX_REQUIRE(Hp * (int64_t)Wpw < (1LL << 31), "padded time step too large");
  const int64_t per_t = Hp * (int64_t)Wpw;
  if (env_variant == 4 && R >= 1) {  // shared-memory tile variant; falls through to the direct kernel when a tile does not fit
    const MorphPlan pl = morph_make_plan(R);
    const int TH = morph_tile_rows(Wpw, R, pl.nlev, 200 * 1024);
    if (TH > 0) {
      const size_t smem = (size_t)(1 + pl.nlev) * (TH + 2 * R) * Wpw * 4;
      cudaError_t e = cudaFuncSetAttribute(morph_disk_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(morph_disk_tile_kernel)");
      for (int64_t t0 = 0; t0 < T; t0 += 65535) {
        const dim3 grid((unsigned)((Hp + TH - 1) / TH), (unsigned)std::min<int64_t>(65535, T - t0));
        morph_disk_tile_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(in + t0 * per_t, out + t0 * per_t, (int)Hp, Wpw,
                                                                          morph_tailmask((int)Wp), pl, erode ? 1 : 0, TH);
        MAREX_LAUNCH_CHECK("morph_disk_tile_kernel");
      }
      return MAREX_OK;
    }
  }
  for (int64_t t0 = 0; t0 < T; t0 += 65535) {  // gridDim.y <= 65535
    const dim3 grid((unsigned)((per_t + 255) / 256), (unsigned)std::min<int64_t>(65535, T - t0));
    morph_disk_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in + t0 * per_t, out + t0 * per_t, (int)Hp, Wpw,
                                                              morph_tailmask((int)Wp), d, erode ? 1 : 0, variant);
    MAREX_LAUNCH_CHECK("morph_disk_kernel");
  }
  return MAREX_OK;
}

extern "C" int32_t marex_morph_disk_levels(int32_t R) { return (R < 1 || R > MORPH_MAX_R) ? 0 : morph_make_plan(R).nlev; }

extern "C" int marex_morph_disk_sep(const uint32_t* in, uint32_t* out, int64_t T, int64_t Hp, int64_t Wp, int32_t R,
                                    int32_t erode, uint32_t* scratch, int64_t scratch_words, void* stream) {
  MAREX_REQUIRE(in && out && scratch && in != out && T > 0 && Hp > 0 && Wp > 0, "bad arguments");
  if (R < 1 || R > MORPH_MAX_R) return fail(MAREX_ERR_UNSUPPORTED, "R_fill must be in 1..32");
  const MorphPlan pl = morph_make_plan(R);
  const int Wpw = (int)((Wp + 31) >> 5);
  MAREX_REQUIRE(Hp * (int64_t)Wpw < (1LL << 31), "padded time step too large");
  const int64_t per_t = Hp * (int64_t)Wpw;
  const unsigned gx = (unsigned)((per_t + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int e = erode ? 1 : 0;
  const uint32_t tail = morph_tailmask((int)Wp);
  const int rc = morph_disk_separable_chunks(
      T, per_t, pl.nlev, scratch_words,
      [&](int64_t t0, int64_t n, int64_t lvl_stride) -> int {
        morph_disk_h_kernel<<<dim3(gx, (unsigned)n), 256, 0, st>>>(in + t0 * per_t, scratch, lvl_stride, (int)Hp, Wpw, pl, e);
        MAREX_LAUNCH_CHECK("morph_disk_h_kernel");
        return MAREX_OK;
      },
      [&](int64_t t0, int64_t n, int64_t lvl_stride) -> int {
        morph_disk_v_kernel<<<dim3(gx, (unsigned)n), 256, 0, st>>>(in + t0 * per_t, scratch, lvl_stride, out + t0 * per_t,
                                                                    (int)Hp, Wpw, tail, pl, e);
        MAREX_LAUNCH_CHECK("morph_disk_v_kernel");
        return MAREX_OK;
      });
  if (rc == -1) return fail(MAREX_ERR_INVALID_ARG, "scratch smaller than marex_morph_disk_levels(R) padded time steps");
  return rc;
}

extern "C" int marex_morph_time(const uint32_t* in, int64_t T_in, int64_t words, uint32_t* out, int64_t T_out, int32_t off,
                                int32_t K, int32_t erode, void* stream) {
  MAREX_REQUIRE(in && out && in != out && T_in > 0 && T_out > 0 && words > 0 && K >= 1, "bad arguments");
  int64_t blocks;
  if (grid_for(T_out * words, 256, &blocks)) return fail(MAREX_ERR_INVALID_ARG, "mask too large for one launch");
  morph_time_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, T_in, words, out, T_out, off, K, erode ? 1 : 0);
  MAREX_LAUNCH_CHECK("morph_time_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_extract(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch,
                                   int64_t src_row_stride, int64_t src_origin, const uint32_t* mask_bits, int64_t T,
                                   int64_t ny, int64_t nx, uint8_t* events, int64_t events_pitch, uint32_t* bits,
                                   int64_t bits_pitch, unsigned long long* count, void* stream) {
  MAREX_REQUIRE((src_bytes != nullptr) != (src_bits != nullptr), "exactly one of src_bytes / src_bits");
  MAREX_REQUIRE(T > 0 && ny > 0 && nx > 0 && (events || bits || count), "bad arguments");
  const int64_t N = ny * nx;
  MAREX_REQUIRE(!events || events_pitch >= N, "events_pitch < N");
  MAREX_REQUIRE(!bits || bits_pitch >= (N + 31) / 32, "bits_pitch < ceil(N / 32)");
  const MorphSrc src = make_src(src_bytes, src_bits, src_t_pitch, src_row_stride, src_origin, mask_bits);
  const int64_t items = T * ((N + 31) / 32);
  if (src_bits) {
    const int64_t blocks = std::min<int64_t>((items + 255) / 256, (int64_t)sm_count() * 32);
    morph_extract_words_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, T, (int)ny, (int)nx, events,
                                                                                     events_pitch, bits, bits_pitch, count);
    MAREX_LAUNCH_CHECK("morph_extract_words_kernel");
    return MAREX_OK;
  }
  const int64_t blocks = std::min<int64_t>((items + 7) / 8, (int64_t)sm_count() * 64);
  morph_extract_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, T, (int)ny, (int)nx, events, events_pitch,
                                                                             bits, bits_pitch, count);
  MAREX_LAUNCH_CHECK("morph_extract_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_pack_u8(const uint8_t* events, int64_t T, int64_t N, int64_t pitch, uint32_t* bits, int64_t bits_pitch,
                                   void* stream) {
  MAREX_REQUIRE(events && bits && T > 0 && N > 0 && pitch >= N && bits_pitch >= (N + 31) / 32, "bad arguments");
  const int64_t items = T * ((N + 31) / 32);
  const int64_t blocks = std::min<int64_t>((items + 255) / 256, (int64_t)sm_count() * 32);
  morph_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(events, T, N, pitch, bits, bits_pitch);
  MAREX_LAUNCH_CHECK("morph_pack_kernel");
  return MAREX_OK;
}

extern "C" int64_t marex_morph_tpack_words(int64_t T) { return T > 0 ? (T + 31) / 32 + 2 : -1; }

extern "C" int marex_morph_tpack(const uint8_t* src_bytes, const uint32_t* src_bits, int64_t src_t_pitch, int64_t T,
                                 int64_t N, uint32_t* packed, void* stream) {
  MAREX_REQUIRE((src_bytes != nullptr) != (src_bits != nullptr), "exactly one of src_bytes / src_bits");
  MAREX_REQUIRE(packed && T > 0 && N > 0, "bad shape");
  const int64_t Tw = marex_morph_tpack_words(T);
  MAREX_REQUIRE(Tw <= 65535, "time axis too long for one launch");
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)Tw);
  morph_tpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(make_src(src_bytes, src_bits, src_t_pitch, 0, 0, nullptr), T, N,
                                                             (int)Tw, packed);
  MAREX_LAUNCH_CHECK("morph_tpack_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_nbr(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, const int32_t* nbr, int32_t nv,
                               const uint8_t* mask, int32_t erode, int32_t set_land, void* stream) {
  MAREX_REQUIRE(in && out && in != out && nbr && T > 0 && N > 0 && nv >= 0, "bad arguments");
  MAREX_REQUIRE(!set_land || mask, "set_land needs the mask");
  const int64_t Tw = marex_morph_tpack_words(T);
  int64_t blocks;
  if (grid_for(N * Tw, 256, &blocks)) return fail(MAREX_ERR_INVALID_ARG, "mask too large for one launch");
  morph_nbr_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, N, (int)Tw, nbr, nv, mask, erode ? 1 : 0,
                                                                         set_land ? 1 : 0);
  MAREX_LAUNCH_CHECK("morph_nbr_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_tshift(const uint32_t* in, uint32_t* out, int64_t T, int64_t N, int32_t half, int32_t erode,
                                  int32_t clip, void* stream) {
  MAREX_REQUIRE(in && out && in != out && T > 0 && N > 0, "bad arguments");
  if (half < 0 || half > 16) return fail(MAREX_ERR_UNSUPPORTED, "T_fill must be in 0..32");
  const int64_t Tw = marex_morph_tpack_words(T);
  int64_t blocks;
  if (grid_for(N * Tw, 256, &blocks)) return fail(MAREX_ERR_INVALID_ARG, "mask too large for one launch");
  morph_tshift_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, N, (int)Tw, T, half, erode ? 1 : 0,
                                                                            clip ? 1 : 0);
  MAREX_LAUNCH_CHECK("morph_tshift_kernel");
  return MAREX_OK;
}

extern "C" int marex_morph_tunpack(const uint32_t* packed, int64_t T, int64_t N, const uint8_t* mask, uint8_t* events,
                                   int64_t events_pitch, uint32_t* bits, int64_t bits_pitch, unsigned long long* count,
                                   void* stream) {
  MAREX_REQUIRE(packed && T > 0 && N > 0 && (events || bits || count), "bad arguments");
  MAREX_REQUIRE(!events || events_pitch >= N, "events_pitch < N");
  MAREX_REQUIRE(!bits || bits_pitch >= (N + 31) / 32, "bits_pitch < ceil(N / 32)");
  const int64_t Tw = marex_morph_tpack_words(T);
  MAREX_REQUIRE(Tw <= 65535, "time axis too long for one launch");
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)(Tw - 2));
  morph_tunpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, T, N, (int)Tw, mask, events, events_pitch, bits,
                                                               bits_pitch, count);
  MAREX_LAUNCH_CHECK("morph_tunpack_kernel");
  return MAREX_OK;
}
