// Stage 2b of the SLIM hot path on B200: multi-level bilinear lookup into the correlation pyramid.
//
// Replaces CorrBlock.__call__ (liso/slim/model/raft_code/corr.py:23-46) + bilinear_sampler
// (raft_code/utils.py:15-29) + F.grid_sample(align_corners=True, zeros padding):
//   out[b, l*(2r+1)^2 + i*(2r+1) + j, y, x] = bilinear(pyr_l[b, (y,x)], (cx / 2^l + i - r, cy / 2^l + j - r))
// i.e. the FIRST window index offsets x (RAFT's transposed window), OOB taps contribute zero.
//
// B200 design: a gather kernel bound by DRAM sector fetches.  One CTA owns 32 consecutive source pixels of one
// sample and one lane owns one pixel (lane == pixel in every phase, so the fp32 output is stored with one
// coalesced 128-byte row per warp instruction and shared memory is addressed [..][lane], conflict-free):
//   phase 1  sample positions per (pixel, level, axis, offset), with the reference's exact normalise /
//            un-normalise fp32 arithmetic; per-offset bilinear weights with the zero padding folded in
//   phase 2  every (pixel, level, window row) fetches its (2r+2)-element row segment of the pyramid with
//            16-byte loads (all loads of the CTA are independent and in flight together) into shared memory;
//            in the pyramid layout (include/slimb200.h) 4 neighbouring pixels x 8 columns share a 64-byte unit, so the
//            lanes 4k..4k+3 of a load are served by the same DRAM burst
//   phase 3  49 outputs per (pixel, level) from the (2r+2)^2 window in shared memory: 4 taps x weights
// Window rows that straddle rounding (floor of offset o != floor of offset 0 + o, |prob| ~ 1e-6) take a slow,
// fully predicated global-memory path so that results always follow the reference arithmetic.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int LK_PIX = 32;
constexpr int LK_THREADS = 256;
constexpr int LK_WARPS = LK_THREADS / 32;
constexpr int PW = SLIMB200_PANEL_COLS;

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int EPC = 4;  // elements per 16-byte chunk
  __device__ static __forceinline__ float ld(const float* p) { return __ldg(p); }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int EPC = 8;
  __device__ static __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
};

// sample position in level pixels: bilinear_sampler's normalisation (utils.py:19-20) followed by grid_sample's
// un-normalisation ((g + 1) / 2) * (size - 1), all in fp32 with IEEE division
__device__ __forceinline__ float sample_pos(float c, float inv, int offs, int size) {
  const float pos = __fadd_rn(c * inv, (float)offs);
  const float sm1 = (float)(size - 1);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, pos), sm1), 1.f);
  float ip = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), sm1);  // x / 2 == x * 0.5 exactly
  if (!(fabsf(ip) < 1e7f)) ip = -1e7f;                             // NaN / inf / far away: every tap is outside
  return ip;
}

// element (b, source pixel i, column col) of the pyramid (include/slimb200.h): 16 KB half-tiles of 128 rows x 64
// columns in which 4 neighbouring rows x 8 columns form one 64-byte unit.  `rows` = rows_padded of the layout.
template <typename T>
__device__ __forceinline__ size_t panel_index(int n_panels, int rows, int b, int i, int col) {
  const int c = col % PW;
  const size_t half_tile = ((size_t)(b * n_panels + (col / PW)) * (rows >> 7) + (i >> 7)) * 2 + (c >> 6);
  return half_tile * 8192 + (size_t)(((i & 127) >> 2) * 256 + ((c & 63) >> 3) * 32 + (i & 3) * 8 + (c & 7));
}

// fully predicated 4-tap sample straight from global memory (rare path)
template <typename T>
__device__ __noinline__ float sample_slow(const T* __restrict__ pyr, int n_panels, int rows, int b, int i, int W, int H, int off,
                                          float ix, float iy) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  const float dx = __fsub_rn(ix, fx), dy = __fsub_rn(iy, fy);
  const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
  float val = 0.f;
  if (yin0 && xin0) val += Elem<T>::ld(pyr + panel_index<T>(n_panels, rows, b, i, off + y0 * W + x0)) * __fmul_rn(ex, ey);
  if (yin0 && xin1) val += Elem<T>::ld(pyr + panel_index<T>(n_panels, rows, b, i, off + y0 * W + x0 + 1)) * __fmul_rn(dx, ey);
  if (yin1 && xin0) val += Elem<T>::ld(pyr + panel_index<T>(n_panels, rows, b, i, off + (y0 + 1) * W + x0)) * __fmul_rn(ex, dy);
  if (yin1 && xin1) val += Elem<T>::ld(pyr + panel_index<T>(n_panels, rows, b, i, off + (y0 + 1) * W + x0 + 1)) * __fmul_rn(dx, dy);
  return val;
}

template <typename T, int R>
struct Cfg {
  static constexpr int WIN = 2 * R + 1;
  static constexpr int W1 = WIN + 1;                                         // window rows / cols fetched
  static constexpr int EPC = Elem<T>::EPC;
  static constexpr int NCHUNK = (EPC - 1 + W1 + EPC - 1) / EPC;              // 16-byte chunks covering any W1-element segment
  static constexpr int NW = NCHUNK * 4;                                      // 32-bit words per segment
  // dynamic shared memory in 4-byte words for `levels` levels
  __host__ __device__ static constexpr int raw_words(int levels) { return levels * W1 * NW * LK_PIX; }
  __host__ __device__ static constexpr int tab_words(int levels) { return levels * 2 * WIN * LK_PIX; }
  __host__ __device__ static constexpr int smem_bytes(int levels) {
    return (raw_words(levels) + 3 * tab_words(levels) + 3 * levels * LK_PIX) * 4;
  }
};

// two neighbouring elements (e, e + 1) of a segment held as words [..][lane] in shared memory
template <typename T>
__device__ __forceinline__ void tap_pair(const uint32_t* seg_words, int e, float& v0, float& v1);
template <>
__device__ __forceinline__ void tap_pair<float>(const uint32_t* seg_words, int e, float& v0, float& v1) {
  v0 = __uint_as_float(seg_words[e * LK_PIX]);
  v1 = __uint_as_float(seg_words[(e + 1) * LK_PIX]);
}
template <>
__device__ __forceinline__ void tap_pair<__nv_bfloat16>(const uint32_t* seg_words, int e, float& v0, float& v1) {
  const int w = e >> 1;
  const uint32_t lo = seg_words[w * LK_PIX], hi = seg_words[(w + 1) * LK_PIX];
  const uint32_t t = __funnelshift_r(lo, hi, (e & 1) * 16);  // [bf16 e | bf16 e+1]
  v0 = __uint_as_float(t << 16);
  v1 = __uint_as_float(t & 0xffff0000u);
}

template <typename T, int R>
__global__ void __launch_bounds__(LK_THREADS) k_corr_lookup(const T* __restrict__ pyr, const slimb200_corr_layout L,
                                                            const float* __restrict__ coords, float* __restrict__ out) {
  using C = Cfg<T, R>;
  constexpr int WIN = C::WIN, W1 = C::W1, EPC = C::EPC, NCHUNK = C::NCHUNK, NW = C::NW;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int levels = L.levels;
  uint32_t* s_raw = s_dyn;                                                    // [level][row][word][lane]
  float* s_pos = reinterpret_cast<float*>(s_dyn + C::raw_words(levels));      // [level][axis][offset][lane] sample position
  float* s_w0 = s_pos + C::tab_words(levels);                                 //   weight of the floor tap (0 if outside)
  float* s_w1 = s_w0 + C::tab_words(levels);                                  //   weight of the floor + 1 tap
  int* s_xb = reinterpret_cast<int*>(s_w1 + C::tab_words(levels));            // [level][lane] window origin x
  int* s_yb = s_xb + levels * LK_PIX;                                         // [level][lane] window origin y
  int* s_ok = s_yb + levels * LK_PIX;                                         // [level][lane] offsets consistent with origin
  __shared__ float s_xy[2][LK_PIX];
  __shared__ int s_lw[SLIMB200_MAX_LEVELS], s_lh[SLIMB200_MAX_LEVELS], s_lo[SLIMB200_MAX_LEVELS];  // no dynamic indexing of L

  const int nf = L.h * L.w;
  const int prow = L.rows_padded;
  const int n_panels = L.n_panels;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * LK_PIX;
  const int lane = lane_id(), warp = warp_id();
  const int pix = i0 + lane;
  const bool live = pix < nf;
  const int n_ch = levels * WIN * WIN;

  if (threadIdx.x < 2 * LK_PIX) {
    const int ch = threadIdx.x / LK_PIX;
    s_xy[ch][lane] = live ? __ldg(coords + ((size_t)b * 2 + ch) * nf + pix) : 0.f;
  }
  if (threadIdx.x == 64) {
    s_lw[0] = L.level_w[0]; s_lw[1] = L.level_w[1]; s_lw[2] = L.level_w[2]; s_lw[3] = L.level_w[3];
    s_lh[0] = L.level_h[0]; s_lh[1] = L.level_h[1]; s_lh[2] = L.level_h[2]; s_lh[3] = L.level_h[3];
    s_lo[0] = L.level_offset[0]; s_lo[1] = L.level_offset[1]; s_lo[2] = L.level_offset[2]; s_lo[3] = L.level_offset[3];
  }
  __syncthreads();

  // ---- phase 1: positions and weights, one (level, axis, offset) row of the tables per warp iteration ----
  for (int e = warp; e < levels * 2 * WIN; e += LK_WARPS) {
    const int l = e / (2 * WIN), rem = e - l * 2 * WIN;
    const int axis = rem / WIN, o = rem - axis * WIN;
    const int size = axis == 0 ? s_lw[l] : s_lh[l];
    const float inv = 1.0f / (float)(1 << l);  // coords / 2**l is exact
    const float ip = sample_pos(s_xy[axis][lane], inv, o - R, size);
    const float f = floorf(ip);
    const int i0p = (int)f;
    const float w_hi = __fsub_rn(ip, f);                  // weight of tap floor + 1  (ix - ix_nw)
    const float w_lo = __fsub_rn(__fadd_rn(f, 1.f), ip);  // weight of tap floor      (ix_se - ix)
    s_pos[e * LK_PIX + lane] = ip;
    s_w0[e * LK_PIX + lane] = ((unsigned)i0p < (unsigned)size) ? w_lo : 0.f;
    s_w1[e * LK_PIX + lane] = ((unsigned)(i0p + 1) < (unsigned)size) ? w_hi : 0.f;
  }
  __syncthreads();
  for (int l = warp; l < levels; l += LK_WARPS) {
    const int xb = (int)floorf(s_pos[((l * 2 + 0) * WIN) * LK_PIX + lane]);
    const int yb = (int)floorf(s_pos[((l * 2 + 1) * WIN) * LK_PIX + lane]);
    bool ok = true;
#pragma unroll
    for (int o = 1; o < WIN; ++o) {
      ok = ok && ((int)floorf(s_pos[((l * 2 + 0) * WIN + o) * LK_PIX + lane]) == xb + o);
      ok = ok && ((int)floorf(s_pos[((l * 2 + 1) * WIN + o) * LK_PIX + lane]) == yb + o);
    }
    s_xb[l * LK_PIX + lane] = xb;
    s_yb[l * LK_PIX + lane] = yb;
    s_ok[l * LK_PIX + lane] = ok ? 1 : 0;
  }
  __syncthreads();

  // ---- phase 2: fetch the (W1 x W1) window of every (pixel, level): one row segment per thread iteration ----
  for (int seg = warp; seg < levels * W1; seg += LK_WARPS) {
    const int l = seg / W1, ry = seg - l * W1;
    const int W = s_lw[l], H = s_lh[l], off = s_lo[l];
    const int xb = s_xb[l * LK_PIX + lane], yb = s_yb[l * LK_PIX + lane];
    const int y = yb + ry;
    const bool row_ok = live && (unsigned)y < (unsigned)H && s_ok[l * LK_PIX + lane];
    const int row0 = off + y * W;
    const int a_start = row0 + xb;                      // absolute column of window column 0 (may be < 0)
    const int ca = a_start & ~(EPC - 1);                // chunk-aligned start (two's complement floor)
    const int lo = row0 + max(xb, 0), hi = row0 + min(xb + W1, W);  // valid absolute columns [lo, hi)
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const int col = ca + c * EPC;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && col < hi && col + EPC > lo)  // (implies 0 <= col < n_cols)
        v = __ldg(reinterpret_cast<const uint4*>(pyr + panel_index<T>(n_panels, prow, b, pix, col)));
      uint32_t* dst = s_raw + ((size_t)(seg * NW + c * 4) * LK_PIX + lane);
      dst[0] = v.x;
      dst[LK_PIX] = v.y;
      dst[2 * LK_PIX] = v.z;
      dst[3 * LK_PIX] = v.w;
    }
  }
  __syncthreads();

  // ---- phase 3: one (level, i) column of the output window per warp iteration, WIN outputs (j) each ----
  for (int u = warp; u < levels * WIN; u += LK_WARPS) {
    const int l = u / WIN, i = u - l * WIN;
    const int W = s_lw[l], off = s_lo[l];
    float* dst = out + ((size_t)b * n_ch + (size_t)l * WIN * WIN + (size_t)i * WIN) * nf + pix;
    if (s_ok[l * LK_PIX + lane]) {
      const int xb = s_xb[l * LK_PIX + lane], yb = s_yb[l * LK_PIX + lane];
      const float wx0 = s_w0[((l * 2 + 0) * WIN + i) * LK_PIX + lane], wx1 = s_w1[((l * 2 + 0) * WIN + i) * LK_PIX + lane];
      // the (i, i + 1) column pair of every window row; outputs j and j + 1 share a row
      float tap0[W1], tap1[W1];
#pragma unroll
      for (int r = 0; r < W1; ++r) {
        const int s = (off + (yb + r) * W + xb) & (EPC - 1);
        tap_pair<T>(s_raw + ((size_t)((l * W1 + r) * NW) * LK_PIX + lane), s + i, tap0[r], tap1[r]);
      }
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        const float wy0 = s_w0[((l * 2 + 1) * WIN + j) * LK_PIX + lane], wy1 = s_w1[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
        float val = tap0[j] * __fmul_rn(wx0, wy0);             // nw
        val += tap1[j] * __fmul_rn(wx1, wy0);                  // ne
        val += tap0[j + 1] * __fmul_rn(wx0, wy1);              // sw
        val += tap1[j + 1] * __fmul_rn(wx1, wy1);              // se
        if (live) dst[(size_t)j * nf] = val;
      }
    } else {
      const float ix = s_pos[((l * 2 + 0) * WIN + i) * LK_PIX + lane];
      for (int j = 0; j < WIN; ++j) {
        const float iy = s_pos[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
        const float val = live ? sample_slow<T>(pyr, n_panels, prow, b, pix, W, s_lh[l], off, ix, iy) : 0.f;
        if (live) dst[(size_t)j * nf] = val;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Specialisation for the SLIM configuration (radius 3: 7 x 7 window, 8 x 8 fetched): same phases as the generic
// kernel, but the fetched row segments are re-aligned to window column 0 while they are staged, so phase 3 works
// on registers with compile-time offsets, and the interpolation is separable (8 x 7 horizontal + 7 x 7 vertical
// blends per level instead of 49 x 4 taps).  Warp w serves level w / 2; its half w % 2 fetches window rows
// 4 (w % 2) .. +3 in phase 2 and produces window columns i in {0..3} or {4..6} in phase 3.
// ------------------------------------------------------------------------------------------
template <typename T>
struct V3 {
  static constexpr int R = 3, WIN = 7, W1 = 8;
  static constexpr int EPC = Elem<T>::EPC;
  static constexpr int NCHUNK = Cfg<T, 3>::NCHUNK;       // 2 (bf16) / 3 (fp32) 16-byte chunks per row segment
  static constexpr int WPR = W1 * (int)sizeof(T) / 4;    // 32-bit words per aligned window row: 4 / 8
  // shared-memory window: one extra row and one extra word per row (window element 8 [and 9]) for windows whose
  // per-offset floors are not all `origin + offset` (see the "shifted" mode in the kernel)
  static constexpr int SROWS = W1 + 1, SWPR = WPR + 1;
  __host__ __device__ static constexpr int win_words(int levels) { return levels * SROWS * SWPR * LK_PIX; }
  __host__ __device__ static constexpr int tab_words(int levels) { return levels * 2 * WIN * LK_PIX; }
  __host__ __device__ static constexpr int smem_bytes(int levels, bool nhwc) {
    return (win_words(levels) + 3 * tab_words(levels) + 3 * levels * LK_PIX + (nhwc ? LK_PIX * (levels * WIN * WIN + 1) : 0)) * 4;
  }
};

// window element c (0..7) of an aligned row held in registers
template <typename T, int C>
__device__ __forceinline__ float win_elem(const uint32_t* w);
template <>
__device__ __forceinline__ float win_elem<float, 0>(const uint32_t* w) { return __uint_as_float(w[0]); }
#define SLIMB200_WIN_ELEM_F32(C) \
  template <>                    \
  __device__ __forceinline__ float win_elem<float, C>(const uint32_t* w) { return __uint_as_float(w[C]); }
SLIMB200_WIN_ELEM_F32(1) SLIMB200_WIN_ELEM_F32(2) SLIMB200_WIN_ELEM_F32(3) SLIMB200_WIN_ELEM_F32(4)
SLIMB200_WIN_ELEM_F32(5) SLIMB200_WIN_ELEM_F32(6) SLIMB200_WIN_ELEM_F32(7)
#define SLIMB200_WIN_ELEM_BF16(C)                                                          \
  template <>                                                                              \
  __device__ __forceinline__ float win_elem<__nv_bfloat16, C>(const uint32_t* w) {         \
    return __uint_as_float((C & 1) ? (w[C >> 1] & 0xffff0000u) : (w[C >> 1] << 16));       \
  }
SLIMB200_WIN_ELEM_BF16(0) SLIMB200_WIN_ELEM_BF16(1) SLIMB200_WIN_ELEM_BF16(2) SLIMB200_WIN_ELEM_BF16(3)
SLIMB200_WIN_ELEM_BF16(4) SLIMB200_WIN_ELEM_BF16(5) SLIMB200_WIN_ELEM_BF16(6) SLIMB200_WIN_ELEM_BF16(7)

// shift `s` elements out of the NCHUNK * 4 loaded words so that window column 0 lands in word 0
template <typename T>
__device__ __forceinline__ void realign(const uint32_t* ld, int s, uint32_t* out);
template <>
__device__ __forceinline__ void realign<float>(const uint32_t* ld, int s, uint32_t* out) {  // 12 words in, s in 0..3, 9 out
  uint32_t t[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) t[k] = (s & 2) ? ld[k + 2] : ld[k];
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = (s & 1) ? t[k + 1] : t[k];
}
template <>
__device__ __forceinline__ void realign<__nv_bfloat16>(const uint32_t* ld, int s, uint32_t* out) {  // 8 words in, s in 0..7, 5 out
  const int ws = s >> 1;                                                                               // (elements 0..8 valid)
  uint32_t t[8], x[6];
#pragma unroll
  for (int k = 0; k < 7; ++k) t[k] = (ws & 1) ? ld[k + 1] : ld[k];
  t[7] = (ws & 1) ? 0u : ld[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) x[k] = (ws & 2) ? (k + 2 < 8 ? t[k + 2] : 0u) : t[k];
  const int sh = (s & 1) * 16;
#pragma unroll
  for (int k = 0; k < 5; ++k) out[k] = __funnelshift_r(x[k], x[k + 1], sh);
}

template <typename T, int I0, int I1>
__device__ __forceinline__ void v3_columns(const uint32_t (*win)[V3<T>::WPR], const float* wy0, const float* wy1,
                                           const float* s_w0x, const float* s_w1x, int lane, float* dst, size_t kstride,
                                           bool live) {
  constexpr int WIN = 7, W1 = 8;
  // I0..I1-1 are compile-time window columns
  auto one = [&](auto ic) {
    constexpr int I = decltype(ic)::value;
    const float wx0 = s_w0x[I * LK_PIX + lane], wx1 = s_w1x[I * LK_PIX + lane];
    float h[W1];
#pragma unroll
    for (int r = 0; r < W1; ++r) h[r] = fmaf(win_elem<T, I + 1>(win[r]), wx1, win_elem<T, I>(win[r]) * wx0);
    if (live) {
#pragma unroll
      for (int j = 0; j < WIN; ++j) dst[((size_t)I * WIN + j) * kstride] = fmaf(h[j + 1], wy1[j], h[j] * wy0[j]);
    }
  };
  if constexpr (I0 == 0) {
    one(std::integral_constant<int, 0>{});
    one(std::integral_constant<int, 1>{});
    one(std::integral_constant<int, 2>{});
    one(std::integral_constant<int, 3>{});
  } else {
    one(std::integral_constant<int, 4>{});
    one(std::integral_constant<int, 5>{});
    one(std::integral_constant<int, 6>{});
  }
}

// one window row of one pixel: NCHUNK 16-byte loads (predicated on the valid column range), returns the shift that
// re-aligns the row to window column 0
template <typename T>
__device__ __forceinline__ int v3_fetch_row(const T* __restrict__ pyr, int n_panels, int rows, int b, int pix, int off, int W, int H,
                                            int xb, int y, int n_elems, bool okp, uint32_t* ld) {
  constexpr int EPC = V3<T>::EPC, NCHUNK = V3<T>::NCHUNK;
  const bool row_ok = okp && (unsigned)y < (unsigned)H;
  const int row0 = off + y * W;
  const int a_start = row0 + xb;
  const int ca = a_start & ~(EPC - 1);
  const int lo = row0 + max(xb, 0), hi = row0 + min(xb + n_elems, W);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int col = ca + c * EPC;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok && col < hi && col + EPC > lo)  // read once per lookup, 878 MB per 8 samples: streaming (evict-first) loads
      v = __ldcs(reinterpret_cast<const uint4*>(pyr + panel_index<T>(n_panels, rows, b, pix, col)));
    ld[c * 4 + 0] = v.x;
    ld[c * 4 + 1] = v.y;
    ld[c * 4 + 2] = v.z;
    ld[c * 4 + 3] = v.w;
  }
  return a_start - ca;
}

// window element e (0..8) of row r from the shared-memory window (run-time indices: the "shifted" mode)
template <typename T>
__device__ __forceinline__ float v3_smem_elem(const uint32_t* s_win, int l, int r, int e, int lane);
template <>
__device__ __forceinline__ float v3_smem_elem<float>(const uint32_t* s_win, int l, int r, int e, int lane) {
  return __uint_as_float(s_win[((size_t)((l * V3<float>::SROWS + r) * V3<float>::SWPR + e)) * LK_PIX + lane]);
}
template <>
__device__ __forceinline__ float v3_smem_elem<__nv_bfloat16>(const uint32_t* s_win, int l, int r, int e, int lane) {
  using V = V3<__nv_bfloat16>;
  const uint32_t w = s_win[((size_t)((l * V::SROWS + r) * V::SWPR + (e >> 1))) * LK_PIX + lane];
  return __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
}

// NHWC: the output is channels-last (batch, h, w, channels): the CTA's 32 pixels x 196 channels form one contiguous
// block, staged through shared memory and written with fully coalesced rows.
//
// Window modes per (pixel, level), from the per-offset floors f_o of the sample positions (origin = min_o(f_o - o)):
//   1 "regular"  f_o == origin + o for all 7 offsets on both axes (any non-integer position): 8 x 8 window, registers
//   2 "shifted"  f_o - o - origin in {0, 1}: positions that sit on integers (the first GRU iteration: coords1 is the
//                pixel grid itself) come back from the normalise / un-normalise round trip a few ulp above OR below the
//                integer, offset by offset.  Same fetch plus window row / column 8, taps addressed per offset in smem
//   0 "slow"     anything else: predicated 4-tap loads from global memory
template <typename T, bool NHWC>
__global__ void __launch_bounds__(LK_THREADS, 3) k_corr_lookup_r3(const T* __restrict__ pyr, const slimb200_corr_layout L,
                                                                  const float* __restrict__ coords, float* __restrict__ out) {
  using V = V3<T>;
  constexpr int R = 3, WIN = 7, W1 = 8, NCHUNK = V::NCHUNK, WPR = V::WPR, SROWS = V::SROWS, SWPR = V::SWPR;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int levels = L.levels;
  uint32_t* s_win = s_dyn;                                                    // [level][row 0..8][word][lane], aligned rows
  float* s_pos = reinterpret_cast<float*>(s_dyn + V::win_words(levels));      // [level][axis][offset][lane]
  float* s_w0 = s_pos + V::tab_words(levels);
  float* s_w1 = s_w0 + V::tab_words(levels);
  int* s_xb = reinterpret_cast<int*>(s_w1 + V::tab_words(levels));
  int* s_yb = s_xb + levels * LK_PIX;
  int* s_ok = s_yb + levels * LK_PIX;                                         // window mode
  float* s_out = reinterpret_cast<float*>(s_ok + levels * LK_PIX);            // NHWC only: [pixel][n_ch + 1]
  __shared__ float s_xy[2][LK_PIX];
  __shared__ int s_lw[SLIMB200_MAX_LEVELS], s_lh[SLIMB200_MAX_LEVELS], s_lo[SLIMB200_MAX_LEVELS];

  const int nf = L.h * L.w;
  const int prow = L.rows_padded;
  const int n_panels = L.n_panels;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * LK_PIX;
  const int lane = lane_id(), warp = warp_id();
  const int pix = i0 + lane;
  const bool live = pix < nf;
  const int n_ch = levels * WIN * WIN;
  const int out_pad = n_ch + 1;

  if (threadIdx.x < 2 * LK_PIX) {
    const int ch = threadIdx.x / LK_PIX;
    s_xy[ch][lane] = live ? __ldg(coords + ((size_t)b * 2 + ch) * nf + pix) : 0.f;
  }
  if (threadIdx.x == 64) {
    s_lw[0] = L.level_w[0]; s_lw[1] = L.level_w[1]; s_lw[2] = L.level_w[2]; s_lw[3] = L.level_w[3];
    s_lh[0] = L.level_h[0]; s_lh[1] = L.level_h[1]; s_lh[2] = L.level_h[2]; s_lh[3] = L.level_h[3];
    s_lo[0] = L.level_offset[0]; s_lo[1] = L.level_offset[1]; s_lo[2] = L.level_offset[2]; s_lo[3] = L.level_offset[3];
  }
  __syncthreads();

  // ---- phase 1: positions and masked weights ----
  for (int e = warp; e < levels * 2 * WIN; e += LK_WARPS) {
    const int l = e / (2 * WIN), rem = e - l * 2 * WIN;
    const int axis = rem / WIN, o = rem - axis * WIN;
    const int size = axis == 0 ? s_lw[l] : s_lh[l];
    const float inv = 1.0f / (float)(1 << l);
    const float ip = sample_pos(s_xy[axis][lane], inv, o - R, size);
    const float f = floorf(ip);
    const int i0p = (int)f;
    const float w_hi = __fsub_rn(ip, f);
    const float w_lo = __fsub_rn(__fadd_rn(f, 1.f), ip);
    s_pos[e * LK_PIX + lane] = ip;
    s_w0[e * LK_PIX + lane] = ((unsigned)i0p < (unsigned)size) ? w_lo : 0.f;
    s_w1[e * LK_PIX + lane] = ((unsigned)(i0p + 1) < (unsigned)size) ? w_hi : 0.f;
  }
  __syncthreads();
  for (int l = warp; l < levels; l += LK_WARPS) {
    int fx[WIN], fy[WIN];
    int xb = 0x7fffffff, yb = 0x7fffffff;
#pragma unroll
    for (int o = 0; o < WIN; ++o) {
      fx[o] = (int)floorf(s_pos[((l * 2 + 0) * WIN + o) * LK_PIX + lane]) - o;
      fy[o] = (int)floorf(s_pos[((l * 2 + 1) * WIN + o) * LK_PIX + lane]) - o;
      xb = min(xb, fx[o]);
      yb = min(yb, fy[o]);
    }
    bool regular = true, shifted = true;
#pragma unroll
    for (int o = 0; o < WIN; ++o) {
      regular = regular && fx[o] == xb && fy[o] == yb;
      shifted = shifted && fx[o] - xb <= 1 && fy[o] - yb <= 1;
    }
    s_xb[l * LK_PIX + lane] = xb;
    s_yb[l * LK_PIX + lane] = yb;
    s_ok[l * LK_PIX + lane] = regular ? 1 : (shifted ? 2 : 0);
  }
  __syncthreads();

  // ---- phase 2: warp (level, half) fetches 4 window rows per pixel, re-aligned to window column 0 ----
  for (int u = warp; u < levels * 2; u += LK_WARPS) {
    const int l = u >> 1, half = u & 1;
    const int W = s_lw[l], H = s_lh[l], off = s_lo[l];
    const int xb = s_xb[l * LK_PIX + lane], yb = s_yb[l * LK_PIX + lane];
    const int mode = s_ok[l * LK_PIX + lane];
    const bool okp = live && mode != 0;
    const int n_elems = W1 + (mode == 2 ? 1 : 0);
    uint32_t ld[4][NCHUNK * 4];
    int sft[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
      sft[rr] = v3_fetch_row<T>(pyr, n_panels, prow, b, pix, off, W, H, xb, yb + half * 4 + rr, n_elems, okp, ld[rr]);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      uint32_t al[SWPR];
      realign<T>(ld[rr], sft[rr], al);
      uint32_t* dst = s_win + ((size_t)((l * SROWS + half * 4 + rr) * SWPR) * LK_PIX + lane);
#pragma unroll
      for (int k = 0; k < SWPR; ++k) dst[k * LK_PIX] = al[k];
    }
    if (half == 1 && __any_sync(0xffffffffu, mode == 2)) {  // window row 8 for the shifted windows of this warp
      uint32_t al[SWPR];
      const int sf = v3_fetch_row<T>(pyr, n_panels, prow, b, pix, off, W, H, xb, yb + 8, n_elems, live && mode == 2, ld[0]);
      realign<T>(ld[0], sf, al);
      uint32_t* dst = s_win + ((size_t)((l * SROWS + 8) * SWPR) * LK_PIX + lane);
#pragma unroll
      for (int k = 0; k < SWPR; ++k) dst[k * LK_PIX] = al[k];
    }
  }
  __syncthreads();

  // ---- phase 3: warp (level, half) blends window columns {0..3} / {4..6} ----
  for (int u = warp; u < levels * 2; u += LK_WARPS) {
    const int l = u >> 1, half = u & 1;
    float* dst = NHWC ? s_out + lane * out_pad + l * WIN * WIN : out + ((size_t)b * n_ch + (size_t)l * WIN * WIN) * nf + pix;
    const size_t kstride = NHWC ? 1 : (size_t)nf;
    const int mode = s_ok[l * LK_PIX + lane];
    // warp-uniform choice: as soon as one lane has a shifted window, every non-slow lane takes the smem-tap path
    // (a regular window is a shifted one with all shifts zero), so the two paths are never both executed
    const bool any_shifted = __any_sync(0xffffffffu, mode == 2);
    if (mode != 0 && !any_shifted) {
      uint32_t win[W1][WPR];
#pragma unroll
      for (int r = 0; r < W1; ++r)
#pragma unroll
        for (int k = 0; k < WPR; ++k) win[r][k] = s_win[((size_t)((l * SROWS + r) * SWPR + k)) * LK_PIX + lane];
      float wy0[WIN], wy1[WIN];
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        wy0[j] = s_w0[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
        wy1[j] = s_w1[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
      }
      const float* w0x = s_w0 + ((l * 2 + 0) * WIN) * LK_PIX;
      const float* w1x = s_w1 + ((l * 2 + 0) * WIN) * LK_PIX;
      if (half == 0)
        v3_columns<T, 0, 4>(win, wy0, wy1, w0x, w1x, lane, dst, kstride, live);
      else
        v3_columns<T, 4, 7>(win, wy0, wy1, w0x, w1x, lane, dst, kstride, live);
    } else if (mode != 0) {
      const int xb = s_xb[l * LK_PIX + lane], yb = s_yb[l * LK_PIX + lane];
      int ry[WIN];
      float wy0[WIN], wy1[WIN];
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        ry[j] = (int)floorf(s_pos[((l * 2 + 1) * WIN + j) * LK_PIX + lane]) - yb;  // j or j + 1
        wy0[j] = s_w0[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
        wy1[j] = s_w1[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
      }
      const int ib = half ? 4 : 0, ie = half ? WIN : 4;
      for (int i = ib; i < ie; ++i) {
        const int ex = (int)floorf(s_pos[((l * 2 + 0) * WIN + i) * LK_PIX + lane]) - xb;  // i or i + 1
        const float wx0 = s_w0[((l * 2 + 0) * WIN + i) * LK_PIX + lane], wx1 = s_w1[((l * 2 + 0) * WIN + i) * LK_PIX + lane];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
          // same association as the regular path: horizontal blends of the two rows, then the vertical blend
          const float h0 = fmaf(v3_smem_elem<T>(s_win, l, ry[j], ex + 1, lane), wx1, v3_smem_elem<T>(s_win, l, ry[j], ex, lane) * wx0);
          const float h1 = fmaf(v3_smem_elem<T>(s_win, l, ry[j] + 1, ex + 1, lane), wx1,
                                v3_smem_elem<T>(s_win, l, ry[j] + 1, ex, lane) * wx0);
          if (live) dst[((size_t)i * WIN + j) * kstride] = fmaf(h1, wy1[j], h0 * wy0[j]);
        }
      }
    } else {
      const int W = s_lw[l], H = s_lh[l], off = s_lo[l];
      const int ib = half ? 4 : 0, ie = half ? WIN : 4;
      for (int i = ib; i < ie; ++i) {
        const float ix = s_pos[((l * 2 + 0) * WIN + i) * LK_PIX + lane];
        for (int j = 0; j < WIN; ++j) {
          const float iy = s_pos[((l * 2 + 1) * WIN + j) * LK_PIX + lane];
          if (live) dst[((size_t)i * WIN + j) * kstride] = sample_slow<T>(pyr, n_panels, prow, b, pix, W, H, off, ix, iy);
        }
      }
    }
  }
  if (NHWC) {
    __syncthreads();
    const int n_pix = min(LK_PIX, nf - i0);
    float* blk = out + ((size_t)b * nf + i0) * n_ch;  // 32 pixels x n_ch floats, contiguous
    for (int pp = warp; pp < n_pix; pp += LK_WARPS)
      for (int k = lane; k < n_ch; k += 32) blk[(size_t)pp * n_ch + k] = s_out[pp * out_pad + k];
  }
}

template <typename T, bool NHWC>
int launch_lookup_r3(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, cudaStream_t stream) {
  using V = V3<T>;
  const int nf = L->h * L->w;
  const int smem = V::smem_bytes(L->levels, NHWC);
  SLIMB200_DEVICE(dev, n_sm);
  (void)n_sm;
  static bool attr_set[SLIMB200_MAX_DEVICES] = {false};
  if (!attr_set[dev]) {
    SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup_r3<T, NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           V::smem_bytes(SLIMB200_MAX_LEVELS, NHWC)));
    attr_set[dev] = true;
  }
  dim3 grid((nf + LK_PIX - 1) / LK_PIX, L->batch);
  SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                  (k_corr_lookup_r3<T, NHWC><<<grid, LK_THREADS, smem, stream>>>(static_cast<const T*>(pyramid), *L, coords, out)));
  return SLIMB200_OK;
}

template <typename T, int R>
int launch_lookup(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, cudaStream_t stream) {
  using C = Cfg<T, R>;
  const int nf = L->h * L->w;
  const int smem = C::smem_bytes(L->levels);
  SLIMB200_DEVICE(dev, n_sm);
  (void)n_sm;
  static bool attr_set[SLIMB200_MAX_DEVICES] = {false};
  if (!attr_set[dev]) {
    SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           C::smem_bytes(SLIMB200_MAX_LEVELS)));
    attr_set[dev] = true;
  }
  dim3 grid((nf + LK_PIX - 1) / LK_PIX, L->batch);
  SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                  (k_corr_lookup<T, R><<<grid, LK_THREADS, smem, stream>>>(static_cast<const T*>(pyramid), *L, coords, out)));
  return SLIMB200_OK;
}

template <typename T>
int dispatch_radius(int radius, int out_layout, const void* pyramid, const slimb200_corr_layout* L, const float* coords,
                    float* out, cudaStream_t stream) {
  if (out_layout == SLIMB200_CANVAS_NHWC) {
    if (radius != 3) return SLIMB200_E_UNSUPPORTED;  // channels-last output exists for the SLIM configuration only
    return launch_lookup_r3<T, true>(pyramid, L, coords, out, stream);
  }
  switch (radius) {
    case 0: return launch_lookup<T, 0>(pyramid, L, coords, out, stream);
    case 1: return launch_lookup<T, 1>(pyramid, L, coords, out, stream);
    case 2: return launch_lookup<T, 2>(pyramid, L, coords, out, stream);
    case 3: return launch_lookup_r3<T, false>(pyramid, L, coords, out, stream);
    case 4: return launch_lookup<T, 4>(pyramid, L, coords, out, stream);
    default: return SLIMB200_E_UNSUPPORTED;
  }
}

int g_lookup_generation = 1;

}  // namespace

// csrc/corr_lookup2.cu, csrc/corr_lookup3.cu
int slimb200_lookup_v2_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, int out_layout,
                              cudaStream_t stream);
int slimb200_lookup_v3_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, int out_layout,
                              cudaStream_t stream);
int slimb200_lookup_probe_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, cudaStream_t stream);

extern "C" int slimb200_lookup_generation(int32_t generation) {
  const int prev = g_lookup_generation;
  if (generation >= 0) g_lookup_generation = generation;
  return prev;
}

extern "C" int slimb200_corr_lookup(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                                    const float* coords, int32_t radius, float* out, int32_t out_layout, void* stream_) {
  if (out_layout != SLIMB200_CANVAS_NCHW && out_layout != SLIMB200_CANVAS_NHWC) return SLIMB200_E_INVALID;
  if (!pyramid || !L || !coords || !out) return SLIMB200_E_INVALID;
  if (radius < 0 || radius > 4 || L->levels < 1 || L->levels > SLIMB200_MAX_LEVELS) return SLIMB200_E_UNSUPPORTED;
  if (L->n_panels * PW != L->pitch || L->n_panels < 1) return SLIMB200_E_INVALID;
  if (L->rows_padded < L->h * L->w || (L->rows_padded & 127)) return SLIMB200_E_INVALID;
  if (reinterpret_cast<uintptr_t>(pyramid) & 15) return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16 && radius == 3 && g_lookup_generation == 9)  // memory-side probe (kbench only)
    return slimb200_lookup_probe_launch(pyramid, L, coords, out, stream);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16 && radius == 3 && g_lookup_generation >= 2)
    return slimb200_lookup_v3_launch(pyramid, L, coords, out, out_layout, stream);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16 && radius == 3 && g_lookup_generation == 1)
    return slimb200_lookup_v2_launch(pyramid, L, coords, out, out_layout, stream);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16) return dispatch_radius<__nv_bfloat16>(radius, out_layout, pyramid, L, coords, out, stream);
  if (pyramid_dtype == SLIMB200_DTYPE_F32) return dispatch_radius<float>(radius, out_layout, pyramid, L, coords, out, stream);
  return SLIMB200_E_UNSUPPORTED;
}
