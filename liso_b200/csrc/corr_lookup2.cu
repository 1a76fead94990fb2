// Stage 2b of the SLIM hot path on B200, second generation: the radius-3 lookup into the bf16 correlation pyramid, alone
// or FUSED with the 1x1 convolution that consumes it.
//
// Replaces CorrBlock.__call__ (liso/slim/model/raft_code/corr.py:23-46) + bilinear_sampler (raft_code/utils.py:15-29)
// and, in the fused entry, SmallMotionEncoder.conv_stat_corr1 + ReLU (liso/slim/model/update.py:49,71):
//   c = relu(conv1x1(lookup(coords), W (N, L*49), bias))
// so that the (B, 196, h, w) lookup tensor (40 MB per call at KITTI size, batch 8) never reaches HBM.
//
// Gather core (both kernels): ONE THREAD PER (source pixel, level), lane = pixel.  Everything lives in registers:
//   1. the 7 + 7 sample positions with the reference's normalise / un-normalise fp32 round trip (IEEE division), their
//      floors and zero-padding-masked bilinear weights
//   2. the 8 x 8 window: 16 independent 16-byte streaming loads (two per window row; in the pyramid layout of
//      include/slimb200.h four neighbouring pixels x 8 columns share a 64-byte unit, so lanes 4k..4k+3 share DRAM bursts),
//      each row re-aligned to window column 0 with two select stages + a funnel shift
//   3. separable blend, window column by column: 8 horizontal blends, 7 vertical ones, results handed to a sink
// Windows whose per-offset floors scatter by one around an integer position (every pixel of the first GRU iteration)
// take a 3-tap variant of the same code on a 9 x 9 window: taps (o, o + 1, o + 2) with weights (w0, w1, 0) or
// (0, w0, w1) -- the zero weight adds an exact zero, so the result equals the 2-tap blend at the shifted taps.
// Anything else (non-finite / absurd coordinates) takes predicated 4-tap loads.
//
// Sinks: NCHW (coalesced 128-byte rows straight from registers), channels-last (staged per tile in shared memory), or
// the A operand of a tcgen05 MMA:
//
// k_lookup_conv_tf32: persistent CTAs, 512 threads = 128 pixels x 4 levels per tile.  The 196 window values of a pixel
// (rounded to tf32) are written as row `pixel` of a K-major, 128-byte-swizzled A tile in shared memory (K = 4 levels x
// 56: 49 values + 7 zero pads, so every thread owns 14 whole 16-byte chunks and the stores are conflict-free); the
// weights (N x 224, tf32, packed once per weight tensor by k_lookup_conv_pack) arrive in the same layout with one bulk
// copy and stay for the whole kernel.  One elected thread issues 28 tcgen05.mma.kind::tf32 (M = 128, N, K = 8) into one
// of two TMEM accumulators; four warps read the PREVIOUS tile's accumulator back (tcgen05.ld), add the bias, apply the
// ReLU and store the rows, while the tensor core works and the other warps already compute the next tile's positions.
#include <cuda_bf16.h>

#include <climits>

#include "common.cuh"
#include "ptx.cuh"

namespace {

using namespace slimb200_ptx;

constexpr int R = 3, WIN = 7;
constexpr int KPL = 56;  // K slots per level in the fused A operand: 49 window values + 7 zero pads = 14 chunks of 4
constexpr unsigned FULL = 0xffffffffu;

struct LookupGeo {
  int nf, n_panels, pitch, m_tiles, levels, batch;
  int lw[SLIMB200_MAX_LEVELS], lh[SLIMB200_MAX_LEVELS], lo[SLIMB200_MAX_LEVELS];
};

__device__ __forceinline__ int pick4(const int (&a)[SLIMB200_MAX_LEVELS], int l) {
  return l == 0 ? a[0] : (l == 1 ? a[1] : (l == 2 ? a[2] : a[3]));  // (no dynamic indexing of kernel parameters)
}

// sample position in level pixels: bilinear_sampler's normalisation (utils.py:19-20) followed by grid_sample's
// un-normalisation ((g + 1) / 2) * (size - 1), all in fp32 with IEEE division
//
// The division runs without the generic IEEE sequence: the divisor size - 1 is a per-level constant, so its correctly
// rounded reciprocal `rinv` = __frcp_rn(size - 1) is computed once and the quotient is q0 = a * rinv refined by two
// residual steps r = fma(-b, q, a), q += r * rinv -- the correctly rounded a / b (Markstein; the same steps the
// hardware sequence takes after its reciprocal refinement; tests/test_host_logic.py checks the sequence in exact
// arithmetic).  Non-finite or zero divisors fall out as non-finite positions, i.e. "every tap outside", like before.
__device__ __forceinline__ float div_by_const(float a, float b, float rinv) {
  float q = __fmul_rn(a, rinv);
  q = __fmaf_rn(__fmaf_rn(-b, q, a), rinv, q);
  q = __fmaf_rn(__fmaf_rn(-b, q, a), rinv, q);
  return q;
}

__device__ __forceinline__ float sample_pos2(float c, float inv, int offs, float sm1, float rinv) {
  const float pos = __fadd_rn(c * inv, (float)offs);
  const float g = __fsub_rn(div_by_const(__fmul_rn(2.f, pos), sm1, rinv), 1.f);
  float ip = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), sm1);  // x / 2 == x * 0.5 exactly
  if (!(fabsf(ip) < 1e7f)) ip = -1e7f;                             // NaN / inf / far away: every tap is outside
  return ip;
}

// element offset of pyramid column `col` relative to the (sample, source pixel) base of the thread (include/slimb200.h)
__device__ __forceinline__ int col_offset(int col, int panel_stride) {
  return (col >> 7) * panel_stride + ((col >> 6) & 1) * 8192 + ((col >> 3) & 7) * 32 + (col & 7);
}

// masked weights, window origin and per-offset shift bits of one axis
__device__ __forceinline__ void axis_taps(float c, float inv, int size, float (&w0)[WIN], float (&w1)[WIN], int& origin,
                                          unsigned& shift_bits, bool& ok) {
  int f[WIN];
  origin = INT_MAX;
  const float sm1 = (float)(size - 1);
  const float rinv = __frcp_rn(sm1);
#pragma unroll
  for (int o = 0; o < WIN; ++o) {
    const float ip = sample_pos2(c, inv, o - R, sm1, rinv);
    const float fl = floorf(ip);
    const int i0 = (int)fl;
    const float w_hi = __fsub_rn(ip, fl);                   // weight of tap floor + 1  (ix - ix_nw)
    const float w_lo = __fsub_rn(__fadd_rn(fl, 1.f), ip);   // weight of tap floor      (ix_se - ix)
    w0[o] = ((unsigned)i0 < (unsigned)size) ? w_lo : 0.f;
    w1[o] = ((unsigned)(i0 + 1) < (unsigned)size) ? w_hi : 0.f;
    f[o] = i0 - o;
    origin = min(origin, f[o]);
  }
  shift_bits = 0u;
  ok = true;
#pragma unroll
  for (int o = 0; o < WIN; ++o) {
    const int d = f[o] - origin;
    ok = ok && d <= 1;
    shift_bits |= (unsigned)(d & 1) << o;
  }
}

// one window row: two 16-byte streaming loads (the volume is read once per lookup).  The loads are unconditional: the
// column is clamped into the padded pitch of the pixel's own rows, and whatever (finite) value is fetched from outside
// the level only ever meets a zero weight.  Returns the element shift of window column 0.
__device__ __forceinline__ int fetch_row(const __nv_bfloat16* __restrict__ base, int panel_stride, int pitch, int a_start,
                                         uint32_t (&raw)[8]) {
  const int ca = a_start & ~7;  // (two's complement floor)
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = min(max(ca + 8 * c, 0), pitch - 8);
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(base + col_offset(col, panel_stride)));
    raw[c * 4 + 0] = v.x;
    raw[c * 4 + 1] = v.y;
    raw[c * 4 + 2] = v.z;
    raw[c * 4 + 3] = v.w;
  }
  return a_start - ca;
}

// shift `s` (0..7) bf16 elements out of the 8 loaded words: window column 0 lands in the low half of out[0]
template <int NW>
__device__ __forceinline__ void realign(const uint32_t (&raw)[8], int s, uint32_t (&out)[NW]) {
  const int ws = s >> 1;
  uint32_t t[8], x[6];
#pragma unroll
  for (int k = 0; k < 7; ++k) t[k] = (ws & 1) ? raw[k + 1] : raw[k];
  t[7] = (ws & 1) ? 0u : raw[7];
#pragma unroll
  for (int k = 0; k < NW + 1; ++k) x[k] = (ws & 2) ? t[k + 2] : t[k];
  const int sh = (s & 1) * 16;
#pragma unroll
  for (int k = 0; k < NW; ++k) out[k] = __funnelshift_r(x[k], x[k + 1], sh);
}

template <int NW>
__device__ __forceinline__ float wel(const uint32_t (&w)[NW], int c) {  // window element c of an aligned row
  return __uint_as_float((c & 1) ? (w[c >> 1] & 0xffff0000u) : (w[c >> 1] << 16));
}

// fully predicated 4-tap sample straight from global memory (rare path)
__device__ __noinline__ float sample_slow2(const __nv_bfloat16* __restrict__ base, int panel_stride, int W, int H, int off, float ix,
                                           float iy) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  const float dx = __fsub_rn(ix, fx), dy = __fsub_rn(iy, fy);
  const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
  auto ld = [&](int y, int x) { return __bfloat162float(__ldg(base + col_offset(off + y * W + x, panel_stride))); };
  // horizontal blends first, like the fast paths
  float h0 = 0.f, h1 = 0.f;
  if (yin0) h0 = fmaf(xin1 ? ld(y0, x0 + 1) : 0.f, xin1 ? dx : 0.f, (xin0 ? ld(y0, x0) : 0.f) * (xin0 ? ex : 0.f));
  if (yin1) h1 = fmaf(xin1 ? ld(y0 + 1, x0 + 1) : 0.f, xin1 ? dx : 0.f, (xin0 ? ld(y0 + 1, x0) : 0.f) * (xin0 ? ex : 0.f));
  return fmaf(h1, yin1 ? dy : 0.f, h0 * (yin0 ? ey : 0.f));
}

// The 49 window values of one (pixel, level), handed to `sink.emit(k, value)` with k = i * 7 + j in ascending order
// (i offsets x, j offsets y: the reference's transposed window, corr.py:29-41).
template <class Sink>
__device__ __forceinline__ void lookup_pixel_level(const __nv_bfloat16* __restrict__ base, int panel_stride, int pitch, int W, int H,
                                                   int off, float cx, float cy, float inv, bool live, Sink& sink) {
  float wx0[WIN], wx1[WIN], wy0[WIN], wy1[WIN];
  int xb, yb;
  unsigned sx, sy;
  bool okx, oky;
  axis_taps(cx, inv, W, wx0, wx1, xb, sx, okx);
  axis_taps(cy, inv, H, wy0, wy1, yb, sy, oky);
  const int mode = (okx && oky) ? ((sx | sy) ? 2 : 1) : 0;
  const bool any_slow = __any_sync(FULL, live && mode == 0);
  const bool any_shift = __any_sync(FULL, live && mode == 2);

  if (!any_slow && !any_shift) {
    // ---- regular windows: 8 x 8, two taps per axis ----
    uint32_t win[8][4];
    {
      uint32_t raw[8][8];
      int sft[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) sft[r] = fetch_row(base, panel_stride, pitch, off + (yb + r) * W + xb, raw[r]);
#pragma unroll
      for (int r = 0; r < 8; ++r) realign<4>(raw[r], sft[r], win[r]);
    }
    sink.begin();
    float e0[8], e1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) e0[r] = wel<4>(win[r], 0);
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      float h[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        e1[r] = wel<4>(win[r], i + 1);
        h[r] = fmaf(e1[r], wx1[i], e0[r] * wx0[i]);
        e0[r] = e1[r];
      }
#pragma unroll
      for (int j = 0; j < WIN; ++j) sink.emit(i * WIN + j, fmaf(h[j + 1], wy1[j], h[j] * wy0[j]));
    }
    return;
  }
  if (mode != 0 || !live) {
    // ---- shifted windows: 9 x 9, three taps per axis, one weight of the three is zero ----
    uint32_t win[9][5];
    {
      uint32_t raw[9][8];
      int sft[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) sft[r] = fetch_row(base, panel_stride, pitch, off + (yb + r) * W + xb, raw[r]);
#pragma unroll
      for (int r = 0; r < 9; ++r) realign<5>(raw[r], sft[r], win[r]);
    }
    sink.begin();
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const bool s = (sx >> i) & 1u;
      const float a = s ? 0.f : wx0[i], bq = s ? wx0[i] : wx1[i], c = s ? wx1[i] : 0.f;
      float h[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) h[r] = fmaf(wel<5>(win[r], i + 2), c, fmaf(wel<5>(win[r], i + 1), bq, wel<5>(win[r], i) * a));
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        const bool t = (sy >> j) & 1u;
        const float ay = t ? 0.f : wy0[j], by = t ? wy0[j] : wy1[j], cyw = t ? wy1[j] : 0.f;
        sink.emit(i * WIN + j, fmaf(h[j + 2], cyw, fmaf(h[j + 1], by, h[j] * ay)));
      }
    }
    return;
  }
  // ---- anything else ----
  const float swm1 = (float)(W - 1), shm1 = (float)(H - 1);
  const float rw = __frcp_rn(swm1), rh = __frcp_rn(shm1);
  sink.begin();
#pragma unroll 1
  for (int i = 0; i < WIN; ++i) {
    const float ix = sample_pos2(cx, inv, i - R, swm1, rw);
    float v[WIN];
#pragma unroll 1
    for (int j = 0; j < WIN; ++j) v[j] = sample_slow2(base, panel_stride, W, H, off, ix, sample_pos2(cy, inv, j - R, shm1, rh));
    // (the sink wants compile-time k: a switch over the column)
    switch (i) {
#define SLIMB200_SLOW_COL(I)                                    \
  case I:                                                       \
    _Pragma("unroll") for (int j = 0; j < WIN; ++j) sink.emit(I * WIN + j, v[j]); \
    break;
      SLIMB200_SLOW_COL(0) SLIMB200_SLOW_COL(1) SLIMB200_SLOW_COL(2) SLIMB200_SLOW_COL(3) SLIMB200_SLOW_COL(4)
      SLIMB200_SLOW_COL(5) SLIMB200_SLOW_COL(6)
#undef SLIMB200_SLOW_COL
    }
  }
}

__device__ __forceinline__ size_t pixel_base(const LookupGeo& G, int b, int pix) {
  return ((size_t)b * G.n_panels * G.m_tiles + (size_t)(pix >> 7)) * 2 * 8192 + (size_t)(((pix & 127) >> 2) * 256 + (pix & 3) * 8);
}

// ------------------------------------------------------------------------------------------ stand-alone lookup
struct SinkGlobalStrided {  // NCHW: channel k of this (pixel, level) at p + k * stride; a warp stores 128-byte rows
  float* p;                 // (k arrives in ascending order: a running pointer instead of 49 64-bit multiplies)
  size_t stride;
  bool live;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void emit(int, float v) {
    if (live) *p = v;
    p += stride;
  }
};
struct SinkSmem {  // channels-last staging: p[k]
  float* p;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void emit(int k, float v) { p[k] = v; }
};

constexpr int V2_PIX = 64;
constexpr int V2_THREADS = V2_PIX * SLIMB200_MAX_LEVELS;  // 256
constexpr int V2_PITCH = SLIMB200_MAX_LEVELS * WIN * WIN + 1;  // odd pitch: conflict-free scalar stores

template <bool NHWC>
__global__ void __launch_bounds__(V2_THREADS, 2) k_corr_lookup_v2(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G,
                                                                  const float* __restrict__ coords, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_stage[];  // NHWC only: [pixel][V2_PITCH]
  const int lane = lane_id(), warp = warp_id();
  const int level = warp >> 1;
  const int prow = (warp & 1) * 32 + lane;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * V2_PIX;
  const int pix = i0 + prow;
  const int n_ch = G.levels * WIN * WIN;
  if (level < G.levels) {
    const bool live = pix < G.nf;
    const int W = pick4(G.lw, level), H = pick4(G.lh, level), off = pick4(G.lo, level);
    const float cx = live ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
    const float cy = live ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
    const float inv = 1.0f / (float)(1 << level);  // coords / 2**l is exact
    const __nv_bfloat16* base = pyr + pixel_base(G, b, live ? pix : 0);
    const int panel_stride = G.m_tiles * 2 * 8192;
    if (NHWC) {
      SinkSmem sink{s_stage + prow * V2_PITCH + level * WIN * WIN};
      lookup_pixel_level(base, panel_stride, G.pitch, W, H, off, cx, cy, inv, live, sink);
    } else {
      SinkGlobalStrided sink{out + ((size_t)b * n_ch + (size_t)level * WIN * WIN) * G.nf + pix, (size_t)G.nf, live};
      lookup_pixel_level(base, panel_stride, G.pitch, W, H, off, cx, cy, inv, live, sink);
    }
  }
  if (NHWC) {
    __syncthreads();
    // the tile is one contiguous block of the channels-last tensor: n_pix * n_ch floats
    const int n_pix = min(V2_PIX, G.nf - i0);
    float* blk = out + ((size_t)b * G.nf + i0) * n_ch;
    for (int pp = warp; pp < n_pix; pp += V2_THREADS / 32)
      for (int k = lane; k < n_ch; k += 32) blk[(size_t)pp * n_ch + k] = s_stage[pp * V2_PITCH + k];
  }
}

// ------------------------------------------------------------------------------------------ fused lookup + 1x1 conv
constexpr int F_PIX = 128;                  // pixels per tile = MMA M
constexpr int F_THREADS = F_PIX * 4;        // one thread per (pixel, level)
constexpr int F_LEVELS = 4;
constexpr int F_K = F_LEVELS * KPL;         // 224
constexpr int F_KBLK = 32;                  // tf32 elements per 128-byte swizzle row
constexpr int F_KBLOCKS = F_K / F_KBLK;     // 7
constexpr int F_UMMA_K = 8;                 // tf32: 32 bytes of K per instruction
constexpr uint32_t F_A_KBLK_BYTES = F_PIX * 128;            // 16 KB
constexpr uint32_t F_A_BYTES = F_KBLOCKS * F_A_KBLK_BYTES;  // 112 KB
constexpr int F_MAX_N = 128;
constexpr uint32_t F_ACC_COLS = 128;        // TMEM columns per accumulator (N <= 128 fp32 columns)
constexpr uint32_t F_TMEM_COLS = 2 * F_ACC_COLS;  // two accumulators: the MMAs of tile t overlap the epilogue of tile t - 1
// packed weights (slimb200_corr_lookup_conv_pack): the B operand exactly as it sits in shared memory -- 7 K blocks of
// N rows x 128 bytes (32 tf32 values, 128-byte swizzle) -- followed by the N fp32 biases
__host__ __device__ constexpr uint32_t packed_w_bytes(int n) { return (uint32_t)F_KBLOCKS * (uint32_t)n * 128u; }
__host__ __device__ constexpr uint32_t packed_bytes(int n) { return packed_w_bytes(n) + (uint32_t)n * 4u; }
__host__ __device__ constexpr uint32_t fused_smem_bytes(int n) {
  return F_A_BYTES + packed_bytes(n) + 64u /*barriers + tmem ptr*/ + 1024u /*align*/;
}
static_assert(fused_smem_bytes(F_MAX_N) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// slot k' = l * 56 + c of row n holds W[n][l * 49 + c] (c < 49) or 0, rounded to tf32
__global__ void __launch_bounds__(256) k_lookup_conv_pack(const float* __restrict__ weight, const float* __restrict__ bias, int N,
                                                          uint8_t* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < N * F_K) {
    const int n = idx / F_K, kk = idx - n * F_K;
    const int l = kk / KPL, c = kk - l * KPL;
    const float v = c < WIN * WIN ? __ldg(weight + (size_t)n * (F_LEVELS * WIN * WIN) + l * WIN * WIN + c) : 0.f;
    const uint32_t off = (uint32_t)(kk >> 5) * ((uint32_t)N * 128u) + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                         (((uint32_t)((kk & 31) >> 2) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(kk & 3) * 4u;
    *reinterpret_cast<uint32_t*>(packed + off) = to_tf32(v);
  }
  if (idx < N) reinterpret_cast<float*>(packed + packed_w_bytes(N))[idx] = bias ? __ldg(bias + idx) : 0.f;
}

// A-operand sink: 14 chunks of 4 tf32 values per (pixel, level) into the 128-byte-swizzled K-major tile
struct SinkA {
  uint32_t a_row;     // smem address of this pixel's row inside K block 0 (row / 8 * 1024 + row % 8 * 128)
  uint32_t swz;       // row % 8
  int q0;             // first 16-byte chunk of this level: level * 14
  uint32_t wait_bar;  // mbarrier of the MMAs that still read the A tile (0: none)
  uint32_t wait_parity;
  uint32_t pend[4];
  __device__ __forceinline__ void begin() {  // right before the first store: the previous tile's MMAs must be through
    if (wait_bar) mbar_wait(wait_bar, wait_parity);
  }
  __device__ __forceinline__ void flush(int c, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    const uint32_t q = (uint32_t)(q0 + c);
    const uint32_t addr = a_row + (q >> 3) * F_A_KBLK_BYTES + (((q & 7u) ^ swz) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
  }
  __device__ __forceinline__ void emit(int k, float v) {
    pend[k & 3] = to_tf32(v);
    if ((k & 3) == 3) flush(k >> 2, pend[0], pend[1], pend[2], pend[3]);
    if (k == WIN * WIN - 1) {  // k = 48 is element 0 of chunk 12; chunk 13 is padding
      flush(12, pend[0], 0u, 0u, 0u);
      flush(13, 0u, 0u, 0u, 0u);
    }
  }
};

// accumulator rows -> + bias -> ReLU -> global, one thread per pixel row (warps 0..3 = TMEM lane quarters 0..3)
__device__ __forceinline__ void fused_epilogue(uint32_t tmem_acc, int warp, int lane, const float* s_bias, int N, int relu,
                                               float* __restrict__ tile_out, int out_pitch, int rows_live) {
  const int row = warp * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16);
  float* const dst = tile_out + (size_t)row * out_pitch;
  for (int cb = 0; cb < (N >> 5); ++cb) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + (uint32_t)(cb * 32), v);
    tmem_ld_wait();
    if (row < rows_live) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 o;
        o.x = __uint_as_float(v[c * 4 + 0]) + s_bias[cb * 32 + c * 4 + 0];
        o.y = __uint_as_float(v[c * 4 + 1]) + s_bias[cb * 32 + c * 4 + 1];
        o.z = __uint_as_float(v[c * 4 + 2]) + s_bias[cb * 32 + c * 4 + 2];
        o.w = __uint_as_float(v[c * 4 + 3]) + s_bias[cb * 32 + c * 4 + 3];
        if (relu) {
          o.x = fmaxf(o.x, 0.f);
          o.y = fmaxf(o.y, 0.f);
          o.z = fmaxf(o.z, 0.f);
          o.w = fmaxf(o.w, 0.f);
        }
        *reinterpret_cast<float4*>(dst + cb * 32 + c * 4) = o;
      }
    }
  }
}

// Per tile: every thread gathers + blends its (pixel, level) and writes 14 chunks of the A tile; ONE __syncthreads;
// one thread issues the 28 MMAs of the tile into accumulator (tile & 1); warps 0..3 then write out the PREVIOUS tile
// (its MMAs finished long ago) while the tensor core works, and everybody moves on to the positions / loads of the next
// tile -- whose first A store waits for this tile's MMAs (an mbarrier that has normally fired by then).
__global__ void __launch_bounds__(F_THREADS, 1)
k_lookup_conv_tf32(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G, const float* __restrict__ coords,
                   const uint8_t* __restrict__ packed, float* __restrict__ out, int out_pitch, int N, int relu, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_base + F_A_BYTES;
  const uint32_t w_kblk_bytes = (uint32_t)N * 128u;
  const float* const s_bias = reinterpret_cast<const float*>(gen_base + F_A_BYTES + packed_w_bytes(N));
  const uint32_t bar0 = smem_w + packed_bytes(N);  // 8-byte aligned (N % 32 == 0)
  auto mma_bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t w_bar = bar0 + 16u;
  const uint32_t tmem_ptr_smem = bar0 + 24u;

  const int lane = lane_id(), warp = warp_id();
  if (threadIdx.x == 0) {
    mbar_init(mma_bar(0), 1);
    mbar_init(mma_bar(1), 1);
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // weights + biases: one bulk copy of the packed image (L2 -> shared memory), under the first tile's gather
    mbar_expect_tx(w_bar, packed_bytes(N));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_w), "l"(packed),
                 "r"(packed_bytes(N)), "r"(w_bar)
                 : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "r"(F_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(F_PIX >> 4) << 24);

  const int level = warp >> 2;
  const int prow = (warp & 3) * 32 + lane;
  const int W = pick4(G.lw, level), H = pick4(G.lh, level), off = pick4(G.lo, level);
  const float inv = 1.0f / (float)(1 << level);
  const int panel_stride = G.m_tiles * 2 * 8192;
  const uint32_t a_row = smem_a + (uint32_t)(prow >> 3) * 1024u + (uint32_t)(prow & 7) * 128u;

  int it = 0;              // tiles this CTA has started
  int prev_b = 0, prev_mt = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / G.m_tiles, mt = tile - b * G.m_tiles;
    const int pix = mt * F_PIX + prow;
    const bool live = pix < G.nf;
    {
      const float cx = live ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
      const float cy = live ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
      SinkA sink{a_row, (uint32_t)(prow & 7), level * (KPL / 4), it > 0 ? mma_bar((it - 1) & 1) : 0u,
                 (uint32_t)((it - 1) >> 1) & 1u, {0u, 0u, 0u, 0u}};
      lookup_pixel_level(pyr + pixel_base(G, b, live ? pix : 0), panel_stride, G.pitch, W, H, off, cx, cy, inv, live, sink);
    }
    fence_proxy_async_smem();  // the A rows were written through the generic proxy, the MMA reads through the async one
    __syncthreads();
    if (warp == 0) {
      if (it == 0) mbar_wait(w_bar, 0);  // the packed weights have landed
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(it & 1) * F_ACC_COLS;
#pragma unroll
        for (int kb = 0; kb < F_KBLOCKS; ++kb) {
          const uint64_t adesc = make_smem_desc_sw128(smem_a + (uint32_t)kb * F_A_KBLK_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(smem_w + (uint32_t)kb * w_kblk_bytes);
#pragma unroll
          for (int k = 0; k < F_KBLK / F_UMMA_K; ++k)  // + k * 8 elements * 4 B = 32 B (>> 4 = 2) inside the swizzle row
            umma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(mma_bar(it & 1));
      }
      __syncwarp();
    }
    if (warp < 4 && it > 0) {
      // the previous tile: its MMAs were complete before this tile's A stores began
      mbar_wait(w_bar, 0);
      mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
      tcgen05_fence_after();
      fused_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * F_ACC_COLS, warp, lane, s_bias, N, relu,
                     out + ((size_t)prev_b * G.nf + (size_t)prev_mt * F_PIX) * (size_t)out_pitch, out_pitch,
                     min(F_PIX, G.nf - prev_mt * F_PIX));
      tcgen05_fence_before();  // (ordered before the MMAs of tile it + 1 by the next __syncthreads)
    }
    prev_b = b;
    prev_mt = mt;
  }
  if (warp < 4 && it > 0) {  // drain: the last tile
    mbar_wait(w_bar, 0);
    mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
    tcgen05_fence_after();
    fused_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * F_ACC_COLS, warp, lane, s_bias, N, relu,
                   out + ((size_t)prev_b * G.nf + (size_t)prev_mt * F_PIX) * (size_t)out_pitch, out_pitch,
                   min(F_PIX, G.nf - prev_mt * F_PIX));
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F_TMEM_COLS) : "memory");
  }
}

int make_geo(const slimb200_corr_layout* L, LookupGeo* G) {
  if (L->n_panels * SLIMB200_PANEL_COLS != L->pitch || L->n_panels < 1) return SLIMB200_E_INVALID;
  if (L->rows_padded < L->h * L->w || (L->rows_padded & 127)) return SLIMB200_E_INVALID;
  if (L->levels < 1 || L->levels > SLIMB200_MAX_LEVELS) return SLIMB200_E_UNSUPPORTED;
  G->nf = L->h * L->w;
  G->n_panels = L->n_panels;
  G->pitch = L->pitch;
  G->m_tiles = L->rows_padded >> 7;
  G->levels = L->levels;
  G->batch = L->batch;
  for (int l = 0; l < SLIMB200_MAX_LEVELS; ++l) {
    G->lw[l] = l < L->levels ? L->level_w[l] : 1;
    G->lh[l] = l < L->levels ? L->level_h[l] : 1;
    G->lo[l] = l < L->levels ? L->level_offset[l] : 0;
  }
  // 32-bit element offsets inside one sample's panels, and (row index * width) of far-away windows
  if ((long long)L->n_panels * G->m_tiles * 2 * 8192 > 0x7fffffffLL || L->w > 4096 || L->h > 4096) return SLIMB200_E_UNSUPPORTED;
  return SLIMB200_OK;
}

}  // namespace

// radius-3 lookup on a bf16 pyramid, gather core of this file (called by slimb200_corr_lookup)
int slimb200_lookup_v2_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, int out_layout,
                              cudaStream_t stream) {
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  dim3 grid((G.nf + V2_PIX - 1) / V2_PIX, L->batch);
  if (out_layout == SLIMB200_CANVAS_NHWC) {
    const int smem = V2_PIX * V2_PITCH * 4;
    static bool attr_set = false;
    if (!attr_set) {
      SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v2<true><<<grid, V2_THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  } else {
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v2<false><<<grid, V2_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  }
  return SLIMB200_OK;
}

extern "C" size_t slimb200_corr_lookup_conv_packed_bytes(int32_t c_out) {
  return (c_out < 32 || c_out > F_MAX_N || (c_out & 31)) ? 0 : packed_bytes(c_out);
}

extern "C" int slimb200_corr_lookup_conv_pack(const float* weight, const float* bias, int32_t levels, int32_t radius, int32_t c_out,
                                              void* packed, void* stream_) {
  if (!weight || !packed) return SLIMB200_E_INVALID;
  if (levels != F_LEVELS || radius != R || c_out < 32 || c_out > F_MAX_N || (c_out & 31)) return SLIMB200_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(packed) & 15) return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = c_out * F_K;
  SLIMB200_LAUNCH(SLIMB200_K_LOOKUP_CONV_PACK, stream,
                  (k_lookup_conv_pack<<<(n + 255) / 256, 256, 0, stream>>>(weight, bias, c_out, static_cast<uint8_t*>(packed))));
  return SLIMB200_OK;
}

extern "C" int slimb200_corr_lookup_conv(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                                         const float* coords, int32_t radius, const void* packed, int32_t c_out, int32_t relu,
                                         float* out, int32_t out_pitch, void* stream_) {
  if (!pyramid || !L || !coords || !packed || !out) return SLIMB200_E_INVALID;
  if (pyramid_dtype != SLIMB200_DTYPE_BF16 || radius != R || L->levels != F_LEVELS) return SLIMB200_E_UNSUPPORTED;
  if (c_out < 32 || c_out > F_MAX_N || (c_out & 31)) return SLIMB200_E_UNSUPPORTED;
  if (out_pitch < c_out || (out_pitch & 3)) return SLIMB200_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(pyramid) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(packed) & 15))
    return SLIMB200_E_ALIGNMENT;
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    SLIMB200_CUDA_TRY(cudaGetDevice(&dev));
    SLIMB200_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_lookup_conv_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)fused_smem_bytes(F_MAX_N)));
  }
  const int n_tiles = L->batch * G.m_tiles;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  SLIMB200_LAUNCH(SLIMB200_K_LOOKUP_CONV, stream,
                  (k_lookup_conv_tf32<<<grid, F_THREADS, fused_smem_bytes(c_out), stream>>>(
                      static_cast<const __nv_bfloat16*>(pyramid), G, coords, static_cast<const uint8_t*>(packed), out, out_pitch,
                      c_out, relu, n_tiles)));
  return SLIMB200_OK;
}
