// Device helpers shared by the second- and third-generation lookup kernels (csrc/corr_lookup2.cu, corr_lookup3.cu):
// sample positions with the reference's fp32 round trip, pyramid addressing, window-row fetch and re-alignment.
#pragma once
#include <cuda_bf16.h>

#include <climits>

#include "common.cuh"

namespace slimb200_lookup {

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
static __device__ __noinline__ float sample_slow2(const __nv_bfloat16* __restrict__ base, int panel_stride, int W, int H, int off, float ix,
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

__device__ __forceinline__ size_t pixel_base(const LookupGeo& G, int b, int pix) {
  return ((size_t)b * G.n_panels * G.m_tiles + (size_t)(pix >> 7)) * 2 * 8192 + (size_t)(((pix & 127) >> 2) * 256 + (pix & 3) * 8);
}

inline int make_geo(const slimb200_corr_layout* L, LookupGeo* G) {
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


}  // namespace slimb200_lookup
