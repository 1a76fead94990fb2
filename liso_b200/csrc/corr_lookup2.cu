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
#include "lookup_core.cuh"

namespace {

using namespace slimb200_lookup;

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
constexpr int V2_WPITCH = WIN * WIN;  // 49 floats per (pixel, level): odd pitch, conflict-free scalar stores

template <bool NHWC>
__global__ void __launch_bounds__(V2_THREADS, 2) k_corr_lookup_v2(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G,
                                                                  const float* __restrict__ coords, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_stage[];  // NHWC only: [warp][32 units][49]
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
      // channels-last: the 49 values of a (pixel, level) are 196 contiguous bytes of the pixel's row.  Every warp transposes
      // its 32 units through its own shared-memory patch (odd pitch: conflict-free) and writes them out itself -- no
      // CTA-wide barrier, so the warps of a CTA drift apart and one warp's loads overlap another's blend
      float* patch = s_stage + warp * (32 * V2_WPITCH);
      SinkSmem sink{patch + lane * V2_WPITCH};
      lookup_pixel_level(base, panel_stride, G.pitch, W, H, off, cx, cy, inv, live, sink);
      __syncwarp();
      const int p0 = i0 + (warp & 1) * 32;
      const int n_pix = min(32, G.nf - p0);
      float* dst = out + ((size_t)b * G.nf + p0) * n_ch + level * (WIN * WIN);
      for (int pp = 0; pp < n_pix; ++pp) {
        dst[(size_t)pp * n_ch + lane] = patch[pp * V2_WPITCH + lane];
        if (lane < WIN * WIN - 32) dst[(size_t)pp * n_ch + 32 + lane] = patch[pp * V2_WPITCH + 32 + lane];
      }
    } else {
      SinkGlobalStrided sink{out + ((size_t)b * n_ch + (size_t)level * WIN * WIN) * G.nf + pix, (size_t)G.nf, live};
      lookup_pixel_level(base, panel_stride, G.pitch, W, H, off, cx, cy, inv, live, sink);
    }
  }
}

// Measurement probe (tools/kbench.py, generation 9): the memory side of the lookup alone -- positions, the 16 window-row
// loads of every (pixel, level), and one word per unit written -- no re-alignment, no blend, no 196-channel store.
__global__ void __launch_bounds__(V2_THREADS, 2) k_lookup_probe(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G,
                                                               const float* __restrict__ coords, float* __restrict__ out) {
  const int lane = lane_id(), warp = warp_id();
  const int level = warp >> 1;
  const int pix = blockIdx.x * V2_PIX + (warp & 1) * 32 + lane;
  const int b = blockIdx.y;
  if (level >= G.levels) return;
  const bool live = pix < G.nf;
  const int W = pick4(G.lw, level), H = pick4(G.lh, level), off = pick4(G.lo, level);
  const float cx = live ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
  const float cy = live ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
  const float inv = 1.0f / (float)(1 << level);
  const __nv_bfloat16* base = pyr + pixel_base(G, b, live ? pix : 0);
  float wx0[WIN], wx1[WIN], wy0[WIN], wy1[WIN];
  int xb, yb;
  unsigned sx, sy;
  bool okx, oky;
  axis_taps(cx, inv, W, wx0, wx1, xb, sx, okx);
  axis_taps(cy, inv, H, wy0, wy1, yb, sy, oky);
  uint32_t acc = 0u;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t raw[8];
    fetch_row(base, G.m_tiles * 2 * 8192, G.pitch, off + (yb + r) * W + xb, raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= raw[k];
  }
  if (live) out[((size_t)b * G.nf + pix) * 4 + level] = __uint_as_float(acc & 0x3fffffffu) + wx0[0] + wy1[6];
}

}  // namespace

int slimb200_lookup_probe_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out,
                                 cudaStream_t stream) {
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  dim3 grid((G.nf + V2_PIX - 1) / V2_PIX, L->batch);
  SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                  (k_lookup_probe<<<grid, V2_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  return SLIMB200_OK;
}

// radius-3 lookup on a bf16 pyramid, gather core of this file (called by slimb200_corr_lookup)
int slimb200_lookup_v2_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, int out_layout,
                              cudaStream_t stream) {
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  dim3 grid((G.nf + V2_PIX - 1) / V2_PIX, L->batch);
  if (out_layout == SLIMB200_CANVAS_NHWC) {
    const int smem = (V2_THREADS / 32) * 32 * V2_WPITCH * 4;
    SLIMB200_DEVICE(dev, n_sm);
    (void)n_sm;
    static bool attr_set[SLIMB200_MAX_DEVICES] = {false};
    if (!attr_set[dev]) {
      SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set[dev] = true;
    }
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v2<true><<<grid, V2_THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  } else {
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v2<false><<<grid, V2_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  }
  return SLIMB200_OK;
}

