// Stage 2b of the SLIM hot path on B200: multi-level bilinear lookup in the correlation pyramid,
// re-run on every GRU iteration.
//
// Replaces CorrBlock.__call__ (liso/slim/model/raft_code/corr.py:23-46) and bilinear_sampler
// (liso/slim/model/raft_code/utils.py:15-29) = F.grid_sample(align_corners=True, zeros padding):
//   out[b, l*49 + i*7 + j, y0, x0] = bilinear(level_l[b, (y0,x0)], x = cx/2^l + (i-r), y = cy/2^l + (j-r))
// Note the RAFT transposition: the FIRST window axis offsets x (corr.py:31-41).
//
// One launch covers all levels and writes the (B, L*(2r+1)^2, h, w) fp32 tensor directly (the
// reference needs 4 x (CPU meshgrid + H2D + grid_sample) + cat + permute + contiguous).
// A CTA owns 32 consecutive source pixels; the window taps of one pixel sit in one pyramid row
// (17 KB at 80x80), results are transposed through shared memory so that every channel row is
// written as one full 128-byte line.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int LK_PIX = 32;
constexpr int LK_THREADS = 256;
constexpr int LK_SMEM_MAX = (256 * 33 + 32 * 4 * 2 * 9) * 4;
constexpr int LK_MAX_CH = 256;     // one thread per output channel: levels * (2r+1)^2 <= 256

template <typename T>
__device__ __forceinline__ float ld_val(const T* p);
template <>
__device__ __forceinline__ float ld_val<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float ld_val<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}

// One thread owns one output channel k = (level, i, j) and walks the CTA's 32 pixels, so the
// per-channel constants (level geometry, window offsets) stay in registers and the lanes of a warp
// read neighbouring taps of the SAME pyramid row.  The pixel-space sample positions (the
// normalise / un-normalise round trip of bilinear_sampler + grid_sample) are computed once per
// (pixel, level, axis, offset) into shared memory instead of once per output.
template <typename T>
__global__ void __launch_bounds__(LK_THREADS, 4) k_corr_lookup(const T* __restrict__ pyr, const slimb200_corr_layout L,
                                                               const float* __restrict__ coords, int radius,
                                                               float* __restrict__ out) {
  extern __shared__ float s_dyn[];  // [n_ch][LK_PIX + 1] results, then [LK_PIX][levels][2][win] positions
  __shared__ float s_xy[2][LK_PIX];
  __shared__ int s_lw[SLIMB200_MAX_LEVELS], s_lh[SLIMB200_MAX_LEVELS], s_lo[SLIMB200_MAX_LEVELS];
  const int nf = L.h * L.w;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * LK_PIX;
  const int win = 2 * radius + 1, win2 = win * win;
  const int n_ch = L.levels * win2;
  float* s_out = s_dyn;
  float* s_pos = s_dyn + n_ch * (LK_PIX + 1);
  if (threadIdx.x < 2 * LK_PIX) {
    const int ch = threadIdx.x / LK_PIX, p = threadIdx.x % LK_PIX;
    s_xy[ch][p] = (i0 + p < nf) ? __ldg(coords + ((size_t)b * 2 + ch) * nf + i0 + p) : 0.f;
  }
  if (threadIdx.x < SLIMB200_MAX_LEVELS) {
    s_lw[threadIdx.x] = L.level_w[threadIdx.x];
    s_lh[threadIdx.x] = L.level_h[threadIdx.x];
    s_lo[threadIdx.x] = L.level_offset[threadIdx.x];
  }
  __syncthreads();
  // ---- sample positions: ix = (((2*xs/(W-1) - 1) + 1) / 2) * (W-1), xs = x / 2^l + (i - r) ------------
  const int per_pix = L.levels * 2 * win;
  for (int e = threadIdx.x; e < LK_PIX * per_pix; e += LK_THREADS) {
    const int p = e / per_pix, rem = e - p * per_pix;
    const int l = rem / (2 * win), rem2 = rem - l * 2 * win;
    const int axis = rem2 / win, o = rem2 - axis * win;
    const float inv = 1.0f / (float)(1 << l);  // coords / 2**l is exact
    const float pos = __fadd_rn(s_xy[axis][p] * inv, (float)(o - radius));
    const float sm1 = (float)((axis == 0 ? s_lw[l] : s_lh[l]) - 1);
    const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, pos), sm1), 1.f);            // utils.py:19-20
    float ip = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), sm1);                    // grid_sample un-normalise
    if (!(fabsf(ip) < 1e7f)) ip = -1e7f;                                             // NaN / inf / far away: all taps outside
    s_pos[e] = ip;
  }
  __syncthreads();
  // ---- one output channel per thread, 32 pixels ---------------------------------------------------
  const int k = threadIdx.x;
  if (k < n_ch) {
    const int l = k / win2, rem = k - l * win2;
    const int wa = rem / win, wb = rem - wa * win;
    const int W = s_lw[l], H = s_lh[l];
    const T* base = pyr + ((size_t)b * nf + i0) * L.pitch + s_lo[l];
    const float* px = s_pos + (l * 2 + 0) * win + wa;
    const float* py = s_pos + (l * 2 + 1) * win + wb;
    const int n_pix = min(LK_PIX, nf - i0);
#pragma unroll 4
    for (int p = 0; p < n_pix; ++p) {
      const float ix = px[p * per_pix], iy = py[p * per_pix];
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);  // se - i
      const float dx = __fsub_rn(ix, fx), dy = __fsub_rn(iy, fy);                                  // i - nw
      const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
      const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
      const T* row = base + (size_t)p * L.pitch + y0 * W + x0;
      float val = 0.f;
      if (yin0 && xin0) val += ld_val(row) * __fmul_rn(ex, ey);          // nw
      if (yin0 && xin1) val += ld_val(row + 1) * __fmul_rn(dx, ey);      // ne
      if (yin1 && xin0) val += ld_val(row + W) * __fmul_rn(ex, dy);      // sw
      if (yin1 && xin1) val += ld_val(row + W + 1) * __fmul_rn(dx, dy);  // se
      s_out[k * (LK_PIX + 1) + p] = val;
    }
  }
  __syncthreads();
  const int lane = lane_id(), warp = warp_id();
  if (i0 + lane < nf) {
    for (int kk = warp; kk < n_ch; kk += LK_THREADS / 32)
      out[((size_t)b * n_ch + kk) * nf + i0 + lane] = s_out[kk * (LK_PIX + 1) + lane];
  }
}

}  // namespace

extern "C" int slimb200_corr_lookup(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                                    const float* coords, int32_t radius, float* out, void* stream_) {
  if (!pyramid || !L || !coords || !out) return SLIMB200_E_INVALID;
  if (radius < 0 || radius > 4 || L->levels < 1 || L->levels > SLIMB200_MAX_LEVELS) return SLIMB200_E_UNSUPPORTED;
  for (int l = 0; l < L->levels; ++l)
    if (L->level_h[l] < 2 || L->level_w[l] < 2) return SLIMB200_E_UNSUPPORTED;  // (size - 1) normalisation
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nf = L->h * L->w;
  const int n_ch = L->levels * (2 * radius + 1) * (2 * radius + 1);
  if (n_ch > LK_MAX_CH) return SLIMB200_E_UNSUPPORTED;
  const size_t smem = ((size_t)n_ch * (LK_PIX + 1) + (size_t)LK_PIX * L->levels * 2 * (2 * radius + 1)) * sizeof(float);
  dim3 grid((nf + LK_PIX - 1) / LK_PIX, L->batch);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16) {
    static bool attr_set = false;
    if (!attr_set) {
      SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             LK_SMEM_MAX));
      attr_set = true;
    }
    slimb200_prof_pre(SLIMB200_K_CORR_LOOKUP, stream);
    k_corr_lookup<__nv_bfloat16><<<grid, LK_THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), *L,
                                                                     coords, radius, out);
    slimb200_prof_post(SLIMB200_K_CORR_LOOKUP, stream);
  } else if (pyramid_dtype == SLIMB200_DTYPE_F32) {
    static bool attr_set = false;
    if (!attr_set) {
      SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             LK_SMEM_MAX));
      attr_set = true;
    }
    slimb200_prof_pre(SLIMB200_K_CORR_LOOKUP, stream);
    k_corr_lookup<float><<<grid, LK_THREADS, smem, stream>>>(static_cast<const float*>(pyramid), *L, coords, radius,
                                                             out);
    slimb200_prof_post(SLIMB200_K_CORR_LOOKUP, stream);
  } else {
    return SLIMB200_E_UNSUPPORTED;
  }
  SLIMB200_LAUNCH_CHECK();
  return SLIMB200_OK;
}

extern "C" const char* slimb200_strerror(int code) {
  switch (code) {
    case SLIMB200_OK: return "success";
    case SLIMB200_E_INVALID: return "slimb200: invalid argument";
    case SLIMB200_E_UNSUPPORTED: return "slimb200: unsupported shape or dtype";
    case SLIMB200_E_WORKSPACE: return "slimb200: workspace too small";
    case SLIMB200_E_ALIGNMENT: return "slimb200: misaligned pointer or pitch";
    case SLIMB200_E_DRIVER: return "slimb200: CUDA driver entry point unavailable";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "slimb200: unknown error";
  }
}

extern "C" int slimb200_version(void) { return SLIMB200_VERSION; }
