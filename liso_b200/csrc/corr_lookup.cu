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
constexpr int LK_MAX_CH = 4 * 81;  // up to 4 levels, radius <= 4

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

template <typename T>
__global__ void __launch_bounds__(LK_THREADS) k_corr_lookup(const T* __restrict__ pyr, const slimb200_corr_layout L,
                                                            const float* __restrict__ coords, int radius,
                                                            float* __restrict__ out) {
  extern __shared__ float s_out[];  // [n_ch][LK_PIX + 1]
  __shared__ float s_xy[2][LK_PIX];
  const int nf = L.h * L.w;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * LK_PIX;
  const int win = 2 * radius + 1, win2 = win * win;
  const int n_ch = L.levels * win2;
  if (threadIdx.x < 2 * LK_PIX) {
    const int ch = threadIdx.x / LK_PIX, p = threadIdx.x % LK_PIX;
    s_xy[ch][p] = (i0 + p < nf) ? __ldg(coords + ((size_t)b * 2 + ch) * nf + i0 + p) : 0.f;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < LK_PIX * n_ch; q += LK_THREADS) {
    const int p = q / n_ch, k = q - p * n_ch;
    const int i = i0 + p;
    float val = 0.f;
    if (i < nf) {
      const int l = k / win2, rem = k - l * win2;
      const int wa = rem / win, wb = rem - wa * win;
      const float inv = 1.0f / (float)(1 << l);  // coords / 2**l, exact
      const float xs = __fadd_rn(s_xy[0][p] * inv, (float)(wa - radius));
      const float ys = __fadd_rn(s_xy[1][p] * inv, (float)(wb - radius));
      const int W = L.level_w[l], H = L.level_h[l];
      const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
      // bilinear_sampler normalisation (utils.py:19-20) ...
      const float xg = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, xs), wm1), 1.f);
      const float yg = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, ys), hm1), 1.f);
      // ... undone by grid_sample(align_corners=True): ((g + 1) / 2) * (size - 1)
      const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(xg, 1.f), 2.f), wm1);
      const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(yg, 1.f), 2.f), hm1);
      if (fabsf(ix) < 1e7f && fabsf(iy) < 1e7f) {
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);  // se - i
        const float dx = __fsub_rn(ix, fx), dy = __fsub_rn(iy, fy);                                  // i - nw
        const T* row = pyr + ((size_t)b * nf + i) * L.pitch + L.level_offset[l];
        const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
        const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
        if (yin0 && xin0) val += ld_val(row + y0 * W + x0) * __fmul_rn(ex, ey);            // nw
        if (yin0 && xin1) val += ld_val(row + y0 * W + x0 + 1) * __fmul_rn(dx, ey);        // ne
        if (yin1 && xin0) val += ld_val(row + (y0 + 1) * W + x0) * __fmul_rn(ex, dy);      // sw
        if (yin1 && xin1) val += ld_val(row + (y0 + 1) * W + x0 + 1) * __fmul_rn(dx, dy);  // se
      }
    }
    s_out[k * (LK_PIX + 1) + p] = val;
  }
  __syncthreads();
  const int lane = lane_id(), warp = warp_id();
  if (i0 + lane < nf) {
    for (int k = warp; k < n_ch; k += LK_THREADS / 32)
      out[((size_t)b * n_ch + k) * nf + i0 + lane] = s_out[k * (LK_PIX + 1) + lane];
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
  const size_t smem = (size_t)n_ch * (LK_PIX + 1) * sizeof(float);
  dim3 grid((nf + LK_PIX - 1) / LK_PIX, L->batch);
  if (pyramid_dtype == SLIMB200_DTYPE_BF16) {
    static bool attr_set = false;
    if (!attr_set) {
      SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_corr_lookup<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             LK_MAX_CH * (LK_PIX + 1) * (int)sizeof(float)));
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
                                             LK_MAX_CH * (LK_PIX + 1) * (int)sizeof(float)));
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
