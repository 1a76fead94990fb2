// Glue kernel for the channels-last feature encoder: affine InstanceNorm2d (+ optional ReLU) on an NHWC tensor.
//
// Replaces `norm_fn = "instance_affine"` + ReLU of the SLIM feature extractor (liso/slim/model/extractor.py:5-68,
// 211-297: nn.InstanceNorm2d(C, eps=1e-3, affine=True) followed by nn.ReLU), which PyTorch runs as a copy to NCHW,
// cuDNN's batch-norm kernel on (1, B*C, H, W), a copy back and a clamp.  Three launches, two passes over the data:
//   k_in_stats     per (sample, slab of pixels): shifted sums (shift = the slab's first pixel) per channel, lanes = channel
//                  groups of 4 (every load is a coalesced float4 of a pixel's channel vector), lanes added in fp64 ->
//                  (count, mean, M2) per slab; about one wave of CTAs whatever the tensor size
//   k_in_finalize  per (sample, channel): exact pooling of the slab partials in fp64 (no division per partial) ->
//                  scale = gamma / sqrt(var + eps), shift = beta - mean * scale (biased variance, like F.instance_norm)
//   k_in_apply     out = max(x * scale + shift, 0) as float4; optionally the residual join of the block is fused in:
//                  out = relu(residual + relu(x * scale + shift))
// The convolutions themselves stay stock cuDNN.
#include "common.cuh"

namespace {

constexpr int IN_THREADS = 256;
constexpr int IN_MAX_SLABS = 512;

struct InArgs {
  const float* x;
  const float* gamma;
  const float* beta;
  const float* residual;  // optional (same shape as x): added after the normalisation (+ inner ReLU)
  float* out;
  float* partial;   // [batch][slabs][C][3] (count, mean, M2)
  float* scale_shift;  // [batch][C][2]
  int batch, hw, C, slabs, relu;
  int xpitch;  // floats between consecutive pixels of x (== C for a packed tensor, larger for a channel slice)
  float eps;
};

// One warp merges the slab partials of one (sample, channel): exact pooled mean and M2 in fp64 without a division per
// partial -- N = sum n_s, mean = sum n_s mean_s / N, M2 = sum [M2_s + n_s (mean_s - mean)^2] -- lanes take strided
// subsets, plain butterfly sums; biased variance like F.instance_norm.
__device__ __forceinline__ void in_finalize_channel(const InArgs& a, int b, int c, int lane) {
  const float* base = a.partial + ((size_t)b * a.slabs * a.C + c) * 3;
  const size_t stride = (size_t)a.C * 3;
  double n = 0.0, sm = 0.0;
  for (int s = lane; s < a.slabs; s += 32) {
    const double nb = (double)__ldcg(base + s * stride), mb = (double)__ldcg(base + s * stride + 1);
    n += nb;
    sm = fma(nb, mb, sm);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, d);
    sm += __shfl_xor_sync(0xffffffffu, sm, d);
  }
  const double mu = n > 0.0 ? sm / n : 0.0;
  double M2 = 0.0;
  for (int s = lane; s < a.slabs; s += 32) {
    const double nb = (double)__ldcg(base + s * stride), dl = (double)__ldcg(base + s * stride + 1) - mu;
    M2 += (double)__ldcg(base + s * stride + 2) + nb * dl * dl;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) M2 += __shfl_xor_sync(0xffffffffu, M2, d);
  if (lane == 0) {
    const double var = n > 0.0 ? M2 / n : 0.0;
    const float scale = a.gamma[c] * (float)(1.0 / sqrt(var + (double)a.eps));
    a.scale_shift[((size_t)b * a.C + c) * 2] = scale;
    a.scale_shift[((size_t)b * a.C + c) * 2 + 1] = a.beta[c] - (float)mu * scale;
  }
}

__global__ void __launch_bounds__(IN_THREADS, 6) k_in_stats(const InArgs a) {
  __shared__ float s_s1[IN_THREADS][4], s_s2[IN_THREADS][4];
  const int G = a.C >> 2;               // float4 groups per pixel
  const int P = IN_THREADS / G;         // pixels per pass
  const int b = blockIdx.y, slab = blockIdx.x;
  const int g = threadIdx.x % G, p = threadIdx.x / G;
  const int per_slab = (a.hw + a.slabs - 1) / a.slabs;
  const int lo = slab * per_slab, hi = min(a.hw, lo + per_slab);
  // shifted sums: the shift of a channel is its value at the slab's first pixel (shared by all pixel-lanes, so their
  // sums simply add); the sums stay small and s2 - s1^2/n does not cancel.  Two FMAs per element, 8 loads in flight.
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float K[4] = {0.f, 0.f, 0.f, 0.f};
  if (lo < hi && p < P) {
    const int XG = a.xpitch >> 2;
    const float4* src = reinterpret_cast<const float4*>(a.x + (size_t)b * a.hw * a.xpitch) + g;
    const float4 k4 = __ldg(src + (size_t)lo * XG);
    K[0] = k4.x; K[1] = k4.y; K[2] = k4.z; K[3] = k4.w;
#pragma unroll 8
    for (int i = lo + p; i < hi; i += P) {
      const float4 v = __ldg(src + (size_t)i * XG);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = xs[k] - K[k];
        s1[k] += d;
        s2[k] = fmaf(d, d, s2[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_s1[threadIdx.x][k] = s1[k];
    s_s2[threadIdx.x][k] = s2[k];
  }
  __syncthreads();
  // thread (g, k) adds the P pixel-lanes of its channel in fixed order (fp64) and converts to (count, mean, M2)
  if (threadIdx.x < a.C) {
    const int c = threadIdx.x, gg = c >> 2, k = c & 3;
    double S1 = 0.0, S2 = 0.0;
    for (int q = 0; q < P; ++q) {
      S1 += (double)s_s1[q * G + gg][k];
      S2 += (double)s_s2[q * G + gg][k];
    }
    const int n = max(hi - lo, 0);
    float* o = a.partial + (((size_t)b * a.slabs + slab) * a.C + c) * 3;
    double mean = 0.0, M2 = 0.0;
    if (n > 0) {
      // the thread that loaded channel c's shift is (p = 0, g = gg): same value in every lane of the group
      const double Kc = (double)__ldg(a.x + ((size_t)b * a.hw + lo) * a.xpitch + c);
      mean = Kc + S1 / n;
      M2 = fmax(S2 - S1 * S1 / n, 0.0);
    }
    o[0] = (float)n;
    o[1] = (float)mean;
    o[2] = (float)M2;
  }
}

// one warp per (sample, channel)
__global__ void __launch_bounds__(IN_THREADS) k_in_finalize(const InArgs a) {
  const int i = (blockIdx.x * IN_THREADS + threadIdx.x) >> 5;
  if (i >= a.batch * a.C) return;
  in_finalize_channel(a, i / a.C, i % a.C, threadIdx.x & 31);
}

// out = [relu_outer]([relu_inner](x * scale + shift) + residual), float4.  relu bit 0 = inner (the norm's own ReLU),
// bit 1 = outer (the residual join relu(x + y) of extractor.py:57-68)
__global__ void __launch_bounds__(IN_THREADS) k_in_apply(const InArgs a) {
  const int G = a.C >> 2;
  const size_t per_sample = (size_t)a.hw * G;
  const int b = blockIdx.y;
  const int XG = a.xpitch >> 2;
  const float4* src = reinterpret_cast<const float4*>(a.x) + (size_t)b * a.hw * XG;
  float4* dst = reinterpret_cast<float4*>(a.out) + (size_t)b * per_sample;
  const float4* res = a.residual ? reinterpret_cast<const float4*>(a.residual) + (size_t)b * per_sample : nullptr;
  const float* ss = a.scale_shift + (size_t)b * a.C * 2;
  for (size_t i = (size_t)blockIdx.x * IN_THREADS + threadIdx.x; i < per_sample; i += (size_t)gridDim.x * IN_THREADS) {
    const int g = (int)(i % G);
    const float4 v = __ldg(XG == G ? src + i : src + (i / G) * XG + g);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(ss + g * 8));      // scale0 shift0 scale1 shift1
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(ss + g * 8 + 4));  // scale2 shift2 scale3 shift3
    float4 o;
    o.x = fmaf(v.x, s0.x, s0.y);
    o.y = fmaf(v.y, s0.z, s0.w);
    o.z = fmaf(v.z, s1.x, s1.y);
    o.w = fmaf(v.w, s1.z, s1.w);
    if (a.relu & 1) {
      o.x = fmaxf(o.x, 0.f);
      o.y = fmaxf(o.y, 0.f);
      o.z = fmaxf(o.z, 0.f);
      o.w = fmaxf(o.w, 0.f);
    }
    if (res) {
      const float4 r = __ldg(res + i);
      o.x = __fadd_rn(r.x, o.x);
      o.y = __fadd_rn(r.y, o.y);
      o.z = __fadd_rn(r.z, o.z);
      o.w = __fadd_rn(r.w, o.w);
      if (a.relu & 2) {
        o.x = fmaxf(o.x, 0.f);
        o.y = fmaxf(o.y, 0.f);
        o.z = fmaxf(o.z, 0.f);
        o.w = fmaxf(o.w, 0.f);
      }
    }
    dst[i] = o;
  }
}

// slabs per sample: one full wave of CTAs (148 SMs x 6 resident) over the batch, at least 32 pixels per slab
int slabs_for(int batch, int hw) {
  int s = (148 * 6) / batch;
  if (s > hw / 32) s = hw / 32;
  return s < 1 ? 1 : (s > IN_MAX_SLABS ? IN_MAX_SLABS : s);
}

}  // namespace

extern "C" size_t slimb200_instnorm_workspace_bytes(int32_t batch, int32_t channels, int32_t hw) {
  if (batch < 1 || channels < 4 || hw < 1) return 0;
  WorkspaceCarver w(nullptr);
  w.take<float>((size_t)batch * slabs_for(batch, hw) * channels * 3);
  w.take<float>((size_t)batch * channels * 2);
  return w.used();
}

extern "C" int slimb200_instnorm_nhwc_slice(const float* x, int32_t x_pitch, const float* gamma, const float* beta, float eps,
                                            int32_t batch, int32_t height, int32_t width, int32_t channels, int32_t relu,
                                            const float* residual, float* out, void* workspace, size_t workspace_bytes,
                                            void* stream_) {
  if (!x || !gamma || !beta || !out || !workspace || batch < 1 || height < 1 || width < 1) return SLIMB200_E_INVALID;
  if (channels < 4 || (channels & 3) || channels > IN_THREADS || x_pitch < channels || (x_pitch & 3)) return SLIMB200_E_UNSUPPORTED;
  if (x_pitch != channels && out == x) return SLIMB200_E_INVALID;  // a slice cannot be normalised in place (out is packed)
  const long long hw = (long long)height * width;
  if (hw > 0x7fffffffLL) return SLIMB200_E_UNSUPPORTED;
  if (workspace_bytes < slimb200_instnorm_workspace_bytes(batch, channels, (int)hw)) return SLIMB200_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(residual) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  InArgs a{};
  a.x = x;
  a.xpitch = x_pitch;
  a.gamma = gamma;
  a.beta = beta;
  a.residual = residual;
  a.out = out;
  a.batch = batch;
  a.hw = (int)hw;
  a.C = channels;
  a.slabs = slabs_for(batch, (int)hw);
  a.relu = relu;
  a.eps = eps;
  WorkspaceCarver w(workspace);
  a.partial = w.take<float>((size_t)batch * a.slabs * channels * 3);
  a.scale_shift = w.take<float>((size_t)batch * channels * 2);
  SLIMB200_LAUNCH(SLIMB200_K_IN_STATS, stream, (k_in_stats<<<dim3(a.slabs, batch), IN_THREADS, 0, stream>>>(a)));
  SLIMB200_LAUNCH(SLIMB200_K_IN_FINALIZE, stream,
                  (k_in_finalize<<<(batch * channels * 32 + IN_THREADS - 1) / IN_THREADS, IN_THREADS, 0, stream>>>(a)));
  const size_t per_sample = (size_t)hw * (channels >> 2);
  const unsigned gx = (unsigned)((per_sample + IN_THREADS * 4 - 1) / (IN_THREADS * 4) < 148 * 4 ? (per_sample + IN_THREADS * 4 - 1) / (IN_THREADS * 4) : 148 * 4);
  SLIMB200_LAUNCH(SLIMB200_K_IN_APPLY, stream, (k_in_apply<<<dim3(gx, batch), IN_THREADS, 0, stream>>>(a)));
  return SLIMB200_OK;
}

extern "C" int slimb200_instnorm_nhwc(const float* x, const float* gamma, const float* beta, float eps, int32_t batch,
                                      int32_t height, int32_t width, int32_t channels, int32_t relu, const float* residual,
                                      float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  return slimb200_instnorm_nhwc_slice(x, channels, gamma, beta, eps, batch, height, width, channels, relu, residual, out, workspace,
                                      workspace_bytes, stream_);
}
