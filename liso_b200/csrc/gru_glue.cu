// SURVEY 8(f).2 glue: the element-wise work between the stock convolutions of the ConvGRU update block, on
// channels-last (NHWC) fp32 tensors.  The convolutions themselves stay stock cuDNN.
//
// Reference (liso/slim/model/update.py):
//   ConvGRU.forward :30-38        hx = cat[h, x]; z = sigmoid(convz(hx)); r = sigmoid(convr(hx));
//                                 q = tanh(convq(cat[r * h, x])); h = (1 - z) * h + z * q
//   SmallMotionEncoder.forward :70-93   cat[cor, flo, logits] -> conv -> cat[out, logits, flow]
//   SmallUpdateBlock.forward :130-150   cat[inp, motion_features]
//   refinement loop (raft_mod.py:188-212)  coords1 += delta_flow; logits += delta_logits; flow = coords1 - coords0
// PyTorch runs this as 5 concatenation copies + 9 element-wise launches + 5 tiny adds per GRU iteration (12 iterations per
// pair).  Here the two 304-channel convolution inputs [h | x] and [r*h | x] are persistent buffers whose channel
// slots the producers write directly:
//   k_nhwc_pack      concatenates up to 4 packed NHWC sources along channels into a channel slot of up to 2
//                    destinations (cat[cor, flo, logits] for the motion conv; [out | logits | flow] -> both GRU inputs)
//   k_gru_gate_zr    z = sigmoid(zr[:, :96] + bz) -> packed; r = sigmoid(zr[:, 96:] + br); r * h -> slot 0 of [r*h | x]
//   k_gru_gate_out   q = tanh(qraw + bq); h' = (1 - z) * h + z * q -> slot 0 of [h | x] (in place) and a packed copy
//                    for the two heads
//   k_iter_update    bias add of both head outputs, coords1 / logits update and flow = coords1 - coords0 in one launch
//   k_add_relu       relu(x + y) (residual join of the context encoder, extractor.py:57-68)
// Arithmetic order follows the PyTorch expressions (separately rounded mul / add, IEEE division, expf / tanhf), so
// the results equal the stock element-wise kernels'.
#include "common.cuh"

namespace {

constexpr int GL_THREADS = 256;

inline unsigned grid_for(size_t items, int per_thread = 1) {
  size_t blocks = (items + (size_t)GL_THREADS * per_thread - 1) / ((size_t)GL_THREADS * per_thread);
  const size_t cap = 148 * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

struct PackArgs {
  const float4* src[4];
  int src_g[4];      // float4 groups per pixel of each source
  int src_pitch_g[4];  // row pitch of each source in float4 groups (== src_g for a packed tensor, larger for a channel slice)
  int n_src, total_g;
  float4* dst[2];
  int dst_off_g[2], dst_pitch_g[2];
  int n_dst;
  long long pixels;
};

__global__ void __launch_bounds__(GL_THREADS) k_nhwc_pack(const PackArgs a) {
  const size_t total = (size_t)a.pixels * a.total_g;
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * GL_THREADS) {
    const size_t p = i / a.total_g;
    const int gg = (int)(i - p * a.total_g);
    int g = gg, sg = a.src_g[0], spitch = a.src_pitch_g[0];
    const float4* sp = a.src[0];
#pragma unroll
    for (int s = 1; s < 4; ++s) {
      if (s < a.n_src && g >= sg) {  // past the current source: move on to source s
        g -= sg;
        sp = a.src[s];
        sg = a.src_g[s];
        spitch = a.src_pitch_g[s];
      }
    }
    const float4 v = __ldg(sp + p * spitch + g);
    a.dst[0][p * a.dst_pitch_g[0] + a.dst_off_g[0] + gg] = v;
    if (a.n_dst > 1) a.dst[1][p * a.dst_pitch_g[1] + a.dst_off_g[1] + gg] = v;
  }
}

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// zr: packed (pixels, 2 * Ch) raw convolution output (update gate channels first); hx / rhx: pitch in floats
__global__ void __launch_bounds__(GL_THREADS) k_gru_gate_zr(const float* __restrict__ zr, const float* __restrict__ bias_zr,
                                                            const float* __restrict__ hx, int hx_pitch, float* __restrict__ z_out,
                                                            float* __restrict__ rhx, int rhx_pitch, int Ch, long long pixels) {
  const int G = Ch >> 2;
  const size_t total = (size_t)pixels * G;
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * GL_THREADS) {
    const size_t p = i / G;
    const int c = (int)(i - p * G) * 4;
    const float4 zv = __ldg(reinterpret_cast<const float4*>(zr + p * 2 * Ch + c));
    const float4 rv = __ldg(reinterpret_cast<const float4*>(zr + p * 2 * Ch + Ch + c));
    const float4 bz = __ldg(reinterpret_cast<const float4*>(bias_zr + c));
    const float4 br = __ldg(reinterpret_cast<const float4*>(bias_zr + Ch + c));
    const float4 h = *reinterpret_cast<const float4*>(hx + p * hx_pitch + c);
    float4 z, rh;
    z.x = sigmoid_ref(__fadd_rn(zv.x, bz.x));
    z.y = sigmoid_ref(__fadd_rn(zv.y, bz.y));
    z.z = sigmoid_ref(__fadd_rn(zv.z, bz.z));
    z.w = sigmoid_ref(__fadd_rn(zv.w, bz.w));
    rh.x = __fmul_rn(sigmoid_ref(__fadd_rn(rv.x, br.x)), h.x);
    rh.y = __fmul_rn(sigmoid_ref(__fadd_rn(rv.y, br.y)), h.y);
    rh.z = __fmul_rn(sigmoid_ref(__fadd_rn(rv.z, br.z)), h.z);
    rh.w = __fmul_rn(sigmoid_ref(__fadd_rn(rv.w, br.w)), h.w);
    *reinterpret_cast<float4*>(z_out + p * Ch + c) = z;
    *reinterpret_cast<float4*>(rhx + p * rhx_pitch + c) = rh;
  }
}

__device__ __forceinline__ float gru_blend(float z, float h, float qraw, float b) {
  const float q = tanhf(__fadd_rn(qraw, b));
  return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, z), h), __fmul_rn(z, q));  // (1 - z) * h + z * q
}

__global__ void __launch_bounds__(GL_THREADS) k_gru_gate_out(const float* __restrict__ qraw, const float* __restrict__ bias_q,
                                                             const float* __restrict__ z, float* __restrict__ hx, int hx_pitch,
                                                             float* __restrict__ h_out, int Ch, long long pixels) {
  const int G = Ch >> 2;
  const size_t total = (size_t)pixels * G;
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * GL_THREADS) {
    const size_t p = i / G;
    const int c = (int)(i - p * G) * 4;
    const float4 qv = __ldg(reinterpret_cast<const float4*>(qraw + p * Ch + c));
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + p * Ch + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias_q + c));
    const float4 h = *reinterpret_cast<const float4*>(hx + p * hx_pitch + c);
    float4 o;
    o.x = gru_blend(zv.x, h.x, qv.x, b.x);
    o.y = gru_blend(zv.y, h.y, qv.y, b.y);
    o.z = gru_blend(zv.z, h.z, qv.z, b.z);
    o.w = gru_blend(zv.w, h.w, qv.w, b.w);
    *reinterpret_cast<float4*>(hx + p * hx_pitch + c) = o;
    *reinterpret_cast<float4*>(h_out + p * Ch + c) = o;
  }
}

struct IterArgs {
  const float* dflow;    // (batch, 2, h, w) raw head output, element (b, c, pix) at b * bs_f + c * cs_f + pix * ps_f
  const float* dlogits;  // (batch, nl, h, w) likewise with bs_l / cs_l / ps_l
  const float* bias_f;
  const float* bias_l;
  float* coords1;  // (batch, 2, h, w) planar
  float* flow;     // (batch, 2, h, w) planar
  float* logits;   // (batch, nl, h, w) planar
  float* stacked;  // optional (batch, |stacked_ch| >= 2 + nl, h, w) copy [flow | logits | untouched padding channels]:
  int stacked_ch;  // > 0 planar, < 0 channels-last (what the stock convolution that reads it wants)
  const float* taps;  // alternative source of the raw head outputs (see slimb200_iter_update_taps), else NULL
  int ksize;
  int batch, h, w, nl;
  long long bs_f, cs_f, ps_f, bs_l, cs_l, ps_l;
};

__global__ void __launch_bounds__(GL_THREADS) k_iter_update(const IterArgs a) {
  const int hw = a.h * a.w;
  const long long total = (long long)a.batch * hw;
  const int sch = a.stacked_ch < 0 ? -a.stacked_ch : a.stacked_ch;
  const size_t s_cs = a.stacked_ch < 0 ? 1 : (size_t)hw, s_ps = a.stacked_ch < 0 ? (size_t)sch : 1;  // channel / pixel strides
  for (long long i = (long long)blockIdx.x * GL_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * GL_THREADS) {
    const int b = (int)(i / hw), pix = (int)(i - (long long)b * hw);
    const int row = pix / a.w, col = pix - row * a.w;
    float* sp = a.stacked ? a.stacked + (size_t)b * sch * hw + (size_t)pix * s_ps : nullptr;
    float raw[2 + 16];
    if (a.taps) {
      // the k x k output convolution of both heads as ONE 1x1 convolution to (k*k taps) x (2 + nl) channels + this sum
      // of the taps over the window (zero padding): out[p] = sum_t taps[p + offset(t)][t]
      const int nc = 2 + a.nl, pad = a.ksize >> 1;
#pragma unroll
      for (int c = 0; c < 2 + 16; ++c) raw[c] = 0.f;
      for (int ky = 0; ky < a.ksize; ++ky) {
        const int r = row + ky - pad;
        if ((unsigned)r >= (unsigned)a.h) continue;
        for (int kx = 0; kx < a.ksize; ++kx) {
          const int cc = col + kx - pad;
          if ((unsigned)cc >= (unsigned)a.w) continue;
          const float* src = a.taps + (((size_t)b * hw + (size_t)r * a.w + cc) * (a.ksize * a.ksize) + (ky * a.ksize + kx)) * nc;
#pragma unroll
          for (int c = 0; c < 2 + 16; ++c)
            if (c < nc) raw[c] = __fadd_rn(raw[c], __ldg(src + c));
        }
      }
    } else {
      raw[0] = __ldg(a.dflow + b * a.bs_f + pix * a.ps_f);
      raw[1] = __ldg(a.dflow + b * a.bs_f + a.cs_f + pix * a.ps_f);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < a.nl) raw[2 + c] = __ldg(a.dlogits + b * a.bs_l + c * a.cs_l + pix * a.ps_l);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float d = __fadd_rn(raw[c], __ldg(a.bias_f + c));
      float* cp = a.coords1 + ((size_t)b * 2 + c) * hw + pix;
      const float nc = __fadd_rn(*cp, d);
      *cp = nc;
      const float fl = __fsub_rn(nc, (float)(c == 0 ? col : row));  // coords0: ch0 = x, ch1 = y
      a.flow[((size_t)b * 2 + c) * hw + pix] = fl;
      if (sp) sp[c * s_cs] = fl;
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c >= a.nl) break;
      const float d = __fadd_rn(raw[2 + c], __ldg(a.bias_l + c));
      float* lp = a.logits + ((size_t)b * a.nl + c) * hw + pix;
      const float nv = __fadd_rn(*lp, d);
      *lp = nv;
      if (sp) sp[(2 + c) * s_cs] = nv;
    }
  }
}

__global__ void __launch_bounds__(GL_THREADS) k_add_relu(const float4* __restrict__ x, const float4* __restrict__ y,
                                                         float4* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * GL_THREADS) {
    const float4 a = __ldg(x + i), b = __ldg(y + i);
    float4 o;
    o.x = fmaxf(__fadd_rn(a.x, b.x), 0.f);
    o.y = fmaxf(__fadd_rn(a.y, b.y), 0.f);
    o.z = fmaxf(__fadd_rn(a.z, b.z), 0.f);
    o.w = fmaxf(__fadd_rn(a.w, b.w), 0.f);
    out[i] = o;
  }
}

// relu((x + bias_x[c]) + y) on channels-last data: the projection shortcut of a residual block computed without its bias
// (ATen adds a convolution's bias in a separate pass over the tensor), rounded like conv-with-bias followed by relu(x + y)
__global__ void __launch_bounds__(GL_THREADS) k_add_bias_relu(const float4* __restrict__ x, const float4* __restrict__ bias_x,
                                                              int c4, const float4* __restrict__ y, float4* __restrict__ out,
                                                              size_t n4) {
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * GL_THREADS) {
    const float4 a = __ldg(x + i), b = __ldg(y + i), c = __ldg(bias_x + (i % (size_t)c4));
    float4 o;
    o.x = fmaxf(__fadd_rn(__fadd_rn(a.x, c.x), b.x), 0.f);
    o.y = fmaxf(__fadd_rn(__fadd_rn(a.y, c.y), b.y), 0.f);
    o.z = fmaxf(__fadd_rn(__fadd_rn(a.z, c.z), b.z), 0.f);
    o.w = fmaxf(__fadd_rn(__fadd_rn(a.w, c.w), b.w), 0.f);
    out[i] = o;
  }
}

// context encoder tail (raft_mod.py:170-173): net = tanh(raw[:, :ch] + b), inp = relu(raw[:, ch:] + b) from the bias-free
// output of cnet's last convolution -- one pass instead of ATen's bias add, split, tanh and clamp
__global__ void __launch_bounds__(GL_THREADS) k_ctx_split(const float4* __restrict__ raw, const float4* __restrict__ bias, int ch4,
                                                          int cx4, float4* __restrict__ net, float4* __restrict__ inp, size_t pixels) {
  const int c4 = ch4 + cx4;
  const size_t n4 = pixels * (size_t)c4;
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * GL_THREADS) {
    const size_t pix = i / (size_t)c4;
    const int g = (int)(i - pix * (size_t)c4);
    const float4 a = __ldg(raw + i), b = __ldg(bias + g);
    float4 v = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    if (g < ch4) {
      net[pix * (size_t)ch4 + g] = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    } else {
      inp[pix * (size_t)cx4 + (g - ch4)] = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
  }
}

// relu(x[:, c0 : c0 + C] + bias) of a channels-last tensor with `pitch` channels per pixel, packed: one half of a stacked
// convolution (two parallel convolutions evaluated as one, without bias) handed to its consumer
__global__ void __launch_bounds__(GL_THREADS) k_bias_relu_slice(const float4* __restrict__ x, int pitch4, const float4* __restrict__ bias,
                                                                int c4, float4* __restrict__ out, size_t pixels) {
  const size_t n4 = pixels * (size_t)c4;
  for (size_t i = (size_t)blockIdx.x * GL_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * GL_THREADS) {
    const size_t pix = i / (size_t)c4;
    const int g = (int)(i - pix * (size_t)c4);
    const float4 a = __ldg(x + pix * (size_t)pitch4 + g), b = __ldg(bias + g);
    out[i] = make_float4(fmaxf(__fadd_rn(a.x, b.x), 0.f), fmaxf(__fadd_rn(a.y, b.y), 0.f), fmaxf(__fadd_rn(a.z, b.z), 0.f),
                         fmaxf(__fadd_rn(a.w, b.w), 0.f));
  }
}

inline bool mis16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

}  // namespace

extern "C" int slimb200_nhwc_pack(const float* const* src /*host[n_src]*/, const int32_t* src_channels /*host[n_src]*/,
                                  const int32_t* src_pitch /*host[n_src] or NULL = packed*/, int32_t n_src, float* const* dst /*host[n_dst]*/, const int32_t* dst_channel_offset,
                                  const int32_t* dst_pitch, int32_t n_dst, int64_t pixels, void* stream_) {
  if (!src || !src_channels || !dst || !dst_channel_offset || !dst_pitch || pixels < 0) return SLIMB200_E_INVALID;
  if (n_src < 1 || n_src > 4 || n_dst < 1 || n_dst > 2) return SLIMB200_E_UNSUPPORTED;
  PackArgs a{};
  a.n_src = n_src;
  a.n_dst = n_dst;
  a.pixels = pixels;
  for (int i = 0; i < 4; ++i) {
    a.src[i] = nullptr;
    a.src_g[i] = 0;
    a.src_pitch_g[i] = 0;
  }
  int total = 0;
  for (int i = 0; i < n_src; ++i) {
    if (!src[i] || src_channels[i] < 4) return SLIMB200_E_INVALID;
    if (src_channels[i] & 3) return SLIMB200_E_UNSUPPORTED;
    if (mis16(src[i])) return SLIMB200_E_ALIGNMENT;
    a.src[i] = reinterpret_cast<const float4*>(src[i]);
    a.src_g[i] = src_channels[i] >> 2;
    const int pitch = src_pitch ? src_pitch[i] : src_channels[i];
    if (pitch < src_channels[i]) return SLIMB200_E_INVALID;
    if (pitch & 3) return SLIMB200_E_UNSUPPORTED;
    a.src_pitch_g[i] = pitch >> 2;
    total += src_channels[i];
  }
  a.total_g = total >> 2;
  for (int i = 0; i < n_dst; ++i) {
    if (!dst[i] || dst_channel_offset[i] < 0 || dst_pitch[i] < dst_channel_offset[i] + total) return SLIMB200_E_INVALID;
    if ((dst_channel_offset[i] & 3) || (dst_pitch[i] & 3)) return SLIMB200_E_UNSUPPORTED;
    if (mis16(dst[i])) return SLIMB200_E_ALIGNMENT;
    a.dst[i] = reinterpret_cast<float4*>(dst[i]);
    a.dst_off_g[i] = dst_channel_offset[i] >> 2;
    a.dst_pitch_g[i] = dst_pitch[i] >> 2;
  }
  if (pixels == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_NHWC_PACK, stream,
                  (k_nhwc_pack<<<grid_for((size_t)pixels * a.total_g, 2), GL_THREADS, 0, stream>>>(a)));
  return SLIMB200_OK;
}

extern "C" int slimb200_gru_gate_zr(const float* zr_raw, const float* bias_zr, const float* hx, int32_t hx_pitch, float* z_out,
                                    float* rhx, int32_t rhx_pitch, int32_t hidden, int64_t pixels, void* stream_) {
  if (!zr_raw || !bias_zr || !hx || !z_out || !rhx || pixels < 0) return SLIMB200_E_INVALID;
  if (hidden < 4 || (hidden & 3) || (hx_pitch & 3) || (rhx_pitch & 3) || hx_pitch < hidden || rhx_pitch < hidden)
    return SLIMB200_E_UNSUPPORTED;
  if (mis16(zr_raw) || mis16(bias_zr) || mis16(hx) || mis16(z_out) || mis16(rhx)) return SLIMB200_E_ALIGNMENT;
  if (pixels == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_GRU_GATE_ZR, stream,
                  (k_gru_gate_zr<<<grid_for((size_t)pixels * (hidden >> 2), 2), GL_THREADS, 0, stream>>>(
                      zr_raw, bias_zr, hx, hx_pitch, z_out, rhx, rhx_pitch, hidden, pixels)));
  return SLIMB200_OK;
}

extern "C" int slimb200_gru_gate_out(const float* q_raw, const float* bias_q, const float* z, float* hx, int32_t hx_pitch,
                                     float* h_out, int32_t hidden, int64_t pixels, void* stream_) {
  if (!q_raw || !bias_q || !z || !hx || !h_out || pixels < 0) return SLIMB200_E_INVALID;
  if (hidden < 4 || (hidden & 3) || (hx_pitch & 3) || hx_pitch < hidden) return SLIMB200_E_UNSUPPORTED;
  if (mis16(q_raw) || mis16(bias_q) || mis16(z) || mis16(hx) || mis16(h_out)) return SLIMB200_E_ALIGNMENT;
  if (pixels == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_GRU_GATE_OUT, stream,
                  (k_gru_gate_out<<<grid_for((size_t)pixels * (hidden >> 2), 2), GL_THREADS, 0, stream>>>(
                      q_raw, bias_q, z, hx, hx_pitch, h_out, hidden, pixels)));
  return SLIMB200_OK;
}

extern "C" int slimb200_iter_update(const float* dflow_raw, int64_t dflow_batch_stride, int64_t dflow_channel_stride,
                                    int64_t dflow_pixel_stride, const float* bias_flow, const float* dlogits_raw,
                                    int64_t dlogits_batch_stride, int64_t dlogits_channel_stride, int64_t dlogits_pixel_stride,
                                    const float* bias_logits, int32_t n_logits, int32_t batch, int32_t h, int32_t w,
                                    float* coords1, float* flow, float* logits, float* stacked, int32_t stacked_channels,
                                    void* stream_) {
  if (!dflow_raw || !bias_flow || !dlogits_raw || !bias_logits || !coords1 || !flow || !logits) return SLIMB200_E_INVALID;
  if (stacked && (stacked_channels < 0 ? -stacked_channels : stacked_channels) < 2 + n_logits) return SLIMB200_E_INVALID;
  if (batch < 1 || h < 1 || w < 1 || n_logits < 1 || n_logits > 16) return SLIMB200_E_INVALID;
  IterArgs a{};
  a.dflow = dflow_raw;
  a.dlogits = dlogits_raw;
  a.bias_f = bias_flow;
  a.bias_l = bias_logits;
  a.coords1 = coords1;
  a.flow = flow;
  a.logits = logits;
  a.stacked = stacked;
  a.stacked_ch = stacked_channels;
  a.bs_f = dflow_batch_stride;
  a.bs_l = dlogits_batch_stride;
  a.batch = batch;
  a.h = h;
  a.w = w;
  a.nl = n_logits;
  a.cs_f = dflow_channel_stride;
  a.ps_f = dflow_pixel_stride;
  a.cs_l = dlogits_channel_stride;
  a.ps_l = dlogits_pixel_stride;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_ITER_UPDATE, stream,
                  (k_iter_update<<<grid_for((size_t)batch * h * w), GL_THREADS, 0, stream>>>(a)));
  return SLIMB200_OK;
}

extern "C" int slimb200_iter_update_taps(const float* taps, int32_t ksize, const float* bias_flow, const float* bias_logits,
                                         int32_t n_logits, int32_t batch, int32_t h, int32_t w, float* coords1, float* flow,
                                         float* logits, float* stacked, int32_t stacked_channels, void* stream_) {
  if (!taps || !bias_flow || !bias_logits || !coords1 || !flow || !logits) return SLIMB200_E_INVALID;
  if (stacked && (stacked_channels < 0 ? -stacked_channels : stacked_channels) < 2 + n_logits) return SLIMB200_E_INVALID;
  if (batch < 1 || h < 1 || w < 1 || n_logits < 1 || n_logits > 16 || ksize < 1 || !(ksize & 1) || ksize > 7) return SLIMB200_E_INVALID;
  IterArgs a{};
  a.taps = taps;
  a.ksize = ksize;
  a.bias_f = bias_flow;
  a.bias_l = bias_logits;
  a.coords1 = coords1;
  a.flow = flow;
  a.logits = logits;
  a.stacked = stacked;
  a.stacked_ch = stacked_channels;
  a.batch = batch;
  a.h = h;
  a.w = w;
  a.nl = n_logits;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_ITER_UPDATE, stream,
                  (k_iter_update<<<grid_for((size_t)batch * h * w), GL_THREADS, 0, stream>>>(a)));
  return SLIMB200_OK;
}

extern "C" int slimb200_add_relu(const float* x, const float* y, float* out, int64_t n, void* stream_) {
  if (!x || !y || !out || n < 0) return SLIMB200_E_INVALID;
  if (n & 3) return SLIMB200_E_UNSUPPORTED;
  if (mis16(x) || mis16(y) || mis16(out)) return SLIMB200_E_ALIGNMENT;
  if (n == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_ADD_RELU, stream,
                  (k_add_relu<<<grid_for((size_t)n / 4, 4), GL_THREADS, 0, stream>>>(
                      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(out),
                      (size_t)n / 4)));
  return SLIMB200_OK;
}

extern "C" int slimb200_add_bias_relu(const float* x, const float* bias_x, int32_t channels, const float* y, float* out, int64_t n,
                                      void* stream_) {
  if (!x || !bias_x || !y || !out || n < 0 || channels < 4) return SLIMB200_E_INVALID;
  if ((n & 3) || (channels & 3) || n % channels) return SLIMB200_E_UNSUPPORTED;
  if (mis16(x) || mis16(y) || mis16(out) || mis16(bias_x)) return SLIMB200_E_ALIGNMENT;
  if (n == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_ADD_RELU, stream,
                  (k_add_bias_relu<<<grid_for((size_t)n / 4, 4), GL_THREADS, 0, stream>>>(
                      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(bias_x), channels / 4,
                      reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(out), (size_t)n / 4)));
  return SLIMB200_OK;
}

extern "C" int slimb200_ctx_split(const float* raw, const float* bias, int32_t hidden, int32_t context, int64_t pixels, float* net,
                                  float* inp, void* stream_) {
  if (!raw || !bias || !net || !inp || pixels < 0 || hidden < 4 || context < 4) return SLIMB200_E_INVALID;
  if ((hidden & 3) || (context & 3)) return SLIMB200_E_UNSUPPORTED;
  if (mis16(raw) || mis16(bias) || mis16(net) || mis16(inp)) return SLIMB200_E_ALIGNMENT;
  if (pixels == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_CTX_SPLIT, stream,
                  (k_ctx_split<<<grid_for((size_t)pixels * (hidden + context) / 4, 4), GL_THREADS, 0, stream>>>(
                      reinterpret_cast<const float4*>(raw), reinterpret_cast<const float4*>(bias), hidden / 4, context / 4,
                      reinterpret_cast<float4*>(net), reinterpret_cast<float4*>(inp), (size_t)pixels)));
  return SLIMB200_OK;
}

extern "C" int slimb200_bias_relu_slice(const float* x, int32_t x_pitch, const float* bias, int32_t channels, int64_t pixels,
                                        float* out, void* stream_) {
  if (!x || !bias || !out || pixels < 0 || channels < 4 || x_pitch < channels) return SLIMB200_E_INVALID;
  if ((channels & 3) || (x_pitch & 3)) return SLIMB200_E_UNSUPPORTED;
  if (mis16(x) || mis16(bias) || mis16(out)) return SLIMB200_E_ALIGNMENT;
  if (pixels == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIMB200_LAUNCH(SLIMB200_K_BIAS_RELU_SLICE, stream,
                  (k_bias_relu_slice<<<grid_for((size_t)pixels * channels / 4, 4), GL_THREADS, 0, stream>>>(
                      reinterpret_cast<const float4*>(x), x_pitch / 4, reinterpret_cast<const float4*>(bias), channels / 4,
                      reinterpret_cast<float4*>(out), (size_t)pixels)));
  return SLIMB200_OK;
}
