// SURVEY 8(f).1 -- the SLIM output decoder on B200: BEV class/flow maps, BEV -> point gather and the weighted
// Kabsch "static aggregation", fused into five launches with no host synchronisation.
//
// Replaces, for the released output_modification (liso_config.yml:303-310):
//   HeadDecoder.forward                               liso/slim/model/head_decoder.py:410-496, 517-717
//   batched_grid_data_to_pointwise_data               liso/slim/slim_loss/static_aggregation.py:8-31
//   compute_batched_bev_static_aggregated_flow        static_aggregation.py:34-110
//   weighted_pc_alignment (no epsilon)                slim_loss/weighted_pc_alignment.py:10-80
//   symmetric_orthogonalization (U @ Vh, no det fix)  liso/torch_symm_ortho/__init__.py:68-69
// The reference runs ~60 element-wise launches per call plus, per sample, boolean-mask indexing (host sync), a
// fp64 3x3 SVD and host-side asserts; it is called 12 times per frame pair.
//
//   k_decode_min      global min of the live static/dynamic logits (ground logit "off" = min - 100)
//   k_decode_bev      one thread per BEV cell: masks, 3-way softmax, class decisions, masked flows, aggregated flow;
//                     one packed (B,H,W,16) row per cell, staged in smem and written back linearly
//   k_decode_points   one thread per point: gather of the BEV values, Kabsch weight
//   k_kabsch_moments  streaming fp64 moments (several points per thread, block partials in fixed order)
//   k_kabsch_finalize one CTA per sample: sum the partials, 3x3 one-sided Jacobi SVD in fp64, R = U V^T, T (4x4)
//   k_decode_aggr     (T - I) * cell centre for every cell and every point
#include "common.cuh"

namespace {

constexpr int BEV_C = SLIMB200_DECODE_BEV_CHANNELS;    // 20
constexpr int PT_C = SLIMB200_DECODE_POINT_CHANNELS;   // 14
constexpr int N_MOM = 33;  // 1 + 3 + 3 + 9 weighted, the same 16 unweighted (for the +1e-7 case), count(w > 0)
constexpr int PT_THREADS = 256;
constexpr int PT_PER_THREAD = 8;  // points per thread before the (expensive, fp64) block reduction

__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Same-address atomics serialise in L2 (~1 per 2 clocks): reduce per CTA and only touch the global word when this
// CTA can still lower it (the running minimum settles after a handful of CTAs).
__device__ __forceinline__ void block_min_to_global(unsigned best, unsigned* __restrict__ key) {
  __shared__ unsigned s_min[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0 && key) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) best = min(best, s_min[k]);
    if (best < *reinterpret_cast<volatile unsigned*>(key)) atomicMin(key, best);
  }
}

__global__ void __launch_bounds__(256) k_decode_min(const float* __restrict__ net_out, size_t n_cells, unsigned* __restrict__ out_key) {
  unsigned best = 0xffffffffu;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (size_t)gridDim.x * blockDim.x) {
    const float4 lo = __ldg(reinterpret_cast<const float4*>(net_out + i * 8));  // logits 0..3
    best = min(best, min(f2key(lo.y), f2key(lo.z)));
  }
  block_min_to_global(best, out_key);
}

// ------------------------------------------------------------------------------------------
// SURVEY 8(f).2 glue: RAFT.predict_single_flow_map_and_classes, raft_mod.py:216-257 -- per GRU iteration
//   upflow_n / uplogits_n   (raft_code/utils.py:50-60: F.interpolate bilinear, align_corners=True, flow * n)
//   change_flow_convention_from_raft2usfl (raft_mod.py:262-266: flip (x, y) -> (row, col), * metres per pixel)
//   HeadDecoder.concat2network_output (head_decoder.py:36-64: cat [logits, static, dynamic] -> (B, H, W, 8))
// as ONE pass that writes the (B, H, W, 8) network output (32 contiguous bytes per pixel) and, on the way, takes
// the global min of the static / dynamic logits that the decoder needs (k_decode_min disappears).
// Interpolation follows ATen's upsample_bilinear2d: src = dst * (in - 1) / (out - 1), taps (i, i + (i < in - 1)).
// ------------------------------------------------------------------------------------------
constexpr int RO_THREADS = 128;

__device__ __forceinline__ float sel3(float a, float b, float c, int i) { return i == 0 ? a : (i == 1 ? b : c); }

// One thread = one output column x and RO_ROWS <= n consecutive output rows: the rows share the column taps / weights
// and, since (RO_ROWS - 1) * (h - 1) / (H - 1) < 1, touch at most 3 source rows, so (n = 8) the 6 maps cost 36 loads
// per 8 output cells instead of 192 and the horizontal blends are shared.  Same association as ATen:
// hl0 * (wl0 * a + wl1 * b) + hl1 * (wl0 * c + wl1 * d).
template <int RO_ROWS>
__global__ void __launch_bounds__(RO_THREADS) k_raft_output(const float* __restrict__ flow, const float* __restrict__ logits,
                                                            int batch, int h, int w, int n, float res_rows, float res_cols,
                                                            float* __restrict__ net_out, unsigned* __restrict__ min_key) {
  const int H = h * n, W = w * n;
  const int groups = (H + RO_ROWS - 1) / RO_ROWS;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = (blockIdx.y % groups) * RO_ROWS, b = blockIdx.y / groups;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int x0 = x - lane;  // first pixel of the warp
  unsigned best = 0xffffffffu;
  const float rh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int hb = (int)(rh * (float)y0);  // first source row of the group
  float t[6][3];                         // horizontally blended source rows hb, hb + 1, hb + 2 of the 6 maps
  if (x < W) {
    const float wr = rw * (float)x;
    const int w1 = (int)wr;
    const int wp = w1 < w - 1 ? 1 : 0;
    const float wl1 = wr - (float)w1, wl0 = 1.f - wl1;
    const size_t plane = (size_t)h * w;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const float* img = c < 4 ? logits + ((size_t)b * 4 + c) * plane : flow + ((size_t)b * 2 + (c - 4)) * plane;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hr = min(hb + r, h - 1);
        t[c][r] = wl0 * __ldg(img + (size_t)hr * w + w1) + wl1 * __ldg(img + (size_t)hr * w + w1 + wp);
      }
    }
  }
#pragma unroll
  for (int dy = 0; dy < RO_ROWS; ++dy) {
    const int y = y0 + dy;
    if (y >= H) break;  // warp-uniform
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (x < W) {
      const float hr = rh * (float)y;
      const int h1 = (int)hr;
      const int hp = h1 < h - 1 ? 1 : 0;
      const float hl1 = hr - (float)h1, hl0 = 1.f - hl1;
      const int i0 = h1 - hb, i1 = i0 + hp;  // 0..2
      float o[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) o[c] = hl0 * sel3(t[c][0], t[c][1], t[c][2], i0) + hl1 * sel3(t[c][0], t[c][1], t[c][2], i1);
      const float fx = __fmul_rn((float)n, o[4]), fy = __fmul_rn((float)n, o[5]);              // upflow_n
      const float f_row = __fmul_rn(fy, res_rows), f_col = __fmul_rn(fx, res_cols);            // flip, then * res
      v0 = make_float4(o[0], o[1], o[2], o[3]);
      v1 = make_float4(f_row, f_col, f_row, f_col);
      best = min(best, min(f2key(o[1]), f2key(o[2])));
    }
    // a warp owns 32 consecutive pixels = 1 KB of output: exchange so that each store instruction writes 512
    // contiguous bytes (lane l stores float4 number l resp. 32 + l of the warp's 64)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int src = half * 16 + (lane >> 1);
      const bool second = lane & 1;
      float4 t0, t1;
      t0.x = __shfl_sync(0xffffffffu, v0.x, src); t0.y = __shfl_sync(0xffffffffu, v0.y, src);
      t0.z = __shfl_sync(0xffffffffu, v0.z, src); t0.w = __shfl_sync(0xffffffffu, v0.w, src);
      t1.x = __shfl_sync(0xffffffffu, v1.x, src); t1.y = __shfl_sync(0xffffffffu, v1.y, src);
      t1.z = __shfl_sync(0xffffffffu, v1.z, src); t1.w = __shfl_sync(0xffffffffu, v1.w, src);
      const float4 tv = second ? t1 : t0;
      if (x0 + src < W) reinterpret_cast<float4*>(net_out + (((size_t)b * H + y) * W + x0) * 8)[half * 32 + lane] = tv;
    }
  }
  block_min_to_global(best, min_key);
}

struct DecodeArgs {
  const float* net_out;
  const uint8_t* filled;
  const float* pc;
  const int32_t* coors;
  const uint8_t* valid;
  const float* thr;
  slimb200_decode_params p;
  float* bev;
  float* bev_aggr;
  uint8_t* bev_cls;
  float* pts;
  double* trafo;
  uint8_t* not_enough;
  unsigned* min_key;
  double* partials;  // [batch][blocks_per_sample][N_MOM]
  float* pt_weight;  // [batch][n_points] Kabsch weight per point, -1 = invalid point
  int blocks_per_sample;
};

__global__ void __launch_bounds__(256) k_decode_bev(const DecodeArgs a) {
  // a CTA's 256 cells form one contiguous 20 KB block of the packed output: stage the rows in shared memory
  // (80-byte smem pitch: the 128-bit stores of a quarter-warp hit 32 distinct banks) and write them back linearly
  constexpr int S_PITCH = BEV_C + 4;
  __shared__ __align__(16) float s_row[256 * S_PITCH];
  __shared__ __align__(4) uint8_t s_cls[256 * 3];
  const size_t n_cells = (size_t)a.p.batch * a.p.H * a.p.W;
  const size_t base = (size_t)blockIdx.x * 256;
  const size_t i = base + threadIdx.x;
  if (i < n_cells) {
    const float4 lg = __ldg(reinterpret_cast<const float4*>(a.net_out + i * 8));
    const float4 fl = __ldg(reinterpret_cast<const float4*>(a.net_out + i * 8 + 4));
    const bool filled = a.filled[i] != 0;
    const float ground_off = key2f(*a.min_key) - 100.0f;
    // mask non-filled pillars (head_decoder.py:568-609)
    const float l_st = filled ? lg.y : 0.f;
    const float l_dy = filled ? lg.z : -100.f;
    const float l_gr = filled ? ground_off : -100.f;
    const float sfx = filled ? fl.x : 0.f, sfy = filled ? fl.y : 0.f;
    const float dfx = filled ? fl.z : 0.f, dfy = filled ? fl.w : 0.f;
    // softmax over (static, dynamic, ground)
    const float m = fmaxf(l_st, fmaxf(l_dy, l_gr));
    const float e0 = expf(l_st - m), e1 = expf(l_dy - m), e2 = expf(l_gr - m);
    const float s = e0 + e1 + e2;
    const float p_st = e0 / s, p_dy = e1 / s, p_gr = e2 / s;
    const float thr = __ldg(a.thr);
    const bool is_dyn = p_dy >= thr;
    const bool is_sta = (p_st >= p_gr) && !is_dyn;
    const bool is_gr = !(is_sta || is_dyn);
    const float g = 1.0f - p_gr;
    const float agx = is_sta ? sfx : dfx * g, agy = is_sta ? sfy : dfy * g, agz = is_sta ? 0.f : 0.f * g;
    float4* o = reinterpret_cast<float4*>(s_row + threadIdx.x * S_PITCH);
    o[0] = make_float4(-100.f, l_st, l_dy, l_gr);      // disappearing | class_logits (static, dynamic, ground)
    o[1] = make_float4(p_st, p_dy, p_gr, sfx);         // class_probs | static3.x
    o[2] = make_float4(sfy, 0.f, dfx, dfy);            // static3.yz | dynamic3.xy
    o[3] = make_float4(0.f, agx, agy, agz);            // dynamic3.z | aggregated3
    s_cls[threadIdx.x * 3 + 0] = is_dyn;
    s_cls[threadIdx.x * 3 + 1] = is_sta;
    s_cls[threadIdx.x * 3 + 2] = is_gr;
  }
  __syncthreads();
  const int n_here = (int)min((size_t)256, n_cells - base);
  float4* dst = reinterpret_cast<float4*>(a.bev + base * BEV_C);
  const float4* src = reinterpret_cast<const float4*>(s_row);
  for (int k = threadIdx.x; k < n_here * (BEV_C / 4); k += 256) dst[k] = src[(k >> 2) * (S_PITCH / 4) + (k & 3)];
  static_assert(BEV_C == 16, "row copy assumes 4 float4 per cell");
  if (n_here == 256) {
    if (threadIdx.x < 192) reinterpret_cast<uint32_t*>(a.bev_cls + base * 3)[threadIdx.x] = reinterpret_cast<const uint32_t*>(s_cls)[threadIdx.x];
  } else {
    for (int k = threadIdx.x; k < n_here * 3; k += 256) a.bev_cls[base * 3 + k] = s_cls[k];
  }
}

// gather: one thread per point, fully parallel; the 256 x 14 output floats of a CTA are contiguous and are written back
// linearly from shared memory; the Kabsch weight goes to a scratch array (-1 marks invalid points)
__global__ void __launch_bounds__(PT_THREADS) k_decode_points(const DecodeArgs a) {
  __shared__ __align__(16) float s_pts[PT_THREADS * PT_C];
  const int b = blockIdx.y;
  const int n = a.p.n_points;
  const int j0 = blockIdx.x * PT_THREADS;
  const int j = j0 + threadIdx.x;
  float* so = s_pts + threadIdx.x * PT_C;
#pragma unroll
  for (int k = 0; k < PT_C; ++k) so[k] = 0.f;
  if (j < n) {
    const size_t pi = (size_t)b * n + j;
    float w = -1.f;
    if (a.valid[pi] != 0) {
      const int r = a.coors[pi * 2] / a.p.final_scale, c = a.coors[pi * 2 + 1] / a.p.final_scale;
      const size_t cell = ((size_t)b * a.p.H + r) * a.p.W + c;
      const float* row = a.bev + cell * BEV_C;
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(row + 4));    // staticness dynamicness groundness static.x
      const float4 q2 = __ldg(reinterpret_cast<const float4*>(row + 8));    // static.y static.z dynamic.x dynamic.y
      const float4 q3 = __ldg(reinterpret_cast<const float4*>(row + 12));   // dynamic.z aggregated.xyz
      so[0] = q1.w; so[1] = q2.x; so[2] = q2.y;      // static3
      so[3] = q2.z; so[4] = q2.w; so[5] = q3.x;      // dynamic3
      so[6] = q1.y;                                  // dynamicness
      so[7] = q1.x;                                  // staticness
      so[8] = q3.y; so[9] = q3.z; so[10] = q3.w;     // aggregated3
      if (a.p.static_aggregation) w = a.filled[cell] ? q1.x : 0.f;  // staticness * filled (static_aggregation.py:66-71)
    }
    if (a.p.static_aggregation) a.pt_weight[pi] = w;
  }
  __syncthreads();
  const int n_here = min(PT_THREADS, n - j0);
  float2* dst = reinterpret_cast<float2*>(a.pts + ((size_t)b * n + j0) * PT_C);  // 56-byte rows: 8-byte aligned
  const float2* src = reinterpret_cast<const float2*>(s_pts);
  for (int k = threadIdx.x; k < n_here * (PT_C / 2); k += PT_THREADS) dst[k] = src[k];
}

// fp64 Kabsch moments: streaming pass over (weight, point, gathered static flow), PT_PER_THREAD points per thread,
// fixed-order reduction lanes -> warps -> block partial; blocks are summed in order by k_kabsch_finalize
__global__ void __launch_bounds__(PT_THREADS) k_kabsch_moments(const DecodeArgs a) {
  __shared__ double s_part[PT_THREADS / 32][N_MOM];
  const int b = blockIdx.y;
  const int n = a.p.n_points;
  double mom[N_MOM];
#pragma unroll
  for (int k = 0; k < N_MOM; ++k) mom[k] = 0.0;
#pragma unroll 4
  for (int it = 0; it < PT_PER_THREAD; ++it) {
    const int j = (blockIdx.x * PT_PER_THREAD + it) * PT_THREADS + threadIdx.x;
    if (j < n) {
      const size_t pi = (size_t)b * n + j;
      const float w = __ldg(a.pt_weight + pi);
      if (w >= 0.f) {
        const float* q = a.pc + pi * a.p.pc_stride;
        const float* f = a.pts + pi * PT_C;
        const float x0 = __ldg(q), y0 = __ldg(q + 1), z0 = __ldg(q + 2);
        const float x1 = x0 + f[0], y1 = y0 + f[1], z1 = z0 + f[2];  // fp32 add like the reference
        const double p0[3] = {x0, y0, z0}, p1[3] = {x1, y1, z1};
        const double wd = (double)w;
        mom[0] += wd;
        mom[16] += 1.0;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          mom[1 + u] += wd * p0[u];
          mom[4 + u] += wd * p1[u];
          mom[17 + u] += p0[u];
          mom[20 + u] += p1[u];
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            mom[7 + u * 3 + v] += wd * p1[u] * p0[v];  // S[u][v] = sum w * y_u * x_v
            mom[23 + u * 3 + v] += p1[u] * p0[v];
          }
        }
        mom[32] += w > 0.f ? 1.0 : 0.0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N_MOM; ++k) {
    double v = mom[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < N_MOM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < PT_THREADS / 32; ++w) v += s_part[w][threadIdx.x];
    a.partials[((size_t)b * a.blocks_per_sample + blockIdx.x) * N_MOM + threadIdx.x] = v;
  }
}

// One-sided Jacobi (Hestenes) SVD of a 3x3 matrix in fp64: A V = U diag(sigma); returns R = U V^T (the
// orthogonal polar factor; symmetric_orthogonalization returns exactly U @ Vh, without determinant fix).
__device__ void polar_uvt(const double S[3][3], double R[3][3]) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i][j] = S[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        if (gamma == 0.0) continue;
        off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double U[3][3], sig[3];
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  const double smax = fmax(sig[0], fmax(sig[1], sig[2]));
  int n_ok = 0;
  bool ok[3];
  for (int j = 0; j < 3; ++j) {
    ok[j] = sig[j] > smax * 1e-14 && sig[j] > 0.0;
    n_ok += ok[j];
    for (int i = 0; i < 3; ++i) U[i][j] = ok[j] ? A[i][j] / sig[j] : 0.0;
  }
  if (n_ok == 2) {  // rank 2: complete the basis with the cross product (the null direction is not unique anyway)
    int z = !ok[0] ? 0 : (!ok[1] ? 1 : 2);
    const int u = (z + 1) % 3, v = (z + 2) % 3;
    U[0][z] = U[1][u] * U[2][v] - U[2][u] * U[1][v];
    U[1][z] = U[2][u] * U[0][v] - U[0][u] * U[2][v];
    U[2][z] = U[0][u] * U[1][v] - U[1][u] * U[0][v];
  } else if (n_ok < 2) {  // rank <= 1: fall back to the identity rotation
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        U[i][j] = i == j ? 1.0 : 0.0;
        V[i][j] = i == j ? 1.0 : 0.0;
      }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
}

__global__ void __launch_bounds__(64) k_kabsch_finalize(const DecodeArgs a) {
  __shared__ double s_m[N_MOM];
  const int b = blockIdx.x;
  if (threadIdx.x < N_MOM) {
    double v = 0.0;
    const double* src = a.partials + (size_t)b * a.blocks_per_sample * N_MOM + threadIdx.x;
    for (int k = 0; k < a.blocks_per_sample; ++k) v += src[(size_t)k * N_MOM];
    s_m[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const bool not_enough = s_m[32] < 3.0;  // (weights > 0).sum() < 3 -> weights += 1e-7 (weighted_pc_alignment.py:31-35)
  const double eps = not_enough ? 1e-7 : 0.0;
  const double cum = s_m[0] + eps * s_m[16];
  double mx[3], my[3], S[3][3], R[3][3];
  for (int u = 0; u < 3; ++u) {
    mx[u] = (s_m[1 + u] + eps * s_m[17 + u]) / cum;
    my[u] = (s_m[4 + u] + eps * s_m[20 + u]) / cum;
  }
  for (int u = 0; u < 3; ++u)
    for (int v = 0; v < 3; ++v) S[u][v] = (s_m[7 + u * 3 + v] + eps * s_m[23 + u * 3 + v]) / cum - my[u] * mx[v];
  polar_uvt(S, R);
  double* T = a.trafo + (size_t)b * 16;
  for (int u = 0; u < 3; ++u) {
    for (int v = 0; v < 3; ++v) T[u * 4 + v] = R[u][v];
    T[u * 4 + 3] = my[u] - (R[u][0] * mx[0] + R[u][1] * mx[1] + R[u][2] * mx[2]);
  }
  T[12] = 0.0;
  T[13] = 0.0;
  T[14] = 0.0;
  T[15] = 1.0;
  a.not_enough[b] = not_enough ? 1 : 0;
}

// static_aggr_flow of a cell: ((T - I) [xc, yc, 0, 1])[:2] in fp64, then float (static_aggregation.py:88-103);
// cell centres as head_decoder.py:498-514: (idx + 0.5) / shape * (max - min) + min
__device__ __forceinline__ float2 aggr_flow_of_cell(const double* __restrict__ T, const slimb200_decode_params& p, int r, int c) {
  // (the reference divides by the shape; multiplying by the fp64 reciprocal differs by <= 1 ulp of fp64, invisible
  // after the cast to float)
  const double xc = ((double)r + 0.5) * (1.0 / (double)p.H) * (p.ext_max_x - p.ext_min_x) + p.ext_min_x;
  const double yc = ((double)c + 0.5) * (1.0 / (double)p.W) * (p.ext_max_y - p.ext_min_y) + p.ext_min_y;
  const double fx = (T[0] - 1.0) * xc + T[1] * yc + T[3];
  const double fy = T[4] * xc + (T[5] - 1.0) * yc + T[7];
  return make_float2((float)fx, (float)fy);
}

// blockIdx.y < row_groups: AGGR_ROWS consecutive BEV rows of one sample per CTA, thread = column (a few fat CTAs instead
// of one tiny CTA per row: the kernel is a 52 MB streaming write); the remaining blockIdx.y values cover the points
constexpr int AGGR_THREADS = 128, AGGR_ROWS = 16;

__global__ void __launch_bounds__(AGGR_THREADS) k_decode_aggr(const DecodeArgs a, int row_groups) {
  if ((int)blockIdx.y < row_groups) {
    const int c = blockIdx.x * AGGR_THREADS + threadIdx.x;
    if (c >= a.p.W) return;
    const int groups_per_sample = (a.p.H + AGGR_ROWS - 1) / AGGR_ROWS;
    const int b = blockIdx.y / groups_per_sample, r0 = (blockIdx.y - b * groups_per_sample) * AGGR_ROWS;
    const double* T = a.trafo + (size_t)b * 16;
#pragma unroll 4
    for (int r = r0; r < min(r0 + AGGR_ROWS, a.p.H); ++r) {
      const size_t i = ((size_t)b * a.p.H + r) * a.p.W + c;
      const float2 f = aggr_flow_of_cell(T, a.p, r, c);
      const bool filled = a.filled[i] != 0;
      reinterpret_cast<float4*>(a.bev_aggr)[i] = make_float4(f.x, f.y, filled ? f.x : 0.f, filled ? f.y : 0.f);
    }
  } else {
    const size_t n_pts = (size_t)a.p.batch * a.p.n_points;
    const size_t pi = ((size_t)(blockIdx.y - row_groups) * gridDim.x + blockIdx.x) * AGGR_THREADS + threadIdx.x;
    if (pi >= n_pts) return;
    const int b = (int)(pi / (size_t)a.p.n_points);
    float2 f = make_float2(0.f, 0.f);
    if (a.valid[pi]) {
      const int r = a.coors[pi * 2] / a.p.final_scale, c = a.coors[pi * 2 + 1] / a.p.final_scale;
      f = aggr_flow_of_cell(a.trafo + (size_t)b * 16, a.p, r, c);
    }
    float* o = a.pts + pi * PT_C + 11;
    o[0] = f.x;
    o[1] = f.y;
    o[2] = 0.f;
  }
}

int blocks_per_sample(const slimb200_decode_params* p) {
  return (p->n_points + PT_THREADS * PT_PER_THREAD - 1) / (PT_THREADS * PT_PER_THREAD);
}

}  // namespace

extern "C" size_t slimb200_head_decode_workspace_bytes(const slimb200_decode_params* p) {
  if (!p || p->batch < 1 || p->n_points < 0) return 0;
  WorkspaceCarver w(nullptr);
  w.take<unsigned>(64);
  w.take<double>((size_t)p->batch * (blocks_per_sample(p) + 1) * N_MOM);
  w.take<float>((size_t)p->batch * p->n_points + 1);
  return w.used();
}

extern "C" int slimb200_raft_output(const float* flow, const float* logits, int32_t batch, int32_t h, int32_t w, int32_t n,
                                    float res_rows, float res_cols, float* net_out, uint32_t* min_key, void* stream_) {
  if (!flow || !logits || !net_out || batch < 1 || h < 1 || w < 1 || n < 1) return SLIMB200_E_INVALID;
  if (reinterpret_cast<uintptr_t>(net_out) & 15) return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (min_key) SLIMB200_CUDA_TRY(cudaMemsetAsync(min_key, 0xff, sizeof(uint32_t), stream));
  if ((long long)batch * h * n > 0x7fffffffLL) return SLIMB200_E_UNSUPPORTED;
  const int rows = n >= 8 ? 8 : (n >= 4 ? 4 : (n >= 2 ? 2 : 1));  // rows per thread, <= n
  dim3 grid((w * n + RO_THREADS - 1) / RO_THREADS, batch * ((h * n + rows - 1) / rows));
#define SLIMB200_RO_LAUNCH(R)                                                                                        \
  SLIMB200_LAUNCH(SLIMB200_K_RAFT_OUTPUT, stream,                                                                    \
                  (k_raft_output<R><<<grid, RO_THREADS, 0, stream>>>(flow, logits, batch, h, w, n, res_rows, res_cols, \
                                                                      net_out, min_key)))
  if (rows == 8) {
    SLIMB200_RO_LAUNCH(8);
  } else if (rows == 4) {
    SLIMB200_RO_LAUNCH(4);
  } else if (rows == 2) {
    SLIMB200_RO_LAUNCH(2);
  } else {
    SLIMB200_RO_LAUNCH(1);
  }
#undef SLIMB200_RO_LAUNCH
  return SLIMB200_OK;
}

extern "C" int slimb200_head_decode(const float* net_out, const uint32_t* logit_min_key, const uint8_t* filled, const float* pc,
                                    const int32_t* coors, const uint8_t* valid, const float* dyn_threshold,
                                    const slimb200_decode_params* p,
                                    float* bev, float* bev_aggr, uint8_t* bev_classes, float* points, double* trafo,
                                    uint8_t* not_enough, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!net_out || !filled || !dyn_threshold || !p || !bev || !bev_classes || !workspace) return SLIMB200_E_INVALID;
  if (p->batch < 1 || p->H < 1 || p->W < 1 || p->n_points < 0 || p->final_scale < 1) return SLIMB200_E_INVALID;
  if (p->n_points > 0 && (!pc || !coors || !valid || !points || p->pc_stride < 3)) return SLIMB200_E_INVALID;
  if (p->static_aggregation && (!trafo || !not_enough || !bev_aggr)) return SLIMB200_E_INVALID;
  if (reinterpret_cast<uintptr_t>(bev_aggr) & 15) return SLIMB200_E_ALIGNMENT;
  if (workspace_bytes < slimb200_head_decode_workspace_bytes(p)) return SLIMB200_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(net_out) & 15) || (reinterpret_cast<uintptr_t>(bev) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255))
    return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DecodeArgs a{};
  a.net_out = net_out;
  a.filled = filled;
  a.pc = pc;
  a.coors = coors;
  a.valid = valid;
  a.thr = dyn_threshold;
  a.p = *p;
  a.bev = bev;
  a.bev_aggr = bev_aggr;
  a.bev_cls = bev_classes;
  a.pts = points;
  a.trafo = trafo;
  a.not_enough = not_enough;
  WorkspaceCarver w(workspace);
  a.min_key = w.take<unsigned>(64);
  a.blocks_per_sample = blocks_per_sample(p);
  a.partials = w.take<double>((size_t)p->batch * (a.blocks_per_sample + 1) * N_MOM);
  a.pt_weight = w.take<float>((size_t)p->batch * p->n_points + 1);

  const size_t n_cells = (size_t)p->batch * p->H * p->W;
  if (logit_min_key) {
    a.min_key = const_cast<unsigned*>(logit_min_key);  // already taken by slimb200_raft_output
  } else {
    SLIMB200_CUDA_TRY(cudaMemsetAsync(a.min_key, 0xff, sizeof(unsigned), stream));
    const unsigned blocks = (unsigned)((n_cells + 255) / 256 < 148 * 8 ? (n_cells + 255) / 256 : 148 * 8);
    SLIMB200_LAUNCH(SLIMB200_K_DECODE_MIN, stream, (k_decode_min<<<blocks, 256, 0, stream>>>(net_out, n_cells, a.min_key)));
  }
  SLIMB200_LAUNCH(SLIMB200_K_DECODE_BEV, stream, (k_decode_bev<<<(unsigned)((n_cells + 255) / 256), 256, 0, stream>>>(a)));
  if (p->n_points > 0) {
    dim3 g((p->n_points + PT_THREADS - 1) / PT_THREADS, p->batch);
    SLIMB200_LAUNCH(SLIMB200_K_DECODE_POINTS, stream, (k_decode_points<<<g, PT_THREADS, 0, stream>>>(a)));
  }
  if (p->static_aggregation) {
    if (p->n_points > 0) {
      dim3 g(a.blocks_per_sample, p->batch);
      SLIMB200_LAUNCH(SLIMB200_K_KABSCH_MOMENTS, stream, (k_kabsch_moments<<<g, PT_THREADS, 0, stream>>>(a)));
    }
    SLIMB200_LAUNCH(SLIMB200_K_KABSCH, stream, (k_kabsch_finalize<<<p->batch, 64, 0, stream>>>(a)));
    const unsigned gx = (unsigned)((p->W + AGGR_THREADS - 1) / AGGR_THREADS);
    const size_t n_pts = (size_t)p->batch * p->n_points;
    const int point_rows = (int)((n_pts + (size_t)gx * AGGR_THREADS - 1) / ((size_t)gx * AGGR_THREADS));
    const int row_groups = p->batch * ((p->H + AGGR_ROWS - 1) / AGGR_ROWS);
    dim3 g(gx, (unsigned)(row_groups + point_rows));
    SLIMB200_LAUNCH(SLIMB200_K_DECODE_AGGR, stream, (k_decode_aggr<<<g, AGGR_THREADS, 0, stream>>>(a, row_groups)));
  }
  return SLIMB200_OK;
}
