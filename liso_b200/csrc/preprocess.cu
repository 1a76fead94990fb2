// SURVEY 8(f).4 -- dataset-side pre-processing of raw LiDAR scans on the GPU, so that a scan goes host -> device once
// and everything the network and the decoder need is derived there.
//
// Replaces, per frame (paths relative to baurst/liso):
//   infer_ground_label_using_cone                       liso/datasets/torch_dataset_commons.py:133-146
//   remove_ground_points_from_sample                    torch_dataset_commons.py:1164-1184  (label | cone rule)
//   voxelize_sample + voxelize_pcl (a12)                torch_dataset_commons.py:975-987, datasets/nuscenes/analyse_boxes.py:6-26
//   pillarize_bev (keep the in-range points)            torch_dataset_commons.py:1140-1162
//   the padding of the collate function                 torch_dataset_commons.py:380-401 (NaN points, -1 coors, valid mask)
// Result: pcl_ta = points that are not ground AND inside the BEV / height range, in scan order (stable compaction),
// their pillar coordinates, the validity mask, padded to `cap` points per sample; counts stay on the device.
// (The network input "pcl_full_no_ground" needs no compaction at all: slimb200_pillar_encode applies the same
// ground rule to the raw scan, see slimb200_pillar_params.ground_filter.)
//
//   k_pre_count    predicate per point, per-256-point-block counts
//   k_pre_scan     one CTA per sample: exclusive scan of its block counts, sample total
//   k_pre_scatter  stable scatter by ballot rank + padding of the tail
#include "common.cuh"

namespace {

constexpr int PRE_BLOCK = 256;

struct PreArgs {
  const float* scans[SLIMB200_MAX_BATCH];
  const uint8_t* ground_in[SLIMB200_MAX_BATCH];
  int32_t n_pts[SLIMB200_MAX_BATCH];
  int32_t blk_off[SLIMB200_MAX_BATCH + 1];
  int32_t batch, cap;
  slimb200_preprocess_params p;
  int32_t* blk_cnt;
  int32_t* blk_base;
  float* pcl_ta;
  int32_t* coors;
  uint8_t* valid;
  int32_t* counts;
};

// ground rule in float32 like the reference environment evaluates it (float32 cloud, scalars cast to float32):
//   d = sqrt(x*x + y*y); ground = z < z_thr + tan(angle) * d
__device__ __forceinline__ bool is_ground_cone(const slimb200_preprocess_params& p, float x, float y, float z) {
  const float d = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
  return z < __fadd_rn(p.cone_z_threshold, __fmul_rn(p.cone_tan, d));
}

// a12: int32_trunc(((p + R/2) / R) * G) in float64 + strict height filter (same arithmetic as k_pillar_coors_f64)
__device__ __forceinline__ bool bev_coors(const slimb200_preprocess_params& p, float x, float y, float z, int& ix, int& iy) {
  const double cx = __dmul_rn(__ddiv_rn(__dadd_rn((double)x, 0.5 * p.range_x), p.range_x), (double)p.grid_x);
  const double cy = __dmul_rn(__ddiv_rn(__dadd_rn((double)y, 0.5 * p.range_y), p.range_y), (double)p.grid_y);
  const double cz = __dmul_rn(__ddiv_rn(__dadd_rn((double)z, 0.5 * 1000.0), 1000.0), 1.0);
  ix = __double2int_rz(cx);
  iy = __double2int_rz(cy);
  const int iz = __double2int_rz(cz);
  bool ok = ix >= 0 && iy >= 0 && iz >= 0 && ix < p.grid_x && iy < p.grid_y && iz < 1;
  return ok && (p.z_min < z) && (z < p.z_max);
}

__device__ __forceinline__ int sample_of(const PreArgs& a, int pb) {
  int b = 0;
  while (b + 1 < a.batch && pb >= a.blk_off[b + 1]) ++b;
  return b;
}

template <bool SCATTER>
__global__ void __launch_bounds__(PRE_BLOCK) k_pre(const PreArgs a) {
  __shared__ int s_warp[PRE_BLOCK / 32];
  const int pb = blockIdx.x;
  const int b = sample_of(a, pb);
  const int i = (pb - a.blk_off[b]) * PRE_BLOCK + threadIdx.x;
  const int c_in = a.p.c_in;
  bool keep = false;
  float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
  int ix = 0, iy = 0;
  if (i < a.n_pts[b]) {
    const float* q = a.scans[b] + (size_t)i * c_in;
    pt = c_in == 4 ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
    bool ground = is_ground_cone(a.p, pt.x, pt.y, pt.z);
    if (a.ground_in[b]) ground = ground || a.ground_in[b][i] != 0;
    keep = !ground && bev_coors(a.p, pt.x, pt.y, pt.z, ix, iy);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (!SCATTER) {
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < PRE_BLOCK / 32; ++w) s += s_warp[w];
      a.blk_cnt[pb] = s;
    }
  } else if (keep) {
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = a.blk_base[pb] + before + __popc(bal & ((1u << lane) - 1u));
    const size_t o = (size_t)b * a.cap + pos;
    float* dst = a.pcl_ta + o * c_in;
    if (c_in == 4) {
      *reinterpret_cast<float4*>(dst) = pt;
    } else {
      dst[0] = pt.x;
      dst[1] = pt.y;
      dst[2] = pt.z;
    }
    a.coors[o * 2] = ix;
    a.coors[o * 2 + 1] = iy;
    a.valid[o] = 1;
  }
}

// exclusive scan of the block counts of one sample (one CTA per sample), total -> counts[b]
__global__ void __launch_bounds__(1024) k_pre_scan(const PreArgs a) {
  __shared__ int s_tmp[33];
  const int b = blockIdx.x;
  const int n = a.blk_off[b + 1] - a.blk_off[b];
  const int32_t* in = a.blk_cnt + a.blk_off[b];
  int32_t* out = a.blk_base + a.blk_off[b];
  const int t = threadIdx.x;
  const int chunk = (n + 1023) / 1024;
  const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
  int sum = 0;
  for (int j = lo; j < hi; ++j) sum += in[j];
  int inc = sum;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += v;
  }
  if (lane == 31) s_tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = s_tmp[lane];
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += v;
    }
    s_tmp[lane] = winc - w;
    if (lane == 31) s_tmp[32] = winc;
  }
  __syncthreads();
  int run = s_tmp[warp] + inc - sum;
  for (int j = lo; j < hi; ++j) {
    const int v = in[j];
    out[j] = run;
    run += v;
  }
  if (t == 0) a.counts[b] = s_tmp[32];
}

// padding of the tail [count, cap): NaN points, -1 coordinates, valid = 0 (collate, torch_dataset_commons.py:380-401)
__global__ void __launch_bounds__(256) k_pre_pad(const PreArgs a) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= a.cap || j < a.counts[b]) return;
  const size_t o = (size_t)b * a.cap + j;
  const float nan = __int_as_float(0x7fc00000);
  for (int k = 0; k < a.p.c_in; ++k) a.pcl_ta[o * a.p.c_in + k] = nan;
  a.coors[o * 2] = -1;
  a.coors[o * 2 + 1] = -1;
  a.valid[o] = 0;
}

int make_args(const float* const* scans, const int32_t* n_points, int32_t batch, int32_t cap, const slimb200_preprocess_params* p,
              void* workspace, PreArgs* a, size_t* bytes) {
  if (!p || batch < 1 || batch > SLIMB200_MAX_BATCH || cap < 0) return SLIMB200_E_INVALID;
  if (p->c_in != 3 && p->c_in != 4) return SLIMB200_E_UNSUPPORTED;
  *a = PreArgs{};
  a->batch = batch;
  a->cap = cap;
  a->p = *p;
  int n_blk = 0;
  for (int b = 0; b < batch; ++b) {
    const int n = n_points ? n_points[b] : cap;
    if (n < 0 || n > cap) return SLIMB200_E_INVALID;
    a->scans[b] = scans ? scans[b] : nullptr;
    a->n_pts[b] = n;
    a->blk_off[b] = n_blk;
    n_blk += (n + PRE_BLOCK - 1) / PRE_BLOCK;
  }
  a->blk_off[batch] = n_blk;
  WorkspaceCarver w(workspace);
  a->blk_cnt = w.take<int32_t>((size_t)n_blk + 1);
  a->blk_base = w.take<int32_t>((size_t)n_blk + 1);
  *bytes = w.used();
  return SLIMB200_OK;
}

}  // namespace

extern "C" size_t slimb200_preprocess_workspace_bytes(int32_t batch, int32_t cap, const slimb200_preprocess_params* p) {
  PreArgs a;
  size_t bytes = 0;
  if (make_args(nullptr, nullptr, batch, cap, p, nullptr, &a, &bytes) != SLIMB200_OK) return 0;
  return bytes;
}

extern "C" int slimb200_preprocess_points(const float* const* scans, const uint8_t* const* ground_labels, const int32_t* n_points,
                                          int32_t batch, int32_t cap, const slimb200_preprocess_params* p, float* pcl_ta,
                                          int32_t* pillar_coors, uint8_t* valid, int32_t* counts, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  if (!scans || !n_points || !pcl_ta || !pillar_coors || !valid || !counts || !workspace) return SLIMB200_E_INVALID;
  PreArgs a;
  size_t bytes = 0;
  const int rc = make_args(scans, n_points, batch, cap, p, workspace, &a, &bytes);
  if (rc != SLIMB200_OK) return rc;
  if (bytes > workspace_bytes) return SLIMB200_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return SLIMB200_E_ALIGNMENT;
  for (int b = 0; b < batch; ++b) {
    if (n_points[b] > 0 && !scans[b]) return SLIMB200_E_INVALID;
    if (p->c_in == 4 && ((reinterpret_cast<uintptr_t>(scans[b]) & 15) || (reinterpret_cast<uintptr_t>(pcl_ta) & 15)))
      return SLIMB200_E_ALIGNMENT;
    a.ground_in[b] = ground_labels ? ground_labels[b] : nullptr;
  }
  a.pcl_ta = pcl_ta;
  a.coors = pillar_coors;
  a.valid = valid;
  a.counts = counts;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n_blk = a.blk_off[batch];
  if (n_blk > 0) SLIMB200_LAUNCH(SLIMB200_K_PRE_COUNT, stream, (k_pre<false><<<n_blk, PRE_BLOCK, 0, stream>>>(a)));
  SLIMB200_LAUNCH(SLIMB200_K_PRE_SCAN, stream, (k_pre_scan<<<batch, 1024, 0, stream>>>(a)));
  if (n_blk > 0) SLIMB200_LAUNCH(SLIMB200_K_PRE_SCATTER, stream, (k_pre<true><<<n_blk, PRE_BLOCK, 0, stream>>>(a)));
  if (cap > 0) SLIMB200_LAUNCH(SLIMB200_K_PRE_PAD, stream, (k_pre_pad<<<dim3((cap + 255) / 256, batch), 256, 0, stream>>>(a)));
  return SLIMB200_OK;
}
