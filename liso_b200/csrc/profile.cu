// Launch accounting + optional per-kernel CUDA-event timing for bench.py (roofline numbers are
// measured live on the launching stream, never under a profiler).
#include <vector>

#include "common.cuh"

namespace {
struct Rec {
  cudaEvent_t a, b;
  int id;
};
bool g_enabled = false;
std::vector<Rec> g_recs;
size_t g_used = 0;
long long g_count[SLIMB200_N_KERNELS] = {0};
constexpr size_t MAX_RECS = 16384;
}  // namespace

int slimb200_device_info(int* dev, int* n_sm) {
  static int sm_count[SLIMB200_MAX_DEVICES] = {0};
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess) return (int)e;
  if (*dev < 0 || *dev >= SLIMB200_MAX_DEVICES) return SLIMB200_E_UNSUPPORTED;
  if (sm_count[*dev] == 0) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, *dev);
    if (e != cudaSuccess) return (int)e;
    sm_count[*dev] = n;
  }
  *n_sm = sm_count[*dev];
  return 0;
}

void slimb200_prof_pre(int id, cudaStream_t s) {
  g_count[id]++;
  if (!g_enabled || g_used >= MAX_RECS) return;
  if (g_used == g_recs.size()) {
    Rec r;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    g_recs.push_back(r);
  }
  g_recs[g_used].id = id;
  cudaEventRecord(g_recs[g_used].a, s);
}

void slimb200_prof_post(int id, cudaStream_t s) {
  if (!g_enabled || g_used >= g_recs.size() || g_recs[g_used].id != id) return;
  cudaEventRecord(g_recs[g_used].b, s);
  g_used++;
}

extern "C" int slimb200_profile_begin(void) {
  g_used = 0;
  g_enabled = true;
  return SLIMB200_OK;
}

extern "C" int slimb200_profile_end(float* ms_total, int64_t* timed_launches) {
  g_enabled = false;
  if (!ms_total || !timed_launches) return SLIMB200_E_INVALID;
  for (int i = 0; i < SLIMB200_N_KERNELS; ++i) {
    ms_total[i] = 0.f;
    timed_launches[i] = 0;
  }
  for (size_t i = 0; i < g_used; ++i) {
    SLIMB200_CUDA_TRY(cudaEventSynchronize(g_recs[i].b));
    float ms = 0.f;
    SLIMB200_CUDA_TRY(cudaEventElapsedTime(&ms, g_recs[i].a, g_recs[i].b));
    ms_total[g_recs[i].id] += ms;
    timed_launches[g_recs[i].id]++;
  }
  g_used = 0;
  return SLIMB200_OK;
}

extern "C" int64_t slimb200_launch_count(int32_t kernel_id) {
  if (kernel_id < 0) {
    long long s = 0;
    for (int i = 0; i < SLIMB200_N_KERNELS; ++i) s += g_count[i];
    return s;
  }
  return kernel_id < SLIMB200_N_KERNELS ? g_count[kernel_id] : 0;
}

extern "C" const char* slimb200_kernel_name(int32_t kernel_id) {
  static const char* names[SLIMB200_N_KERNELS] = {
      "k_point_keys", "k_scan_local", "k_scan_global", "k_rank_scatter", "k_tile_encode_stats", "k_bn_finalize",
      "k_tile_encode", "k_pillar_nhwc", "k_nchw_to_pixel_major", "k_feat_pack", "k_corr_gemm_tcgen05", "k_corr_lookup", "k_pillar_coors_f64", "k_decode_min", "k_decode_bev", "k_decode_points", "k_kabsch_finalize",
      "k_decode_aggr", "k_raft_output", "k_pre_count", "k_pre_scan", "k_pre_scatter", "k_pre_pad", "k_kabsch_moments", "k_in_stats", "k_in_finalize", "k_in_apply",
      "k_nhwc_pack", "k_gru_gate_zr", "k_gru_gate_out", "k_iter_update", "k_add_relu", "k_lookup_conv_tf32", "k_lookup_conv_pack", "k_deflate_tables", "k_deflate_chunks", "k_deflate_scan", "k_deflate_gather", "k_ctx_split", "k_bias_relu_slice"};
  return (kernel_id >= 0 && kernel_id < SLIMB200_N_KERNELS) ? names[kernel_id] : "?";
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
