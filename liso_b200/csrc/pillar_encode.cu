// Stage 1 of the SLIM hot path on B200: points -> pillar indices -> PFN -> scatter-max canvas.
//
// Replaces (reference paths relative to baurst/liso):
//   mmcv.ops.Voxelization (hard voxelisation; semantics of
//     mmdetection3d/mmdet3d/core/voxel/voxel_generator.py:137-208, deterministic point order)
//   PointsPillarFeatureNetWrapper.voxelize          pcl_to_feature_grid.py:56-84
//   PillarFeatureNet.forward                        pillar_encoder.py:93-159
//   PFNLayer.forward                                voxel_encoders/utils.py:146-182
//   PointPillarsScatter.forward_batch (x2)          pillar_scatter.py:62-102
//
// Data flow (all samples of the batch in every launch):
//   k_point_keys     1 thread / point, 128-bit loads: fp32 floor((p-min)/vs) with true division,
//                    tile-major cell key, per-cell count (atomicAdd) and first point (atomicMax of
//                    the inverted index) in L2-resident int maps
//   k_scan_local     per BEV tile (4 x 32 cells) exclusive prefix of the counts + tile totals;
//                    per 256-point block: number of "first points" (one per occupied cell)
//   k_scan_global    single-CTA scans: tile starts, per-sample pillar ordinals bases, BN fold
//   k_rank_scatter   per point: pillar ordinal in first-appearance order (40000 cap), counting-sort
//                    scatter of (index, xyzi) into cell-contiguous order
//   k_tile_encode    one CTA per BEV tile: warp per occupied cell picks the 20 lowest point indices
//                    (bitonic network in registers), warp-level segmented reduction for the cluster
//                    centre, 10->64 linear + BN + ReLU + max in registers, 64 x tile staged in shared
//                    memory, then the whole tile (zeros included) is written with full 128-byte rows:
//                    canvas bytes are written exactly once and never read.
//
// HBM algorithmic bytes per frame: N*C*4 (points) + 64*H*W*4 (canvas) + H*W*4 (occupancy).
#include "common.cuh"

namespace {

constexpr int TILE_R = 4;     // tile rows (x index)
constexpr int TILE_C = 32;    // tile cols (y index) == one warp == one 128-byte row segment
constexpr int TILE_CELLS = TILE_R * TILE_C;
constexpr int PT_BLOCK = 256;
constexpr int ENC_THREADS = 256;
constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int MAX_COUT = 64;
constexpr int INV_BASE = 0x7fffffff;
constexpr int STATS_MAX_CTAS = 148 * 4;
constexpr int ENC_CTAS_PER_SM = 3;

struct PillarArgs {
  const float* pts[SLIMB200_MAX_BATCH];
  int32_t n_pts[SLIMB200_MAX_BATCH];
  int32_t pt_off[SLIMB200_MAX_BATCH + 1];   // global index of the first point of sample b
  int32_t blk_off[SLIMB200_MAX_BATCH + 1];  // first 256-point block of sample b
  int32_t batch;
  slimb200_pillar_params p;
  int32_t tiles_x, tiles_y, tiles_per_sample, n_tiles;
  // workspace
  int32_t* cell_count;      // [n_tiles * TILE_CELLS]  zeroed per call
  int32_t* cell_inv_first;  // [n_tiles * TILE_CELLS]  zeroed per call; INV_BASE - first point index
  int32_t* cell_fill;       // [n_tiles * TILE_CELLS]  zeroed per call; scatter cursor
  int32_t* cell_prefix;     // exclusive prefix of counts inside the tile
  int32_t* cell_ord;        // pillar ordinal inside the sample (valid where count > 0)
  int32_t* tile_total;      // [n_tiles]
  int32_t* tile_start;      // [n_tiles] pt_off[sample] + exclusive prefix inside the sample: position in the sorted arrays
  int32_t* pt_key;          // [total points] cell key or -1
  int32_t* sorted_idx;      // [total points] per-sample point index, cell-contiguous
  float4* sorted_pts;       // [total points] xyzi, same order
  int32_t* blk_cnt;         // [total point blocks] first-point flags per block
  int32_t* blk_base;        // [total point blocks] exclusive prefix inside the sample
  int32_t* pillar_base;     // [batch + 1] exclusive prefix of kept pillars
  int32_t* sample_kept;     // [batch] kept pillars per sample
  int32_t* ticket;          // zeroed per call: how many per-sample scans have finished
  int4* pillar_info;        // [kept pillars] (cell key, first sorted point, point count, b << 24 | xi << 12 | yi), first-appearance order
  float* bn_ab;             // [2][MAX_COUT] alpha, beta' of the folded BatchNorm
  double* stat_partials;    // [STATS_MAX_CTAS][MAX_COUT][2]
  // parameters / outputs
  const float* linear_weight;
  const float* bn_weight;
  const float* bn_bias;
  float* bn_mean;
  float* bn_var;
  float* canvas;
  float* occupancy;
  int32_t* pillar_counts;
  int32_t* coors_out;
  int32_t* num_points_out;
  float* voxels_out;
  int32_t* pt2pillar_out;
};

__device__ __forceinline__ int sample_of_block(const PillarArgs& a, int pb) {
  int b = 0;
  while (b + 1 < a.batch && pb >= a.blk_off[b + 1]) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------
// fp32 pillar coordinate of one point; mirrors voxel_generator.py:186-191 / mmcv's
// dynamic_voxelize_kernel: c = floor((p - min) / voxel) with IEEE subtraction and division.
// NaN / inf coordinates are rejected (the comparison is written so that NaN fails).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pillar_cell(const slimb200_pillar_params& p, float x, float y, float z,
                                            int& xi, int& yi) {
  const float cx = floorf(__fdiv_rn(__fsub_rn(x, p.range_min[0]), p.voxel_size[0]));
  const float cy = floorf(__fdiv_rn(__fsub_rn(y, p.range_min[1]), p.voxel_size[1]));
  const float cz = floorf(__fdiv_rn(__fsub_rn(z, p.range_min[2]), p.voxel_size[2]));
  const bool ok = (cx >= 0.f && cx < (float)p.grid[0]) && (cy >= 0.f && cy < (float)p.grid[1]) &&
                  (cz >= 0.f && cz < (float)p.grid[2]);
  xi = ok ? (int)cx : 0;
  yi = ok ? (int)cy : 0;
  return ok;
}

__device__ __forceinline__ float4 load_point(const float* pts, int i, int c_in) {
  if (c_in == 4) return __ldg(reinterpret_cast<const float4*>(pts) + i);
  const float* q = pts + (size_t)i * 3;
  return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
}

__global__ void __launch_bounds__(PT_BLOCK) k_point_keys(const PillarArgs a) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * PT_BLOCK + threadIdx.x;
  if (i >= a.n_pts[b]) return;
  const float4 pt = load_point(a.pts[b], i, a.p.c_in);
  int xi, yi;
  bool ok = pillar_cell(a.p, pt.x, pt.y, pt.z, xi, yi);
  if (a.p.ground_filter) {
    // raw scan in: drop what the reference's dataset drops (cone rule in float32, torch_dataset_commons.py:133-146)
    const float d = sqrtf(__fadd_rn(__fmul_rn(pt.x, pt.x), __fmul_rn(pt.y, pt.y)));
    ok = ok && !(pt.z < __fadd_rn(a.p.ground_cone_z, __fmul_rn(a.p.ground_cone_tan, d)));
  }
  int key = -1;
  if (ok) {
    const int tile = (b * a.tiles_x + xi / TILE_R) * a.tiles_y + yi / TILE_C;
    key = tile * TILE_CELLS + (xi % TILE_R) * TILE_C + (yi % TILE_C);
    atomicAdd(a.cell_count + key, 1);
    atomicMax(a.cell_inv_first + key, INV_BASE - i);
  }
  a.pt_key[a.pt_off[b] + i] = key;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scan_local(const PillarArgs a, int n_cell_blocks) {
  const int lane = lane_id(), warp = warp_id();
  if ((int)blockIdx.x < n_cell_blocks) {
    // one warp per tile (128 cells), four consecutive cells per lane: one 16-byte load, one warp scan, one 16-byte store
    const int tile = blockIdx.x * 8 + warp;
    if (tile < a.n_tiles) {
      const int4 c = *reinterpret_cast<const int4*>(a.cell_count + (size_t)tile * TILE_CELLS + lane * 4);
      const int sum = c.x + c.y + c.z + c.w;
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
      }
      const int ex = inc - sum;
      *reinterpret_cast<int4*>(a.cell_prefix + (size_t)tile * TILE_CELLS + lane * 4) = make_int4(ex, ex + c.x, ex + c.x + c.y, ex + c.x + c.y + c.z);
      if (lane == 31) a.tile_total[tile] = inc;
    }
  } else {
    const int pb = blockIdx.x - n_cell_blocks;
    const int b = sample_of_block(a, pb);
    const int i = (pb - a.blk_off[b]) * PT_BLOCK + threadIdx.x;
    bool flag = false;
    if (i < a.n_pts[b]) {
      const int key = a.pt_key[a.pt_off[b] + i];
      flag = key >= 0 && a.cell_inv_first[key] == INV_BASE - i;
    }
    const int c = __syncthreads_count(flag);
    if (threadIdx.x == 0) a.blk_cnt[pb] = c;
  }
}

// exclusive scan of in[0..n) by one 1024-thread CTA; returns the total to every thread
__device__ int block_exclusive_scan_1024(const int32_t* in, int32_t* out, int n, int* s_tmp /*[33]*/) {
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
  __syncthreads();  // protect s_tmp from the previous call
  if (lane == 31) s_tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_tmp[lane];
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
  return s_tmp[32];
}

// One CTA per independent scan (round 2; a single CTA used to walk through the samples one after the other):
//   CTA 0                  eval-mode BatchNorm fold
//   CTA 1 + b              tile starts of sample b: exclusive prefix of its tile totals, based at pt_off[b] -- the sample's
//                          share of the sorted arrays (host-known upper bound of its in-range points), so that no scan
//                          crosses a sample
//   CTA 1 + batch + b      pillar-ordinal bases of sample b's 256-point blocks; the LAST of these CTAs to finish (ticket)
//                          turns the per-sample pillar counts (capped at max_voxels) into pillar_base
__global__ void __launch_bounds__(1024) k_scan_global(const PillarArgs a) {
  __shared__ int s_tmp[33];
  __shared__ int s_last;
  if (blockIdx.x == 0) {
    // eval-mode BatchNorm folded the way ATen does: alpha = gamma / sqrt(var + eps),
    // beta' = beta - mean * alpha, y = x * alpha + beta'
    const int c = threadIdx.x;
    if (!a.p.bn_training && c < a.p.c_out) {
      const float invstd = 1.0f / sqrtf(a.bn_var[c] + a.p.bn_eps);
      const float alpha = a.bn_weight[c] * invstd;
      a.bn_ab[c] = alpha;
      a.bn_ab[MAX_COUT + c] = a.bn_bias[c] - a.bn_mean[c] * alpha;
    }
  } else if ((int)blockIdx.x <= a.batch) {
    const int b = blockIdx.x - 1, t0 = b * a.tiles_per_sample;
    block_exclusive_scan_1024(a.tile_total + t0, a.tile_start + t0, a.tiles_per_sample, s_tmp);
    __syncthreads();
    const int base = a.pt_off[b];
    for (int j = threadIdx.x; j < a.tiles_per_sample; j += 1024) a.tile_start[t0 + j] += base;
  } else {
    const int b = blockIdx.x - 1 - a.batch;
    const int nb = a.blk_off[b + 1] - a.blk_off[b];
    const int total = block_exclusive_scan_1024(a.blk_cnt + a.blk_off[b], a.blk_base + a.blk_off[b], nb, s_tmp);
    if (threadIdx.x == 0) {
      a.sample_kept[b] = min(total, a.p.max_voxels);
      __threadfence();
      s_last = atomicAdd(a.ticket, 1) == a.batch - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      __threadfence();
      int base = 0;
      for (int i = 0; i < a.batch; ++i) {
        a.pillar_base[i] = base;
        if (a.pillar_counts) a.pillar_counts[i] = base;
        base += __ldcg(a.sample_kept + i);
      }
      a.pillar_base[a.batch] = base;
      if (a.pillar_counts) a.pillar_counts[a.batch] = base;
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_BLOCK) k_rank_scatter(const PillarArgs a) {
  __shared__ int s_warp[PT_BLOCK / 32];
  const int pb = blockIdx.x;
  const int b = sample_of_block(a, pb);
  const int i = (pb - a.blk_off[b]) * PT_BLOCK + threadIdx.x;
  const bool live = i < a.n_pts[b];
  int key = -1;
  bool flag = false;
  if (live) {
    key = a.pt_key[a.pt_off[b] + i];
    flag = key >= 0 && a.cell_inv_first[key] == INV_BASE - i;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int lane = lane_id(), warp = warp_id();
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  if (flag) {
    const int ord = a.blk_base[pb] + before + __popc(bal & ((1u << lane) - 1u));
    a.cell_ord[key] = ord;
    if (ord >= a.p.max_voxels) {
      a.cell_count[key] = -a.cell_count[key];  // over the pillar cap: dropped, reads as "not kept" (count <= 0) downstream
    } else {
      const int tile = key / TILE_CELLS, cl = key - tile * TILE_CELLS;
      const int tl = tile - b * a.tiles_per_sample;
      const int tx = tl / a.tiles_y, ty = tl - tx * a.tiles_y;
      const int xi = tx * TILE_R + cl / TILE_C, yi = ty * TILE_C + (cl % TILE_C);
      a.pillar_info[a.pillar_base[b] + ord] =
          make_int4(key, a.tile_start[tile] + a.cell_prefix[key], a.cell_count[key], (b << 24) | (xi << 12) | yi);
    }
  }
  if (live && a.pt2pillar_out) a.pt2pillar_out[a.pt_off[b] + i] = -1;
  if (key >= 0) {
    const int slot = atomicAdd(a.cell_fill + key, 1);
    const int pos = a.tile_start[key / TILE_CELLS] + a.cell_prefix[key] + slot;
    a.sorted_idx[pos] = i;
    a.sorted_pts[pos] = load_point(a.pts[b], i, a.p.c_in);
  }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long bitonic_sort_warp(unsigned long long v) {
  const int lane = lane_id();
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = (lane & k) == 0;
      const bool lower = (lane & j) == 0;
      const unsigned long long mn = v < o ? v : o, mx = v < o ? o : v;
      v = (lower == up) ? mn : mx;
    }
  }
  return v;
}

// Rank the (<= 32) candidate points of one cell by point index without a sorting network: two
// counting passes of n shuffles each.  Returns, for destination lane r < n, the source lane holding
// the r-th smallest index (keys are unique).
__device__ __forceinline__ int rank_order_small(unsigned idx, int n, int lane) {
  int rank = 0;
  for (int j = 0; j < n; ++j) rank += (__shfl_sync(0xffffffffu, idx, j) < idx) ? 1 : 0;
  if (lane >= n) rank = 32;
  int src = 0;
  for (int j = 0; j < n; ++j)
    if (__shfl_sync(0xffffffffu, rank, j) == lane) src = j;
  return src;
}

constexpr int STAGE_PITCH = TILE_CELLS + 4;  // multiple of 4 floats: 128-bit reads of the stage

template <int MODE>  // 0: encode + write canvas, 1: BatchNorm batch statistics only
__global__ void __launch_bounds__(ENC_THREADS, 3) k_tile_encode(const PillarArgs a) {
  __shared__ __align__(16) float s_stage[MODE == 0 ? MAX_COUT : 1][STAGE_PITCH];
  __shared__ int s_cnt[TILE_CELLS], s_start[TILE_CELLS], s_ord[TILE_CELLS];
  __shared__ unsigned s_mask[TILE_R];
  __shared__ unsigned char s_list[TILE_CELLS];
  __shared__ double s_red[MODE == 1 ? ENC_WARPS * MAX_COUT * 2 : 1];

  const int lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const slimb200_pillar_params& p = a.p;
  const int G0 = p.grid[0], G1 = p.grid[1];
  const int c_out = p.c_out;
  const int max_pts = p.max_points;
  const bool vec_ok = (G1 & 3) == 0;  // rows are 16-byte aligned: 128-bit stores

  // per-lane slice of the PFN weights (loaded once per persistent CTA): channels lane and lane + 32,
  // canonical 10 columns [xc yc zc i | dx dy dz | xc yc zc]; a 3-channel cloud has no intensity column.
  float W[2][10], alpha[2], betap[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c = lane + 32 * q;
    const bool cv = c < c_out;
    const int cf = p.c_in + 6;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      int col = j;
      if (p.c_in == 3) col = j < 3 ? j : (j == 3 ? -1 : j - 1);
      W[q][j] = (cv && col >= 0) ? __ldg(a.linear_weight + c * cf + col) : 0.f;
    }
    alpha[q] = (cv && MODE == 0) ? a.bn_ab[c] : 0.f;
    betap[q] = (cv && MODE == 0) ? a.bn_ab[MAX_COUT + c] : 0.f;
  }
  double acc1[2] = {0.0, 0.0}, acc2[2] = {0.0, 0.0};

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_sample;
    const int tl = tile - b * a.tiles_per_sample;
    const int tx = tl / a.tiles_y, ty = tl - tx * a.tiles_y;
    const int total = a.tile_total[tile];
    if (total != 0) {
      if (tid < TILE_CELLS) {
        const int cell = tile * TILE_CELLS + tid;
        const int cnt = a.cell_count[cell];
        const int ord = cnt > 0 ? a.cell_ord[cell] : 0;
        const bool kept = cnt > 0 && ord < p.max_voxels;
        const unsigned m = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) s_mask[warp] = m;
        s_cnt[tid] = cnt;
        s_start[tid] = a.tile_start[tile] + a.cell_prefix[cell];
        s_ord[tid] = ord;
      }
      __syncthreads();
      int nocc = 0;
      {
        int before = 0;
#pragma unroll
        for (int r = 0; r < TILE_R; ++r) {
          const int pc = __popc(s_mask[r]);
          if (r < warp) before += pc;
          nocc += pc;
        }
        if (tid < TILE_CELLS && ((s_mask[warp] >> lane) & 1u))
          s_list[before + __popc(s_mask[warp] & ((1u << lane) - 1u))] = (unsigned char)tid;
      }
      __syncthreads();

      for (int it = warp; it < nocc; it += ENC_WARPS) {
        const int cl = s_list[it];
        const int n_all = s_cnt[cl], st = s_start[cl];
        const int xi = tx * TILE_R + cl / TILE_C, yi = ty * TILE_C + (cl % TILE_C);
        // --- the max_points lowest point indices of the cell, ascending ------------------
        float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
        int pidx = 0;
        const int n = n_all < max_pts ? n_all : max_pts;
        if (n_all <= 32) {
          unsigned idx = 0xffffffffu;
          if (lane < n_all) {
            idx = (unsigned)a.sorted_idx[st + lane];
            pt = a.sorted_pts[st + lane];
          }
          const int src = rank_order_small(idx, n_all, lane);
          pidx = (int)__shfl_sync(0xffffffffu, idx, src);
          pt.x = __shfl_sync(0xffffffffu, pt.x, src);
          pt.y = __shfl_sync(0xffffffffu, pt.y, src);
          pt.z = __shfl_sync(0xffffffffu, pt.z, src);
          pt.w = __shfl_sync(0xffffffffu, pt.w, src);
        } else {
          // heavy cell: stream the candidates through a bitonic network, keep the lowest max_points
          unsigned long long key = ((unsigned long long)(unsigned)a.sorted_idx[st + lane] << 32) | (unsigned)lane;
          key = bitonic_sort_warp(key);
          const int keep = max_pts;  // <= 24: lanes [keep, 32) take new candidates
          for (int base = 32; base < n_all; base += 32 - keep) {
            if (lane >= keep) {
              const int j = base + lane - keep;
              key = j < n_all ? (((unsigned long long)(unsigned)a.sorted_idx[st + j] << 32) | (unsigned)j) : ~0ull;
            }
            key = bitonic_sort_warp(key);
          }
          pidx = (int)(unsigned)(key >> 32);
          if (lane < n) pt = a.sorted_pts[st + (int)(unsigned)(key & 0xffffffffull)];
        }
        const bool act = lane < n;
        // --- cluster centre: sum over the slots / num_points (pillar_encoder.py:108-113) ---
        float sx = act ? pt.x : 0.f, sy = act ? pt.y : 0.f, sz = act ? pt.z : 0.f;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          sx += __shfl_xor_sync(0xffffffffu, sx, d);
          sy += __shfl_xor_sync(0xffffffffu, sy, d);
          sz += __shfl_xor_sync(0xffffffffu, sz, d);
        }
        const float fn = (float)n;
        const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
        // --- voxel-centre offsets, legacy aliasing + swapped index (pillar_encoder.py:129-139):
        //     x uses coors[:,3] (= y index), y uses coors[:,2] (= x index), z index is 0
        float f[7];
        f[0] = __fsub_rn(pt.x, __fadd_rn(__fmul_rn((float)yi, p.vx), p.x_offset));
        f[1] = __fsub_rn(pt.y, __fadd_rn(__fmul_rn((float)xi, p.vy), p.y_offset));
        f[2] = __fsub_rn(pt.z, __fadd_rn(__fmul_rn(0.f, p.vz), p.z_offset));
        f[3] = pt.w;
        f[4] = __fsub_rn(pt.x, mx);
        f[5] = __fsub_rn(pt.y, my);
        f[6] = __fsub_rn(pt.z, mz);
        // --- Linear(10->64) + BN + ReLU + max over the slots, 2 channels per lane ------------
        float best[2] = {0.f, 0.f};
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
        for (int k = 0; k < n; ++k) {
          float g[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) g[j] = __shfl_sync(0xffffffffu, f[j], k);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float x = W[q][0] * g[0];
            x = fmaf(W[q][1], g[1], x);
            x = fmaf(W[q][2], g[2], x);
            x = fmaf(W[q][3], g[3], x);
            x = fmaf(W[q][4], g[4], x);
            x = fmaf(W[q][5], g[5], x);
            x = fmaf(W[q][6], g[6], x);
            x = fmaf(W[q][7], g[0], x);
            x = fmaf(W[q][8], g[1], x);
            x = fmaf(W[q][9], g[2], x);
            if (MODE == 0) {
              best[q] = fmaxf(best[q], fmaf(x, alpha[q], betap[q]));  // ReLU folded: best starts at 0
            } else {
              s1[q] += x;
              s2[q] = fmaf(x, x, s2[q]);
            }
          }
        }
        if (MODE == 0) {
          // padded (all-zero) rows take part in the max: BN(0) = beta' (voxel_encoders/utils.py:161-169)
          if (n < max_pts) {
            best[0] = fmaxf(best[0], betap[0]);
            best[1] = fmaxf(best[1], betap[1]);
          }
          s_stage[lane % MAX_COUT][cl] = best[0];
          if (c_out > 32) s_stage[(lane + 32) % MAX_COUT][cl] = best[1];
          if (a.coors_out || a.num_points_out || a.voxels_out || a.pt2pillar_out) {
            const int row = a.pillar_base[b] + s_ord[cl];
            if (lane == 0 && a.coors_out) {
              int4 cc = make_int4(b, 0, xi, yi);
              *reinterpret_cast<int4*>(a.coors_out + (size_t)row * 4) = cc;
            }
            if (lane == 0 && a.num_points_out) a.num_points_out[row] = n;
            if (a.pt2pillar_out && act) a.pt2pillar_out[a.pt_off[b] + pidx] = row;
            if (a.voxels_out && lane < max_pts) {
              float* v = a.voxels_out + ((size_t)row * max_pts + lane) * p.c_in;
              v[0] = act ? pt.x : 0.f;
              v[1] = act ? pt.y : 0.f;
              v[2] = act ? pt.z : 0.f;
              if (p.c_in == 4) v[3] = act ? pt.w : 0.f;
            }
          }
        } else {
          acc1[0] += (double)s1[0];
          acc1[1] += (double)s1[1];
          acc2[0] += (double)s2[0];
          acc2[1] += (double)s2[1];
        }
      }
      __syncthreads();
    }

    if (MODE == 0) {
      // ---- write the whole tile, zeros included; every store instruction covers full 128-byte rows ----
      const size_t plane = (size_t)G0 * G1;
      float* const cbase = a.canvas + (size_t)b * c_out * plane;
      float* const obase = a.occupancy + (size_t)b * plane;
      if (vec_ok) {
        // 8 lanes x 16 B = one row segment; a warp instruction = the 4 rows of one channel
        const int r = (tid >> 3) & 3, j = tid & 7;
        const int xi = tx * TILE_R + r, yi = ty * TILE_C + 4 * j;
        const bool ok = xi < G0 && yi < G1;
        const unsigned bits = total != 0 ? (s_mask[r] >> (4 * j)) & 15u : 0u;
        const size_t off = (size_t)xi * G1 + yi;
        if (ok) {
          for (int c = tid >> 5; c < c_out; c += ENC_WARPS) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bits) {
              v = *reinterpret_cast<const float4*>(&s_stage[c][r * TILE_C + 4 * j]);
              v.x = (bits & 1u) ? v.x : 0.f;
              v.y = (bits & 2u) ? v.y : 0.f;
              v.z = (bits & 4u) ? v.z : 0.f;
              v.w = (bits & 8u) ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(cbase + c * plane + off) = v;
          }
          if ((tid >> 5) == ENC_WARPS - 1) {
            const float4 o = make_float4((bits & 1u) ? 1.f : 0.f, (bits & 2u) ? 1.f : 0.f, (bits & 4u) ? 1.f : 0.f,
                                         (bits & 8u) ? 1.f : 0.f);
            *reinterpret_cast<float4*>(obase + off) = o;
          }
        }
      } else {
        const int yi = ty * TILE_C + lane;
        const bool col_ok = yi < G1;
        for (int it = warp; it < (c_out + 1) * TILE_R; it += ENC_WARPS) {
          const int c = it / TILE_R, r = it - c * TILE_R;
          const int xi = tx * TILE_R + r;
          if (xi >= G0 || !col_ok) continue;
          const bool occ = total != 0 && ((s_mask[r] >> lane) & 1u);
          if (c < c_out) {
            cbase[c * plane + (size_t)xi * G1 + yi] = occ ? s_stage[c][r * TILE_C + lane] : 0.f;
          } else {
            obase[(size_t)xi * G1 + yi] = occ ? 1.f : 0.f;
          }
        }
      }
      __syncthreads();
    }
  }

  if (MODE == 1) {
    // per-CTA partial sums in double, fixed order -> deterministic batch statistics
    for (int q = 0; q < 2; ++q) {
      s_red[(warp * MAX_COUT + lane + 32 * q) * 2 + 0] = acc1[q];
      s_red[(warp * MAX_COUT + lane + 32 * q) * 2 + 1] = acc2[q];
    }
    __syncthreads();
    if (tid < MAX_COUT * 2) {
      double s = 0.0;
      for (int w = 0; w < ENC_WARPS; ++w) s += s_red[w * MAX_COUT * 2 + tid];
      a.stat_partials[(size_t)blockIdx.x * MAX_COUT * 2 + tid] = s;
    }
  }
}


// ------------------------------------------------------------------------------------------
// Channels-last canvas (canvas_layout = NHWC): a cell is c_out contiguous floats, so an occupied pillar is two
// coalesced 128-byte stores straight from the registers that hold its 64 maxima and the empty cells are a
// contiguous zero stream.  Every warp of the persistent grid interleaves two independent work lists, so the
// (latency-bound) pillar arithmetic runs under the (bandwidth-bound) zero fill and both are perfectly balanced
// whatever the spatial distribution of the points:
//   pillars  a.pillar_info[0 .. n_pillars): rank <= 32 candidates by point index (20 lowest kept), cluster mean by
//            warp reduction, 10 -> 64 linear + folded BN + ReLU + max in registers, 2 channels per lane
//   chunks   32 consecutive cells of a canvas row (32 * c_out * 4 contiguous bytes) each, enumerated row-major so
//            that consecutive warps stream consecutive pieces: the occupancy ballot of the chunk is cut into
//            maximal runs of empty cells and every run is ONE bulk copy (cp.async.bulk / UBLKCP) from an 8 KB
//            zero buffer in shared memory -- no store instructions, no registers; plus the 128-byte occupancy row
// Both lists are software-pipelined (next chunk's counts, next pillar's points, the descriptor after next).
// Every canvas byte is written exactly once and never read.
// ------------------------------------------------------------------------------------------
constexpr int NH_THREADS = 256;
constexpr int NH_WARPS = NH_THREADS / 32;
constexpr int NH_CTAS_PER_SM = 3;

// 32-bit bitonic network over the first `width` (power of two) lanes; keys unique
__device__ __forceinline__ unsigned bitonic_sort_u32(unsigned v, int width, int lane) {
  for (int k = 2; k <= width; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
      v = take_min ? min(v, o) : max(v, o);
    }
  }
  return v;
}

__device__ __forceinline__ void bulk_store_zero(void* dst, uint32_t zero_smem, int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(zero_smem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(NH_THREADS, NH_CTAS_PER_SM) k_pillar_nhwc(const PillarArgs a) {
  // per-warp staging: the (<= 24) kept points of the current pillar in slot order, then their 7 features
  __shared__ __align__(16) float s_feat[NH_WARPS][24][8];
  __shared__ __align__(128) float4 s_zero[TILE_C * MAX_COUT / 4];  // one chunk of zeros (8 KB), source of the bulk stores
  const int lane = lane_id(), warp = warp_id();
  for (int i = threadIdx.x; i < TILE_C * MAX_COUT / 4; i += NH_THREADS) s_zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zero_smem = (uint32_t)__cvta_generic_to_shared(s_zero);
  const slimb200_pillar_params& p = a.p;
  const int G0 = p.grid[0], G1 = p.grid[1];
  const int c_out = p.c_out;
  const int max_pts = p.max_points;
  const int gw = blockIdx.x * NH_WARPS + warp;
  const int tw = gridDim.x * NH_WARPS;

  // Linear(10 -> 64) with the legacy aliasing folded in: input columns 7..9 repeat columns 0..2
  // (pillar_encoder.py:129-147), so x = sum_{j<3} (W_j + W_{j+7}) g_j + sum_{3<=j<7} W_j g_j: 7 FMAs per channel.
  float Wc[2][7], alpha[2], betap[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c = lane + 32 * q;
    const bool cv = c < c_out;
    const int cf = p.c_in + 6;
    const float* wr = a.linear_weight + c * cf;
    const int sh = p.c_in == 3 ? 1 : 0;  // a 3-channel cloud has no intensity column
#pragma unroll
    for (int j = 0; j < 3; ++j) Wc[q][j] = cv ? __ldg(wr + j) + __ldg(wr + 7 + j - sh) : 0.f;
    Wc[q][3] = (cv && !sh) ? __ldg(wr + 3) : 0.f;
#pragma unroll
    for (int j = 4; j < 7; ++j) Wc[q][j] = cv ? __ldg(wr + j - sh) : 0.f;
    alpha[q] = cv ? a.bn_ab[c] : 0.f;
    betap[q] = cv ? a.bn_ab[MAX_COUT + c] : 0.f;
  }

  const int n_pillars = a.pillar_base[a.batch];
  const int n_chunks = a.n_tiles * TILE_R;  // = batch * rows_pad * tiles_y
  const int my_chunks = gw < n_chunks ? (n_chunks - gw + tw - 1) / tw : 0;
  const int my_pillars = gw < n_pillars ? (n_pillars - gw + tw - 1) / tw : 0;
  const int q4_log = 31 - __clz(c_out >> 2);  // float4 per cell = c_out / 4, a power of two
  int ci = 0, pi = 0;
  const bool want_extra = a.coors_out || a.num_points_out || a.voxels_out || a.pt2pillar_out;
  // software pipeline (the kernel is bound by load latency otherwise): the occupancy counts of the next chunk,
  // the descriptor of the pillar after next and the points of the next pillar are always in flight
  // chunks are enumerated row-major (sample, row, column tile), so consecutive warps stream consecutive 8 KB pieces
  const int rows_pad = a.tiles_x * TILE_R;
  const int d_ty = tw % a.tiles_y, d_rid = tw / a.tiles_y;
  const int d_xi = d_rid % rows_pad, d_b = d_rid / rows_pad;
  int zy = gw % a.tiles_y, zx = (gw / a.tiles_y) % rows_pad, zb = (gw / a.tiles_y) / rows_pad;
  int cnt_next = my_chunks > 0
                     ? a.cell_count[((zb * a.tiles_x + (zx >> 2)) * a.tiles_y + zy) * TILE_CELLS + (zx & 3) * TILE_C + lane]
                     : 0;
  int4 info = my_pillars > 0 ? a.pillar_info[gw] : make_int4(0, 0, 0, 0);
  int4 info_next = my_pillars > 1 ? a.pillar_info[gw + tw] : make_int4(0, 0, 0, 0);
  unsigned idx_cur = 0xffffffffu;
  float4 pt_cur = make_float4(0.f, 0.f, 0.f, 0.f);
  if (my_pillars > 0 && lane < info.z) {
    idx_cur = (unsigned)a.sorted_idx[info.y + lane];
    pt_cur = a.sorted_pts[info.y + lane];
  }

  while (ci < my_chunks || pi < my_pillars) {
    const bool do_chunk = pi >= my_pillars || (ci < my_chunks && (long long)ci * my_pillars <= (long long)pi * my_chunks);
    if (do_chunk) {
      // ---------------- zero fill of one tile row (chunk) ----------------
      ++ci;
      const int b = zb, xi = zx, yi0 = zy * TILE_C;
      const int cnt = cnt_next;
      if (ci < my_chunks) {
        // advance (sample, row, column tile) by tw chunks without dividing, prefetch its occupancy counts
        zy += d_ty;
        if (zy >= a.tiles_y) { zy -= a.tiles_y; ++zx; }
        zx += d_xi;
        if (zx >= rows_pad) { zx -= rows_pad; ++zb; }
        zb += d_b;
        cnt_next = a.cell_count[((zb * a.tiles_x + (zx >> 2)) * a.tiles_y + zy) * TILE_CELLS + (zx & 3) * TILE_C + lane];
      }
      const bool kept = cnt > 0;  // cells over the pillar cap carry a negative count (k_rank_scatter)
      const unsigned mask = __ballot_sync(0xffffffffu, kept);
      if (xi < G0) {
        const size_t row = ((size_t)b * G0 + xi) * G1 + yi0;
        if (yi0 + lane < G1) a.occupancy[row + lane] = kept ? 1.f : 0.f;
        if (lane == 0) {
          // every maximal run of empty cells is one bulk copy (UBLKCP) of zeros from shared memory
          const int n_cols = min(TILE_C, G1 - yi0);
          unsigned empty = ~mask & (n_cols == 32 ? 0xffffffffu : ((1u << n_cols) - 1u));
          const int cell_bytes = c_out * 4;
          char* dst = reinterpret_cast<char*>(a.canvas + row * c_out);
          while (empty) {
            const int s0 = __ffs(empty) - 1;
            const unsigned rest = ~(empty >> s0);           // first zero bit above s0 ends the run
            const int len = rest ? __ffs(rest) - 1 : 32 - s0;
            bulk_store_zero(dst + s0 * cell_bytes, zero_smem, len * cell_bytes);
            empty = (len + s0 >= 32) ? 0u : (empty & ~((1u << (s0 + len)) - 1u));
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else {
      // ---------------- one pillar ----------------
      const int row = gw + pi * tw;
      ++pi;
      const int st = info.y, n_all = info.z;
      const int b = (unsigned)info.w >> 24, xi = (info.w >> 12) & 0xfff, yi = info.w & 0xfff;
      const unsigned idx_mine = idx_cur;
      const float4 mine = pt_cur;
      // rotate the pipeline: points of the next pillar, descriptor of the one after
      info = info_next;
      idx_cur = 0xffffffffu;
      if (pi < my_pillars && lane < info.z) {
        idx_cur = (unsigned)a.sorted_idx[info.y + lane];
        pt_cur = a.sorted_pts[info.y + lane];
      }
      if (pi + 1 < my_pillars) info_next = a.pillar_info[gw + (pi + 1) * tw];
      const int n = n_all < max_pts ? n_all : max_pts;
      // --- the max_points lowest point indices of the cell in ascending order -> slots 0..n-1 ---
      float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
      int pidx = 0;
      if (n_all == 1) {
        if (lane == 0) {
          pidx = (int)idx_mine;
          pt = mine;
        }
      } else {
        if (n_all <= 32) {
          unsigned key = lane < n_all ? ((idx_mine << 5) | (unsigned)lane) : 0xffffffffu;  // point index < 2^27
          const int width = n_all <= 2 ? 2 : (n_all <= 4 ? 4 : (n_all <= 8 ? 8 : (n_all <= 16 ? 16 : 32)));
          key = bitonic_sort_u32(key, width, lane);
          const int src = key & 31u;
          pidx = (int)(key >> 5);
          pt.x = __shfl_sync(0xffffffffu, mine.x, src);
          pt.y = __shfl_sync(0xffffffffu, mine.y, src);
          pt.z = __shfl_sync(0xffffffffu, mine.z, src);
          pt.w = __shfl_sync(0xffffffffu, mine.w, src);
        } else {
          // heavy cell: stream the candidates through a bitonic network, keep the lowest max_points
          unsigned long long k64 = ((unsigned long long)idx_mine << 32) | (unsigned)lane;
          k64 = bitonic_sort_warp(k64);
          const int keep = max_pts;  // <= 24: lanes [keep, 32) take new candidates
          for (int base = 32; base < n_all; base += 32 - keep) {
            if (lane >= keep) {
              const int j = base + lane - keep;
              k64 = j < n_all ? (((unsigned long long)(unsigned)a.sorted_idx[st + j] << 32) | (unsigned)j) : ~0ull;
            }
            k64 = bitonic_sort_warp(k64);
          }
          pidx = (int)(unsigned)(k64 >> 32);
          if (lane < n) pt = a.sorted_pts[st + (int)(unsigned)(k64 & 0xffffffffull)];
        }
      }
      const bool act = lane < n;
      // --- cluster centre: sum over the slots / num_points (pillar_encoder.py:108-113) ---
      float sx = act ? pt.x : 0.f, sy = act ? pt.y : 0.f, sz = act ? pt.z : 0.f;
      if (n_all > 1) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          sx += __shfl_xor_sync(0xffffffffu, sx, d);
          sy += __shfl_xor_sync(0xffffffffu, sy, d);
          sz += __shfl_xor_sync(0xffffffffu, sz, d);
        }
      } else {
        sx = __shfl_sync(0xffffffffu, sx, 0);
        sy = __shfl_sync(0xffffffffu, sy, 0);
        sz = __shfl_sync(0xffffffffu, sz, 0);
      }
      const float fn = (float)n;
      const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
      // --- voxel-centre offsets, legacy aliasing + swapped index (pillar_encoder.py:129-139):
      //     x uses coors[:,3] (= y index), y uses coors[:,2] (= x index), z index is 0
      if (lane < n) {
        float4 fa, fb;
        fa.x = __fsub_rn(pt.x, __fadd_rn(__fmul_rn((float)yi, p.vx), p.x_offset));
        fa.y = __fsub_rn(pt.y, __fadd_rn(__fmul_rn((float)xi, p.vy), p.y_offset));
        fa.z = __fsub_rn(pt.z, __fadd_rn(__fmul_rn(0.f, p.vz), p.z_offset));
        fa.w = pt.w;
        fb.x = __fsub_rn(pt.x, mx);
        fb.y = __fsub_rn(pt.y, my);
        fb.z = __fsub_rn(pt.z, mz);
        fb.w = 0.f;
        *reinterpret_cast<float4*>(&s_feat[warp][lane][0]) = fa;
        *reinterpret_cast<float4*>(&s_feat[warp][lane][4]) = fb;
      }
      __syncwarp();
      // --- Linear + BN + ReLU + max over the slots, 2 channels per lane; features broadcast from smem ---
      float best[2] = {0.f, 0.f};
#pragma unroll 2
      for (int k = 0; k < n; ++k) {
        const float4 ga = *reinterpret_cast<const float4*>(&s_feat[warp][k][0]);
        const float4 gb = *reinterpret_cast<const float4*>(&s_feat[warp][k][4]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float x = Wc[q][0] * ga.x;
          x = fmaf(Wc[q][1], ga.y, x);
          x = fmaf(Wc[q][2], ga.z, x);
          x = fmaf(Wc[q][3], ga.w, x);
          x = fmaf(Wc[q][4], gb.x, x);
          x = fmaf(Wc[q][5], gb.y, x);
          x = fmaf(Wc[q][6], gb.z, x);
          best[q] = fmaxf(best[q], fmaf(x, alpha[q], betap[q]));  // ReLU folded: best starts at 0
        }
      }
      __syncwarp();  // s_feat is rewritten by the next pillar
      if (n < max_pts) {  // padded zero rows take part in the max: BN(0) = beta'
        best[0] = fmaxf(best[0], betap[0]);
        best[1] = fmaxf(best[1], betap[1]);
      }
      float* dst = a.canvas + (((size_t)b * G0 + xi) * G1 + yi) * c_out;
      if (lane < c_out) dst[lane] = best[0];
      if (lane + 32 < c_out) dst[lane + 32] = best[1];
      if (want_extra) {
        if (lane == 0 && a.coors_out) *reinterpret_cast<int4*>(a.coors_out + (size_t)row * 4) = make_int4(b, 0, xi, yi);
        if (lane == 0 && a.num_points_out) a.num_points_out[row] = n;
        if (a.pt2pillar_out && act) a.pt2pillar_out[a.pt_off[b] + pidx] = row;
        if (a.voxels_out && lane < max_pts) {
          float* v = a.voxels_out + ((size_t)row * max_pts + lane) * p.c_in;
          v[0] = act ? pt.x : 0.f;
          v[1] = act ? pt.y : 0.f;
          v[2] = act ? pt.z : 0.f;
          if (p.c_in == 4) v[3] = act ? pt.w : 0.f;
        }
      }
    }
  }
  // the zero buffer must stay allocated until every bulk copy has read it
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncthreads();
}

// BatchNorm1d in training mode: statistics over all (kept pillars x max_points) rows, padded zero
// rows included (they add nothing to the sums but count in the denominator); biased variance for
// the normalisation, unbiased for the running update (momentum 0.01).  Q3/Q4 of SURVEY.md.
__global__ void k_bn_finalize(const PillarArgs a, int n_partials) {
  const int c = threadIdx.x;
  if (c >= a.p.c_out) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < n_partials; ++i) {
    s1 += a.stat_partials[((size_t)i * MAX_COUT + c) * 2 + 0];
    s2 += a.stat_partials[((size_t)i * MAX_COUT + c) * 2 + 1];
  }
  const double rows = (double)a.pillar_base[a.batch] * (double)a.p.max_points;
  const double mean = rows > 0 ? s1 / rows : 0.0;
  double var = rows > 0 ? s2 / rows - mean * mean : 0.0;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)a.p.bn_eps));
  const float alpha = a.bn_weight[c] * invstd;
  a.bn_ab[c] = alpha;
  a.bn_ab[MAX_COUT + c] = a.bn_bias[c] - (float)mean * alpha;
  const float mom = a.p.bn_momentum;
  const double unbiased = rows > 1 ? var * rows / (rows - 1.0) : var;
  a.bn_mean[c] = (1.f - mom) * a.bn_mean[c] + mom * (float)mean;
  a.bn_var[c] = (1.f - mom) * a.bn_var[c] + mom * (float)unbiased;
}

// ------------------------------------------------------------------------------------------
__global__ void k_pillar_coors_f64(const float* __restrict__ pts, int64_t n, int c_in, double rx, double ry,
                                   int gx, int gy, float zmin, float zmax, int32_t* __restrict__ coors,
                                   uint8_t* __restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = pts + i * c_in;
  const float x = q[0], y = q[1], z = q[2];
  // ((p + 0.5 * R) / R) * G in float64, then astype(int32) = truncation toward zero
  // (analyse_boxes.py:6-17); third axis: R = 1000, G = 1.  No FMA contraction.
  const double cx = __dmul_rn(__ddiv_rn(__dadd_rn((double)x, 0.5 * rx), rx), (double)gx);
  const double cy = __dmul_rn(__ddiv_rn(__dadd_rn((double)y, 0.5 * ry), ry), (double)gy);
  const double cz = __dmul_rn(__ddiv_rn(__dadd_rn((double)z, 0.5 * 1000.0), 1000.0), 1.0);
  const int ix = __double2int_rz(cx), iy = __double2int_rz(cy), iz = __double2int_rz(cz);
  bool ok = ix >= 0 && iy >= 0 && iz >= 0 && ix < gx && iy < gy && iz < 1;
  ok = ok && (zmin < z) && (z < zmax);
  coors[i * 2 + 0] = ix;
  coors[i * 2 + 1] = iy;
  valid[i] = ok ? 1 : 0;
}

struct Plan {
  PillarArgs a;
  size_t bytes;
  size_t zero_bytes;  // leading region that must be zeroed per call
  int n_cell_blocks, n_pt_blocks, max_pts_per_sample;
};

int make_plan(const float* const* points, const int32_t* n_points, int32_t batch, const slimb200_pillar_params* p,
              int64_t total_points_hint, void* workspace, Plan* plan) {
  if (!p || batch < 1 || batch > SLIMB200_MAX_BATCH) return SLIMB200_E_INVALID;
  if (p->c_in != 3 && p->c_in != 4) return SLIMB200_E_UNSUPPORTED;
  if (p->c_out < 1 || p->c_out > MAX_COUT) return SLIMB200_E_UNSUPPORTED;
  if (p->max_points < 1 || p->max_points > 24) return SLIMB200_E_UNSUPPORTED;
  if (p->grid[0] < 1 || p->grid[1] < 1 || p->grid[2] != 1) return SLIMB200_E_UNSUPPORTED;
  if (p->canvas_layout != SLIMB200_CANVAS_NCHW && p->canvas_layout != SLIMB200_CANVAS_NHWC) return SLIMB200_E_INVALID;
  // channels-last cells are written as a power-of-two number of float4; grid indices are packed in 12 bits
  if (p->canvas_layout == SLIMB200_CANVAS_NHWC &&
      ((p->c_out & 3) || ((p->c_out >> 2) & ((p->c_out >> 2) - 1)) || p->grid[0] > 4096 || p->grid[1] > 4096 || batch > 128))
    return SLIMB200_E_UNSUPPORTED;
  PillarArgs& a = plan->a;
  a = PillarArgs{};
  a.batch = batch;
  a.p = *p;
  a.tiles_x = (p->grid[0] + TILE_R - 1) / TILE_R;
  a.tiles_y = (p->grid[1] + TILE_C - 1) / TILE_C;
  a.tiles_per_sample = a.tiles_x * a.tiles_y;
  const int64_t n_tiles64 = (int64_t)a.tiles_per_sample * batch;
  if (n_tiles64 * TILE_CELLS > 0x7fffffffLL) return SLIMB200_E_UNSUPPORTED;
  a.n_tiles = (int)n_tiles64;
  int64_t total = 0;
  int n_blk = 0, max_n = 0;
  for (int b = 0; b < batch; ++b) {
    const int64_t n = n_points ? n_points[b] : (total_points_hint + batch - 1) / batch;
    if (n < 0) return SLIMB200_E_INVALID;
    a.pts[b] = points ? points[b] : nullptr;
    a.n_pts[b] = (int)n;
    a.pt_off[b] = (int)total;
    a.blk_off[b] = n_blk;
    total += n;
    n_blk += (int)((n + PT_BLOCK - 1) / PT_BLOCK);
    if (n > max_n) max_n = (int)n;
  }
  if (total > 0x7fffffffLL) return SLIMB200_E_UNSUPPORTED;
  a.pt_off[batch] = (int)total;
  a.blk_off[batch] = n_blk;
  // when sizing only (no n_points) leave slack for per-sample block rounding
  const size_t n_blk_cap = (size_t)n_blk + (n_points ? 0 : batch);
  const size_t n_cells = (size_t)a.n_tiles * TILE_CELLS;
  const size_t n_tot = (size_t)total + (n_points ? 0 : batch);
  WorkspaceCarver w(workspace);
  a.cell_count = w.take<int32_t>(n_cells);
  a.cell_inv_first = w.take<int32_t>(n_cells);
  a.cell_fill = w.take<int32_t>(n_cells);
  a.ticket = w.take<int32_t>(64);
  plan->zero_bytes = w.used();
  a.cell_prefix = w.take<int32_t>(n_cells);
  a.cell_ord = w.take<int32_t>(n_cells);
  a.tile_total = w.take<int32_t>(a.n_tiles);
  a.tile_start = w.take<int32_t>(a.n_tiles);
  a.pt_key = w.take<int32_t>(n_tot);
  a.sorted_idx = w.take<int32_t>(n_tot);
  a.sorted_pts = w.take<float4>(n_tot);
  a.blk_cnt = w.take<int32_t>(n_blk_cap);
  a.blk_base = w.take<int32_t>(n_blk_cap);
  a.pillar_base = w.take<int32_t>(SLIMB200_MAX_BATCH + 1);
  a.sample_kept = w.take<int32_t>(SLIMB200_MAX_BATCH);
  {
    const size_t cap = (size_t)batch * (size_t)p->max_voxels;
    a.pillar_info = w.take<int4>((n_tot < cap ? n_tot : cap) + 1);
  }
  a.bn_ab = w.take<float>(2 * MAX_COUT);
  a.stat_partials = w.take<double>((size_t)STATS_MAX_CTAS * MAX_COUT * 2);
  plan->bytes = w.used();
  plan->n_cell_blocks = (a.n_tiles + 7) / 8;  // k_scan_local: one warp per tile
  plan->n_pt_blocks = n_blk;
  plan->max_pts_per_sample = max_n;
  return SLIMB200_OK;
}

}  // namespace

extern "C" size_t slimb200_pillar_workspace_bytes(int32_t batch, int64_t total_points,
                                                  const slimb200_pillar_params* p) {
  Plan plan;
  if (total_points < 0) return 0;
  if (make_plan(nullptr, nullptr, batch, p, total_points, nullptr, &plan) != SLIMB200_OK) return 0;
  return plan.bytes;
}

extern "C" int slimb200_pillar_encode(const float* const* points, const int32_t* n_points, int32_t batch,
                                      const slimb200_pillar_params* p, const float* linear_weight,
                                      const float* bn_weight, const float* bn_bias, float* bn_running_mean,
                                      float* bn_running_var, float* canvas, float* occupancy,
                                      int32_t* pillar_counts, int32_t* coors_out, int32_t* num_points_out,
                                      float* voxels_out, int32_t* pt2pillar_out, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
  if (!points || !n_points || !p || !linear_weight || !bn_weight || !bn_bias || !bn_running_mean ||
      !bn_running_var || !canvas || !occupancy || !workspace)
    return SLIMB200_E_INVALID;
  Plan plan;
  const int rc = make_plan(points, n_points, batch, p, 0, workspace, &plan);
  if (rc != SLIMB200_OK) return rc;
  if (plan.bytes > workspace_bytes) return SLIMB200_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SLIMB200_E_ALIGNMENT;
  for (int b = 0; b < batch; ++b) {
    if (n_points[b] > 0 && !points[b]) return SLIMB200_E_INVALID;
    if (p->c_in == 4 && (reinterpret_cast<uintptr_t>(points[b]) & 15) != 0) return SLIMB200_E_ALIGNMENT;
  }
  PillarArgs& a = plan.a;
  a.linear_weight = linear_weight;
  a.bn_weight = bn_weight;
  a.bn_bias = bn_bias;
  a.bn_mean = bn_running_mean;
  a.bn_var = bn_running_var;
  a.canvas = canvas;
  a.occupancy = occupancy;
  a.pillar_counts = pillar_counts;
  a.coors_out = coors_out;
  a.num_points_out = num_points_out;
  a.voxels_out = voxels_out;
  a.pt2pillar_out = pt2pillar_out;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  SLIMB200_CUDA_TRY(cudaMemsetAsync(workspace, 0, plan.zero_bytes, stream));
  if (plan.max_pts_per_sample > 0) {
    dim3 g((plan.max_pts_per_sample + PT_BLOCK - 1) / PT_BLOCK, batch);
    SLIMB200_LAUNCH(SLIMB200_K_POINT_KEYS, stream, (k_point_keys<<<g, PT_BLOCK, 0, stream>>>(a)));
  }
  SLIMB200_LAUNCH(SLIMB200_K_SCAN_LOCAL, stream,
                  (k_scan_local<<<plan.n_cell_blocks + plan.n_pt_blocks, 256, 0, stream>>>(a, plan.n_cell_blocks)));
  SLIMB200_LAUNCH(SLIMB200_K_SCAN_GLOBAL, stream, (k_scan_global<<<1 + 2 * batch, 1024, 0, stream>>>(a)));
  if (plan.n_pt_blocks > 0) {
    SLIMB200_LAUNCH(SLIMB200_K_RANK_SCATTER, stream, (k_rank_scatter<<<plan.n_pt_blocks, PT_BLOCK, 0, stream>>>(a)));
  }
  if (p->bn_training) {
    const int n_ctas = a.n_tiles < STATS_MAX_CTAS ? a.n_tiles : STATS_MAX_CTAS;
    SLIMB200_LAUNCH(SLIMB200_K_TILE_ENCODE_STATS, stream, (k_tile_encode<1><<<n_ctas, ENC_THREADS, 0, stream>>>(a)));
    SLIMB200_LAUNCH(SLIMB200_K_BN_FINALIZE, stream, (k_bn_finalize<<<1, MAX_COUT, 0, stream>>>(a, n_ctas)));
  }
  {
    SLIMB200_DEVICE(dev, n_sm);
    (void)dev;
    if (p->canvas_layout == SLIMB200_CANVAS_NHWC) {
      SLIMB200_LAUNCH(SLIMB200_K_PILLAR_NHWC, stream,
                      (k_pillar_nhwc<<<n_sm * NH_CTAS_PER_SM, NH_THREADS, 0, stream>>>(a)));
    } else {
      const int n_ctas = a.n_tiles < n_sm * ENC_CTAS_PER_SM ? a.n_tiles : n_sm * ENC_CTAS_PER_SM;
      SLIMB200_LAUNCH(SLIMB200_K_TILE_ENCODE, stream, (k_tile_encode<0><<<n_ctas, ENC_THREADS, 0, stream>>>(a)));
    }
  }
  return SLIMB200_OK;
}

extern "C" int slimb200_pillar_coors_f64(const float* pts, int64_t n, int32_t c_in, double range_x, double range_y,
                                         int32_t grid_x, int32_t grid_y, float z_min, float z_max, int32_t* coors,
                                         uint8_t* valid, void* stream_) {
  if (n < 0 || c_in < 3 || (n > 0 && (!pts || !coors || !valid))) return SLIMB200_E_INVALID;
  if (n == 0) return SLIMB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int threads = 256;
  const int64_t blocks = (n + threads - 1) / threads;
  SLIMB200_LAUNCH(SLIMB200_K_PILLAR_COORS, stream,
                  (k_pillar_coors_f64<<<(unsigned)blocks, threads, 0, stream>>>(pts, n, c_in, range_x, range_y, grid_x,
                                                                               grid_y, z_min, z_max, coors, valid)));
  return SLIMB200_OK;
}
