/*
 * slimb200 -- C ABI of the B200-native (sm_100a) SLIM scene-flow hot path.
 *
 * Drop-in boundary for baurst/liso (paths relative to the reference tree):
 *   stage 1  liso/networks/pcl_to_feature_grid/pcl_to_feature_grid.py:56-107
 *            (mmcv.ops.Voxelization -> PillarFeatureNet -> PointPillarsScatter x2)
 *   stage 2  liso/slim/model/raft_code/corr.py:6-56 (CorrBlock: all-pairs volume, pyramid, lookup)
 *   dataset  liso/datasets/nuscenes/analyse_boxes.py:6-26 (point -> pillar map used by HeadDecoder)
 *
 * Conventions
 *   - plain pointers and sizes only; every device buffer (inputs, outputs, workspace) is owned by
 *     the caller; the library never allocates or frees device memory and keeps no pointer after
 *     return.  Host arrays (marked "host") are read before the call returns.
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no device sync) and are
 *     CUDA-graph capturable.
 *   - return value: 0 success; negative SLIMB200_E_*; positive = cudaError_t of a failed launch.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SLIMB200_H_
#define SLIMB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIMB200_VERSION 210

enum {
  SLIMB200_OK = 0,
  SLIMB200_E_INVALID = -1,      /* bad argument (null pointer, negative size, ...) */
  SLIMB200_E_UNSUPPORTED = -2,  /* shape / dtype outside what the kernels implement */
  SLIMB200_E_WORKSPACE = -3,    /* workspace too small */
  SLIMB200_E_ALIGNMENT = -4,    /* pointer or pitch not aligned as required */
  SLIMB200_E_DRIVER = -5        /* CUDA driver entry point (tensor map encode) unavailable */
};

#define SLIMB200_MAX_BATCH 64
#define SLIMB200_MAX_LEVELS 4
#define SLIMB200_PANEL_COLS 128

/* ------------------------------------------------------------------------------------------
 * Stage 1: pillar encoder.  Replaces PointsPillarFeatureNetWrapper.extract_pts_feat
 * (pcl_to_feature_grid.py:86-102): hard voxelisation (mmcv-full==1.7.1 hard_voxelize_forward,
 * semantics of mmdet3d/core/voxel/voxel_generator.py:137-208, deterministic point-index order),
 * PillarFeatureNet.forward (mmdet3d/models/voxel_encoders/pillar_encoder.py:93-159), PFNLayer
 * (voxel_encoders/utils.py:146-182) and both PointPillarsScatter calls
 * (middle_encoders/pillar_scatter.py:62-102).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  /* fp32 casts of point_cloud_range[:3] and voxel_size, exactly what mmcv hands its kernels */
  float range_min[3];
  float voxel_size[3];
  int32_t grid[3];        /* round((max-min)/voxel): x, y, z cell counts; grid[2] must be 1 */
  int32_t max_points;     /* 20   (pcl_to_feature_grid.py:25) */
  int32_t max_voxels;     /* 40000 (pcl_to_feature_grid.py:27) */
  /* PillarFeatureNet constants (pillar_encoder.py:86-92), fp32 casts of the python floats */
  float vx, vy, vz;
  float x_offset, y_offset, z_offset;
  int32_t c_in;           /* 3 or 4 point channels */
  int32_t c_out;          /* PFN units, 1..64 (64 // channel_reduction_factor) */
  int32_t bn_training;    /* 0: running stats; 1: batch statistics (+ running-stat update) */
  float bn_eps;           /* 1e-3 */
  float bn_momentum;      /* 0.01 */
  int32_t ground_filter;  /* 1: drop ground points with the cone rule below before voxelising, i.e. take the RAW scan
                             ("pcl_full_w_ground") and behave as if "pcl_full_no_ground" had been passed
                             (torch_dataset_commons.py:133-146, 1164-1184); point order and hence pillar order are the
                             same as for the compacted cloud */
  float ground_cone_z;    /* cone_z_threshold__m = data.ground_height_map.ground_threshold (liso_config.yml:113-114) */
  float ground_cone_tan;  /* tan(cone_angle__deg = 0.8 deg) as float32 */
  int32_t canvas_layout;  /* SLIMB200_CANVAS_NCHW (the reference's contiguous layout) or SLIMB200_CANVAS_NHWC
                             (same logical tensor in channels-last memory format, what cuDNN's convs consume);
                             NHWC needs c_out % 4 == 0 */
} slimb200_pillar_params;

enum { SLIMB200_CANVAS_NCHW = 0, SLIMB200_CANVAS_NHWC = 1 };

size_t slimb200_pillar_workspace_bytes(int32_t batch, int64_t total_points,
                                       const slimb200_pillar_params* p);

/*
 * points[b]            device (n_points[b], c_in) f32 row-major, 16-byte aligned when c_in == 4
 * linear_weight        device (c_out, c_in + 6) f32     pfn_layers[0].linear.weight
 * bn_weight..bn_var    device (c_out) f32               pfn_layers[0].norm.{weight,bias,running_mean,running_var}
 *                      (running_mean / running_var are updated in place when bn_training)
 * canvas               device (batch, c_out, grid[0], grid[1]) f32; row = x index, col = y index; stored NCHW or,
 *                      with canvas_layout = NHWC, as (batch, grid[0], grid[1], c_out)
 * occupancy            device (batch, 1, grid[0], grid[1]) f32
 * optional outputs (may be NULL):
 *   pillar_counts      device int32[batch + 1]: exclusive prefix of kept pillars per sample
 *   coors_out          device int32 (>= sum kept pillars, 4): (b, z=0, x index, y index) rows in
 *                      first-appearance order per sample (== voxelize() of the reference, :56-84)
 *   num_points_out     device int32 (>= sum kept pillars)
 *   voxels_out         device f32 (>= sum kept pillars, max_points, c_in), zero padded
 *   pt2pillar_out      device int32 (total_points): row in coors_out of every stored point, else -1
 */
int slimb200_pillar_encode(const float* const* points /*host[batch]*/,
                           const int32_t* n_points /*host[batch]*/, int32_t batch,
                           const slimb200_pillar_params* p, const float* linear_weight,
                           const float* bn_weight, const float* bn_bias, float* bn_running_mean,
                           float* bn_running_var, float* canvas, float* occupancy,
                           int32_t* pillar_counts, int32_t* coors_out, int32_t* num_points_out,
                           float* voxels_out, int32_t* pt2pillar_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Dataset-side point -> pillar map (voxelize_pcl + voxelize_sample, analyse_boxes.py:6-26,
 * torch_dataset_commons.py:975-987): coors = int32_trunc(((p + R/2) / R) * G) in float64, valid =
 * in range && zmin < z < zmax.  pts (n, c_in) f32; coors (n, 2) int32; valid (n) uint8. */
int slimb200_pillar_coors_f64(const float* pts, int64_t n, int32_t c_in, double range_x,
                              double range_y, int32_t grid_x, int32_t grid_y, float z_min,
                              float z_max, int32_t* coors, uint8_t* valid, void* stream);

/* dataset-side pre-processing on the GPU (raw scan in, decoder inputs out).  Replaces
 * infer_ground_label_using_cone + remove_ground_points_from_sample + voxelize_sample/voxelize_pcl + pillarize_bev
 * (torch_dataset_commons.py:133-146, 975-987, 1140-1184; analyse_boxes.py:6-26) and the NaN / -1 / mask padding of
 * the collate function (torch_dataset_commons.py:380-401).
 *   scans[b]          device (n_points[b], c_in) f32 raw scan      ground_labels[b] optional device u8 (n_points[b])
 *   pcl_ta            device (batch, cap, c_in) f32: the non-ground, in-range points in scan order, then NaN
 *   pillar_coors      device (batch, cap, 2) i32 (then -1)          valid   device (batch, cap) u8
 *   counts            device (batch) i32: kept points per sample (stays on the device; nothing is synchronised) */
typedef struct {
  float cone_z_threshold; /* ground rule, evaluated in float32: z < cone_z_threshold + cone_tan * sqrt(x^2 + y^2) */
  float cone_tan;
  double range_x, range_y; /* a12: bev_range_m */
  int32_t grid_x, grid_y;  /* img_grid_size */
  float z_min, z_max;      /* pillar_height_range_m, strict */
  int32_t c_in;
  int32_t reserved;
} slimb200_preprocess_params;

size_t slimb200_preprocess_workspace_bytes(int32_t batch, int32_t cap, const slimb200_preprocess_params* p);
int slimb200_preprocess_points(const float* const* scans /*host[batch]*/, const uint8_t* const* ground_labels /*host[batch] or NULL*/,
                               const int32_t* n_points /*host[batch]*/, int32_t batch, int32_t cap,
                               const slimb200_preprocess_params* p, float* pcl_ta, int32_t* pillar_coors, uint8_t* valid,
                               int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2: all-pairs correlation pyramid + lookup.  Replaces CorrBlock.__init__ / CorrBlock.corr
 * (corr.py:7-21,48-56) and CorrBlock.__call__ + bilinear_sampler (corr.py:23-46,
 * raft_code/utils.py:15-29).
 *
 * Pyramid storage ("panel" layout, chosen for the two kernels that touch it):
 *   the Ncols = sum_l h_l * w_l columns (all levels side by side, level_offset[l] = sum_{k<l} h_k * w_k,
 *   h_0 = h, w_0 = w, h_{l+1} = h_l / 2 floor) are cut into n_panels = ceil(Ncols / 128) panels of 128, the
 *   h * w source pixels (rows) into m_tiles = ceil(h * w / 128) tiles of 128 (rows_padded = 128 * m_tiles).
 *   Every 128 x 128 GEMM tile is two contiguous 16 KB half-tiles (64 columns each) -- contiguous blocks store at
 *   ~6.3 TB/s on B200, 128-byte pieces of a pitched row at 4.7.  Inside a half-tile FOUR neighbouring source pixels
 *   x EIGHT consecutive columns form one 64-byte unit: neighbouring pixels look up nearly the same window, so one
 *   64-byte DRAM burst serves four lanes of the lookup (a pixel-row-major tile spends a burst per pixel and row).
 *   Element (b, source pixel i, column j), with c = j % 128:
 *       half-tile  t = ((b * n_panels + j / 128) * m_tiles + i / 128) * 2 + c / 64
 *       pyramid[t * 8192 + ((i % 128) / 4) * 256 + ((c % 64) / 8) * 32 + (i % 4) * 8 + c % 8]
 *   Columns >= Ncols of the last panel and rows >= h * w of the last tile are zero.  `pitch` = n_panels * 128.
 * Level l as the reference exposes it (corr_pyramid[l], shape (B*h*w, 1, h_l, w_l)) is columns
 * [level_offset[l], level_offset[l] + h_l * w_l) of every row (liso_b200/slim/corr.py materialises it lazily).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, h, w, levels;
  int32_t level_h[SLIMB200_MAX_LEVELS];
  int32_t level_w[SLIMB200_MAX_LEVELS];
  int32_t level_offset[SLIMB200_MAX_LEVELS];
  int32_t n_cols;   /* sum h_l * w_l */
  int32_t pitch;    /* n_panels * 128: columns per source pixel including the zero padding */
  int32_t n_panels; /* ceil(n_cols / 128) */
  int32_t rows_padded; /* 128 * ceil(h * w / 128): source-pixel rows per panel including the zero padding */
} slimb200_corr_layout;

enum { SLIMB200_DTYPE_F32 = 0, SLIMB200_DTYPE_BF16 = 1 };

/* fills `out` for (batch, dim, h, w, levels); returns 0 or a negative error */
int slimb200_corr_layout_init(int32_t batch, int32_t dim, int32_t h, int32_t w, int32_t levels,
                              slimb200_corr_layout* out);
size_t slimb200_corr_workspace_bytes(const slimb200_corr_layout* L);
size_t slimb200_corr_pyramid_bytes(const slimb200_corr_layout* L, int32_t store_dtype);

/* fmap1, fmap2: device (batch, dim, h, w) f32 (dim == 128), 16-byte aligned; fmap_layout says how they are stored:
 *   SLIMB200_CANVAS_NCHW contiguous, or SLIMB200_CANVAS_NHWC = channels-last (batch, h, w, dim), which is what a
 *   channels-last fnet emits and needs no transposition.
 * store_dtype: SLIMB200_DTYPE_BF16 only (the north star stores the volume in bf16; anything else returns
 *   SLIMB200_E_UNSUPPORTED -- slimb200_corr_pyramid_bytes / slimb200_corr_lookup accept SLIMB200_DTYPE_F32 for pyramids the
 *   caller packs itself in the same tiled layout, which the parity tests do).
 * pyramid: device bf16, slimb200_corr_pyramid_bytes() bytes, 128-byte aligned.
 * Computes pyramid[b][i][j] = bf16( sum_d f1[b,d,i] * pool_l(f2)[b,d,j] / sqrt(dim) ) with bf16
 * operands and fp32 accumulation on the tcgen05 tensor cores (pooling folded into the operand:
 * avg_pool(corr) == corr(avg_pool(f2)) by linearity). */
int slimb200_corr_build(const float* fmap1, const float* fmap2, int32_t fmap_layout,
                        const slimb200_corr_layout* L, int32_t store_dtype, void* pyramid, void* workspace,
                        size_t workspace_bytes, void* stream);

/* coords: device (batch, 2, h, w) f32, channel 0 = x (column), 1 = y (row).
 * out:    device (batch, levels * (2r+1)^2, h, w) f32 contiguous, channel k = l*(2r+1)^2 + i*(2r+1) + j
 *         sampled at (x / 2^l + i - r, y / 2^l + j - r), bilinear, zeros outside, align_corners.
 *         out_layout = SLIMB200_CANVAS_NCHW, or SLIMB200_CANVAS_NHWC for the same tensor in channels-last
 *         memory format (batch, h, w, channels) -- radius 3 only -- which the 1x1 conv consuming it prefers. */
int slimb200_corr_lookup(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                         const float* coords, int32_t radius, float* out, int32_t out_layout, void* stream);

/* the lookup FUSED with the 1x1 convolution that consumes it, SmallMotionEncoder.conv_stat_corr1 (+ ReLU)
 * (liso/slim/model/update.py:49,71), so that the (batch, levels*49, h, w) lookup tensor never reaches HBM:
 *   out[b, y, x, n] = act( bias[n] + sum_k weight[n, k] * lookup[b, k, y, x] ),  k = l*49 + i*7 + j as above.
 * The window values and the weights are rounded to tf32 (cvt.rna), products accumulate in fp32 on the tcgen05 tensor
 * cores -- the precision cuDNN uses for this convolution when TF32 is allowed (|err| <= 2^-10 * sum_k |w_nk| |v_k|).
 * bf16 pyramid, radius 3 and 4 levels only; c_out in {32, 64, 96}.
 *
 * corr_lookup_conv_pack: once per weight tensor -- weight: device (c_out, levels*49) f32 = conv weight
 *   (c_out, levels*49, 1, 1); bias: device (c_out) f32 or NULL; packed: device, slimb200_corr_lookup_conv_packed_bytes(c_out)
 *   bytes, 16-byte aligned: the tf32 B operand in its shared-memory layout (7 K blocks of c_out rows x 128 bytes,
 *   128-byte swizzle, zero padded to 56 K slots per level; one image per K-slot order of the two fused kernels) followed
 *   by the biases.
 * corr_lookup_conv: out: device, channels-last rows: pixel (b, y, x) at out + ((b*h + y)*w + x) * out_pitch, c_out floats
 *   each; out_pitch >= c_out floats, multiple of 4 (a channel slice of a wider channels-last tensor works); 16-byte
 *   aligned.  relu: 0 / 1. */
size_t slimb200_corr_lookup_conv_packed_bytes(int32_t c_out);
int slimb200_corr_lookup_conv_pack(const float* weight, const float* bias, int32_t levels, int32_t radius, int32_t c_out,
                                   void* packed, void* stream);
int slimb200_corr_lookup_conv(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                              const float* coords, int32_t radius, const void* packed, int32_t c_out, int32_t relu,
                              float* out, int32_t out_pitch, void* stream);

/* Tuning hook (tools/kbench.py) for the radius-3 lookup on bf16 pyramids: 0 = first-generation kernel (lane = pixel, window
 * staged in shared memory), 1 (default) = one thread per (pixel, level), registers only (csrc/corr_lookup2.cu), 2 = one
 * thread per window row (csrc/corr_lookup3.cu; the fused lookup + convolution is built on it).  Returns the previous value; negative values only query. */
int slimb200_lookup_generation(int32_t generation);
/* Same for the fused lookup + convolution: 3 = row-per-thread gather, A tile in shared memory (csrc/corr_lookup3.cu);
 * 4 (default) = (pixel, level)-per-thread gather with cp.async landing slots, A tile in tensor memory (csrc/corr_lookup4.cu). */
int slimb200_lookup_conv_generation(int32_t generation);

/* ------------------------------------------------------------------------------------------
 * output decoder.  Replaces HeadDecoder.forward (liso/slim/model/head_decoder.py:410-496,
 * 517-717) for the released output_modification, batched_grid_data_to_pointwise_data and
 * compute_batched_bev_static_aggregated_flow (slim_loss/static_aggregation.py:8-110), weighted_pc_alignment
 * (slim_loss/weighted_pc_alignment.py:10-80, no epsilon) and symmetric_orthogonalization
 * (torch_symm_ortho/__init__.py:68-69: U @ Vh, no determinant fix).  No host synchronisation.
 *
 * bev (batch, H, W, 16) f32, one packed 64-byte row per cell (the reference's tensors are channel slices of it):
 *    0      disappearing_logit (-100)            1:4    class_logits = static | dynamic | ground logit (masked)
 *    4:7    class_probs = staticness | dynamicness | groundness
 *    7:10   static flow (x, y, 0), masked         10:13  dynamic flow (x, y, 0), masked
 *    13:16  aggregated flow
 * bev_aggr (batch, H, W, 4) f32, only with static_aggregation: 0:2 static_aggr_flow, 2:4 masked_static_aggr_flow
 * bev_classes (batch, H, W, 3) u8: is_dynamic | is_static | is_ground
 * points (batch, n_points, 14) f32: 0:3 static flow, 3:6 dynamic flow, 6 dynamicness, 7 staticness,
 *    8:11 aggregated flow, 11:14 static_aggr_flow (x, y, 0); all zero for invalid points
 * trafo (batch, 4, 4) f64 row-major, not_enough (batch) u8: only with static_aggregation
 * ---------------------------------------------------------------------------------------- */
#define SLIMB200_DECODE_BEV_CHANNELS 16
#define SLIMB200_DECODE_POINT_CHANNELS 14
typedef struct {
  int32_t batch, H, W;
  int32_t n_points;            /* padded points per sample */
  int32_t pc_stride;           /* floats per point in `pc` (>= 3) */
  int32_t final_scale;         /* pillar coordinate divisor (u_net.final_scale, 1) */
  int32_t static_aggregation;  /* 0: skip the weighted Kabsch part (bev_aggr untouched, point channels 11:14 zero) */
  int32_t reserved;
  double ext_min_x, ext_min_y, ext_max_x, ext_max_y; /* bev_extent (head_decoder.py:498-514) */
} slimb200_decode_params;

size_t slimb200_head_decode_workspace_bytes(const slimb200_decode_params* p);

/* net_out (batch, H, W, 8) f32 contiguous: logits 0:4 (1 = static, 2 = dynamic), static flow 4:6, dynamic flow 6:8
 * filled (batch, H, W) u8; pc (batch, n_points, pc_stride) f32; coors (batch, n_points, 2) i32; valid (batch, n_points) u8
 * dyn_threshold: DEVICE pointer to one float (MovingAverageThreshold.value(), no host read) */
int slimb200_head_decode(const float* net_out, const uint32_t* logit_min_key /* from slimb200_raft_output, or NULL */,
                         const uint8_t* filled, const float* pc, const int32_t* coors,
                         const uint8_t* valid, const float* dyn_threshold, const slimb200_decode_params* p,
                         float* bev, float* bev_aggr, uint8_t* bev_classes, float* points, double* trafo,
                         uint8_t* not_enough, void* workspace, size_t workspace_bytes, void* stream);

/* SURVEY 8(f).2 glue, once per GRU iteration (raft_mod.py:216-257): upflow_n / uplogits_n (bilinear, align_corners,
 * raft_code/utils.py:50-60), flip (x, y) -> (row, col) and * metres per pixel (raft_mod.py:262-266), and
 * HeadDecoder.concat2network_output (head_decoder.py:36-64) in one pass.
 * flow (batch, 2, h, w) f32 = coords1 - coords0, logits (batch, 4, h, w) f32, both NCHW contiguous;
 * net_out (batch, n*h, n*w, 8) f32 = [logits 0:4 | static flow (row, col) | dynamic flow (row, col)];
 * logit_min_key (optional, one device word): order-preserving encoding of min(logits[:, 1:3]) over the whole batch,
 * to be handed to slimb200_head_decode. */
int slimb200_raft_output(const float* flow, const float* logits, int32_t batch, int32_t h, int32_t w, int32_t n,
                         float res_rows, float res_cols, float* net_out, uint32_t* logit_min_key, void* stream);

/* Glue for the channels-last feature encoder: affine InstanceNorm2d (eps, biased variance) + optional ReLU on an
 * NHWC tensor -- `norm_fn = "instance_affine"` + ReLU of liso/slim/model/extractor.py:5-68,211-297 -- in three
 * launches and two passes over the data (PyTorch: copy to NCHW, cuDNN batch-norm on (1, B*C, H, W), copy back, clamp).
 * x, out: device (batch, height, width, channels) f32, 16-byte aligned (out may alias x); channels % 4 == 0, <= 256.
 * relu: bit 0 = ReLU right after the normalisation; residual (optional, same shape) is added after that, and bit 1 =
 * ReLU after the addition -- the residual join `relu(x + y)` of extractor.py:57-68 fused into the same pass. */
size_t slimb200_instnorm_workspace_bytes(int32_t batch, int32_t channels, int32_t hw);
int slimb200_instnorm_nhwc(const float* x, const float* gamma, const float* beta, float eps, int32_t batch,
                           int32_t height, int32_t width, int32_t channels, int32_t relu, const float* residual,
                           float* out, void* workspace, size_t workspace_bytes, void* stream);
/* The same for a channel SLICE of a wider channels-last tensor: x points at the slice's first channel, x_pitch = channels
 * per pixel of the whole tensor (x_pitch == channels: identical to slimb200_instnorm_nhwc); out is packed (channels per
 * pixel) and must not alias x when x_pitch != channels.  Used for one half of two parallel convolutions evaluated as one. */
int slimb200_instnorm_nhwc_slice(const float* x, int32_t x_pitch, const float* gamma, const float* beta, float eps, int32_t batch,
                                 int32_t height, int32_t width, int32_t channels, int32_t relu, const float* residual, float* out,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* out (pixels, channels) packed = relu(x[:, slice] + bias): the other half (a convolution followed by ReLU, extractor.py:
 * 262-266 with norm_fn "none"), evaluated without its bias inside the stacked convolution. */
int slimb200_bias_relu_slice(const float* x, int32_t x_pitch, const float* bias, int32_t channels, int64_t pixels, float* out,
                             void* stream);

/* SURVEY 8(f).2 glue between the stock convolutions of the ConvGRU update block (liso/slim/model/update.py:23-38,
 * 70-93,130-150; raft_mod.py:188-212), all on channels-last fp32 tensors given as (pixels = batch*h*w, channels) rows.
 * The two 304-channel GRU convolution inputs [h | x] and [r*h | x] are persistent caller buffers; these entry points
 * write their channel slots directly instead of torch.cat + separate element-wise launches.  Channel counts, channel
 * offsets and pitches are in floats and must be multiples of 4; pointers 16-byte aligned.
 *
 * nhwc_pack: concatenate n_src (<= 4) sources along channels and store the result at channel offset
 *   dst_channel_offset[d] of each of the n_dst (<= 2) destinations with row pitch dst_pitch[d].  A source is src_channels[s]
 *   channels starting at src[s] in rows of pitch src_pitch[s] (NULL: packed, pitch == channels; a channel slice of a wider
 *   tensor has the pointer advanced to its first channel).  src / src_channels / src_pitch / dst / dst_channel_offset /
 *   dst_pitch are HOST arrays.
 * gru_gate_zr: zr_raw (pixels, 2*hidden) = conv output WITHOUT bias of the stacked update|reset gate convolution,
 *   bias_zr (2*hidden); z_out (pixels, hidden) = sigmoid(zr[:, :hidden] + b); rhx[:, :hidden] = sigmoid(zr[:, hidden:] + b) * hx[:, :hidden].
 * gru_gate_out: q_raw (pixels, hidden) conv output without bias; hx[:, :hidden] <- (1 - z) * h + z * tanh(q_raw + b),
 *   in place, and the same rows packed into h_out (pixels, hidden) for the two heads.
 * iter_update: raw head outputs (no bias) of FlowOrClassificationHead (update.py:6-20) addressed as
 *   base + b*batch_stride + c*channel_stride + pix*pixel_stride (NCHW: (C*h*w, h*w, 1); channels-last: (C*h*w, 1, C);
 *   channel slices of one stacked head output work too);
 *   coords1 (batch,2,h,w) += dflow + bias; logits (batch,n_logits,h,w) += dlogits + bias; flow = coords1 - coords_grid
 *   (channel 0 = column, channel 1 = row: raft_code/utils.py:32-37); coords1 / flow / logits are NCHW contiguous;
 *   stacked (optional) receives the NCHW concatenation [flow | logits] in the first 2 + n_logits channels of a
 *   (batch, stacked_channels, h, w) buffer; further (padding) channels are left untouched.
 * add_relu: out = relu(x + y) over n floats (residual join of extractor.py:57-68; out may alias x or y). */
int slimb200_nhwc_pack(const float* const* src, const int32_t* src_channels, const int32_t* src_pitch, int32_t n_src,
                       float* const* dst, const int32_t* dst_channel_offset, const int32_t* dst_pitch, int32_t n_dst,
                       int64_t pixels, void* stream);
int slimb200_gru_gate_zr(const float* zr_raw, const float* bias_zr, const float* hx, int32_t hx_pitch, float* z_out,
                         float* rhx, int32_t rhx_pitch, int32_t hidden, int64_t pixels, void* stream);
int slimb200_gru_gate_out(const float* q_raw, const float* bias_q, const float* z, float* hx, int32_t hx_pitch,
                          float* h_out, int32_t hidden, int64_t pixels, void* stream);
/* stacked (optional): copy of [flow | logits] for the stacked motion-encoder convolution, (batch, |stacked_channels|, h, w);
 * stacked_channels > 0: planar (NCHW), < 0: channels-last with -stacked_channels channels per pixel. */
int slimb200_iter_update(const float* dflow_raw, int64_t dflow_batch_stride, int64_t dflow_channel_stride,
                         int64_t dflow_pixel_stride, const float* bias_flow, const float* dlogits_raw,
                         int64_t dlogits_batch_stride, int64_t dlogits_channel_stride, int64_t dlogits_pixel_stride,
                         const float* bias_logits, int32_t n_logits, int32_t batch, int32_t h, int32_t w, float* coords1,
                         float* flow, float* logits, float* stacked, int32_t stacked_channels, void* stream);
/* Same update with the k x k output convolution of both heads (stride 1, zero padding k/2) evaluated as ONE 1x1
 * convolution to k*k "taps" plus the sum of the taps over the window, done here: taps (batch, h, w, k*k*(2 + n_logits))
 * f32 channels-last, channel = (ky*k + kx)*(2 + n_logits) + c with c = [dflow 0:2 | dlogits], i.e. the 1x1 weight is
 * W1[(ky*k + kx)*(2 + n_logits) + c][cin] = Wconv[c][cin][ky][kx]; raw[p] = sum_(ky,kx) taps[p + (ky - k/2, kx - k/2)][ky*k + kx]
 * in fp32, taps outside the map skipped.  (cuDNN's 3x3 convolution to 6 channels takes as long as the one to 256.) */
int slimb200_iter_update_taps(const float* taps, int32_t ksize, const float* bias_flow, const float* bias_logits,
                              int32_t n_logits, int32_t batch, int32_t h, int32_t w, float* coords1, float* flow,
                              float* logits, float* stacked, int32_t stacked_channels, void* stream);
int slimb200_add_relu(const float* x, const float* y, float* out, int64_t n, void* stream);
/* out = relu((x + bias_x[c]) + y), channels-last data with `channels` (multiple of 4) innermost: the projection shortcut
 * of a residual block (extractor.py:44-68) convolved WITHOUT its bias, which joins here instead of in a pass of its own. */
/* Context encoder tail (raft_mod.py:170-173): raw = cnet's last convolution WITHOUT its bias, channels-last (pixels,
 * hidden + context); net = tanh(raw[:, :hidden] + bias), inp = relu(raw[:, hidden:] + bias), both packed channels-last. */
int slimb200_ctx_split(const float* raw, const float* bias, int32_t hidden, int32_t context, int64_t pixels, float* net,
                       float* inp, void* stream);
int slimb200_add_bias_relu(const float* x, const float* bias_x, int32_t channels, const float* y, float* out, int64_t n,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Export writer.  Replaces the zlib pass of np.savez_compressed in
 * slim_inference_and_save_result (liso/slim/experiment.py:459-471): every saved fp32 map becomes a raw
 * DEFLATE stream (RFC 1951, fixed Huffman code, zero-word runs as distance-1 matches) plus the CRC-32
 * remainder of its bytes, computed on the device; the host frames the streams as zip members
 * (liso_b200/slim/npz_stream.py) that np.load reads like the reference's files
 * (torch_dataset_commons.py:614-616).
 *
 * A member = one array of one sample: n_words 4-byte words, word k at
 * src[(k / words_per_cell) * cell_stride + k % words_per_cell] (a channel slice of a packed
 * channels-last buffer, or cell_stride == words_per_cell for a contiguous array).
 *   plan    host only: fills first_chunk of every member; sizes of the workspace and of the largest
 *           possible output (members are cut into SLIMB200_DEFLATE_CHUNK_BYTES chunks).
 *   init    fills the caller-owned device table (SLIMB200_DEFLATE_TABLE_BYTES) the CRC needs; once per device.
 *   encode  members_dev: the planned member array copied to the device.  out: the streams of all
 *           members back to back.  member_out: (n_members + 1) x 4 uint32 on the device:
 *           row i = {offset into out, bytes, crc remainder R, chunks}; row n_members = {total bytes,
 *           overflow flag (out_capacity too small: out is incomplete), 0, 0}.
 *           crc32(prefix | member bytes) = crc32(prefix | zeros of the member's length) ^ R.
 *           Each stream is complete (last block has BFINAL); a stored block holding the .npy header may
 *           be put in front of it.
 * ---------------------------------------------------------------------------------------- */
#define SLIMB200_DEFLATE_CHUNK_BYTES 8192
#define SLIMB200_DEFLATE_SLOT_BYTES 9232
#define SLIMB200_DEFLATE_MAX_CHUNKS 8192 /* per member: 64 MB */
#define SLIMB200_DEFLATE_TABLE_BYTES ((256 + 2049 + 8192) * 4)
typedef struct {
  const void* src;        /* device, 4-byte aligned */
  int32_t words_per_cell; /* >= 1 */
  int32_t cell_stride;    /* in words, >= words_per_cell */
  uint32_t n_words;       /* > 0 */
  uint32_t first_chunk;   /* filled by slimb200_deflate_plan */
  uint32_t crc_geo;       /* filled by slimb200_deflate_plan: sum_{i < n_words} x^(32 i) mod P (CRC of a constant fill) */
  uint32_t reserved;
} slimb200_deflate_member;
int slimb200_deflate_plan(slimb200_deflate_member* members /*host[n_members]*/, int32_t n_members, int64_t* total_chunks,
                          size_t* workspace_bytes, size_t* out_bound);
int slimb200_deflate_init(void* tables, void* stream);
int slimb200_deflate_encode(const slimb200_deflate_member* members_dev, int32_t n_members, int64_t total_chunks,
                            const void* tables, void* workspace, size_t workspace_bytes, void* out, size_t out_capacity,
                            uint32_t* member_out, void* stream);

const char* slimb200_strerror(int code);
int slimb200_version(void);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): every kernel launch of the library is counted; between
 * profile_begin / profile_end each launch is additionally bracketed by CUDA events on its stream.
 * profile_end synchronises on the recorded events and returns per-kernel totals.
 * ---------------------------------------------------------------------------------------- */
enum {
  SLIMB200_K_POINT_KEYS = 0,
  SLIMB200_K_SCAN_LOCAL,
  SLIMB200_K_SCAN_GLOBAL,
  SLIMB200_K_RANK_SCATTER,
  SLIMB200_K_TILE_ENCODE_STATS,
  SLIMB200_K_BN_FINALIZE,
  SLIMB200_K_TILE_ENCODE,
  SLIMB200_K_PILLAR_NHWC,
  SLIMB200_K_FEAT_TRANSPOSE,
  SLIMB200_K_FEAT_PACK,
  SLIMB200_K_CORR_GEMM,
  SLIMB200_K_CORR_LOOKUP,
  SLIMB200_K_PILLAR_COORS,
  SLIMB200_K_DECODE_MIN,
  SLIMB200_K_DECODE_BEV,
  SLIMB200_K_DECODE_POINTS,
  SLIMB200_K_KABSCH,
  SLIMB200_K_DECODE_AGGR,
  SLIMB200_K_RAFT_OUTPUT,
  SLIMB200_K_PRE_COUNT,
  SLIMB200_K_PRE_SCAN,
  SLIMB200_K_PRE_SCATTER,
  SLIMB200_K_PRE_PAD,
  SLIMB200_K_KABSCH_MOMENTS,
  SLIMB200_K_IN_STATS,
  SLIMB200_K_IN_FINALIZE,
  SLIMB200_K_IN_APPLY,
  SLIMB200_K_NHWC_PACK,
  SLIMB200_K_GRU_GATE_ZR,
  SLIMB200_K_GRU_GATE_OUT,
  SLIMB200_K_ITER_UPDATE,
  SLIMB200_K_ADD_RELU,
  SLIMB200_K_LOOKUP_CONV,
  SLIMB200_K_LOOKUP_CONV_PACK,
  SLIMB200_K_DEFLATE_TABLES,
  SLIMB200_K_DEFLATE_CHUNKS,
  SLIMB200_K_DEFLATE_SCAN,
  SLIMB200_K_DEFLATE_GATHER,
  SLIMB200_K_CTX_SPLIT,
  SLIMB200_K_BIAS_RELU_SLICE,
  SLIMB200_N_KERNELS
};
int slimb200_profile_begin(void);
int slimb200_profile_end(float* ms_total /*host[SLIMB200_N_KERNELS]*/,
                         int64_t* timed_launches /*host[SLIMB200_N_KERNELS]*/);
int64_t slimb200_launch_count(int32_t kernel_id /* <0: all kernels */);
const char* slimb200_kernel_name(int32_t kernel_id);

#ifdef __cplusplus
}
#endif
#endif /* SLIMB200_H_ */
