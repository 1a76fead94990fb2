// Stage 2b of the SLIM hot path on B200, third generation: the radius-3 lookup into the bf16 correlation pyramid with ONE
// THREAD PER WINDOW ROW, alone or FUSED with the 1x1 convolution that consumes it.
//
// Replaces CorrBlock.__call__ (liso/slim/model/raft_code/corr.py:23-46) + bilinear_sampler (raft_code/utils.py:15-29)
// and, in the fused entry, SmallMotionEncoder.conv_stat_corr1 + ReLU (liso/slim/model/update.py:49,71):
//   c = relu(conv1x1(lookup(coords), W (N, L*49), bias))
// so that the (B, 196, h, w) lookup tensor (40 MB per call at KITTI size, batch 8) never reaches HBM.
//
// Why rows: the lookup is a latency-bound gather (ncu, round 2: 84 MB of DRAM reads per launch at 40 % of the DRAM
// rate, "long scoreboard" the top stall at 16 resident warps per SM).  A thread that owns a whole (pixel, level) window
// needs 128 registers (64 of them the landing zone of its 16 loads), i.e. 16 warps per SM.  Here a warp owns FOUR
// neighbouring pixels (one 64-byte unit of the pyramid layout, include/slimb200.h) x EIGHT window rows, lane = row * 4 +
// pixel, and a thread holds one window row: 2 loads, ~48 registers, 32+ warps per SM.
//   1. lane (p, r) computes the sample position of window offset r on BOTH axes (7 of 8 lanes; the reference's
//      normalise / un-normalise fp32 round trip, division by the precomputed reciprocal), floors and masked weights;
//      the window origin is the minimum over the 8 lanes of the pixel (3 butterfly shuffles per axis)
//   2. the x weights of all 7 offsets reach every lane of the pixel by shuffles
//   3. the lane fetches ITS window row (two 16-byte streaming loads, lanes 4r..4r+3 share the DRAM burst), re-aligns it
//      and forms the 7 horizontal blends h[r][i]
//   4. the vertical blend of output row j needs h[j] and h[j + 1]: one shuffle per value from the lane 4 above
//   5. lane (p, j) hands out[i][j], i = 0..6 to a sink
// Windows whose per-offset floors scatter by one around an integer position (every pixel of the first GRU iteration)
// take the 3-tap variant (weights (w0, w1, 0) or (0, w0, w1): the zero adds an exact zero) on a 9 x 9 window; pixels
// with non-finite / absurd coordinates take predicated 4-tap loads.
//
// k_lookup_conv_tf32: persistent CTAs, 512 threads = 64 pixels x 8 rows, two passes per 128-pixel tile, the window loads
// of four (pass, level) units in flight per thread while an earlier unit is blended (software pipeline).  Lane
// (p, j) writes its 7 values (rounded to tf32) + one zero as two 16-byte chunks of row `pixel` of a K-major,
// 128-byte-swizzled A tile in shared memory (K slot = level * 56 + j * 8 + i); the weights (N x 224 tf32, packed once
// per weight tensor by k_lookup_conv_pack in the same slot order) arrive with one bulk copy and stay for the whole
// kernel.  One elected thread issues 28 tcgen05.mma.kind::tf32 (M = 128, N, K = 8) into one of two TMEM accumulators;
// four warps read the PREVIOUS tile's accumulator back (tcgen05.ld), add the bias, apply the ReLU and store the rows
// while the tensor core works and the other warps already gather the next tile.
#include "lookup_core.cuh"
#include "ptx.cuh"

namespace {

using namespace slimb200_lookup;
using namespace slimb200_ptx;

constexpr int ROWS = 8;       // window rows fetched per (pixel, level) = lanes per pixel
constexpr int WARP_PIX = 4;   // pixels per warp: one 64-byte unit of the pyramid layout

// sample position of ONE window offset + floor + masked weights
__device__ __forceinline__ void one_tap(float c, float inv, int o, int size, float sm1, float rinv, float& w0, float& w1, int& f) {
  const float ip = sample_pos2(c, inv, o - R, sm1, rinv);
  const float fl = floorf(ip);
  const int i0 = (int)fl;
  const float w_hi = __fsub_rn(ip, fl);
  const float w_lo = __fsub_rn(__fadd_rn(fl, 1.f), ip);
  w0 = ((unsigned)i0 < (unsigned)size) ? w_lo : 0.f;
  w1 = ((unsigned)(i0 + 1) < (unsigned)size) ? w_hi : 0.f;
  f = i0 - o;
}

__device__ __forceinline__ int group_min(int v) {  // over the 8 lanes (rows) of a pixel
  v = min(v, __shfl_xor_sync(FULL, v, 4));
  v = min(v, __shfl_xor_sync(FULL, v, 8));
  v = min(v, __shfl_xor_sync(FULL, v, 16));
  return v;
}

// One level of 4 pixels x 8 rows (a whole warp, every lane must call), split in two so that the window loads of several
// (pixel, level) units can be in flight while an earlier one is blended (the fused kernel keeps four):
//   rows_prepare  positions, floors, weights, window origin, the lane's two 16-byte loads (nothing waits for them)
//   rows_finish   re-alignment, x weights of all offsets (shuffles), horizontal + vertical blend, sink
struct RowState {
  uint32_t raw[8];         // the lane's window row as loaded
  float wx0, wx1, wy0, wy1;  // masked weights of window offset r on both axes (this lane's offset)
  int sft;                 // element shift of window column 0 inside raw
  int row0;                // pyramid column of window element (0, 0)
  int dx, dy;              // floor - origin of this lane's offset: 0 regular, 1 shifted, else slow (meaningless for r == 7)
};

struct LevelGeo {
  int W, H, off;
  float inv, swm1, shm1, rw, rh;  // 1 / 2^level, size - 1 and the correctly rounded reciprocals
};

__device__ __forceinline__ LevelGeo level_geo(const LookupGeo& G, int level) {
  LevelGeo g;
  g.W = pick4(G.lw, level);
  g.H = pick4(G.lh, level);
  g.off = pick4(G.lo, level);
  g.inv = 1.0f / (float)(1 << level);  // coords / 2**l is exact
  g.swm1 = (float)(g.W - 1);
  g.shm1 = (float)(g.H - 1);
  g.rw = __frcp_rn(g.swm1);
  g.rh = __frcp_rn(g.shm1);
  return g;
}

__device__ __forceinline__ void rows_prepare(RowState& st, const __nv_bfloat16* __restrict__ base, int panel_stride, int pitch,
                                             const LevelGeo& g, float cx, float cy) {
  const int r = lane_id() >> 2;
  int fx, fy;
  one_tap(cx, g.inv, r, g.W, g.swm1, g.rw, st.wx0, st.wx1, fx);
  one_tap(cy, g.inv, r, g.H, g.shm1, g.rh, st.wy0, st.wy1, fy);
  if (r == ROWS - 1) fx = fy = INT_MAX;  // offsets 0..6 only
  const int xb = group_min(fx), yb = group_min(fy);
  st.dx = fx - xb;
  st.dy = fy - yb;
  // window element (0, 0); origins beyond +-2^18 (every tap outside, zero weights) are clamped so that the index stays an int
  st.row0 = g.off + max(min(yb, 1 << 18), -(1 << 18)) * g.W + max(min(xb, 1 << 18), -(1 << 18));
  st.sft = fetch_row(base, panel_stride, pitch, st.row0 + r * g.W, st.raw);
}

// Lane (p, j), j < 7, ends with sink.emit_row(j, o) where o[i] = out[i][j]: the value of channel i * 7 + j
// (i offsets x, j offsets y: the reference's transposed window, corr.py:29-41).
template <class Sink>
__device__ __forceinline__ void rows_finish(const RowState& st, const __nv_bfloat16* __restrict__ base, int panel_stride, int pitch,
                                            const LevelGeo& g, float cx, float cy, bool live, Sink& sink) {
  const int lane = lane_id();
  const int p = lane & 3, r = lane >> 2;
  const bool rowlane = r < ROWS - 1;
  const unsigned bad = __ballot_sync(FULL, live && rowlane && ((unsigned)st.dx > 1u || (unsigned)st.dy > 1u));
  const unsigned shx = __ballot_sync(FULL, rowlane && st.dx == 1), shy = __ballot_sync(FULL, rowlane && st.dy == 1);
  // x weights of all 7 offsets: offset i lives in lane (p, i)
  float ax0[WIN], ax1[WIN];
#pragma unroll
  for (int i = 0; i < WIN; ++i) {
    ax0[i] = __shfl_sync(FULL, st.wx0, i * 4 + p);
    ax1[i] = __shfl_sync(FULL, st.wx1, i * 4 + p);
  }
  float o[WIN];
  if (bad == 0u && (shx | shy) == 0u) {
    // ---- regular windows: 8 x 8, two taps per axis ----
    uint32_t win[4];
    realign<4>(st.raw, st.sft, win);
    sink.begin();
    float h[WIN];
    float e0 = wel<4>(win, 0);
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const float e1 = wel<4>(win, i + 1);
      h[i] = fmaf(e1, ax1[i], e0 * ax0[i]);
      e0 = e1;
    }
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const float hn = __shfl_down_sync(FULL, h[i], 4);  // row r + 1 of the same pixel
      o[i] = fmaf(hn, st.wy1, h[i] * st.wy0);
    }
  } else {
    // ---- shifted windows: 9 x 9, three taps per axis, one weight of the three is zero ----
    uint32_t raw8[8], win[5], win8[5];
    const int sft8 = fetch_row(base, panel_stride, pitch, st.row0 + 8 * g.W, raw8);  // window row 8 (needed by row 6 only)
    realign<5>(st.raw, st.sft, win);
    realign<5>(raw8, sft8, win8);
    sink.begin();
    float h[WIN], h8[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const bool s = (shx >> (i * 4 + p)) & 1u;
      const float a = s ? 0.f : ax0[i], bq = s ? ax0[i] : ax1[i], c = s ? ax1[i] : 0.f;
      h[i] = fmaf(wel<5>(win, i + 2), c, fmaf(wel<5>(win, i + 1), bq, wel<5>(win, i) * a));
      h8[i] = fmaf(wel<5>(win8, i + 2), c, fmaf(wel<5>(win8, i + 1), bq, wel<5>(win8, i) * a));
    }
    const bool t = st.dy == 1;
    const float ay = t ? 0.f : st.wy0, by = t ? st.wy0 : st.wy1, cyw = t ? st.wy1 : 0.f;
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const float hn1 = __shfl_down_sync(FULL, h[i], 4);
      float hn2 = __shfl_down_sync(FULL, h[i], 8);
      if (r == ROWS - 2) hn2 = h8[i];
      o[i] = fmaf(hn2, cyw, fmaf(hn1, by, h[i] * ay));
    }
    const bool pixel_slow = (bad & (0x11111111u << p)) != 0u;
    if (pixel_slow && live && rowlane) {  // ---- anything else ----
      const float iy = sample_pos2(cy, g.inv, r - R, g.shm1, g.rh);
#pragma unroll
      for (int i = 0; i < WIN; ++i)
        o[i] = sample_slow2(base, panel_stride, g.W, g.H, g.off, sample_pos2(cx, g.inv, i - R, g.swm1, g.rw), iy);
    }
  }
  if (rowlane) sink.emit_row(r, o);
}

// ------------------------------------------------------------------------------------------ stand-alone lookup
// CTA = 32 pixels (8 warps x 4 pixels), the levels in turn; the 32 x n_ch tile is staged in shared memory and leaves
// with coalesced rows: NCHW = one 128-byte row of 32 pixels per channel, channels-last = one contiguous block.
constexpr int V3_PIX = 32;
constexpr int V3_THREADS = V3_PIX * ROWS;  // 256
constexpr int V3_NCH = SLIMB200_MAX_LEVELS * WIN * WIN;
constexpr int V3_PITCH_C = 36;             // NCHW staging [channel][36]: bank = 4 * j + p for lane (p, j)
constexpr int V3_PITCH_P = V3_NCH + 1;     // channels-last staging [pixel][197]

template <bool NHWC>
struct SinkStage {
  float* s;  // NCHW: s + level * 49 * 36 + pixel;  NHWC: s + pixel * 197 + level * 49
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void emit_row(int j, const float (&o)[WIN]) {
#pragma unroll
    for (int i = 0; i < WIN; ++i) s[NHWC ? (i * WIN + j) : (i * WIN + j) * V3_PITCH_C] = o[i];
  }
};

template <bool NHWC>
__global__ void __launch_bounds__(V3_THREADS, 4) k_corr_lookup_v3(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G,
                                                                  const float* __restrict__ coords, float* __restrict__ out) {
  __shared__ __align__(16) float s_stage[NHWC ? V3_PIX * V3_PITCH_P : V3_NCH * V3_PITCH_C];
  const int lane = lane_id(), warp = warp_id();
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * V3_PIX;
  const int pl = warp * WARP_PIX + (lane & 3);  // pixel of the tile
  const int pix = i0 + pl;
  const bool live = pix < G.nf;
  const float cx = live ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
  const float cy = live ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
  const __nv_bfloat16* base = pyr + pixel_base(G, b, live ? pix : 0);
  const int panel_stride = G.m_tiles * 2 * 8192;
  const int n_ch = G.levels * WIN * WIN;
#pragma unroll 1
  for (int level = 0; level < G.levels; ++level) {
    const LevelGeo g = level_geo(G, level);
    SinkStage<NHWC> sink{NHWC ? s_stage + pl * V3_PITCH_P + level * WIN * WIN : s_stage + level * WIN * WIN * V3_PITCH_C + pl};
    RowState st;
    rows_prepare(st, base, panel_stride, G.pitch, g, cx, cy);
    rows_finish(st, base, panel_stride, G.pitch, g, cx, cy, live, sink);
  }
  __syncthreads();
  const int n_pix = min(V3_PIX, G.nf - i0);
  if (NHWC) {
    float* blk = out + ((size_t)b * G.nf + i0) * n_ch;  // n_pix * n_ch contiguous floats
    for (int pp = warp; pp < n_pix; pp += V3_THREADS / 32)
      for (int k = lane; k < n_ch; k += 32) blk[(size_t)pp * n_ch + k] = s_stage[pp * V3_PITCH_P + k];
  } else {
    float* dst = out + (size_t)b * n_ch * G.nf + i0;
    for (int k = warp; k < n_ch; k += V3_THREADS / 32)
      if (lane < n_pix) dst[(size_t)k * G.nf + lane] = s_stage[k * V3_PITCH_C + lane];
  }
}

// ------------------------------------------------------------------------------------------ fused lookup + 1x1 conv
constexpr int F_PIX = 128;                  // pixels per tile = MMA M
constexpr int F_PASSES = 2;                 // a tile is gathered in two passes of 64 pixels
constexpr int F_PASS_PIX = F_PIX / F_PASSES;
constexpr int F_THREADS = F_PASS_PIX * ROWS;  // 512: one thread per (pixel of the pass, window row)
constexpr int F_LEVELS = 4;
constexpr int F_K = F_LEVELS * KPL;         // 224
constexpr int F_KBLK = 32;                  // tf32 elements per 128-byte swizzle row
constexpr int F_KBLOCKS = F_K / F_KBLK;     // 7
constexpr int F_UMMA_K = 8;                 // tf32: 32 bytes of K per instruction
constexpr uint32_t F_A_KBLK_BYTES = F_PIX * 128;            // 16 KB
constexpr uint32_t F_A_BYTES = F_KBLOCKS * F_A_KBLK_BYTES;  // 112 KB
constexpr int F_MAX_N = 96;
constexpr uint32_t F_ACC_COLS = 128;        // TMEM columns per accumulator (N <= 96 fp32 columns)
constexpr uint32_t F_TMEM_COLS = 2 * F_ACC_COLS;  // two accumulators: the MMAs of tile t overlap the epilogue of tile t - 1
// packed weights (slimb200_corr_lookup_conv_pack): the B operand exactly as it sits in shared memory -- 7 K blocks of
// N rows x 128 bytes (32 tf32 values, 128-byte swizzle) -- in the K-slot order of each fused kernel (image 0: level * 56 +
// j * 8 + i for the row-per-thread kernel of this file, image 1: level * 56 + i * 7 + j for csrc/corr_lookup4.cu),
// followed by the N fp32 biases
__host__ __device__ constexpr uint32_t packed_w_bytes(int n) { return (uint32_t)F_KBLOCKS * (uint32_t)n * 128u; }
__host__ __device__ constexpr uint32_t packed_bytes(int n) { return 2u * packed_w_bytes(n) + (uint32_t)n * 4u; }
__host__ __device__ constexpr uint32_t fused_smem_bytes(int n) {
  return F_A_BYTES + packed_w_bytes(n) + (uint32_t)n * 4u + 64u /*barriers + tmem ptr*/ + 1024u /*align*/;
}
static_assert(fused_smem_bytes(F_MAX_N) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// image 0: K slot l * 56 + j * 8 + i of row n holds W[n][l * 49 + i * 7 + j] (i < 7) or 0 (i == 7); image 1: K slot
// l * 56 + c holds W[n][l * 49 + c] (c < 49) or 0; both rounded to tf32
__global__ void __launch_bounds__(256) k_lookup_conv_pack(const float* __restrict__ weight, const float* __restrict__ bias, int N,
                                                          uint8_t* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < N * F_K) {
    const int n = idx / F_K, kk = idx - n * F_K;
    const int l = kk / KPL, c = kk - l * KPL;
    const int j = c >> 3, i = c & 7;
    const float* wrow = weight + (size_t)n * (F_LEVELS * WIN * WIN) + l * WIN * WIN;
    const float v0 = i < WIN ? __ldg(wrow + i * WIN + j) : 0.f;
    const float v1 = c < WIN * WIN ? __ldg(wrow + c) : 0.f;
    const uint32_t off = (uint32_t)(kk >> 5) * ((uint32_t)N * 128u) + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                         (((uint32_t)((kk & 31) >> 2) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(kk & 3) * 4u;
    *reinterpret_cast<uint32_t*>(packed + off) = to_tf32(v0);
    *reinterpret_cast<uint32_t*>(packed + packed_w_bytes(N) + off) = to_tf32(v1);
  }
  if (idx < N) reinterpret_cast<float*>(packed + 2u * packed_w_bytes(N))[idx] = bias ? __ldg(bias + idx) : 0.f;
}

// A-operand sink: lane (p, j) owns chunks 2j, 2j + 1 of the level's 14 chunks in row `pixel` of the swizzled K-major tile
struct SinkA {
  uint32_t a_row;     // smem address of this pixel's row inside K block 0 (row / 8 * 1024 + row % 8 * 128)
  uint32_t swz;       // row % 8
  int q0;             // first 16-byte chunk of this level: level * 14
  uint32_t wait_bar;  // mbarrier of the MMAs that still read the A tile (0: none)
  uint32_t wait_parity;
  __device__ __forceinline__ void begin() {  // right before the first store: the previous tile's MMAs must be through
    if (wait_bar) mbar_wait(wait_bar, wait_parity);
  }
  __device__ __forceinline__ void chunk(uint32_t q, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    const uint32_t addr = a_row + (q >> 3) * F_A_KBLK_BYTES + (((q & 7u) ^ swz) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
  }
  __device__ __forceinline__ void emit_row(int j, const float (&o)[WIN]) {
    const uint32_t q = (uint32_t)(q0 + 2 * j);
    chunk(q, to_tf32(o[0]), to_tf32(o[1]), to_tf32(o[2]), to_tf32(o[3]));
    chunk(q + 1u, to_tf32(o[4]), to_tf32(o[5]), to_tf32(o[6]), 0u);
  }
};

// accumulator rows -> + bias -> ReLU -> global, one thread per pixel row (warps 0..3 = TMEM lane quarters 0..3)
__device__ __forceinline__ void fused_epilogue(uint32_t tmem_acc, int warp, int lane, const float* s_bias, int N, int relu,
                                               float* __restrict__ tile_out, int out_pitch, int rows_live) {
  const int row = warp * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16);
  float* const dst = tile_out + (size_t)row * out_pitch;
  for (int cb = 0; cb < (N >> 5); ++cb) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + (uint32_t)(cb * 32), v);
    tmem_ld_wait();
    if (row < rows_live) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 o;
        o.x = __uint_as_float(v[c * 4 + 0]) + s_bias[cb * 32 + c * 4 + 0];
        o.y = __uint_as_float(v[c * 4 + 1]) + s_bias[cb * 32 + c * 4 + 1];
        o.z = __uint_as_float(v[c * 4 + 2]) + s_bias[cb * 32 + c * 4 + 2];
        o.w = __uint_as_float(v[c * 4 + 3]) + s_bias[cb * 32 + c * 4 + 3];
        if (relu) {
          o.x = fmaxf(o.x, 0.f);
          o.y = fmaxf(o.y, 0.f);
          o.z = fmaxf(o.z, 0.f);
          o.w = fmaxf(o.w, 0.f);
        }
        *reinterpret_cast<float4*>(dst + cb * 32 + c * 4) = o;
      }
    }
  }
}

// Software pipeline: a thread works on the "units" (tile, pass, level) of its CTA in order -- 2 passes of 64 pixels x 4
// levels per 128-pixel tile -- and keeps the window loads of FOUR units in flight (one RowState per level): unit u is
// finished (blended, written to the A tile) while the loads of units u + 1 .. u + 3 travel, then the unit that will reuse
// its state (same level, next pass or next tile) is prepared.  The DRAM latency of the gather is therefore paid once at
// kernel start, not once per level.  Per tile: ONE __syncthreads; one thread issues the 28 MMAs of the tile into
// accumulator (tile & 1); warps 0..3 then write out the PREVIOUS tile (its MMAs finished long ago) while the tensor
// core works; the first A store of the next tile waits for this tile's MMAs (an mbarrier that has normally fired).
struct TileCtx {  // what a lane needs to know about its two pixels (one per pass) of a tile
  float cx[F_PASSES], cy[F_PASSES];
  uint32_t base[F_PASSES];  // element offset of the pixel's pyramid rows (fits 32 bits: make_geo checks)
  bool live[F_PASSES];
};

__device__ __forceinline__ TileCtx load_tile_ctx(const LookupGeo& G, const float* __restrict__ coords, int tile, int prow0, int n_tiles) {
  TileCtx c;
  const bool tile_ok = tile < n_tiles;
  const int b = tile_ok ? tile / G.m_tiles : 0, mt = tile_ok ? tile - b * G.m_tiles : 0;
#pragma unroll
  for (int q = 0; q < F_PASSES; ++q) {
    const int pix = mt * F_PIX + q * F_PASS_PIX + prow0;
    c.live[q] = tile_ok && pix < G.nf;
    c.cx[q] = c.live[q] ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
    c.cy[q] = c.live[q] ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
    c.base[q] = (uint32_t)pixel_base(G, b, c.live[q] ? pix : 0);
  }
  return c;
}

__global__ void __launch_bounds__(F_THREADS, 1)
k_lookup_conv_tf32(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G, const float* __restrict__ coords,
                   const uint8_t* __restrict__ packed, float* __restrict__ out, int out_pitch, int N, int relu, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_base + F_A_BYTES;
  const uint32_t w_kblk_bytes = (uint32_t)N * 128u;
  const float* const s_bias = reinterpret_cast<const float*>(gen_base + F_A_BYTES + packed_w_bytes(N));
  const uint32_t bar0 = smem_w + packed_w_bytes(N) + (uint32_t)N * 4u;  // 8-byte aligned (N % 32 == 0)
  auto mma_bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t w_bar = bar0 + 16u;
  const uint32_t tmem_ptr_smem = bar0 + 24u;

  const int lane = lane_id(), warp = warp_id();
  if (threadIdx.x == 0) {
    mbar_init(mma_bar(0), 1);
    mbar_init(mma_bar(1), 1);
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // weights + biases: one bulk copy of the packed image (L2 -> shared memory), under the first tile's gather
    mbar_expect_tx(w_bar, packed_w_bytes(N) + (uint32_t)N * 4u);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_w), "l"(packed),
                 "r"(packed_w_bytes(N)), "r"(w_bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_w + packed_w_bytes(N)),
                 "l"(packed + 2u * packed_w_bytes(N)), "r"((uint32_t)N * 4u), "r"(w_bar)
                 : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "r"(F_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(F_PIX >> 4) << 24);

  const int prow0 = warp * WARP_PIX + (lane & 3);  // pixel (= A row) of pass 0 this lane works for; pass q: + 64 q
  const int panel_stride = G.m_tiles * 2 * 8192;
  const uint32_t a_row0 = smem_a + (uint32_t)(prow0 >> 3) * 1024u + (uint32_t)(prow0 & 7) * 128u;
  const LevelGeo g0 = level_geo(G, 0), g1 = level_geo(G, 1), g2 = level_geo(G, 2), g3 = level_geo(G, 3);
  RowState s0, s1, s2, s3;  // units in flight, one per level

  int tile = blockIdx.x;
  TileCtx cur = load_tile_ctx(G, coords, tile, prow0, n_tiles);
  rows_prepare(s0, pyr + cur.base[0], panel_stride, G.pitch, g0, cur.cx[0], cur.cy[0]);
  rows_prepare(s1, pyr + cur.base[0], panel_stride, G.pitch, g1, cur.cx[0], cur.cy[0]);
  rows_prepare(s2, pyr + cur.base[0], panel_stride, G.pitch, g2, cur.cx[0], cur.cy[0]);
  rows_prepare(s3, pyr + cur.base[0], panel_stride, G.pitch, g3, cur.cx[0], cur.cy[0]);

  int it = 0;  // tiles this CTA has started
  int prev_b = 0, prev_mt = 0;
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / G.m_tiles, mt = tile - b * G.m_tiles;
    const TileCtx nxt = load_tile_ctx(G, coords, tile + gridDim.x, prow0, n_tiles);  // (consumed four units later)
#pragma unroll
    for (int q = 0; q < F_PASSES; ++q) {
      // the unit that reuses a state: same level, next pass of this tile or pass 0 of the next tile (dead lanes of a CTA
      // without a next tile fetch pixel 0 of sample 0: harmless, never finished)
      const bool last = q == F_PASSES - 1;
      const float ncx = last ? nxt.cx[0] : cur.cx[q + 1 < F_PASSES ? q + 1 : 0], ncy = last ? nxt.cy[0] : cur.cy[q + 1 < F_PASSES ? q + 1 : 0];
      const __nv_bfloat16* nbase = pyr + (last ? nxt.base[0] : cur.base[q + 1 < F_PASSES ? q + 1 : 0]);
      const __nv_bfloat16* base = pyr + cur.base[q];
      const uint32_t a_row = a_row0 + (uint32_t)q * (F_PASS_PIX / 8) * 1024u;
      SinkA sink{a_row, (uint32_t)(prow0 & 7), 0, (it > 0 && q == 0) ? mma_bar((it - 1) & 1) : 0u, (uint32_t)((it - 1) >> 1) & 1u};
      rows_finish(s0, base, panel_stride, G.pitch, g0, cur.cx[q], cur.cy[q], cur.live[q], sink);
      rows_prepare(s0, nbase, panel_stride, G.pitch, g0, ncx, ncy);
      sink.wait_bar = 0u;
      sink.q0 = KPL / 4;
      rows_finish(s1, base, panel_stride, G.pitch, g1, cur.cx[q], cur.cy[q], cur.live[q], sink);
      rows_prepare(s1, nbase, panel_stride, G.pitch, g1, ncx, ncy);
      sink.q0 = 2 * (KPL / 4);
      rows_finish(s2, base, panel_stride, G.pitch, g2, cur.cx[q], cur.cy[q], cur.live[q], sink);
      rows_prepare(s2, nbase, panel_stride, G.pitch, g2, ncx, ncy);
      sink.q0 = 3 * (KPL / 4);
      rows_finish(s3, base, panel_stride, G.pitch, g3, cur.cx[q], cur.cy[q], cur.live[q], sink);
      rows_prepare(s3, nbase, panel_stride, G.pitch, g3, ncx, ncy);
    }
    cur = nxt;
    fence_proxy_async_smem();  // the A rows were written through the generic proxy, the MMA reads through the async one
    __syncthreads();
    if (warp == 0) {
      if (it == 0) mbar_wait(w_bar, 0);  // the packed weights have landed
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(it & 1) * F_ACC_COLS;
#pragma unroll
        for (int kb = 0; kb < F_KBLOCKS; ++kb) {
          const uint64_t adesc = make_smem_desc_sw128(smem_a + (uint32_t)kb * F_A_KBLK_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(smem_w + (uint32_t)kb * w_kblk_bytes);
#pragma unroll
          for (int k = 0; k < F_KBLK / F_UMMA_K; ++k)  // + k * 8 elements * 4 B = 32 B (>> 4 = 2) inside the swizzle row
            umma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(mma_bar(it & 1));
      }
      __syncwarp();
    }
    if (warp < 4 && it > 0) {
      // the previous tile: its MMAs were complete before this tile's A stores began
      mbar_wait(w_bar, 0);
      mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
      tcgen05_fence_after();
      fused_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * F_ACC_COLS, warp, lane, s_bias, N, relu,
                     out + ((size_t)prev_b * G.nf + (size_t)prev_mt * F_PIX) * (size_t)out_pitch, out_pitch,
                     min(F_PIX, G.nf - prev_mt * F_PIX));
      tcgen05_fence_before();  // (ordered before the MMAs of tile it + 1 by the next __syncthreads)
    }
    prev_b = b;
    prev_mt = mt;
  }
  if (warp < 4 && it > 0) {  // drain: the last tile
    mbar_wait(w_bar, 0);
    mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
    tcgen05_fence_after();
    fused_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * F_ACC_COLS, warp, lane, s_bias, N, relu,
                   out + ((size_t)prev_b * G.nf + (size_t)prev_mt * F_PIX) * (size_t)out_pitch, out_pitch,
                   min(F_PIX, G.nf - prev_mt * F_PIX));
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F_TMEM_COLS) : "memory");
  }
}

}  // namespace

// radius-3 lookup on a bf16 pyramid, row-per-thread gather (called by slimb200_corr_lookup)
int slimb200_lookup_v3_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, float* out, int out_layout,
                              cudaStream_t stream) {
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  dim3 grid((G.nf + V3_PIX - 1) / V3_PIX, L->batch);
  if (out_layout == SLIMB200_CANVAS_NHWC)
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v3<true><<<grid, V3_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  else
    SLIMB200_LAUNCH(SLIMB200_K_CORR_LOOKUP, stream,
                    (k_corr_lookup_v3<false><<<grid, V3_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(pyramid), G, coords, out)));
  return SLIMB200_OK;
}

// csrc/corr_lookup4.cu
int slimb200_lookup_conv_tmem_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, const void* packed_w,
                                     const float* packed_bias, int c_out, int relu, float* out, int out_pitch, cudaStream_t stream);

namespace {
int g_lookup_conv_generation = 4;
}

extern "C" int slimb200_lookup_conv_generation(int32_t generation) {
  const int prev = g_lookup_conv_generation;
  if (generation >= 0) g_lookup_conv_generation = generation;
  return prev;
}

extern "C" size_t slimb200_corr_lookup_conv_packed_bytes(int32_t c_out) {
  return (c_out < 32 || c_out > F_MAX_N || (c_out & 31)) ? 0 : packed_bytes(c_out);
}

extern "C" int slimb200_corr_lookup_conv_pack(const float* weight, const float* bias, int32_t levels, int32_t radius, int32_t c_out,
                                              void* packed, void* stream_) {
  if (!weight || !packed) return SLIMB200_E_INVALID;
  if (levels != F_LEVELS || radius != R || c_out < 32 || c_out > F_MAX_N || (c_out & 31)) return SLIMB200_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(packed) & 15) return SLIMB200_E_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = c_out * F_K;
  SLIMB200_LAUNCH(SLIMB200_K_LOOKUP_CONV_PACK, stream,
                  (k_lookup_conv_pack<<<(n + 255) / 256, 256, 0, stream>>>(weight, bias, c_out, static_cast<uint8_t*>(packed))));
  return SLIMB200_OK;
}

extern "C" int slimb200_corr_lookup_conv(const void* pyramid, int32_t pyramid_dtype, const slimb200_corr_layout* L,
                                         const float* coords, int32_t radius, const void* packed, int32_t c_out, int32_t relu,
                                         float* out, int32_t out_pitch, void* stream_) {
  if (!pyramid || !L || !coords || !packed || !out) return SLIMB200_E_INVALID;
  if (pyramid_dtype != SLIMB200_DTYPE_BF16 || radius != R || L->levels != F_LEVELS) return SLIMB200_E_UNSUPPORTED;
  if (c_out < 32 || c_out > F_MAX_N || (c_out & 31)) return SLIMB200_E_UNSUPPORTED;
  if (out_pitch < c_out || (out_pitch & 3)) return SLIMB200_E_INVALID;
  // (the kernel keeps 32-bit element offsets into the pyramid)
  if ((unsigned long long)L->batch * L->n_panels * L->rows_padded * SLIMB200_PANEL_COLS > 0xffffffffULL) return SLIMB200_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(pyramid) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(packed) & 15))
    return SLIMB200_E_ALIGNMENT;
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (g_lookup_conv_generation >= 4) {
    const uint8_t* pk = static_cast<const uint8_t*>(packed);
    return slimb200_lookup_conv_tmem_launch(pyramid, L, coords, pk + packed_w_bytes(c_out),
                                            reinterpret_cast<const float*>(pk + 2u * packed_w_bytes(c_out)), c_out, relu, out, out_pitch,
                                            stream);
  }
  SLIMB200_DEVICE(dev, n_sm);
  static bool attr_set[SLIMB200_MAX_DEVICES] = {false};
  if (!attr_set[dev]) {
    SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_lookup_conv_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)fused_smem_bytes(F_MAX_N)));
    attr_set[dev] = true;
  }
  const int n_tiles = L->batch * G.m_tiles;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  SLIMB200_LAUNCH(SLIMB200_K_LOOKUP_CONV, stream,
                  (k_lookup_conv_tf32<<<grid, F_THREADS, fused_smem_bytes(c_out), stream>>>(
                      static_cast<const __nv_bfloat16*>(pyramid), G, coords, static_cast<const uint8_t*>(packed), out, out_pitch,
                      c_out, relu, n_tiles)));
  return SLIMB200_OK;
}
