// The lookup fused with the 1x1 convolution that consumes it, fourth generation: the A operand of the
// tcgen05 MMA lives in TENSOR MEMORY, the window loads land in shared memory through cp.async one tile ahead.
//
//   c = relu(conv1x1(lookup(coords), W (N, 4*49), bias))      corr.py:23-46 + update.py:49,71
//
// Why this shape.  The gather is latency bound (ncu, round 2: "long scoreboard" the top stall, DRAM at 20-40 %), and the
// two earlier fused kernels lose to their own stand-alone gathers because every warp of the CTA walks through the same
// phases together: positions, loads, wait, blend, barrier, MMA.  Here
//   * a thread owns one (pixel, level) unit per 128-pixel tile, like the second-generation gather (the cheapest in
//     instructions: ~1500 per unit against ~3200 for the row-per-thread kernel), 512 threads = 128 pixels x 4 levels
//   * the 16 window-row loads of a unit are cp.async copies (16 bytes, L2 -> shared memory, no register landing zone)
//     into a private 256-byte slot per thread, issued ONE TILE AHEAD in two halves of 4 rows (as soon as the half slot
//     has been read out): while tile t is blended the loads of tile t + 1 travel, so the DRAM latency is paid once per
//     kernel, not once per tile
//   * the 128 x 224 tf32 A tile is never in shared memory: each thread stores its 49 values (+ 7 zeros) to its own
//     TMEM lane with tcgen05.st (lane = pixel, column = K slot level * 56 + i * 7 + j) and the MMA takes A from tensor
//     memory (tcgen05.mma with a TMEM A operand) -- that is what frees the 112 KB the landing slots need
//   * weights: N x 224 tf32 in the 128-byte-swizzled K-major layout, one bulk copy per CTA, resident (84 KB at N = 96)
//   * two fp32 accumulators in TMEM: warps 0..3 write out tile t - 1 (+ bias, ReLU, 16-byte row stores) while the tensor
//     core works on tile t
//   * no CTA-wide barrier per tile: threads arrive on an mbarrier when their A columns are stored and move on; only the
//     MMA-issuing warp waits for the 512 arrivals
// TMEM columns: [0, 96) accumulator 0, [128, 224) accumulator 1, [256, 480) A.
#include "lookup_core.cuh"
#include "ptx.cuh"

namespace {

using namespace slimb200_lookup;
using namespace slimb200_ptx;

constexpr int G_PIX = 128;                  // pixels per tile = MMA M = TMEM lanes
constexpr int G_LEVELS = 4;
constexpr int G_THREADS = G_PIX * G_LEVELS;  // 512: one thread per (pixel, level)
constexpr int G_K = G_LEVELS * KPL;          // 224
constexpr int G_KBLK = 32, G_KBLOCKS = G_K / G_KBLK, G_UMMA_K = 8;
constexpr int G_MAX_N = 96;
constexpr uint32_t G_ACC_STRIDE = 128, G_A_COL0 = 256, G_TMEM_COLS = 512;
constexpr uint32_t G_SLOT_BYTES = 8 * 2 * 16;  // 8 window rows x 2 chunks of 16 bytes per thread
constexpr uint32_t G_ZONE_BYTES = G_THREADS * G_SLOT_BYTES;  // 128 KB
__host__ __device__ constexpr uint32_t g_w_bytes(int n) { return (uint32_t)G_KBLOCKS * (uint32_t)n * 128u; }
__host__ __device__ constexpr uint32_t g_smem_bytes(int n) {
  return G_ZONE_BYTES + g_w_bytes(n) + (uint32_t)n * 4u + 64u /*barriers + tmem ptr*/ + 1024u /*align*/;
}
static_assert(g_smem_bytes(G_MAX_N) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  __syncwarp();  // warp-collective instruction: the lanes may come out of a divergent region (slow-path sampling)
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem descriptor], tf32 operands
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// chunk c of window row r of this thread's slot: + (r * 2 + c) * G_THREADS * 16 (a warp reads 512 contiguous bytes)
__device__ __forceinline__ void load_slot_row(uint32_t my_zone, int r, uint32_t (&raw)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(raw[c * 4 + 0]), "=r"(raw[c * 4 + 1]), "=r"(raw[c * 4 + 2]), "=r"(raw[c * 4 + 3])
                 : "r"(my_zone + (uint32_t)((r * 2 + c) * G_THREADS) * 16u));
}

// masked weights, window origin and per-offset shift bits of one axis in their compact form: the weight of tap floor + 1
// per offset (w_hi = ip - floor; the weight of tap floor is 1 - w_hi EXACTLY: both differences are exact in fp32 and w_hi
// is a multiple of ulp(ip)) and one bit per tap that says whether it lies inside the level
__device__ __forceinline__ void axis_taps_compact(float c, float inv, int size, float (&whi)[WIN], unsigned& in_bits, int& origin,
                                                  unsigned& shift_bits, bool& ok) {
  int f[WIN];
  origin = INT_MAX;
  in_bits = 0u;
  const float sm1 = (float)(size - 1);
  const float rinv = __frcp_rn(sm1);
#pragma unroll
  for (int o = 0; o < WIN; ++o) {
    const float ip = sample_pos2(c, inv, o - R, sm1, rinv);
    const float fl = floorf(ip);
    const int i0 = (int)fl;
    whi[o] = __fsub_rn(ip, fl);
    in_bits |= ((unsigned)i0 < (unsigned)size ? 1u : 0u) << o;
    in_bits |= ((unsigned)(i0 + 1) < (unsigned)size ? 1u : 0u) << (WIN + o);
    f[o] = i0 - o;
    origin = min(origin, f[o]);
  }
  shift_bits = 0u;
  ok = true;
#pragma unroll
  for (int o = 0; o < WIN; ++o) {
    const int d = f[o] - origin;
    ok = ok && d <= 1;
    shift_bits |= (unsigned)(d & 1) << o;
  }
}
__device__ __forceinline__ float tap_w0(float whi, unsigned in_bits, int o) {  // weight of tap floor (ix_se - ix), masked
  return (in_bits >> o) & 1u ? __fsub_rn(1.f, whi) : 0.f;
}
__device__ __forceinline__ float tap_w1(float whi, unsigned in_bits, int o) {  // weight of tap floor + 1 (ix - ix_nw), masked
  return (in_bits >> (WIN + o)) & 1u ? whi : 0.f;
}

struct Unit {  // what a thread keeps about its (pixel, level) of a tile between the load issue and the blend
  float wx[WIN], wy[WIN];  // weight of tap floor + 1 per offset
  unsigned inx, iny;       // tap-inside-the-level bits: bit o = tap floor, bit 7 + o = tap floor + 1
  float cx, cy;
  uint32_t base;   // element offset of the pixel's pyramid rows
  int row0;        // pyramid column of window element (0, 0)
  unsigned sx, sy;
  int mode;        // 1 regular, 2 shifted, 0 slow
  bool live;
};

// TMEM sink: 8 values per tcgen05.st into this thread's lane; k arrives in ascending order; warp-convergent by construction
struct SinkT {
  uint32_t taddr;  // lane base + first column of this level
  uint32_t pend[8];
  __device__ __forceinline__ void emit(int k, float v) {
    pend[k & 7] = to_tf32(v);
    if ((k & 7) == 7) tmem_st_x8(taddr + (uint32_t)(k & ~7), pend);
    if (k == WIN * WIN - 1) {  // k = 48: first value of the last group of 8, the other 7 slots are padding
      const uint32_t last[8] = {pend[0], 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st_x8(taddr + 48u, last);
    }
  }
};

// accumulator rows -> + bias -> ReLU -> global, one thread per pixel row (warps 0..3 = TMEM lane quarters 0..3)
__device__ __forceinline__ void g_epilogue(uint32_t tmem_acc, int warp, int lane, const float* s_bias, int N, int relu,
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

__global__ void __launch_bounds__(G_THREADS, 1)
k_lookup_conv_tmem(const __nv_bfloat16* __restrict__ pyr, const LookupGeo G, const float* __restrict__ coords,
                   const uint8_t* __restrict__ packed_w, const float* __restrict__ packed_bias, float* __restrict__ out, int out_pitch,
                   int N, int relu, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t smem_w = smem_base;                       // weights first: K blocks need 1024-byte alignment
  const uint32_t w_kblk_bytes = (uint32_t)N * 128u;
  const uint32_t smem_bias = smem_w + g_w_bytes(N);
  const float* const s_bias = reinterpret_cast<const float*>(gen_base + g_w_bytes(N));
  const uint32_t smem_zone = smem_bias + (uint32_t)N * 4u;  // 16-byte aligned (N % 32 == 0)
  const uint32_t bar0 = smem_zone + G_ZONE_BYTES;
  auto mma_bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t w_bar = bar0 + 16u;
  const uint32_t full_bar = bar0 + 24u;  // every thread arrives when its A columns of the tile are in tensor memory
  const uint32_t tmem_ptr_smem = bar0 + 32u;

  const int lane = lane_id(), warp = warp_id();
  if (threadIdx.x == 0) {
    mbar_init(mma_bar(0), 1);
    mbar_init(mma_bar(1), 1);
    mbar_init(w_bar, 1);
    mbar_init(full_bar, G_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(w_bar, g_w_bytes(N) + (uint32_t)N * 4u);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_w), "l"(packed_w),
                 "r"(g_w_bytes(N)), "r"(w_bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_bias),
                 "l"(packed_bias), "r"((uint32_t)N * 4u), "r"(w_bar)
                 : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "r"(G_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(G_PIX >> 4) << 24);

  const int level = warp >> 2;
  const int prow = (warp & 3) * 32 + lane;  // pixel of the tile = TMEM lane (lane quarter warp % 4: the one this warp may access)
  const int W = pick4(G.lw, level), H = pick4(G.lh, level), off = pick4(G.lo, level);
  const float inv = 1.0f / (float)(1 << level);
  const int panel_stride = G.m_tiles * 2 * 8192;
  const uint32_t my_zone = smem_zone + (uint32_t)threadIdx.x * 16u;  // chunk c of row r at + (r * 2 + c) * G_THREADS * 16
  const uint32_t my_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + G_A_COL0 + (uint32_t)(level * KPL);

  // coordinates of this thread's pixel in `tile` (prefetched one tile before they are needed)
  struct Coord {
    float cx, cy;
    uint32_t base;
    bool live;
  };
  auto load_coord = [&](int tile) {
    Coord c;
    const bool tile_ok = tile < n_tiles;
    const int b = tile_ok ? tile / G.m_tiles : 0, mt = tile_ok ? tile - b * G.m_tiles : 0;
    const int pix = mt * G_PIX + prow;
    c.live = tile_ok && pix < G.nf;
    c.cx = c.live ? __ldg(coords + ((size_t)b * 2 + 0) * G.nf + pix) : 0.f;
    c.cy = c.live ? __ldg(coords + ((size_t)b * 2 + 1) * G.nf + pix) : 0.f;
    c.base = (uint32_t)pixel_base(G, b, c.live ? pix : 0);
    return c;
  };
  // positions, weights, window origin of a unit
  auto taps = [&](Unit& u, const Coord& c) {
    u.live = c.live;
    u.cx = c.cx;
    u.cy = c.cy;
    u.base = c.base;
    int xb, yb;
    bool okx, oky;
    axis_taps_compact(u.cx, inv, W, u.wx, u.inx, xb, u.sx, okx);
    axis_taps_compact(u.cy, inv, H, u.wy, u.iny, yb, u.sy, oky);
    u.mode = (okx && oky) ? ((u.sx | u.sy) ? 2 : 1) : 0;
    // origins beyond +-2^18 (every tap outside, zero weights) are clamped so that the index stays an int
    u.row0 = off + max(min(yb, 1 << 18), -(1 << 18)) * W + max(min(xb, 1 << 18), -(1 << 18));
  };
  // the cp.async copies of window rows 4 * half .. 4 * half + 3 (one commit group): all 8 addresses first, then the 8
  // copies back to back -- a copy whose address register is recycled right behind it stalls the warp until the LSU has
  // taken the copy (the first version of this kernel spent 60 % of its time there)
  auto issue = [&](const Unit& u, int half) {
    const __nv_bfloat16* base = pyr + u.base;
    const __nv_bfloat16* src[8];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int ca = (u.row0 + (half * 4 + rr) * W) & ~7;
#pragma unroll
      for (int c = 0; c < 2; ++c)  // (whatever finite value lies outside the level meets a zero weight)
        src[rr * 2 + c] = base + col_offset(min(max(ca + 8 * c, 0), G.pitch - 8), panel_stride);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cp_async16(my_zone + (uint32_t)((half * 8 + k) * G_THREADS) * 16u, src[k]);
    cp_async_commit();
  };

  Unit cur;
  int tile = blockIdx.x;
  taps(cur, load_coord(tile));
  issue(cur, 0);
  issue(cur, 1);
  Coord ahead = load_coord(tile + gridDim.x);
  int it = 0;  // tiles this CTA has started
  int prev_b = 0, prev_mt = 0;
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / G.m_tiles, mt = tile - b * G.m_tiles;
    // ---- this tile's window: shared memory -> registers, aligned to window column 0; then the slot is free and the next
    // tile's loads start their journey under this tile's blend; then the blend -> TMEM lane of this pixel (the previous
    // tile's MMAs must have read the A columns by then) ----
    const bool any_slow = __any_sync(FULL, cur.live && cur.mode == 0);
    const bool any_shift = __any_sync(FULL, cur.live && cur.mode == 2);
    Unit nxt;
    taps(nxt, ahead);                                 // (nothing here touches shared memory: it overlaps the wait below)
    ahead = load_coord(tile + 2 * gridDim.x);
    SinkT sink{my_taddr, {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}};
    if (!any_slow && !any_shift) {
      // regular windows: 8 x 8, two taps per axis
      uint32_t win[8][4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        cp_async_wait_but_one();  // this half of the current tile has landed (the newest group may still travel)
#pragma unroll
        for (int r = half * 4; r < half * 4 + 4; ++r) {
          uint32_t raw[8];
          load_slot_row(my_zone, r, raw);
          realign<4>(raw, (cur.row0 + r * W) & 7, win[r]);
        }
        issue(nxt, half);  // the half slot is free: the next tile's rows start their journey under this tile's blend
      }
      if (it > 0) mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
      tcgen05_fence_after();
      float e0[8], e1[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) e0[r] = wel<4>(win[r], 0);
#pragma unroll
      for (int i = 0; i < WIN; ++i) {
        const float wx0 = tap_w0(cur.wx[i], cur.inx, i), wx1 = tap_w1(cur.wx[i], cur.inx, i);
        float h[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          e1[r] = wel<4>(win[r], i + 1);
          h[r] = fmaf(e1[r], wx1, e0[r] * wx0);
          e0[r] = e1[r];
        }
#pragma unroll
        for (int j = 0; j < WIN; ++j)
          sink.emit(i * WIN + j, fmaf(h[j + 1], tap_w1(cur.wy[j], cur.iny, j), h[j] * tap_w0(cur.wy[j], cur.iny, j)));
      }
    } else {
      // 9 x 9 window, three taps per axis, one weight of the three is zero (first GRU iteration / odd coordinates);
      // window row 8 comes straight from global memory
      uint32_t win[9][5];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        cp_async_wait_but_one();
#pragma unroll
        for (int r = half * 4; r < half * 4 + 4; ++r) {
          uint32_t raw[8];
          load_slot_row(my_zone, r, raw);
          realign<5>(raw, (cur.row0 + r * W) & 7, win[r]);
        }
        issue(nxt, half);
      }
      {
        uint32_t raw8[8];
        const int sft8 = fetch_row(pyr + cur.base, panel_stride, G.pitch, cur.row0 + 8 * W, raw8);
        realign<5>(raw8, sft8, win[8]);
      }
      if (it > 0) mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
      tcgen05_fence_after();
      const bool lane_slow = cur.live && cur.mode == 0;
      const float swm1 = (float)(W - 1), shm1 = (float)(H - 1);
      const float rw = __frcp_rn(swm1), rh = __frcp_rn(shm1);
#pragma unroll
      for (int i = 0; i < WIN; ++i) {
        const bool s = (cur.sx >> i) & 1u;
        const float wx0 = tap_w0(cur.wx[i], cur.inx, i), wx1 = tap_w1(cur.wx[i], cur.inx, i);
        const float a = s ? 0.f : wx0, bq = s ? wx0 : wx1, c = s ? wx1 : 0.f;
        float h[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) h[r] = fmaf(wel<5>(win[r], i + 2), c, fmaf(wel<5>(win[r], i + 1), bq, wel<5>(win[r], i) * a));
        const float ix = lane_slow ? sample_pos2(cur.cx, inv, i - R, swm1, rw) : 0.f;
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
          const bool t = (cur.sy >> j) & 1u;
          const float wy0 = tap_w0(cur.wy[j], cur.iny, j), wy1 = tap_w1(cur.wy[j], cur.iny, j);
          const float ay = t ? 0.f : wy0, by = t ? wy0 : wy1, cyw = t ? wy1 : 0.f;
          float v = fmaf(h[j + 2], cyw, fmaf(h[j + 1], by, h[j] * ay));
          if (lane_slow)  // non-finite / absurd coordinates: predicated 4-tap loads (the TMEM store stays convergent)
            v = sample_slow2(pyr + cur.base, panel_stride, W, H, off, ix, sample_pos2(cur.cy, inv, j - R, shm1, rh));
          sink.emit(i * WIN + j, v);
        }
      }
    }
    // No CTA-wide barrier: a thread whose A columns are stored ARRIVES and moves on to the next tile (positions, slot
    // read-out, load issue) while slower warps finish; only warp 0, which issues the MMAs, waits for all 512 arrivals.
    tmem_st_wait();
    tcgen05_fence_before();
    mbar_arrive(full_bar);
    if (warp == 0) {
      if (it == 0) mbar_wait(w_bar, 0);  // the packed weights have landed
      mbar_wait(full_bar, (uint32_t)it & 1u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(it & 1) * G_ACC_STRIDE;
#pragma unroll
        for (int kb = 0; kb < G_KBLOCKS; ++kb) {
          const uint64_t bdesc = make_smem_desc_sw128(smem_w + (uint32_t)kb * w_kblk_bytes);
#pragma unroll
          for (int k = 0; k < G_KBLK / G_UMMA_K; ++k)  // A: 8 columns per step; B: + 32 bytes (>> 4 = 2) inside the swizzle row
            umma_tf32_ts(tmem_d, tmem_base + G_A_COL0 + (uint32_t)((kb * 4 + k) * G_UMMA_K), bdesc + (uint64_t)(2 * k), idesc,
                         (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(mma_bar(it & 1));
      }
      __syncwarp();
    }
    if (warp < 4 && it > 0) {
      // the previous tile: its MMAs were complete before this tile's A stores began
      mbar_wait(w_bar, 0);
      tcgen05_fence_after();
      g_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * G_ACC_STRIDE, warp, lane, s_bias, N, relu,
                 out + ((size_t)prev_b * G.nf + (size_t)prev_mt * G_PIX) * (size_t)out_pitch, out_pitch, min(G_PIX, G.nf - prev_mt * G_PIX));
      tcgen05_fence_before();  // (ordered before the MMAs of tile it + 1 by this thread's next arrival on full_bar)
    }
    prev_b = b;
    prev_mt = mt;
    cur = nxt;
  }
  cp_async_wait_all();  // (the loads issued for a tile that does not exist)
  if (warp < 4 && it > 0) {  // drain: the last tile
    mbar_wait(w_bar, 0);
    mbar_wait(mma_bar((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
    tcgen05_fence_after();
    g_epilogue(tmem_base + (uint32_t)((it - 1) & 1) * G_ACC_STRIDE, warp, lane, s_bias, N, relu,
               out + ((size_t)prev_b * G.nf + (size_t)prev_mt * G_PIX) * (size_t)out_pitch, out_pitch, min(G_PIX, G.nf - prev_mt * G_PIX));
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_TMEM_COLS) : "memory");
  }
}

}  // namespace

// called by slimb200_corr_lookup_conv (csrc/corr_lookup3.cu); packed_w: the B image in K-slot order level * 56 + i * 7 + j
int slimb200_lookup_conv_tmem_launch(const void* pyramid, const slimb200_corr_layout* L, const float* coords, const void* packed_w,
                                     const float* packed_bias, int c_out, int relu, float* out, int out_pitch, cudaStream_t stream) {
  LookupGeo G;
  int rc = make_geo(L, &G);
  if (rc != SLIMB200_OK) return rc;
  if (c_out > G_MAX_N) return SLIMB200_E_UNSUPPORTED;
  SLIMB200_DEVICE(dev, n_sm);
  static bool attr_set[SLIMB200_MAX_DEVICES] = {false};
  if (!attr_set[dev]) {
    SLIMB200_CUDA_TRY(cudaFuncSetAttribute(k_lookup_conv_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_bytes(G_MAX_N)));
    attr_set[dev] = true;
  }
  const int n_tiles = L->batch * G.m_tiles;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  SLIMB200_LAUNCH(SLIMB200_K_LOOKUP_CONV, stream,
                  (k_lookup_conv_tmem<<<grid, G_THREADS, g_smem_bytes(c_out), stream>>>(
                      static_cast<const __nv_bfloat16*>(pyramid), G, coords, static_cast<const uint8_t*>(packed_w), packed_bias, out,
                      out_pitch, c_out, relu, n_tiles)));
  return SLIMB200_OK;
}
