// Export writer on the GPU.  The reference saves the BEV maps of every sample with np.savez_compressed
// (liso/slim/experiment.py:459-471): zlib, on the CPU thread that also drives the GPU.  Here every saved array becomes a
// raw DEFLATE stream (RFC 1951) + the CRC-32 of its bytes ON THE DEVICE; only the compressed bytes cross PCIe and the
// host merely frames them as zip members (liso_b200/slim/npz_stream.py), so that np.load (torch_dataset_commons.py:
// 614-616) reads the files unchanged.
//
// The maps are fp32 and constant wherever the pillar is empty (head_decoder.py:567-609: flow exactly 0, dynamicness the
// softmax of (0, -100, -100) = one denormal): a word that equals the word before it continues a RUN, and a run becomes
// <match distance 4, length <= 258>...; every other word is four literals (a zero word: literal 0 + <distance 1, length 3>),
// all in the FIXED Huffman code (RFC 1951 3.2.6), so that the bit length of every token is known without a histogram pass.
//   * a member (one array of one sample) is cut into chunks of 8 KB; one CTA encodes one chunk as ONE non-final fixed
//     block followed by an empty stored block, which byte-aligns the stream (the zlib "sync flush" marker; the last
//     chunk's stored block carries BFINAL).  Chunks are therefore byte strings that concatenate.
//   * thread = 4 consecutive words.  Runs are found with a block-wide prefix max (last run-breaking word before me) and
//     suffix min (first run-breaking word after me); the tokens of a run are a function of the byte offset from its start,
//     so every thread emits exactly the tokens that START inside its 16 bytes -- no token crosses a thread boundary
//     decision.  Pass 1 counts bits (and keeps up to 128 of them in registers), a block scan turns the counts into bit
//     positions, pass 2 ORs the codes into shared memory (only a thread with more than 128 bits generates its tokens twice).
//   * CRC-32: pure remainders are linear, R(A|B) = R(A) x^(8|B|) + R(B), and R(zero words) = 0.  The kernel folds
//     data ^ background (background = the member's first word repeated; zero almost everywhere, also for the dynamicness
//     maps): a thread with foreground words folds them with the byte table, multiplies by x^(bits behind it in the chunk)
//     (table), the CTA XORs, one thread multiplies by x^(bits behind the chunk in the member) and XORs into the member's
//     accumulator; the member's first chunk adds R(background) = R(word) * sum_i x^(32 i) (the geometric factor comes
//     with the member).  The host adds the constant crc32(npy header | zeros) of the member's shape:
//     crc(header | data) = that ^ R(data).
//   * k_deflate_scan turns chunk sizes into offsets, k_deflate_gather packs the chunks into the output stream.
#include "common.cuh"

namespace {
constexpr int CHUNK_BYTES = SLIMB200_DEFLATE_CHUNK_BYTES;
constexpr int CHUNK_WORDS = CHUNK_BYTES / 4;
constexpr int THREADS = CHUNK_WORDS / 4;  // 512: one uint4 of input per thread
constexpr int WARPS = THREADS / 32;
constexpr int SLOT_BYTES = SLIMB200_DEFLATE_SLOT_BYTES;
constexpr int SLOT_WORDS = SLOT_BYTES / 4;
constexpr uint32_t POLY = 0xEDB88320u;
constexpr int TAB_T = 0;                          // byte table, 256 entries
constexpr int TAB_XPW = 256;                      // x^(32 j), j = 0 .. CHUNK_WORDS
constexpr int TAB_XPC = TAB_XPW + CHUNK_WORDS + 1;  // x^(8 CHUNK_BYTES j), j = 0 .. MAX_CHUNKS - 1
constexpr int TAB_WORDS = TAB_XPC + SLIMB200_DEFLATE_MAX_CHUNKS;
static_assert(TAB_WORDS * 4 == SLIMB200_DEFLATE_TABLE_BYTES, "table size");
// worst case of a chunk: 3 header bits + 9 bits per byte + end-of-block 7 + stored header 3 + padding 7 + LEN/NLEN 32
static_assert((3 + CHUNK_BYTES * 9 + 7 + 3 + 7 + 32 + 7) / 8 + 8 <= SLOT_BYTES, "slot too small");

// GF(2)[x] / P in the reflected representation zlib uses (bit 31 = x^0): a * b mod P
__host__ __device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    p ^= b & (0u - ((a >> (31 - i)) & 1u));
    b = (b >> 1) ^ (POLY & (0u - (b & 1u)));
  }
  return p;
}
__host__ __device__ inline uint32_t gf_xpow(uint64_t n) {  // x^n mod P
  uint32_t p = 0x80000000u, sq = 0x40000000u;
  while (n) {
    if (n & 1) p = gf_mul(sq, p);
    sq = gf_mul(sq, sq);
    n >>= 1;
  }
  return p;
}

__global__ void k_deflate_tables(uint32_t* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TAB_WORDS) return;
  if (i < TAB_XPW) {
    uint32_t c = (uint32_t)i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? POLY ^ (c >> 1) : c >> 1;
    tab[i] = c;
  } else if (i < TAB_XPC) {
    tab[i] = gf_xpow(32ull * (uint64_t)(i - TAB_XPW));
  } else {
    tab[i] = gf_xpow(8ull * CHUNK_BYTES * (uint64_t)(i - TAB_XPC));
  }
}

// ---- fixed Huffman code, bit-reversed for LSB-first packing ---------------------------------------------------------
struct Code {
  uint32_t bits;
  int n;
};
__device__ __forceinline__ Code lit_code(uint32_t v) {  // literal byte
  if (v < 144u) return {__brev(0x30u + v) >> 24, 8};
  return {__brev(0x190u + (v - 144u)) >> 23, 9};
}
__device__ __forceinline__ Code match_code(int len, uint32_t dist_sym) {  // <length, distance>: length code + extra bits + distance
  uint32_t sym, extra = 0;
  int ne = 0;
  if (len == 258) {
    sym = 285;
  } else if (len <= 10) {
    sym = 254 + len;
  } else {
    const uint32_t l = (uint32_t)len - 3u;
    ne = 29 - __clz(l);  // floor(log2 l) - 2
    sym = 257 + 4 * ne + (l >> ne);
    extra = l & ((1u << ne) - 1u);
  }
  Code c;
  if (sym < 280u) {
    c.bits = __brev(sym - 256u) >> 25;
    c.n = 7;
  } else {
    c.bits = __brev(0xC0u + (sym - 280u)) >> 24;
    c.n = 8;
  }
  c.bits |= extra << c.n;
  c.n += ne;
  c.bits |= (__brev(dist_sym) >> 27) << c.n;  // distance symbols 0..3 = distances 1..4, five bits, no extra bits
  c.n += 5;
  return c;
}
constexpr uint32_t DIST_1 = 0, DIST_4 = 3;

// A thread's tokens: at most 4 words x 4 literals x 9 bits = 144 bits (a run costs far less), kept in registers between the
// pass that counts them and the pass that writes them -- tokens are generated once.
struct BitBuffer {
  uint64_t q0 = 0, q1 = 0, q2 = 0;
  uint32_t n = 0;
  __device__ __forceinline__ void put(uint64_t bits, uint32_t len) {  // len <= 40
    const uint32_t sh = n & 63u;
    const uint64_t lo = bits << sh, hi = sh ? bits >> (64u - sh) : 0ull;
    if (n < 64u) {
      q0 |= lo;
      q1 |= hi;
    } else if (n < 128u) {
      q1 |= lo;
      q2 |= hi;
    } else {
      q2 |= lo;
    }
    n += len;
  }
  __device__ __forceinline__ void put(Code c) { put((uint64_t)c.bits, (uint32_t)c.n); }
};

// tokens of a run of `run_bytes` bytes (words equal to the word before the run) that START inside [lo, hi) (byte offsets from
// the run's start, hi - lo <= 16): matches of 258 at offsets 258 k, then the remainder r as one more match; a remainder of 1
// or 2 bytes (too short for a match) borrows 3 bytes from the last full match.  Distance 4 = one word back.
template <class Sink>
__device__ __forceinline__ void run_tokens(Sink& s, int run_bytes, int lo, int hi) {
  const int nfull = run_bytes / 258, r = run_bytes - nfull * 258;
  const int d = (r == 1 || r == 2) ? 3 : 0;
  const int k0 = (lo + 257) / 258;
  // (the two codes almost every run thread emits, as constants: <258, 4> and <255, 4>; tests/test_npz_stream.py pins them)
  if (k0 < nfull && 258 * k0 < hi) s.put((k0 == nfull - 1 && d) ? Code{0x31c23u, 18} : Code{0x18a3u, 13});
  const int q = 258 * nfull - d;
  if (r + d >= 3 && q >= lo && q < hi) s.put(match_code(r + d, DIST_4));
}

// the tokens that start inside this thread's words [i0, i0 + nv); rep bit k: word k equals the word before it;
// zb / za = run words right before / behind my words; lit[v] = code | length << 16 of literal byte v (shared memory)
__device__ __forceinline__ void thread_tokens(BitBuffer& s, const uint32_t (&w)[4], uint32_t rep, int nv, int zb, int za,
                                              const uint32_t* __restrict__ lit) {
  int i = 0;
  while (i < nv) {
    if (!((rep >> i) & 1u)) {
      const uint32_t v = w[i];
      if (v == 0u) {  // 00 00 00 00 as literal + <distance 1, length 3>: 20 bits instead of 32
        const Code m3 = match_code(3, DIST_1);
        s.put((uint64_t)(lit[0] & 0xffffu) | ((uint64_t)m3.bits << 8), 8u + (uint32_t)m3.n);
      } else {
        const uint32_t c0 = lit[v & 255u], c1 = lit[(v >> 8) & 255u], c2 = lit[(v >> 16) & 255u], c3 = lit[v >> 24];
        const uint32_t n0 = c0 >> 16, n1 = c1 >> 16, n2 = c2 >> 16, n3 = c3 >> 16;
        const uint64_t bits = (uint64_t)(c0 & 0xffffu) | ((uint64_t)(c1 & 0xffffu) << n0) | ((uint64_t)(c2 & 0xffffu) << (n0 + n1)) |
                              ((uint64_t)(c3 & 0xffffu) << (n0 + n1 + n2));
        s.put(bits, n0 + n1 + n2 + n3);
      }
      ++i;
    } else {
      const int a = i;
      while (i < nv && ((rep >> i) & 1u)) ++i;
      const int before = a == 0 ? zb : 0, behind = i == nv ? za : 0;
      run_tokens(s, 4 * (before + (i - a) + behind), 4 * before, 4 * (before + (i - a)));
    }
  }
}

__global__ void __launch_bounds__(THREADS)
k_deflate_chunks(const slimb200_deflate_member* __restrict__ members, int n_members, const uint32_t* __restrict__ tab,
                 uint32_t* __restrict__ scratch, uint32_t* __restrict__ chunk_bytes, uint32_t* __restrict__ member_out) {
  __shared__ uint32_t buf[SLOT_WORDS];
  __shared__ uint32_t T[256], LIT[256];
  __shared__ int s_last[WARPS], s_first[WARPS];
  __shared__ uint32_t s_bits[WARPS], s_crc[WARPS], s_edge[WARPS];
  __shared__ uint32_t s_total;
  __shared__ int s_member, s_nfg;
  __shared__ uint4 s_fg[THREADS];      // queue of foreground threads' words ^ background
  __shared__ uint16_t s_fgpos[THREADS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t chunk = blockIdx.x;
  if (wid == 0) {  // the member this chunk belongs to = the last one with first_chunk <= chunk (first_chunk ascends): the
                   // warp looks at 32 members per round
    int cnt = 0;
    for (int base = 0; base < n_members; base += 32) {
      const bool le = base + lane < n_members && members[base + lane].first_chunk <= chunk;
      const int n = __popc(__ballot_sync(0xffffffffu, le));
      cnt += n;
      if (n < 32) break;
    }
    if (lane == 0) {
      s_member = cnt - 1;
      s_nfg = 0;
    }
  }
  if (tid < 256) {
    T[tid] = tab[TAB_T + tid];
    const Code lc = lit_code((uint32_t)tid);
    LIT[tid] = lc.bits | ((uint32_t)lc.n << 16);
  }
  __syncthreads();
  const int mi = s_member;
  const slimb200_deflate_member m = members[mi];
  const uint32_t c = chunk - m.first_chunk, n_chunks = (m.n_words + CHUNK_WORDS - 1) / CHUNK_WORDS;
  const uint32_t w0 = c * CHUNK_WORDS;
  const int nw = (int)min((uint32_t)CHUNK_WORDS, m.n_words - w0);
  const int i0 = 4 * tid, nv = max(0, min(4, nw - i0));
  const uint32_t* src = static_cast<const uint32_t*>(m.src);
  // the member's background word (its first word): the CRC below runs over data ^ background, which is zero almost
  // everywhere -- for the dynamicness maps, whose empty cells hold a non-zero constant, as much as for the flow maps
  const uint32_t bg = __ldg(src);

  // ---- this thread's 4 words
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (nv > 0) {
    const uint32_t g = w0 + i0;
    if (m.cell_stride == m.words_per_cell && nv == 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + g));
      w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
    } else {
      const uint32_t wpc = (uint32_t)m.words_per_cell;
      uint32_t cell = wpc == 1u ? g : (wpc == 2u ? g >> 1 : g / wpc), within = g - cell * wpc;  // (one division at most)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nv) w[k] = __ldg(src + (size_t)cell * (size_t)m.cell_stride + within);
        if (++within == wpc) {
          within = 0;
          ++cell;
        }
      }
    }
  }

  // ---- which of my words repeat the word before them (the chunk's first word never does)
  if (lane == 31) s_edge[wid] = w[3];
  uint32_t prev = __shfl_up_sync(0xffffffffu, w[3], 1);
  __syncthreads();
  if (lane == 0 && wid > 0) prev = s_edge[wid - 1];
  uint32_t rep = 0;
  bool any_fg = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nv && (i0 + k) > 0 && w[k] == (k ? w[k - 1] : prev)) rep |= 1u << k;
    any_fg |= (k < nv && w[k] != bg);
  }
  // ---- run words right before / behind my words: prefix max of the last run-breaking word, suffix min of the first one
  int last = -1, first = nw;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nv && !((rep >> k) & 1u)) {
      last = i0 + k;
      if (first == nw) first = i0 + k;
    }
  int pl = last, sf = first;  // inclusive scans inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, pl, d), b = __shfl_down_sync(0xffffffffu, sf, d);
    if (lane >= d) pl = max(pl, a);
    if (lane + d < 32) sf = min(sf, b);
  }
  if (lane == 31) s_last[wid] = pl;
  if (lane == 0) s_first[wid] = sf;
  int ex_l = __shfl_up_sync(0xffffffffu, pl, 1), ex_f = __shfl_down_sync(0xffffffffu, sf, 1);
  if (lane == 0) ex_l = -1;
  if (lane == 31) ex_f = nw;
  __syncthreads();
  if (wid == 0) {  // second level: exclusive prefix max / suffix min over the warps' aggregates, by warp 0
    int a = lane < WARPS ? s_last[lane] : -1, b = lane < WARPS ? s_first[lane] : nw;
#pragma unroll
    for (int d = 1; d < WARPS; d <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, a, d), y = __shfl_down_sync(0xffffffffu, b, d);
      if (lane >= d) a = max(a, x);
      if (lane + d < WARPS) b = min(b, y);
    }
    const int ea = __shfl_up_sync(0xffffffffu, a, 1), eb = __shfl_down_sync(0xffffffffu, b, 1);
    if (lane < WARPS) {
      s_last[lane] = lane == 0 ? -1 : ea;
      s_first[lane] = lane == WARPS - 1 ? nw : eb;
    }
  }
  __syncthreads();
  ex_l = max(ex_l, s_last[wid]);
  ex_f = min(ex_f, s_first[wid]);
  const int zb = i0 - 1 - ex_l, za = ex_f - (i0 + nv);

  // ---- pass 1: my tokens -> bit count (and the bits themselves when they fit 128) -> bit position
  BitBuffer tk;
  if (nv > 0) thread_tokens(tk, w, rep, nv, zb, za, LIT);
  uint32_t inc = tk.n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += a;
  }
  if (lane == 31) s_bits[wid] = inc;

  // ---- CRC remainder of (my words ^ background), moved to the end of the chunk.  Threads with foreground words are few
  // (a couple per warp): they queue their words, and the queue is worked off by the first threads of the CTA, so that
  // whole warps run the byte-table steps and the shift by x^(bits behind the thread) instead of one or two lanes of every warp
  if (any_fg) {
    const int slot = atomicAdd(&s_nfg, 1);
    s_fg[slot] = make_uint4(w[0] ^ bg, w[1] ^ bg, w[2] ^ bg, w[3] ^ bg);
    s_fgpos[slot] = (uint16_t)(((unsigned)nv << 12) | (unsigned)(nw - (i0 + nv)));  // words, words behind them in the chunk
  }
  __syncthreads();
  uint32_t crc = 0;
  for (int e = tid; e < s_nfg; e += THREADS) {
    const uint4 f = s_fg[e];
    const unsigned meta = s_fgpos[e], nvv = meta >> 12;
    const uint32_t ww[4] = {f.x, f.y, f.z, f.w};
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((unsigned)k < nvv) {
        x ^= ww[k];
        x = T[x & 255u] ^ (x >> 8);
        x = T[x & 255u] ^ (x >> 8);
        x = T[x & 255u] ^ (x >> 8);
        x = T[x & 255u] ^ (x >> 8);
      }
    if (x) crc ^= gf_mul(__ldg(tab + TAB_XPW + (meta & 0xfffu)), x);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, d);
  if (lane == 0) s_crc[wid] = crc;
  __syncthreads();
  if (wid == 0) {  // second level: exclusive prefix sum of the warps' bit counts (+ total), XOR of their CRCs
    const uint32_t v = lane < WARPS ? s_bits[lane] : 0u;
    uint32_t a = v, x = lane < WARPS ? s_crc[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, a, d);
      if (lane >= d) a += y;
      x ^= __shfl_xor_sync(0xffffffffu, x, d);
    }
    if (lane < WARPS) s_bits[lane] = a - v;
    if (lane == 31) {
      s_total = 3u + a;
      s_crc[0] = x;
    }
  }
  __syncthreads();
  const uint32_t pos = 3u + s_bits[wid] + inc - tk.n, total = s_total;
  // block header (BFINAL 0, BTYPE 01 -> bits 0,1,0), end-of-block (7 zero bits), stored-block header (BFINAL of the member's
  // last chunk, BTYPE 00), zero padding to the byte boundary, LEN = 0x0000, NLEN = 0xFFFF
  const uint32_t nbytes = (total + 10 + 7) >> 3;
  const int nwords_out = min((int)((nbytes + 4 + 3) >> 2) + 1, SLOT_WORDS);  // (+1: the gather's funnel shift reads one word ahead)
  for (int i = tid; i < nwords_out; i += THREADS) buf[i] = 0u;  // only what this chunk's stream occupies
  __syncthreads();

  // ---- pass 2: OR the buffered bits into the chunk image
  if (tk.n) {
    const uint32_t sh = pos & 31u;
    const uint32_t wi = pos >> 5;
    const uint32_t piece[6] = {(uint32_t)tk.q0, (uint32_t)(tk.q0 >> 32), (uint32_t)tk.q1, (uint32_t)(tk.q1 >> 32), (uint32_t)tk.q2,
                               (uint32_t)(tk.q2 >> 32)};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (32u * k < tk.n && piece[k]) {
        atomicOr(&buf[wi + k], piece[k] << sh);
        if (sh && (piece[k] >> (32u - sh))) atomicOr(&buf[wi + k + 1], piece[k] >> (32u - sh));
      }
    }
  }
  if (tid == 0) {
    atomicOr(&buf[0], 2u);
    if (c == n_chunks - 1) atomicOr(&buf[(total + 7) >> 5], 1u << ((total + 7) & 31));
    atomicOr(&buf[(nbytes + 2) >> 2], 0xFFu << (8 * ((nbytes + 2) & 3)));
    atomicOr(&buf[(nbytes + 3) >> 2], 0xFFu << (8 * ((nbytes + 3) & 3)));
    chunk_bytes[chunk] = nbytes + 4;
    uint32_t r = s_crc[0];
    if (r && c != n_chunks - 1) {  // bits behind this chunk: (n_chunks - 2 - c) full chunks + the member's last chunk
      const uint32_t last_words = m.n_words - (n_chunks - 1) * CHUNK_WORDS;
      if (last_words == CHUNK_WORDS) {
        r = gf_mul(__ldg(tab + TAB_XPC + (n_chunks - 1 - c)), r);
      } else {
        r = gf_mul(__ldg(tab + TAB_XPC + (n_chunks - 2 - c)), r);
        r = gf_mul(__ldg(tab + TAB_XPW + last_words), r);
      }
    }
    if (c == 0 && bg) {
      // + the remainder of the background itself: n_words copies of bg = R(bg) * sum_i x^(32 i); the geometric factor
      // depends on the member's length only and comes with the member (slimb200_deflate_plan)
      uint32_t rb = bg;
      rb = T[rb & 255u] ^ (rb >> 8);
      rb = T[rb & 255u] ^ (rb >> 8);
      rb = T[rb & 255u] ^ (rb >> 8);
      rb = T[rb & 255u] ^ (rb >> 8);
      r ^= gf_mul(m.crc_geo, rb);
    }
    if (r) atomicXor(&member_out[4 * mi + 2], r);
  }
  __syncthreads();
  uint32_t* dst = scratch + (size_t)chunk * SLOT_WORDS;
  for (int i = tid; i < nwords_out; i += THREADS) dst[i] = buf[i];
}

// chunk sizes -> chunk offsets, member {offset, bytes, (crc), chunks}, total; one CTA
__global__ void __launch_bounds__(1024)
k_deflate_scan(const slimb200_deflate_member* __restrict__ members, int n_members, const uint32_t* __restrict__ chunk_bytes,
               uint32_t total_chunks, uint32_t* __restrict__ chunk_off, uint32_t* __restrict__ member_out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t per = (total_chunks + 1023u) / 1024u;
  const uint32_t b = min(total_chunks, (uint32_t)tid * per), e = min(total_chunks, b + per);
  uint32_t sum = 0;
  for (uint32_t i = b; i < e; ++i) sum += chunk_bytes[i];
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += a;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  uint32_t off = inc - sum;
  for (int k = 0; k < wid; ++k) off += s_warp[k];
  if (tid == 1023) s_total = off + sum;
  for (uint32_t i = b; i < e; ++i) {
    chunk_off[i] = off;
    off += chunk_bytes[i];
  }
  __syncthreads();
  const uint32_t total = s_total;
  for (int mi = tid; mi < n_members; mi += 1024) {
    const uint32_t f = members[mi].first_chunk;
    const uint32_t nxt = mi + 1 < n_members ? chunk_off[members[mi + 1].first_chunk] : total;
    member_out[4 * mi + 0] = chunk_off[f];
    member_out[4 * mi + 1] = nxt - chunk_off[f];
    member_out[4 * mi + 3] = (members[mi].n_words + CHUNK_WORDS - 1) / CHUNK_WORDS;
  }
  if (tid == 0) member_out[4 * n_members + 0] = total;
}

// pack the chunks: aligned 4-byte stores, source words funnel-shifted to the destination's alignment
__global__ void __launch_bounds__(256)
k_deflate_gather(const uint32_t* __restrict__ scratch, const uint32_t* __restrict__ chunk_bytes, const uint32_t* __restrict__ chunk_off,
                 uint8_t* __restrict__ out, size_t out_capacity, uint32_t* __restrict__ overflow) {
  const uint32_t chunk = blockIdx.x, n = chunk_bytes[chunk], off = chunk_off[chunk];
  if ((size_t)off + n > out_capacity) {
    if (threadIdx.x == 0) *overflow = 1u;
    return;
  }
  const uint32_t* src = scratch + (size_t)chunk * SLOT_WORDS;
  const uint8_t* srcb = reinterpret_cast<const uint8_t*>(src);
  uint8_t* dst = out + off;
  const uint32_t head = min(n, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
  const uint32_t nwords = (n - head) >> 2, tail0 = head + 4 * nwords;
  if (threadIdx.x < head) dst[threadIdx.x] = srcb[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x - 32 < n - tail0) dst[tail0 + threadIdx.x - 32] = srcb[tail0 + threadIdx.x - 32];
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = 8 * head;  // dst word j = source bytes head + 4j .. head + 4j + 3
  for (uint32_t j = threadIdx.x; j < nwords; j += 256) dw[j] = __funnelshift_r(src[j], src[j + 1], sh);
}
}  // namespace

extern "C" int slimb200_deflate_plan(slimb200_deflate_member* members, int32_t n_members, int64_t* total_chunks,
                                     size_t* workspace_bytes, size_t* out_bound) {
  if (!members || n_members < 1 || !total_chunks || !workspace_bytes || !out_bound) return SLIMB200_E_INVALID;
  uint64_t chunks = 0;
  for (int i = 0; i < n_members; ++i) {
    slimb200_deflate_member& m = members[i];
    if (!m.src || m.n_words == 0 || m.words_per_cell < 1 || m.cell_stride < m.words_per_cell) return SLIMB200_E_INVALID;
    if ((reinterpret_cast<uintptr_t>(m.src) & 3) != 0) return SLIMB200_E_ALIGNMENT;
    const uint64_t n = ((uint64_t)m.n_words + CHUNK_WORDS - 1) / CHUNK_WORDS;
    if (n > SLIMB200_DEFLATE_MAX_CHUNKS) return SLIMB200_E_UNSUPPORTED;
    m.first_chunk = (uint32_t)chunks;
    chunks += n;
    // G(n_words) = sum_{i < n_words} x^(32 i) mod P by doubling: G(2L) = G(L) (1 + x^(32 L)), G(L + 1) = G(L) x^32 + 1
    uint32_t g = 0u, xl = 0x80000000u;
    const uint32_t x32 = gf_xpow(32);
    for (int bit = 31; bit >= 0; --bit) {
      g ^= gf_mul(g, xl);
      xl = gf_mul(xl, xl);
      if ((m.n_words >> bit) & 1u) {
        g = gf_mul(g, x32) ^ 0x80000000u;
        xl = gf_mul(xl, x32);
      }
    }
    m.crc_geo = g;
    m.reserved = 0;
  }
  if (chunks * (uint64_t)SLOT_BYTES > 0xFFFFFFFFull) return SLIMB200_E_UNSUPPORTED;  // 32-bit stream offsets
  *total_chunks = (int64_t)chunks;
  WorkspaceCarver ws(nullptr);
  ws.take<uint32_t>(chunks * SLOT_WORDS + 1);
  ws.take<uint32_t>(chunks);
  ws.take<uint32_t>(chunks);
  *workspace_bytes = ws.used();
  *out_bound = slimb200_align_up(chunks * (size_t)(SLOT_BYTES - 8), 256);
  return SLIMB200_OK;
}

extern "C" int slimb200_deflate_init(void* tables, void* stream) {
  if (!tables) return SLIMB200_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(tables) & 3) != 0) return SLIMB200_E_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SLIMB200_LAUNCH(SLIMB200_K_DEFLATE_TABLES, s,
                  (k_deflate_tables<<<(TAB_WORDS + 255) / 256, 256, 0, s>>>(static_cast<uint32_t*>(tables))));
  return SLIMB200_OK;
}

extern "C" int slimb200_deflate_encode(const slimb200_deflate_member* members_dev, int32_t n_members, int64_t total_chunks,
                                       const void* tables, void* workspace, size_t workspace_bytes, void* out,
                                       size_t out_capacity, uint32_t* member_out, void* stream) {
  if (!members_dev || n_members < 1 || total_chunks < n_members || !tables || !workspace || !out || !member_out)
    return SLIMB200_E_INVALID;
  if (total_chunks * (int64_t)SLOT_BYTES > 0xFFFFFFFFll) return SLIMB200_E_UNSUPPORTED;
  WorkspaceCarver ws(workspace);
  uint32_t* scratch = ws.take<uint32_t>((size_t)total_chunks * SLOT_WORDS + 1);
  uint32_t* chunk_bytes = ws.take<uint32_t>((size_t)total_chunks);
  uint32_t* chunk_off = ws.take<uint32_t>((size_t)total_chunks);
  if (ws.used() > workspace_bytes) return SLIMB200_E_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SLIMB200_CUDA_TRY(cudaMemsetAsync(member_out, 0, sizeof(uint32_t) * 4 * ((size_t)n_members + 1), s));
  SLIMB200_LAUNCH(SLIMB200_K_DEFLATE_CHUNKS, s,
                  (k_deflate_chunks<<<(unsigned)total_chunks, THREADS, 0, s>>>(members_dev, n_members, static_cast<const uint32_t*>(tables),
                                                                               scratch, chunk_bytes, member_out)));
  SLIMB200_LAUNCH(SLIMB200_K_DEFLATE_SCAN, s,
                  (k_deflate_scan<<<1, 1024, 0, s>>>(members_dev, n_members, chunk_bytes, (uint32_t)total_chunks, chunk_off, member_out)));
  SLIMB200_LAUNCH(SLIMB200_K_DEFLATE_GATHER, s,
                  (k_deflate_gather<<<(unsigned)total_chunks, 256, 0, s>>>(scratch, chunk_bytes, chunk_off, static_cast<uint8_t*>(out),
                                                                           out_capacity, member_out + 4 * (size_t)n_members + 1)));
  return SLIMB200_OK;
}
