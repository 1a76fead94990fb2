// Write-bandwidth ceilings on B200 for the store patterns of the slimb200 kernels (measurement tool, not product).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- (b) plain vector stores, grid-stride ----
__global__ void k_st_v4(float4* __restrict__ dst, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = z;
}
// each CTA writes contiguous chunks of `chunk4` float4 (chunk-major), 4 stores in flight per thread
__global__ void k_st_v4_chunk(float4* __restrict__ dst, size_t n4, int chunk4) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t nchunks = n4 / chunk4;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    float4* p = dst + c * chunk4;
    for (int i = threadIdx.x; i < chunk4; i += blockDim.x) p[i] = z;
  }
}

// ---- (c) 1-D bulk copy smem -> global ----
__global__ void k_bulk_1d(char* __restrict__ dst, size_t bytes, int chunk) {
  extern __shared__ __align__(128) char s[];
  for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<float4*>(s)[i] = make_float4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nchunks = bytes / chunk;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(smem_u32(s)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---- (d) tensor-map stores ----
__global__ void k_tma_3d(const __grid_constant__ CUtensorMap map, int n0, int n1, int n2, int b0, int b1, int b2, int box_bytes) {
  extern __shared__ __align__(1024) char s[];
  for (int i = threadIdx.x; i < box_bytes / 16; i += blockDim.x) reinterpret_cast<float4*>(s)[i] = make_float4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t0 = (n0 + b0 - 1) / b0, t1 = (n1 + b1 - 1) / b1, t2 = (n2 + b2 - 1) / b2;
    const long long total = (long long)t0 * t1 * t2;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      const int i0 = (int)(t % t0), i1 = (int)((t / t0) % t1), i2 = (int)(t / ((long long)t0 * t1));
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map), "r"(smem_u32(s)),
                   "r"(i0 * b0), "r"(i1 * b1), "r"(i2 * b2) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void k_tma_3d_hint(const __grid_constant__ CUtensorMap map, int n0, int n1, int n2, int b0, int b1, int b2, int box_bytes, int order) {
  extern __shared__ __align__(1024) char s[];
  for (int i = threadIdx.x; i < box_bytes / 16; i += blockDim.x) reinterpret_cast<float4*>(s)[i] = make_float4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int t0 = (n0 + b0 - 1) / b0, t1 = (n1 + b1 - 1) / b1, t2 = (n2 + b2 - 1) / b2;
    const long long total = (long long)t0 * t1 * t2;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    for (long long q = 0; q < per; ++q) {
      const long long t = order == 0 ? (long long)blockIdx.x + q * gridDim.x : (long long)blockIdx.x * per + q;
      if (t >= total) break;
      const int i0 = (int)(t % t0), i1 = (int)((t / t0) % t1), i2 = (int)(t / ((long long)t0 * t1));
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(&map), "r"(smem_u32(s)),
                   "r"(i0 * b0), "r"(i1 * b1), "r"(i2 * b2), "l"(pol) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---- (e) canvas pattern with plain stores: CTA per tile (4 rows x 32 cols x 64 ch), 8 lanes x 16B per row ----
__global__ void k_canvas_st(float* __restrict__ canvas, int B, int C, int G0, int G1) {
  const int tiles_x = G0 / 4, tiles_y = G1 / 32;
  const int n_tiles = B * tiles_x * tiles_y;
  const size_t plane = (size_t)G0 * G1;
  const int tid = threadIdx.x;
  const int r = (tid >> 3) & 3, j = tid & 7;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int b = t / (tiles_x * tiles_y), tl = t % (tiles_x * tiles_y);
    const int tx = tl / tiles_y, ty = tl % tiles_y;
    float* base = canvas + (size_t)b * C * plane + (size_t)(tx * 4 + r) * G1 + ty * 32 + 4 * j;
    for (int c = tid >> 5; c < C; c += blockDim.x / 32) *reinterpret_cast<float4*>(base + c * plane) = make_float4(0, 0, 0, 0);
  }
}
// canvas pattern, wider tile: CTA covers 1 row x 640 cols? -> each warp writes 512 B contiguous of one (c,row)
__global__ void k_canvas_rows(float* __restrict__ canvas, int B, int C, int G0, int G1, int rows_per_tile) {
  // tile = rows_per_tile full rows (G1 floats) of all channels; warp w handles (c,row) pairs, lanes cover the row with float4
  const int tiles = B * (G0 / rows_per_tile);
  const size_t plane = (size_t)G0 * G1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int b = t / (G0 / rows_per_tile), x0 = (t % (G0 / rows_per_tile)) * rows_per_tile;
    for (int cr = warp; cr < C * rows_per_tile; cr += nw) {
      const int c = cr / rows_per_tile, rr = cr % rows_per_tile;
      float4* p = reinterpret_cast<float4*>(canvas + ((size_t)b * C + c) * plane + (size_t)(x0 + rr) * G1);
      for (int i = lane; i < G1 / 4; i += 32) p[i] = make_float4(0, 0, 0, 0);
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
float time_it(F f, int reps = 10) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f, tot = 0;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best; tot += ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const size_t bytes = (size_t)8 * 64 * 640 * 640 * 4;  // canvas of 8 frames = 838.9 MB
  char* buf; CK(cudaMalloc(&buf, bytes + (1 << 20)));
  char* buf2; CK(cudaMalloc(&buf2, bytes + (1 << 20)));
  auto rep = [&](const char* name, float ms, size_t by) { printf("%-58s %8.3f ms  %8.1f GB/s\n", name, ms, by / ms / 1e6); fflush(stdout); };
  rep("cudaMemsetAsync", time_it([&] { CK(cudaMemsetAsync(buf, 0, bytes)); }), bytes);
  rep("cudaMemcpyAsync d2d (read+write bytes)", time_it([&] { CK(cudaMemcpyAsync(buf2, buf, bytes, cudaMemcpyDeviceToDevice)); }), 2 * bytes);
  for (int cps : {1, 2, 4, 8, 16})
    for (int th : {256, 1024}) {
      if (cps * th > 2048) continue;
      char nm[128]; snprintf(nm, 128, "st.v4 grid-stride %d CTA/SM x %d thr", cps, th);
      rep(nm, time_it([&] { k_st_v4<<<148 * cps, th>>>((float4*)buf, bytes / 16); }), bytes);
    }
  for (int chunk : {2048, 8192, 32768})
    for (int cps : {2, 4, 8}) {
      char nm[128]; snprintf(nm, 128, "st.v4 chunked %d B/CTA-iter, %d CTA/SM x 256", chunk, cps);
      rep(nm, time_it([&] { k_st_v4_chunk<<<148 * cps, 256>>>((float4*)buf, bytes / 16, chunk / 16); }), bytes);
    }
  for (int chunk : {4096, 16384, 32768, 65536})
    for (int cps : {1, 2, 4}) {
      char nm[128]; snprintf(nm, 128, "bulk 1-D smem->global %d B, %d CTA/SM", chunk, cps);
      CK(cudaFuncSetAttribute(k_bulk_1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      rep(nm, time_it([&] { k_bulk_1d<<<148 * cps, 128, chunk>>>(buf, bytes, chunk); }), bytes);
    }
  // tensor maps
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)p;
  CK(cudaFuncSetAttribute(k_tma_3d, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096));
  CK(cudaFuncSetAttribute(k_tma_3d_hint, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096));
  {  // canvas: dims (G1=640, G0=640, B*C=512) f32
    struct Cfg { int b0, b1, b2; } cfgs[] = {{32, 4, 64}, {64, 2, 64}, {64, 4, 64}, {64, 4, 32}, {128, 2, 64}, {128, 1, 64}};
    for (auto c : cfgs) {
      CUtensorMap map;
      cuuint64_t dims[3] = {640, 640, 512}; cuuint64_t str[2] = {640 * 4, 640 * 640 * 4};
      cuuint32_t box[3] = {(cuuint32_t)c.b0, (cuuint32_t)c.b1, (cuuint32_t)c.b2}; cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      const int bb = c.b0 * c.b1 * c.b2 * 4;
      for (int cps : {1, 2}) {
        if ((bb + 1024) * cps > 220000) continue;
        char nm[128]; snprintf(nm, 128, "TMA canvas box(%d,%d,%d)=%dKB, %d CTA/SM", c.b0, c.b1, c.b2, bb / 1024, cps);
        rep(nm, time_it([&] { k_tma_3d<<<148 * cps, 128, bb + 1024>>>(map, 640, 640, 512, c.b0, c.b1, c.b2, bb); }), bytes);
        snprintf(nm, 128, "TMA canvas box(%d,%d,%d)=%dKB, %d CTA/SM evict_first", c.b0, c.b1, c.b2, bb / 1024, cps);
        rep(nm, time_it([&] { k_tma_3d_hint<<<148 * cps, 128, bb + 1024>>>(map, 640, 640, 512, c.b0, c.b1, c.b2, bb, 0); }), bytes);
      }
    }
  }
  {  // pyramid: dims (8500 cols, 6400 rows, 8) bf16, various pitches
    for (int pitch : {8512, 8576, 8704, 16384}) {
      const size_t pbytes = (size_t)8 * 6400 * pitch * 2;
      char* pyr; CK(cudaMalloc(&pyr, pbytes));
      struct Cfg { int b0, b1; CUtensorMapSwizzle sw; } cfgs[] = {{64, 128, CU_TENSOR_MAP_SWIZZLE_128B}, {256, 64, CU_TENSOR_MAP_SWIZZLE_NONE}};
      for (auto c : cfgs) {
        CUtensorMap map;
        cuuint64_t dims[3] = {8500, 6400, 8}; cuuint64_t str[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)6400 * pitch * 2};
        cuuint32_t box[3] = {(cuuint32_t)c.b0, (cuuint32_t)c.b1, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pyr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int bb = c.b0 * c.b1 * 2;
        for (int order : {0, 1}) {
          char nm[128]; snprintf(nm, 128, "TMA pyramid pitch %d box(%d,%d) order %s evict_first", pitch, c.b0, c.b1, order ? "contig/CTA" : "strided");
          rep(nm, time_it([&] { k_tma_3d_hint<<<148, 128, bb + 1024>>>(map, 8500, 6400, 8, c.b0, c.b1, 1, bb, order); }), (size_t)8 * 6400 * 8500 * 2);
        }
        char nm[128]; snprintf(nm, 128, "TMA pyramid pitch %d box(%d,%d) strided, no hint", pitch, c.b0, c.b1);
        rep(nm, time_it([&] { k_tma_3d<<<148, 128, bb + 1024>>>(map, 8500, 6400, 8, c.b0, c.b1, 1, bb); }), (size_t)8 * 6400 * 8500 * 2);
      }
      CK(cudaFree(pyr));
    }
    // panel layout: [b][panel=67][6400 rows][128 cols] bf16 -> each 128x128 tile is 32 KB contiguous
    {
      const size_t pbytes = (size_t)8 * 67 * 6400 * 128 * 2;
      char* pyr; CK(cudaMalloc(&pyr, pbytes));
      CUtensorMap map;
      cuuint64_t dims[3] = {128, 6400, 8 * 67}; cuuint64_t str[2] = {256, (cuuint64_t)6400 * 256};
      for (int b0 : {64, 128}) {
        cuuint32_t box[3] = {(cuuint32_t)b0, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pyr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         b0 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int bb = b0 * 128 * 2;
        char nm[128]; snprintf(nm, 128, "TMA pyramid PANEL layout box(%d,128)", b0);
        rep(nm, time_it([&] { k_tma_3d<<<148, 128, bb + 1024>>>(map, 128, 6400, 8 * 67, b0, 128, 1, bb); }), pbytes);
        snprintf(nm, 128, "TMA pyramid PANEL layout box(%d,128) evict_first", b0);
        rep(nm, time_it([&] { k_tma_3d_hint<<<148, 128, bb + 1024>>>(map, 128, 6400, 8 * 67, b0, 128, 1, bb, 0); }), pbytes);
      }
      CK(cudaFree(pyr));
    }
  }
  return 0;
}
