// Shared helpers for the slimb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slimb200.h"

#define SLIMB200_LAUNCH_CHECK()                  \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

#define SLIMB200_CUDA_TRY(expr)                  \
  do {                                           \
    cudaError_t e__ = (expr);                    \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

static inline size_t slimb200_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct WorkspaceCarver {
  char* base;
  size_t off;
  explicit WorkspaceCarver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = slimb200_align_up(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return slimb200_align_up(off, 256); }
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// launch accounting / optional event timing (profile.cu)
void slimb200_prof_pre(int id, cudaStream_t s);
void slimb200_prof_post(int id, cudaStream_t s);
#define SLIMB200_LAUNCH(id, stream, ...)  \
  do {                                    \
    slimb200_prof_pre((id), (stream));    \
    __VA_ARGS__;                          \
    slimb200_prof_post((id), (stream));   \
    SLIMB200_LAUNCH_CHECK();              \
  } while (0)
