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

// Per-device launch state: function attributes and SM counts belong to a DEVICE, and one process may drive several GPUs.
// slimb200_device_info (profile.cu): the current device (< SLIMB200_MAX_DEVICES) and its cached SM count.
#define SLIMB200_MAX_DEVICES 64
int slimb200_device_info(int* dev, int* n_sm);
#define SLIMB200_DEVICE(dev, n_sm)                                \
  int dev = 0, n_sm = 0;                                          \
  do {                                                            \
    const int e__ = slimb200_device_info(&dev, &n_sm);            \
    if (e__ != 0) return e__;                                     \
  } while (0)

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
