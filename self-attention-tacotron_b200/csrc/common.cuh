// Shared helpers for libsatk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/satk.h"

namespace satk {

void set_error(const char* fmt, ...);

#define SATK_CHECK_ARG(cond, ...)                      \
  do {                                                 \
    if (!(cond)) {                                     \
      satk::set_error(__VA_ARGS__);                    \
      return SATK_ERR_INVALID;                         \
    }                                                  \
  } while (0)

#define SATK_CUDA(call)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      satk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
      return SATK_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define SATK_LAUNCH_CHECK()                                                                    \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) {                                                                  \
      satk::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SATK_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// accurate-enough tanh: tanhf is ~1 ulp; MUFU.TANH (tanh.approx) is ~5e-4 rel and NOT used on parity paths
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case SATK_ACT_RELU: return fmaxf(x, 0.0f);
    case SATK_ACT_TANH: return tanhf_(x);
    case SATK_ACT_SIGMOID: return sigmoidf_(x);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (all threads get the result); `sh` >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// developer aid: per-phase cycle accounting of thread 0 / CTA 0 (build with EXTRA=-DSATK_PHASE_TIMING)
#ifdef SATK_PHASE_TIMING
static __device__ long long g_phase[16];   // one copy per translation unit
#define PT_DECL long long pt_t = clock64(), pt_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PT(i) { long long pt_n = clock64(); pt_acc[i] += pt_n - pt_t; pt_t = pt_n; }
#define PT_FLUSH(n) if (blockIdx.x == 0 && threadIdx.x == 0) { for (int i_ = 0; i_ < 16; ++i_) satk::g_phase[i_] = pt_acc[i_] / (n); }
#else
#define PT_DECL
#define PT(i)
#define PT_FLUSH(n)
#endif

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace satk
