// Issue rate of scalar FFMA against packed fma.rn.f32x2 (FFMA2) on sm_100a: 16 warps per SM, 16 independent accumulators
// (pairs) per thread, register operands only.  Prints cycles per warp instruction per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void __launch_bounds__(512, 1) k_scalar(float* out, int iters, long long* cyc, float x, float y) {
  float c[16];
  for (int j = 0; j < 16; ++j) c[j] = threadIdx.x * 1e-3f + j;
  float a = x + threadIdx.x * 1e-6f, b = y;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[j]) : "f"(a), "f"(b));
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void __launch_bounds__(512, 1) k_packed(float* out, int iters, long long* cyc, float x, float y) {
  unsigned long long c[16];
  for (int j = 0; j < 16; ++j) { float2 v = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f); c[j] = *reinterpret_cast<unsigned long long*>(&v); }
  float2 av = make_float2(x + threadIdx.x * 1e-6f, x), bv = make_float2(y, y * 0.5f);
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&av), b = *reinterpret_cast<unsigned long long*>(&bv);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) ffma2(c[j], a, b);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 16; ++j) { float2 v = *reinterpret_cast<float2*>(&c[j]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  for (int it : {1000, 4000}) {
    long long h;
    k_scalar<<<148, 512>>>(out, it, cyc, 0.999f, 0.5f); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("scalar FFMA  iters %d: %lld cycles, %.2f cycles per warp instruction per sub-partition (4 warps each x 16 per iter)\n", it, h, (double)h / (it * 16.0 * 4));
    k_packed<<<148, 512>>>(out, it, cyc, 0.999f, 0.5f); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("packed FFMA2 iters %d: %lld cycles, %.2f cycles per warp instruction per sub-partition (2 FMAs per lane each)\n", it, h, (double)h / (it * 16.0 * 4));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
