// Throughput of legacy mma.sync.m16n8k8 TF32 on sm_100a with 16 warps per SM (one CTA of 512 threads per SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, b0 = 0.5f, b1 = 0.25f;
  float c[4][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
                     "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  for (int it : {1000, 4000}) {
    k<<<148, 512>>>(out, it, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("iters %d: %lld cycles, %.2f cycles per mma.sync per SM (16 warps x 4 mma per iter)\n", it, h, (double)h / (it * 4.0 * 16));
  }
  return 0;
}
