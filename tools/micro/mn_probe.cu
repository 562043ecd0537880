// Probe of tcgen05.mma kind::tf32 with MN-major shared-memory operands (128-byte swizzle): which (LBO, SBO) describes a tile stored as
// [MN block of 32 floats][k row][32 floats] (what four TMA boxes of 32 k-rows x 128 B produce)?  A[m,k] and B[n,k] are small integers,
// so C = A.B^T is exact; the kernel prints how many of the 128 x 128 outputs match for each candidate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t ltype = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)ltype << 61;
  return d;
}
// byte offset of element (mn, k) of a 128 x 32 tile: block mn/32 at blk_stride, k row at 128 B, 128-byte swizzle on the address
__device__ __forceinline__ uint32_t off_mn(int mn, int k, uint32_t blk_stride, int swz) {
  uint32_t o = (mn >> 5) * blk_stride + k * 128 + (mn & 31) * 4;
  if (swz == 0) return o ^ (((o >> 7) & 7) << 4);          // Swizzle<3,4,3>: 16-byte chunks x 8 rows
  if (swz == 1) return o ^ (((o >> 7) & 3) << 5);          // Swizzle<2,5,2>: 32-byte chunks x 4 rows (128B_BASE32B)
  return o;                                                // none
}
__global__ void __launch_bounds__(128, 1) probe(float* out, uint32_t lbo, uint32_t sbo, uint32_t kstep, uint32_t major, int nk, uint32_t ltype, int swz) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_sh;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  float* A = reinterpret_cast<float*>(sp);               // 16 KB
  float* Bm = reinterpret_cast<float*>(sp + 16384);      // 16 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 50000; i += 128) A[i] = 0.f;   // the whole allocation: stray reads return zeros
  __syncthreads();
  for (int e = tid; e < 128 * 32; e += 128) {
    const int mn = e & 127, k = e >> 7;
    *reinterpret_cast<float*>(sp + off_mn(mn, k, 4096, swz)) = (float)((mn * 3 + k * 5) % 7 - 3);            // A[m][k]
    *reinterpret_cast<float*>(sp + 16384 + off_mn(mn, k, 4096, swz)) = (float)((mn * 2 + k * 7) % 5 - 2);    // B[n][k]
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_sh)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_sh;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (major << 15);
    for (int k = 0; k < nk; ++k) {
      const uint64_t ad = mkdesc(base + k * kstep, lbo, sbo, ltype), bd = mkdesc(base + 16384 + k * kstep, lbo, sbo, ltype);
      const uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // row = tid (warp w owns TMEM lanes 32w..32w+31)
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) out[tid * 128 + cb * 32 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(128) : "memory");
}
int main() {
  float* out; cudaMalloc(&out, 128 * 128 * 4);
  static float h[128 * 128];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 204800);
  struct { uint32_t lbo, sbo, kstep, major, ltype; int swz; } cand[] = {{4096, 512, 1024, 3, 1, 1}, {4096, 1024, 1024, 3, 1, 1}, {512, 4096, 1024, 3, 1, 1}, {4096, 512, 1024, 3, 1, 0}, {4096, 1024, 1024, 3, 1, 0}, {4096, 512, 512, 3, 1, 1}, {4096, 1024, 1024, 3, 0, 2}, {4096, 128, 1024, 3, 0, 2}, {128, 4096, 1024, 3, 0, 2}, {4096, 1024, 1024, 3, 4, 0}, {4096, 1024, 1024, 3, 6, 0},{0, 1024, 32, 0}, {4096, 1024, 1024, 0}, {4096, 1024, 1024, 3}, {1024, 4096, 1024, 3}, {4096, 1024, 1024, 1}, {4096, 1024, 1024, 2},
                                                        {4096, 128, 1024, 3}, {128, 4096, 1024, 3}, {4096, 1024, 128, 3}, {4096, 1024, 256, 3}};
  for (int nk : {1, 4}) for (auto c : cand) {
    cudaMemset(out, 0, sizeof(h));
    probe<<<1, 128, 204800>>>(out, c.lbo, c.sbo, c.kstep, c.major, nk, c.ltype, c.swz);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int ok = 0, nz = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 8 * nk; ++k) ref += (float)((m * 3 + k * 5) % 7 - 3) * (float)((n * 2 + k * 7) % 5 - 2);
      ok += (h[m * 128 + n] == ref); nz += (h[m * 128 + n] != 0.f);
    }
    printf("nk=%d ltype=%u swz=%d lbo=%u sbo=%u kstep=%u major=%u: %s, %d / 16384 match, %d nonzero; C[0][0..3] = %g %g %g %g\n", nk, c.ltype, c.swz, c.lbo, c.sbo, c.kstep, c.major,
           cudaGetErrorString(e), ok, nz, h[0], h[1], h[2], h[3]);
  }
  return 0;
}
