// Point-to-point synchronisation inside a thread-block cluster: mbarrier transaction barriers fed by
// asynchronous DSMEM copies (cp.async.bulk shared::cta -> shared::cluster) and st.async.
//
// Why not cluster.sync(): barrier.cluster.arrive.release has to drain EVERY outstanding store of the CTA
// (including the fire-and-forget global stores of saved activations) and stalls all 16 CTAs on the
// slowest one, 3-4 times per recurrent step.  A transaction barrier is completed by the arriving data
// itself, orders nothing else, and lets each CTA run as soon as ITS inputs are there.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace satk {
namespace cl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival + expected transaction bytes for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// generic-proxy writes to shared memory -> visible to the async proxy (bulk copy engine)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// local shared memory -> remote CTA's shared memory, completing `bytes` on the remote mbarrier. bytes % 16 == 0.
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                 uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}

// 4-byte remote store that completes 4 bytes on the remote mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t dst_cluster_addr, float v, uint32_t mbar_cluster_addr) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster_addr),
               "r"(__float_as_uint(v)), "r"(mbar_cluster_addr)
               : "memory");
}

// 16-byte remote store (4 floats) completing 16 bytes on the remote mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster_addr, float a, float b, float c, float d, uint32_t mbar_cluster_addr) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(dst_cluster_addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)),
               "r"(mbar_cluster_addr)
               : "memory");
}

// cp.async (LDGSTS): asynchronous global -> shared prefetch rings, several recurrent steps ahead of their use
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int RING = 8;   // prefetch ring slots
constexpr int PFD = 6;    // prefetch distance in steps (>> HBM latency / step time)

// named barrier among a subset of warps
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


// Reduce-scatter of 16 per-lane partial sums over a 16-lane group (xor 8,4,2,1): 15 shuffles instead of the 64
// of a plain butterfly.  On return v[0] of lane L holds the group sum of element (L & 15).
// Shared-memory store the compiler does not treat as a memory operation: ordinary loads may be scheduled across it. Only for
// addresses that no ordinary load/store of the same thread touches before the next barrier (keeps long unrolled
// load -> math -> store chains from being serialised by may-alias analysis).
__device__ __forceinline__ void sts_noalias(float* p, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v));
}

__device__ __forceinline__ float reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
  for (int hs = 8; hs >= 1; hs >>= 1) {
    const bool up = (lane & hs) != 0;
#pragma unroll
    for (int j = 0; j < hs; ++j) {
      const float send = up ? v[j] : v[j + hs];
      const float keep = up ? v[j + hs] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, hs);
    }
  }
  return v[0];
}


// Reductions over a 256-thread group (8 warps) through shared memory + a named barrier; all 256 threads call them.
// `red` needs 16 floats; `wig` = warp index inside the group (0..7).
__device__ __forceinline__ void group_sum2(float& x, float& y, float* red, int wig, int lane, int barid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  if (lane == 0) { red[wig] = x; red[8 + wig] = y; }
  named_bar_sync(barid, 256);
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sx += red[i]; sy += red[8 + i]; }
  named_bar_sync(barid, 256);
  x = sx; y = sy;
}
__device__ __forceinline__ float group_max(float x, float* red, int wig, int lane, int barid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  if (lane == 0) red[wig] = x;
  named_bar_sync(barid, 256);
  float m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  named_bar_sync(barid, 256);
  return m;
}

}  // namespace cl
}  // namespace satk
