// Microbenchmark (B200): all-to-all exchange of a small slice inside a cluster, per-iteration cycles.
// mode 0: cp.async.bulk 256 B per peer; mode 1: st.async b32 (64 x 4 B per peer); mode 2: st.async v4 (16 x 16 B per peer)
// waiters: 0 = all threads spin on the mbarrier, 1 = one warp spins then __syncthreads
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../self-attention-tacotron_b200/csrc/cluster_sync.cuh"
namespace cg = cooperative_groups;
using namespace satk;

template <int CS>
__global__ void k_a2a(long long* out, int iters, int mode, int waiters, int compute) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = cluster.block_rank();
  __shared__ __align__(16) float buf[2][CS][64];
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x;
  if (tid == 0) { cl::mbar_init(&bars[0], 1); cl::mbar_init(&bars[1], 1); cl::fence_mbar_init(); }
  for (int i = tid; i < 2 * CS * 64; i += blockDim.x) (&buf[0][0][0])[i] = 0.f;
  cluster.sync();
  const uint32_t RX = (CS - 1) * 256;
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int cur = it & 1;
    if (tid == 0) cl::mbar_arrive_expect_tx(&bars[cur], RX);
    // "compute": dependent FMA chain of `compute` instructions
    for (int c = 0; c < compute; ++c) acc = fmaf(acc, 1.0001f, 0.5f);
    if (tid < 64) {
      const float v = (float)it + acc * 1e-30f;
      buf[cur][rank][tid] = v;
      if (mode == 0) {
        cl::fence_proxy_async();
        cl::named_bar_sync(1, 64);
        if (tid < CS && tid != rank) {
          const uint32_t src = cl::smem_u32(&buf[cur][rank][0]);
          cl::bulk_copy_to_cta(cl::mapa(src, tid), src, 256, cl::mapa(cl::smem_u32(&bars[cur]), tid));
        }
      } else if (mode == 1) {
        const uint32_t dst = cl::smem_u32(&buf[cur][rank][tid]), bar = cl::smem_u32(&bars[cur]);
#pragma unroll
        for (int r = 0; r < CS; ++r) if (r != rank) cl::st_async_f32(cl::mapa(dst, r), v, cl::mapa(bar, r));
      } else {
        // 16 lanes per row of 4 floats: lane q sends floats [4q, 4q+4) gathered by shuffles
        float v0 = __shfl_sync(0xffffffffu, v, (tid & 7) * 4 + 0), v1 = __shfl_sync(0xffffffffu, v, (tid & 7) * 4 + 1);
        float v2 = __shfl_sync(0xffffffffu, v, (tid & 7) * 4 + 2), v3 = __shfl_sync(0xffffffffu, v, (tid & 7) * 4 + 3);
        if ((tid & 31) < 8) {
          const int q = (tid >> 5) * 8 + (tid & 7);
          const uint32_t dst = cl::smem_u32(&buf[cur][rank][4 * q]), bar = cl::smem_u32(&bars[cur]);
#pragma unroll
          for (int r = 0; r < CS; ++r)
            if (r != rank)
              asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(cl::mapa(dst, r)),
                           "r"(__float_as_uint(v0)), "r"(__float_as_uint(v1)), "r"(__float_as_uint(v2)), "r"(__float_as_uint(v3)), "r"(cl::mapa(bar, r)) : "memory");
        }
      }
    }
    if (waiters == 0) {
      cl::mbar_wait(&bars[cur], (it >> 1) & 1);
    } else {
      if (tid < 32) cl::mbar_wait(&bars[cur], (it >> 1) & 1);
      __syncthreads();
    }
    acc += buf[cur][(rank + 1) % CS][tid & 63];
  }
  long long t1 = clock64();
  cluster.sync();
  if (tid == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (long long)acc; }
}

template <int CS>
void run(long long* out) {
  for (int mode : {0, 1, 2})
    for (int waiters : {0, 1})
      for (int compute : {0, 400}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CS); cfg.blockDim = dim3(256);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (CS > 8) cudaFuncSetAttribute(k_a2a<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_a2a<CS>, out, 1000, mode, waiters, compute);
        if (e != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(e));
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) printf("sync error %s\n", cudaGetErrorString(e));
        printf("CS=%2d mode=%d waiters=%s compute=%3d fma : %lld cycles/iter\n", CS, mode, waiters ? "1warp" : "all  ", compute, out[0]);
      }
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  run<2>(out); run<8>(out); run<16>(out);
  return 0;
}
