// Microbenchmark (B200): cost of cluster.sync() and round-trip latency of DSMEM signalling variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_latency cluster_latency.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../self-attention-tacotron_b200/csrc/cluster_sync.cuh"
namespace cg = cooperative_groups;
using namespace satk;

__global__ void k_cluster_sync(long long* out, int iters, int do_store, float* g) {
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (do_store) g[(blockIdx.x * blockDim.x + threadIdx.x) + (long long)i * gridDim.x * blockDim.x] = (float)i;
    cluster.sync();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

// ping-pong between rank 0 and rank 1 using: mode 0 = bulk copy (256 B) + mbarrier, mode 1 = st.async (64 x 4 B),
// mode 2 = plain remote stores + remote flag (st.release / ld.acquire polling on own smem)
__global__ void k_pingpong(long long* out, int iters, int mode) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = cluster.block_rank();
  __shared__ __align__(16) float buf[2][64];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ volatile int flag;
  const int tid = threadIdx.x;
  if (tid == 0) { cl::mbar_init(&bars[0], 1); cl::mbar_init(&bars[1], 1); cl::fence_mbar_init(); flag = 0; }
  if (tid < 64) { buf[0][tid] = 0.f; buf[1][tid] = 0.f; }
  cluster.sync();
  const int peer = rank ^ 1;
  long long t0 = clock64();
  if (rank < 2) {
    for (int i = 0; i < iters; ++i) {
      // rank 0 sends on even half-steps, rank 1 replies
      for (int half = 0; half < 2; ++half) {
        const bool sender = (half == rank);
        if (sender) {
          if (mode == 0) {
            if (tid < 64) buf[1][tid] = (float)i;
            cl::fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
              uint32_t src = cl::smem_u32(&buf[1][0]);
              cl::bulk_copy_to_cta(cl::mapa(cl::smem_u32(&buf[0][0]), peer), src, 256, cl::mapa(cl::smem_u32(&bars[0]), peer));
            }
          } else if (mode == 1) {
            if (tid < 64) cl::st_async_f32(cl::mapa(cl::smem_u32(&buf[0][tid]), peer), (float)i, cl::mapa(cl::smem_u32(&bars[0]), peer));
          } else {
            if (tid < 64) { float* r = cluster.map_shared_rank(&buf[0][tid], peer); *r = (float)i; }
            __syncthreads();
            if (tid == 0) {
              uint32_t fa = cl::mapa(cl::smem_u32((const void*)&flag), peer);
              asm volatile("fence.acq_rel.cluster;\n\tst.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(fa), "r"(2 * i + half + 1) : "memory");
            }
          }
        } else {
          if (mode <= 1) {
            if (tid == 0) cl::mbar_arrive_expect_tx(&bars[0], 256);
            cl::mbar_wait(&bars[0], i & 1);
          } else {
            if (tid == 0) { while (flag != 2 * i + half + 1) {} asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
            __syncthreads();
          }
        }
      }
    }
  }
  long long t1 = clock64();
  cluster.sync();
  if (threadIdx.x == 0 && rank == 0 && blockIdx.x < 16) out[0] = (t1 - t0) / (2 * iters);
}

template <typename K, typename... A>
void launch(K k, int cs, int nthreads, A... a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs);
  cfg.blockDim = dim3(nthreads);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cs > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, a...);
  if (e != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(e));
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("sync error %s\n", cudaGetErrorString(e));
}

int main() {
  long long* out; cudaMallocManaged(&out, 8);
  float* g; cudaMalloc(&g, 1ull << 30);
  for (int cs : {2, 8, 16})
    for (int nt : {256, 512})
      for (int st : {0, 1}) {
        launch(k_cluster_sync, cs, nt, out, 200, st, g);
        printf("cluster.sync  cs=%2d threads=%3d global_store=%d : %lld cycles\n", cs, nt, st, out[0]);
      }
  for (int cs : {2, 16})
    for (int mode : {0, 1, 2}) {
      launch(k_pingpong, cs, 128, out, 200, mode);
      printf("one-way latency cs=%2d mode=%d (%s): %lld cycles\n", cs, mode, mode == 0 ? "bulk copy 256B" : mode == 1 ? "st.async 64x4B" : "remote st + flag", out[0]);
    }
  return 0;
}
