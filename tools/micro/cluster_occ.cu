// Probe: how many thread-block clusters of each size are co-resident on this GPU (1 CTA/SM-sized shared memory).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
    int smem_list[] = {100 * 1024, 160 * 1024, 227 * 1024};
    int thr_list[] = {256, 512};
    for (int smem : smem_list) for (int thr : thr_list) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cs = 1; cs <= 16; ++cs) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("smem=%dK thr=%d cluster=%d -> max active clusters %d (CTAs %d) %s\n", smem / 1024, thr, cs, n, n * cs,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return 0;
}
