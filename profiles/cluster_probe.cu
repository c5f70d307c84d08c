// How many thread-block clusters of each size fit the GPU at once for a kernel shaped like K6r (512 threads,
// ~165 KB of dynamic shared memory, one CTA per SM)?  Answers where the 16-CTA clusters strand SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o profiles/bin/cluster_probe profiles/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(int *p) { extern __shared__ int s[]; s[threadIdx.x] = 1; if (p) p[0] = s[0]; }
int main() {
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sm);
    const int smem = 165 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c = 1; c <= 16; c *= 2) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %3d (%3d CTAs) %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, 0); printf("cooperative launch %d\n", coop);
    return 0;
}
