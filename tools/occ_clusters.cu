// Max co-resident clusters per cluster size for a 512-thread kernel with a given dynamic shared memory footprint.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 2) k(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; if (p) p[0] = s[0]; }
int main() {
  for (int kb : {56, 74, 86, 108, 113}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    printf("smem %3d KB:", kb);
    for (int c = 1; c <= 16; ++c) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = kb * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      if (e != cudaSuccess) { cudaGetLastError(); printf("  c%d:err", c); } else printf("  c%d:%d(%d)", c, n, n * c);
    }
    printf("\n");
  }
  return 0;
}
