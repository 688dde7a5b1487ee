// How fast can B200 launch/retire CTAs and clusters that each hold ~105 KB of shared memory?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_cluster tools/ubench_cluster.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_empty(float* out, int spin, int nbar) {
  extern __shared__ float sm[];
  if (threadIdx.x == 0) sm[0] = 1.f;
  for (int b = 0; b < nbar; ++b) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = sm[0];
}

float run(int grid, int threads, int smem, int cluster, int spin, int nbar) {
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  float* out = nullptr;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, k_empty, out, spin, nbar);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_empty, out, spin, nbar);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return -1; }
  return ms;
}

int main() {
  const int smem = 105 * 1024;
  printf("grid 12288 CTAs, 105 KB smem, empty body\n");
  for (int threads : {256, 512}) {
    for (int cl : {1, 2, 4, 6, 8}) {
      int grid = 12288 / cl * cl;
      printf("threads %d cluster %d : empty %.3f ms | 3 barriers %.3f ms | spin 20k cyc %.3f ms | spin 20k + 3 barriers %.3f ms\n", threads, cl,
             run(grid, threads, smem, cl, 0, 0), run(grid, threads, smem, cl, 0, 3), run(grid, threads, smem, cl, 20000, 0),
             run(grid, threads, smem, cl, 20000, 3));
    }
  }
  printf("small smem (8 KB), cluster 6: empty %.3f ms\n", run(12288, 256, 8 * 1024, 6, 0, 0));
  return 0;
}
