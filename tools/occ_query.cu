// Queries how many thread-block clusters of size 1/2/4/8 (one ~200 KB CTA per SM) the device can keep resident.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s SMs=%d smemPerBlockOptin=%zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  for (int smem : {100 * 1024, 200 * 1024, 220 * 1024}) {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
      printf("smem %d KB cluster %2d: max active clusters %d (%d SMs) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
