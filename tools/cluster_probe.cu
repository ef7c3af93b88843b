// How many CTAs of a 1-CTA-per-SM kernel (226 KB dynamic smem) can be co-resident for cluster sizes 1/2/4/8?
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 226 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d -> %d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
