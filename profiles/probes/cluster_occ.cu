// probe: how many clusters of size S (one CTA per SM, 200 KB smem) can be co-resident on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int threads : {416, 384, 512})
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t c{};
    c.gridDim = dim3(cs * 16); c.blockDim = dim3(threads); c.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    c.attrs = a; c.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &c);
    printf("threads %d cluster size %2d: max active clusters %d (%d CTAs) %s\n", threads, cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
