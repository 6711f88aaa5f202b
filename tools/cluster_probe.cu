// Probe: how many clusters of size C (640 threads, ~210 KB dynamic smem per CTA) can be co-resident on this GPU?
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(640, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = (int)sm[0];
}
int main() {
  int dev = 0;
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, dev);
  printf("%s SMs %d smem optin %zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
  const int smems[] = {200 * 1024, 215 * 1024, 100 * 1024};
  for (int si = 0; si < 3; ++si) {
    int smem = smems[si];
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c = 1; c <= 16; c *= 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(128);
      cfg.blockDim = dim3(640);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
      printf("smem %d KB cluster %2d: max active clusters %d (%s) -> %d CTAs\n", smem / 1024, c, n, cudaGetErrorString(e), n * c);
      cudaGetLastError();
      int* d = nullptr;
      e = cudaLaunchKernelEx(&cfg, probe_kernel, d);
      cudaError_t e2 = cudaDeviceSynchronize();
      printf("   launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
      cudaGetLastError();
    }
  }
  return 0;
}
