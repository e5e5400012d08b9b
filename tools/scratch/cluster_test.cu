#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) {
  extern __shared__ char sm[];
  unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) out[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (int)r;
  sm[threadIdx.x] = 0;
}
int try_launch(dim3 grid, dim3 cl, size_t smem, cudaStream_t st) {
  int* d; cudaMalloc(&d, 4096 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl.x; at[0].val.clusterDim.y = cl.y; at[0].val.clusterDim.z = cl.z;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h[64]; cudaMemcpy(h, d, 64 * 4, cudaMemcpyDeviceToHost);
  printf("grid (%u,%u,%u) cluster (%u,%u,%u) smem %zu: launch %s sync %s ranks %d %d %d %d %d %d %d %d\n", grid.x, grid.y, grid.z, cl.x, cl.y, cl.z, smem,
         cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  cudaGetLastError(); cudaFree(d); return 0;
}
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  try_launch(dim3(8, 4, 1), dim3(2, 1, 1), 1024, 0);
  try_launch(dim3(7, 4, 1), dim3(1, 2, 1), 1024, 0);
  try_launch(dim3(7, 4, 1), dim3(1, 2, 1), 197888, 0);
  try_launch(dim3(7, 4, 1), dim3(1, 2, 1), 197888, s);
  try_launch(dim3(4, 32, 1), dim3(1, 2, 1), 197888, s);
  try_launch(dim3(4, 7, 1), dim3(2, 1, 1), 197888, s);
  return 0;
}
