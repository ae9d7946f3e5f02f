// launch_limits.cu — what does the driver allow for a 448-thread CTA with a register cap? (tools only; not part of the library)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/launch_limits tools/launch_limits.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int R> __global__ void __maxnreg__(R) k(double* out) {
  extern __shared__ double sm[];
  double acc[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) acc[i] = sm[(threadIdx.x + i) % 64] * i;
#pragma unroll 1
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 96; ++i) acc[i] = fma(acc[i], acc[(i + 7) % 96], sm[(i + j) % 64]);
  double s = 0;
#pragma unroll
  for (int i = 0; i < 96; ++i) s += acc[i];
  out[threadIdx.x] = s;
}
template <int R> void probe(double* d) {
  cudaFuncAttributes a{};
  cudaFuncGetAttributes(&a, k<R>);
  printf("maxnreg %d: numRegs %d maxThreadsPerBlock %d\n", R, a.numRegs, a.maxThreadsPerBlock);
  for (int smem : {0, 100000, 221984, 229000, 232448}) {
    cudaError_t e = cudaFuncSetAttribute(k<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int threads : {416, 448, 480}) {
      int blocks = -1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k<R>, threads, smem);
      k<R><<<1, threads, smem>>>(d);
      cudaError_t l = cudaGetLastError();
      cudaError_t s = cudaDeviceSynchronize();
      printf("  smem %d threads %d: setattr=%s occupancy=%d launch=%s sync=%s\n", smem, threads, cudaGetErrorName(e), blocks, cudaGetErrorName(l),
             cudaGetErrorName(s));
    }
  }
}
int main() {
  double* d;
  cudaMalloc(&d, 8 * 1024);
  cudaDeviceProp p{};
  cudaGetDeviceProperties(&p, 0);
  printf("%s regsPerBlock %d regsPerSM %d smemOptin %zu smemPerSM %zu reserved %zu\n", p.name, p.regsPerBlock, p.regsPerMultiprocessor,
         p.sharedMemPerBlockOptin, p.sharedMemPerMultiprocessor, p.reservedSharedMemPerBlock);
  probe<144>(d);
  probe<136>(d);
  probe<128>(d);
  return 0;
}
