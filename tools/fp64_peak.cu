// Microbenchmark: FP64 pipe peaks on B200 (DFMA vs DMMA shapes), dependent-issue latencies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int CHAINS>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template<int CHAINS>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double d[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { d[i][0] = threadIdx.x; d[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) dmma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int CHAINS, int SHAPE>
__global__ void k_dmma16(double* out, int iters, double a, double b) {
  double d[CHAINS][4];
  double av[8], bv[4];
#pragma unroll
  for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { d[i][0] = threadIdx.x; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (SHAPE == 4) dmma1684(d[i], av, bv[0]);
      if (SHAPE == 8) dmma1688(d[i], av, bv);
      if (SHAPE == 16) dmma16816(d[i], av, bv);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F>
float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

// sustained mode: `fp64_peak sustained [seconds]` keeps the best DMMA configuration running back to back for that long (default 4 s) and
// prints the throughput of consecutive half-second windows: the figure under the board's power cap, for kernels timed inside long steps
int sustained(double seconds) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, threads = 256, iters = 20000;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
  const double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k_dmma884<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  double elapsed = 0.0;
  int window = 0;
  while (elapsed < seconds) {
    int launches = 0;
    CK(cudaEventRecord(e0));
    for (; launches < 190; ++launches) k_dmma884<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);  // about half a second
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    elapsed += ms * 1e-3;
    printf("sustained DMMA m8n8k4 window %d (%.2f s .. %.2f s): %.2f TFLOP/s\n", window++, elapsed - ms * 1e-3, elapsed, fl * launches / ms * 1e-9);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "sustained") return sustained(argc > 2 ? atof(argv[2]) : 4.0);
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
  const int iters = 20000;
  // DFMA throughput: vary warps/SM and chains
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2}) {
      float ms = timeit([&] { k_dfma<8><<<sms * bps, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)threads * sms * bps;
      printf("DFMA chains=8 threads=%d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
    }
  }
  {
    float ms = timeit([&] { k_dfma<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA dependent chain 1 warp/SM: %.3f ms => %.2f ns per dependent DFMA\n", ms, ms * 1e6 / iters);
  }
  // DMMA
  for (int threads : {128, 256, 512}) {
    float ms = timeit([&] { k_dmma884<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms;
    printf("DMMA m8n8k4 chains=8 threads=%d : %.3f ms  %.2f TFLOP/s\n", threads, ms, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma16<8, 4><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 512 * 8 * iters * (double)(threads / 32) * sms;
    printf("DMMA m16n8k4 chains=8 threads=%d : %.3f ms  %.2f TFLOP/s\n", threads, ms, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma16<8, 8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 1024 * 8 * iters * (double)(threads / 32) * sms;
    printf("DMMA m16n8k8 chains=8 threads=%d : %.3f ms  %.2f TFLOP/s\n", threads, ms, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma16<8, 16><<<sms, threads>>>(out, iters / 2, 1.0000001, 1e-9); });
    fl = 2.0 * 2048 * 8 * (iters / 2) * (double)(threads / 32) * sms;
    printf("DMMA m16n8k16 chains=8 threads=%d : %.3f ms  %.2f TFLOP/s\n", threads, ms, fl / ms * 1e-9);
  }
  {
    float ms = timeit([&] { k_dmma884<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 dependent chain 1 warp/SM: %.2f ns per dependent DMMA\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma884<2><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 2 chains 1 warp/SM: %.2f ns per iter\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma884<4><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 4 chains 1 warp/SM: %.2f ns per iter\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma884<8><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 8 chains 1 warp/SM: %.2f ns per iter\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma884<8><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 8 chains 4 warp/SM: %.2f ns per iter\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma16<1, 8><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m16n8k8 dependent chain 1 warp/SM: %.2f ns per dependent DMMA\n", ms * 1e6 / iters);
    ms = timeit([&] { k_dmma16<1, 16><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m16n8k16 dependent chain 1 warp/SM: %.2f ns per dependent DMMA\n", ms * 1e6 / iters);
  }
  // sustained DFMA (about 3 s) to see clocks under load
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int launches = 60;
    for (int i = 0; i < launches; i++) k_dfma<8><<<sms * 2, 512>>>(out, iters * 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * iters * 4 * 512.0 * sms * 2 * launches;
    printf("DFMA sustained %.1f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < launches; i++) k_dmma884<8><<<sms, 512>>>(out, iters * 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    fl = 2.0 * 256 * 8 * iters * 4 * 16.0 * sms * launches;
    printf("DMMA m8n8k4 sustained %.1f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  return 0;
}
