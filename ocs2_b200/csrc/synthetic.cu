// synthetic.cu — seeded, counter-based generator of random-but-stabilisable LQ batches, written straight into the device-resident
// record layout. Bit-identical to oracle/lq_oracle.cpp orc_generate_problem (explicit __dmul_rn/__dadd_rn/fma, same hash), so the
// CPU oracle can regenerate any sampled problem exactly. Family: A = I + dt*Ac (ILQR) or Ac (SLQ) with
// Ac ~ U(-1,1)/sqrt(n); B = dt*Bc or Bc; joint cost W = M'M/(n+m) + 0.1 I (as getRandomCost,
// ocs2_oc/test/include/ocs2_oc/test/testProblemsGeneration.h:45-58) scaled by dt for the discrete model (ILQR.cpp:149-150);
// D = [I | U(-1,1)] (full row rank, cf. generateFullRowRankmatrix), C ~ U(-1,1), e ~ 0.1 U(-1,1); terminal Qf = Mf'Mf/n + 0.1 I.
#include "o2c_common.cuh"

namespace o2c {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double urand(uint64_t seedmix, long long problem, int node, int field, int idx) {
  const uint64_t ctr = ((((uint64_t)problem * 1024ULL + (uint64_t)node) * 16ULL + (uint64_t)field) << 16) + (uint64_t)idx;
  const uint64_t h = mix64(seedmix ^ ctr);
  const double u01 = __dmul_rn((double)(h >> 11), 1.0 / 9007199254740992.0);
  return __dadd_rn(__dmul_rn(2.0, u01), -1.0);
}
enum { F_A = 0, F_B, F_HV, F_M, F_Q, F_R, F_C, F_CC, F_D, F_E, F_MF, F_QF, F_CF, F_X0 };

__global__ void __launch_bounds__(128) generate_kernel(Layout L, int discrete, double* __restrict__ lq, double* __restrict__ term,
                                                       double* __restrict__ x0, uint64_t seedmix, long long first_index, double dt,
                                                       double invSqrtN, double invNm, double invN) {
  extern __shared__ double M[];
  const int n = L.n, m = L.m, nc = L.ncmax, nm = n + m;
  const int prob = blockIdx.x / (L.nodes + 1);
  const int node = blockIdx.x % (L.nodes + 1);
  const long long gp = first_index + prob;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double scale = discrete ? dt : 1.0;
  if (node < L.nodes) {
    double* rec = lq + ((size_t)prob * L.nodes + node) * L.rec;
    for (int idx = tid; idx < nm * nm; idx += nt) M[idx] = urand(seedmix, gp, node, F_M, idx);
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx % n, j = idx / n;
      const double ac = __dmul_rn(urand(seedmix, gp, node, F_A, idx), invSqrtN);
      rec[L.oA + idx] = discrete ? __dadd_rn((i == j) ? 1.0 : 0.0, __dmul_rn(dt, ac)) : ac;
    }
    for (int idx = tid; idx < n * m; idx += nt) rec[L.oB + idx] = __dmul_rn(scale, urand(seedmix, gp, node, F_B, idx));
    for (int i = tid; i < n; i += nt) {
      rec[L.oHv + i] = __dmul_rn(0.01, urand(seedmix, gp, node, F_HV, i));
      rec[L.oq + i] = __dmul_rn(scale, urand(seedmix, gp, node, F_Q, i));
    }
    for (int i = tid; i < m; i += nt) rec[L.or_ + i] = __dmul_rn(scale, urand(seedmix, gp, node, F_R, i));
    if (tid == 0) rec[L.oc] = __dmul_rn(scale, __dmul_rn(0.5, __dadd_rn(urand(seedmix, gp, node, F_C, 0), 1.0)));
    for (int idx = tid; idx < nc * n; idx += nt) rec[L.oC + idx] = urand(seedmix, gp, node, F_CC, idx);
    for (int idx = tid; idx < nc * m; idx += nt) {
      const int i = idx % nc, j = idx / nc;
      rec[L.oD + idx] = (j < nc) ? ((i == j) ? 1.0 : 0.0) : urand(seedmix, gp, node, F_D, idx);
    }
    for (int i = tid; i < nc; i += nt) rec[L.oe + i] = __dmul_rn(0.1, urand(seedmix, gp, node, F_E, i));
    __syncthreads();
    auto W = [&](int a, int b) {
      double acc = 0.0;
      for (int l = 0; l < nm; ++l) acc = fma(M[l + a * nm], M[l + b * nm], acc);
      return fma(acc, invNm, (a == b) ? 0.1 : 0.0);
    };
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx % n, j = idx / n;
      rec[L.oQ + idx] = __dmul_rn(scale, W(min(i, j), max(i, j)));
    }
    for (int idx = tid; idx < m * n; idx += nt) {
      const int i = idx % m, j = idx / m;
      rec[L.oP + idx] = __dmul_rn(scale, W(j, n + i));
    }
    for (int idx = tid; idx < m * m; idx += nt) {
      const int i = idx % m, j = idx / m;
      rec[L.oR + idx] = __dmul_rn(scale, W(n + min(i, j), n + max(i, j)));
    }
  } else {
    double* t = term + (size_t)prob * L.trec;
    for (int idx = tid; idx < n * n; idx += nt) M[idx] = urand(seedmix, gp, 1023, F_MF, idx);
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx % n, j = idx / n;
      const int a = min(i, j), b = max(i, j);
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc = fma(M[l + a * n], M[l + b * n], acc);
      t[L.oQf + idx] = fma(acc, invN, (i == j) ? 0.1 : 0.0);
    }
    for (int i = tid; i < n; i += nt) {
      t[L.oqf + i] = urand(seedmix, gp, 1023, F_QF, i);
      x0[(size_t)prob * n + i] = urand(seedmix, gp, 1023, F_X0, i);
    }
    if (tid == 0) t[L.ocf] = __dmul_rn(0.5, __dadd_rn(urand(seedmix, gp, 1023, F_CF, 0), 1.0));
  }
}

uint64_t host_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

}  // namespace

cudaError_t launch_generate(const Layout& L, int algorithm, double* lq, double* term, double* x0, uint64_t seed, int64_t first_index,
                            double dt, int batch, cudaStream_t stream) {
  const int nm = L.n + L.m;
  const size_t smem = (size_t)nm * nm * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const double invSqrtN = 1.0 / sqrt((double)L.n);
  const double invNm = 1.0 / (double)nm;
  const double invN = 1.0 / (double)L.n;
  // grid.x limit is 2^31-1: batch*(nodes+1) stays far below for every supported size
  const long long blocks = (long long)batch * (L.nodes + 1);
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  generate_kernel<<<(unsigned)blocks, 128, smem, stream>>>(L, algorithm == O2C_ALG_ILQR ? 1 : 0, lq, term, x0, host_mix64(seed),
                                                          (long long)first_index, dt, invSqrtN, invNm, invN);
  return cudaGetLastError();
}

}  // namespace o2c
