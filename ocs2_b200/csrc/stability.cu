// stability.cu — ddp::Settings::checkNumericalStability_ for the value function: checkBeingPSD of every S_k
// (ocs2_core/src/Types.cpp:206-236, called from GaussNewtonDDP::solveSequentialRiccatiEquationsImpl, ocs2_ddp/src/GaussNewtonDDP.cpp:555-579).
// One warp per (problem, node): finite entries, self-adjoint in Eigen's isApprox sense (|S - S'|_F <= 1e-6 min(|S|_F, |S'|_F)), smallest
// eigenvalue >= -epsilon. A Cholesky of the symmetric part settles the common case (positive definite); only a matrix it rejects goes
// through cyclic Jacobi rotations for its smallest eigenvalue (the reference calls Eigen's SelfAdjointEigenSolver, which reads the lower
// triangle; the smallest eigenvalue does not depend on the algorithm beyond rounding). Where the reference throws, the problem's status
// word gets O2C_STATUS_NOT_PSD (O2C_STATUS_NONFINITE for non-finite entries).
#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta) check_psd_kernel(Layout L, const double* __restrict__ sol, int* status, int begin, int count) {
  extern __shared__ __align__(16) double sm[];
  const int n = L.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* A = sm + (size_t)warp * 2 * n * n;  // lower triangle of S, mirrored (what SelfAdjointEigenSolver sees)
  double* C = A + n * n;                      // Cholesky work copy
  const long long items = (long long)count * (L.N + 1);
  for (long long item = (long long)blockIdx.x * kWarpsPerCta + warp; item < items; item += (long long)gridDim.x * kWarpsPerCta) {
    const int prob = begin + (int)(item / (L.N + 1)), node = (int)(item % (L.N + 1));
    const double* S = sol + ((size_t)prob * (L.N + 1) + node) * L.orec + L.oSm;
    double diff2 = 0.0, norm2 = 0.0;
    bool finite = true;
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx % n, j = idx / n;
      const double v = S[idx], vt = S[j + i * n];
      finite = finite && isfinite(v);
      diff2 = fma(v - vt, v - vt, diff2);
      norm2 = fma(v, v, norm2);
      const double lo = (i >= j) ? v : vt;
      A[idx] = lo;
      C[idx] = lo;
    }
    diff2 = warp_sum(diff2);
    norm2 = warp_sum(norm2);
    finite = __all_sync(0xffffffffu, finite);
    __syncwarp();
    int bits = 0;
    if (!finite) {
      bits = O2C_STATUS_NONFINITE;
    } else {
      if (diff2 > 1e-12 * norm2) bits |= O2C_STATUS_NOT_PSD;  // not self-adjoint: |S - S'|^2 > prec^2 min(|S|^2, |S'|^2), prec = 1e-6
      // positive definite <=> the Cholesky runs through (warp-cooperative right-looking, lower triangle)
      bool pd = true;
      for (int j = 0; j < n && pd; ++j) {
        const double d = C[j + j * n];
        if (!(d > 0.0)) {
          pd = false;
          break;
        }
        const double rs = rsqrt(d);
        __syncwarp();
        for (int i = j + 1 + lane; i < n; i += 32) C[i + j * n] *= rs;
        __syncwarp();
        for (int k = j + 1 + (lane >> 3); k < n; k += 4) {
          const double lkj = C[k + j * n];
          for (int i = k + (lane & 7); i < n; i += 8) C[i + k * n] = fma(-C[i + j * n], lkj, C[i + k * n]);
        }
        __syncwarp();
      }
      if (!pd) {
        // smallest eigenvalue by cyclic Jacobi on A (warp-uniform control flow: every lane reads the same pivots)
        for (int sweep = 0; sweep < 60; ++sweep) {
          double off = 0.0, dg = 0.0;
          for (int idx = lane; idx < n * n; idx += 32) {
            const double v = A[idx] * A[idx];
            if (idx % n == idx / n) dg += v; else off += v;
          }
          off = warp_sum(off);
          dg = warp_sum(dg);
          if (off <= 1e-30 * dg || off == 0.0) break;
          for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
              const double apq = A[p + q * n];
              if (apq == 0.0) continue;
              const double theta = (A[q + q * n] - A[p + p * n]) / (2.0 * apq);
              const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
              const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
              __syncwarp();
              for (int k = lane; k < n; k += 32) {  // A <- A J (columns p, q)
                const double akp = A[k + p * n], akq = A[k + q * n];
                A[k + p * n] = c * akp - sn * akq;
                A[k + q * n] = sn * akp + c * akq;
              }
              __syncwarp();
              for (int k = lane; k < n; k += 32) {  // A <- J' A (rows p, q)
                const double apk = A[p + k * n], aqk = A[q + k * n];
                A[p + k * n] = c * apk - sn * aqk;
                A[q + k * n] = sn * apk + c * aqk;
              }
              __syncwarp();
            }
        }
        double mn = 1e300;
        for (int i = lane; i < n; i += 32) mn = fmin(mn, A[i + i * n]);
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (mn < -2.220446049250313e-16) bits |= O2C_STATUS_NOT_PSD;  // Eigen::NumTraits<double>::epsilon()
      }
    }
    if (bits && lane == 0) atomicOr(status + prob, bits);
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_check_psd(const Layout& L, const double* sol, int* status, int begin, int count, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * 2 * (size_t)L.n * L.n * kWarpsPerCta;
  cudaError_t e = cudaFuncSetAttribute(check_psd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int sms = device_sm_count();
  if (sms <= 0) return cudaErrorInvalidDevice;
  const long long items = (long long)count * (L.N + 1);
  const long long ctas = (items + kWarpsPerCta - 1) / kWarpsPerCta;
  const int grid = (int)(ctas < (long long)sms * 8 ? ctas : (long long)sms * 8);
  check_psd_kernel<<<grid, 32 * kWarpsPerCta, smem, stream>>>(L, sol, status, begin, count);
  return cudaGetLastError();
}

}  // namespace o2c
