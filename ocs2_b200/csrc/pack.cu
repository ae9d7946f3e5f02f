// pack.cu — conversion between caller-facing strided struct-of-arrays views (device memory) and the interleaved device records.
#include "o2c_common.cuh"

namespace o2c {
namespace {

__device__ __forceinline__ void copy_in(const FieldDev& f, long long p, long long node, int count, double* dst, int tid, int nt) {
  if (f.ptr == nullptr) {
    for (int i = tid; i < count; i += nt) dst[i] = 0.0;
    return;
  }
  const double* src = f.ptr + p * f.ps + node * f.ns;
  for (int i = tid; i < count; i += nt) dst[i] = src[i];
}
// symmetric dim x dim block from its packed upper triangle (column by column: (i, j), i <= j, at j (j + 1) / 2 + i)
__device__ __forceinline__ void copy_in_sym(const FieldDev& f, long long p, long long node, int dim, double* dst, int tid, int nt) {
  if (f.ptr == nullptr) {
    for (int i = tid; i < dim * dim; i += nt) dst[i] = 0.0;
    return;
  }
  const double* src = f.ptr + p * f.ps + node * f.ns;
  for (int idx = tid; idx < dim * dim; idx += nt) {
    const int i = idx % dim, j = idx / dim;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    dst[idx] = src[hi * (hi + 1) / 2 + lo];
  }
}
__device__ __forceinline__ void copy_out(const FieldDev& f, long long p, long long node, int count, const double* src, int tid, int nt) {
  if (f.ptr == nullptr) return;
  double* dst = f.ptr + p * f.ps + node * f.ns;
  for (int i = tid; i < count; i += nt) dst[i] = src[i];
}

__global__ void __launch_bounds__(128) pack_kernel(Layout L, LqViewDev v, double* lq, double* term, double* x_nom, double* u_nom, int* nc,
                                                   double* x0, int begin, int count) {
  const int span = max(L.nodes, L.N + 1);
  const int lp = blockIdx.x / span;  // index into the view
  const int node = blockIdx.x % span;
  const int prob = begin + lp;
  const int n = L.n, m = L.m, ncm = L.ncmax, tid = threadIdx.x, nt = blockDim.x;
  if (node < L.nodes) {
    double* rec = lq + ((size_t)prob * L.nodes + node) * L.rec;
    copy_in(v.A, lp, node, n * n, rec + L.oA, tid, nt);
    copy_in(v.B, lp, node, n * m, rec + L.oB, tid, nt);
    if (v.sym_packed) {
      copy_in_sym(v.Q, lp, node, n, rec + L.oQ, tid, nt);
      copy_in_sym(v.R, lp, node, m, rec + L.oR, tid, nt);
    } else {
      copy_in(v.Q, lp, node, n * n, rec + L.oQ, tid, nt);
      copy_in(v.R, lp, node, m * m, rec + L.oR, tid, nt);
    }
    copy_in(v.P, lp, node, m * n, rec + L.oP, tid, nt);
    copy_in(v.Hv, lp, node, n, rec + L.oHv, tid, nt);
    copy_in(v.q, lp, node, n, rec + L.oq, tid, nt);
    copy_in(v.r, lp, node, m, rec + L.or_, tid, nt);
    copy_in(v.c, lp, node, 1, rec + L.oc, tid, nt);
    if (ncm > 0) {
      copy_in(v.C, lp, node, ncm * n, rec + L.oC, tid, nt);
      copy_in(v.D, lp, node, ncm * m, rec + L.oD, tid, nt);
      copy_in(v.e, lp, node, ncm, rec + L.oe, tid, nt);
      if (nc != nullptr && tid == 0) nc[(size_t)prob * L.nodes + node] = v.nc ? v.nc[lp * v.nc_ps + node * v.nc_ns] : ncm;
    }
  }
  if (node <= L.N) {
    if (x_nom) copy_in(v.x_nom, lp, node, n, x_nom + ((size_t)prob * (L.N + 1) + node) * n, tid, nt);
    if (u_nom) copy_in(v.u_nom, lp, node, m, u_nom + ((size_t)prob * (L.N + 1) + node) * m, tid, nt);
  }
  if (node == 0) {
    double* t = term + (size_t)prob * L.trec;
    if (v.sym_packed)
      copy_in_sym(v.Qf, lp, 0, n, t + L.oQf, tid, nt);
    else
      copy_in(v.Qf, lp, 0, n * n, t + L.oQf, tid, nt);
    copy_in(v.qf, lp, 0, n, t + L.oqf, tid, nt);
    copy_in(v.cf, lp, 0, 1, t + L.ocf, tid, nt);
    if (v.x0.ptr) copy_in(v.x0, lp, 0, n, x0 + (size_t)prob * n, tid, nt);
  }
}

__global__ void __launch_bounds__(128) unpack_kernel(Layout L, SolViewDev v, const double* sol, const double* xs, const double* us,
                                                     const int* status, int out_nodes, int n_alpha, int batch, int begin, int count) {
  const int span = max(L.N + 1, out_nodes);
  const int lp = blockIdx.x / span;
  const int node = blockIdx.x % span;
  const int prob = begin + lp;
  const int n = L.n, m = L.m, tid = threadIdx.x, nt = blockDim.x;
  if (node <= L.N) {
    const double* rec = sol + ((size_t)prob * (L.N + 1) + node) * L.orec;
    copy_out(v.K, lp, node, m * n, rec + L.oK, tid, nt);
    copy_out(v.dbias, lp, node, m, rec + L.odb, tid, nt);
    copy_out(v.bias, lp, node, m, rec + L.obias, tid, nt);
    copy_out(v.Sm, lp, node, n * n, rec + L.oSm, tid, nt);
    copy_out(v.Sv, lp, node, n, rec + L.oSv, tid, nt);
    copy_out(v.s, lp, node, 1, rec + L.os, tid, nt);
  }
  if (node < out_nodes) {
    for (int a = 0; a < n_alpha; ++a) {
      if (v.x.ptr) {
        const double* src = xs + (((size_t)a * batch + prob) * out_nodes + node) * n;
        double* dst = v.x.ptr + a * v.x_as + (long long)lp * v.x.ps + (long long)node * v.x.ns;
        for (int i = tid; i < n; i += nt) dst[i] = src[i];
      }
      if (v.u.ptr) {
        const double* src = us + (((size_t)a * batch + prob) * out_nodes + node) * m;
        double* dst = v.u.ptr + a * v.u_as + (long long)lp * v.u.ps + (long long)node * v.u.ns;
        for (int i = tid; i < m; i += nt) dst[i] = src[i];
      }
    }
  }
  if (node == 0 && tid == 0 && v.status) v.status[lp] = status[prob];
}

// LinearController::flattenSingle at the controller's own time stamps (ocs2_core/src/control/LinearController.cpp:107-140): one float
// record of m*(n+1) values per node, row i = [uff_i, K_i,:] (row-major), uff = bias + alpha * deltaBias (incrementController applied)
__global__ void __launch_bounds__(128) flatten_kernel(Layout L, const double* __restrict__ sol, float* __restrict__ out, double alpha, int begin,
                                                      int count) {
  const int lp = blockIdx.x / (L.N + 1), node = blockIdx.x % (L.N + 1);
  if (lp >= count) return;
  const int n = L.n, m = L.m, len = m * (n + 1);
  const double* rec = sol + ((size_t)(begin + lp) * (L.N + 1) + node) * L.orec;
  float* dst = out + ((size_t)lp * (L.N + 1) + node) * len;
  for (int idx = threadIdx.x; idx < len; idx += blockDim.x) {
    const int i = idx / (n + 1), j = idx % (n + 1);
    const double v = j == 0 ? rec[L.obias + i] + alpha * rec[L.odb + i] : rec[L.oK + i + m * (j - 1)];
    dst[idx] = static_cast<float>(v);
  }
}

// rk4SensitivityDiscretization (SensitivityIntegratorImpl.cpp:130-169) + ILQR::discreteLQWorker (ILQR.cpp:137-157); one CTA per
// (problem, node), all eight stage matrices in shared memory
__device__ __forceinline__ void cta_gemm_acc(int M, int N, int K, double alpha, const double* A, const double* B, double* C) {
  // C(MxN) += alpha * A(MxK) * B(KxN), column-major, leading dimensions = rows
  for (int idx = threadIdx.x; idx < M * N; idx += blockDim.x) {
    const int i = idx % M, j = idx / M;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(A[i + k * M], B[k + j * K], acc);
    C[idx] += alpha * acc;
  }
  __syncthreads();
}
__global__ void __launch_bounds__(128) discretize_kernel(Layout L, DiscretizeArgs a, double* __restrict__ lq, int begin, int count) {
  extern __shared__ __align__(16) double dsm[];
  const int lp = blockIdx.x / L.N, node = blockIdx.x % L.N;
  if (lp >= count) return;
  const int n = L.n, m = L.m, nn = n * n, nm = n * m;
  double* A[4] = {dsm, dsm + nn, dsm + 2 * nn, dsm + 3 * nn};
  double* B[4] = {dsm + 4 * nn, dsm + 4 * nn + nm, dsm + 4 * nn + 2 * nm, dsm + 4 * nn + 3 * nm};
  double* tmp = dsm + 4 * nn + 4 * nm;
  for (int s = 0; s < 4; ++s) {
    const int src = a.stages == 4 ? s : 0;
    const double* pa = a.dfdx[src].ptr + (long long)lp * a.dfdx[src].ps + (long long)node * a.dfdx[src].ns;
    const double* pb = a.dfdu[src].ptr + (long long)lp * a.dfdu[src].ps + (long long)node * a.dfdu[src].ns;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) A[s][i] = pa[i];
    for (int i = threadIdx.x; i < nm; i += blockDim.x) B[s][i] = pb[i];
  }
  __syncthreads();
  const double dt = a.dt[node];
  double* rec = lq + ((size_t)(begin + lp) * L.nodes + node) * L.rec;
  if (dt == 0.0) {  // zero-length interval: the node keeps its continuous-time model data (ILQR.cpp:123-130)
    for (int i = threadIdx.x; i < nn; i += blockDim.x) rec[L.oA + i] = A[0][i];
    for (int i = threadIdx.x; i < nm; i += blockDim.x) rec[L.oB + i] = B[0][i];
    return;
  }
  const double h2 = dt / 2.0, h6 = dt / 6.0, h3 = dt / 3.0;
  // input sensitivity: dk2/du += dt/2 k2.dfdx dk1/du, dk3/du += dt/2 k3.dfdx dk2/du, dk4/du += dt k4.dfdx dk3/du
  cta_gemm_acc(n, m, n, h2, A[1], B[0], B[1]);
  cta_gemm_acc(n, m, n, h2, A[2], B[1], B[2]);
  cta_gemm_acc(n, m, n, dt, A[3], B[2], B[3]);
  // state sensitivity: one temporary per product to avoid aliasing, as the reference does
  for (int s = 1; s < 4; ++s) {
    for (int i = threadIdx.x; i < nn; i += blockDim.x) tmp[i] = 0.0;
    __syncthreads();
    cta_gemm_acc(n, n, n, s == 3 ? dt : h2, A[s], A[s - 1], tmp);
    for (int i = threadIdx.x; i < nn; i += blockDim.x) A[s][i] += tmp[i];
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
    const double v = h6 * A[0][idx] + h3 * A[1][idx] + h3 * A[2][idx] + h6 * A[3][idx];
    rec[L.oA + idx] = v + ((idx % n == idx / n) ? 1.0 : 0.0);
  }
  for (int idx = threadIdx.x; idx < nm; idx += blockDim.x) rec[L.oB + idx] = h6 * B[0][idx] + h3 * B[1][idx] + h3 * B[2][idx] + h6 * B[3][idx];
  for (int i = threadIdx.x; i < n; i += blockDim.x) rec[L.oHv + i] = 0.0;  // dynamicsBias.setZero (ILQR.cpp:144)
  if (a.scale_cost) {  // modelData.cost *= timeStep (ILQR.cpp:149-150)
    for (int i = threadIdx.x; i < nn; i += blockDim.x) rec[L.oQ + i] *= dt;
    for (int i = threadIdx.x; i < nm; i += blockDim.x) rec[L.oP + i] *= dt;
    for (int i = threadIdx.x; i < m * m; i += blockDim.x) rec[L.oR + i] *= dt;
    for (int i = threadIdx.x; i < n; i += blockDim.x) rec[L.oq + i] *= dt;
    for (int i = threadIdx.x; i < m; i += blockDim.x) rec[L.or_ + i] *= dt;
    if (threadIdx.x == 0) rec[L.oc] *= dt;
  }
}

}  // namespace

cudaError_t launch_discretize(const Layout& L, const DiscretizeArgs& a, double* lq, int begin, int count, cudaStream_t stream) {
  const long long blocks = (long long)count * L.N;
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (blocks == 0) return cudaSuccess;
  const size_t smem = sizeof(double) * (size_t)(5 * L.n * L.n + 4 * L.n * L.m);
  cudaError_t e = cudaFuncSetAttribute(discretize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  discretize_kernel<<<(unsigned)blocks, 128, smem, stream>>>(L, a, lq, begin, count);
  return cudaGetLastError();
}

cudaError_t launch_flatten(const Layout& L, const double* sol, float* out, double alpha, int begin, int count, cudaStream_t stream) {
  const long long blocks = (long long)count * (L.N + 1);
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (blocks == 0) return cudaSuccess;
  flatten_kernel<<<(unsigned)blocks, 128, 0, stream>>>(L, sol, out, alpha, begin, count);
  return cudaGetLastError();
}

cudaError_t launch_pack(const Layout& L, const LqViewDev& v, double* lq, double* term, double* x_nom, double* u_nom, int* nc, double* x0,
                        int begin, int count, cudaStream_t stream) {
  const long long blocks = (long long)count * (L.nodes > L.N + 1 ? L.nodes : L.N + 1);
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (blocks == 0) return cudaSuccess;
  pack_kernel<<<(unsigned)blocks, 128, 0, stream>>>(L, v, lq, term, x_nom, u_nom, nc, x0, begin, count);
  return cudaGetLastError();
}

cudaError_t launch_unpack(const Layout& L, const SolViewDev& v, const double* sol, const double* xs, const double* us, const int* status,
                          int out_nodes, int n_alpha, int batch, int begin, int count, cudaStream_t stream) {
  const long long blocks = (long long)count * (L.N + 1 > out_nodes ? L.N + 1 : out_nodes);
  if (blocks > 2147483647LL) return cudaErrorInvalidValue;
  if (blocks == 0) return cudaSuccess;
  unpack_kernel<<<(unsigned)blocks, 128, 0, stream>>>(L, v, sol, xs, us, status, out_nodes, n_alpha, batch, begin, count);
  return cudaGetLastError();
}

}  // namespace o2c
