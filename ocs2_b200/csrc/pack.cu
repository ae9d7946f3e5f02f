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
    copy_in(v.Q, lp, node, n * n, rec + L.oQ, tid, nt);
    copy_in(v.P, lp, node, m * n, rec + L.oP, tid, nt);
    copy_in(v.R, lp, node, m * m, rec + L.oR, tid, nt);
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

}  // namespace

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
