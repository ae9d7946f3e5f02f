// line_search.cu — batched Armijo line search on the LQ model (the step after the backward pass in the DDP iteration).
//
// LineSearchStrategy::run / lineSearchTask (ocs2_ddp/src/search_strategy/LineSearchStrategy.cpp:125-258) rolls the system out for the
// step lengths alpha_e = maxStepLength * contractionRate^e >= minStepLength on worker threads and keeps the largest one whose merit
// satisfies  merit(alpha) < baselineMerit - armijoCoefficient * alpha * IS(deltaBias)   ("equivalent to a single core line search").
// Here every candidate of every problem is rolled out on the LQ model in ONE launch (o2c_rollout's kernels) and
//   * merit_kernel    evaluates the merit of every (candidate, problem): the LQ-model cost along the rollout
//                     sum_k [ c + q.dx + r.du + 1/2 dx'Q dx + du'P dx + 1/2 du'R du ] + final cost   (one warp per rollout)
//   * select_kernel   integrates |deltaBias|^2 over the controller time stamps by the trapezoidal rule (computeControllerUpdateIS,
//                     DDP_HelperFunctions.cpp:285-291; TrapezoidalIntegration.h:43-58) and applies the Armijo rule per problem.
#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) merit_kernel(Layout L, DeviceBuffers buf, int out_nodes, int n_alpha, int batch, int begin,
                                                                  int count, double* __restrict__ merit /* [max_alphas][batch] */) {
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (task >= (long long)count * n_alpha) return;
  const int ia = (int)(task / count);
  const int prob = begin + (int)(task % count);
  const int n = L.n, m = L.m, N = L.N;
  const double* lqp = buf.lq + (size_t)prob * L.nodes * L.rec;
  const double* xs = buf.xs + ((size_t)ia * batch + prob) * (size_t)out_nodes * n;
  const double* us = buf.us + ((size_t)ia * batch + prob) * (size_t)out_nodes * m;
  const double* xnom = buf.x_nom ? buf.x_nom + (size_t)prob * (N + 1) * n : nullptr;
  const double* unom = buf.u_nom ? buf.u_nom + (size_t)prob * (N + 1) * m : nullptr;
  // every lane accumulates the terms of the rows it owns; one warp reduction at the end
  double part = 0.0;
  for (int k = 0; k < N; ++k) {
    const double* rec = lqp + (size_t)k * L.rec;
    const double* x = xs + (size_t)k * n;
    const double* u = us + (size_t)k * m;
    const double* xn = xnom ? xnom + (size_t)k * n : nullptr;
    const double* un = unom ? unom + (size_t)k * m : nullptr;
    for (int i = lane; i < n; i += 32) {
      const double dxi = x[i] - (xn ? xn[i] : 0.0);
      double qx = 0.0;
      for (int j = 0; j < n; ++j) qx = fma(rec[L.oQ + i + j * n], x[j] - (xn ? xn[j] : 0.0), qx);
      part += dxi * (rec[L.oq + i] + 0.5 * qx);
    }
    const bool jump = buf.event != nullptr && buf.event[(size_t)prob * L.nodes + k] != 0;  // pre-jump cost: state terms only
    for (int i = lane; i < m && !jump; i += 32) {
      const double dui = u[i] - (un ? un[i] : 0.0);
      double px = 0.0, ru = 0.0;
      for (int j = 0; j < n; ++j) px = fma(rec[L.oP + i + j * m], x[j] - (xn ? xn[j] : 0.0), px);
      for (int j = 0; j < m; ++j) ru = fma(rec[L.oR + i + j * m], u[j] - (un ? un[j] : 0.0), ru);
      part += dui * (rec[L.or_ + i] + px + 0.5 * ru);
    }
    if (lane == 0) part += rec[L.oc];
  }
  {
    const double* term = buf.term + (size_t)prob * L.trec;
    const double* x = xs + (size_t)N * n;
    const double* xn = xnom ? xnom + (size_t)N * n : nullptr;
    for (int i = lane; i < n; i += 32) {
      const double dxi = x[i] - (xn ? xn[i] : 0.0);
      double qx = 0.0;
      for (int j = 0; j < n; ++j) qx = fma(term[L.oQf + i + j * n], x[j] - (xn ? xn[j] : 0.0), qx);
      part += dxi * (term[L.oqf + i] + 0.5 * qx);
    }
    if (lane == 0) part += term[L.ocf];
  }
  const double J = warp_sum(part);
  if (lane == 0) merit[(size_t)ia * batch + prob] = J;
}

// one thread per problem: IS(deltaBias), default baseline, Armijo selection
__global__ void select_kernel(Layout L, DeviceBuffers buf, const double* __restrict__ merit, const double* __restrict__ alphas, int n_alpha,
                              int batch, int begin, int count, double armijo, const double* __restrict__ baseline_in /* [count] or null */,
                              double* __restrict__ baseline_out, double* __restrict__ update_is, double* __restrict__ step, int* __restrict__ index) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int prob = begin + t;
  const int m = L.m, N = L.N;
  const double* solp = buf.sol + (size_t)prob * (N + 1) * L.orec;
  double is = 0.0, prev = 0.0;
  for (int k = 0; k <= N; ++k) {
    double sq = 0.0;
    for (int i = 0; i < m; ++i) {
      const double d = solp[(size_t)k * L.orec + L.odb + i];
      sq = fma(d, d, sq);
    }
    if (k >= 1) is += (prev + sq) * (0.5 * (buf.time[k] - buf.time[k - 1]));
    prev = sq;
  }
  double base;
  if (baseline_in) {
    base = baseline_in[t];
  } else {  // cost of the zero-deviation trajectory: the constants of the LQ model
    base = buf.term[(size_t)prob * L.trec + L.ocf];
    const double* lqp = buf.lq + (size_t)prob * L.nodes * L.rec;
    for (int k = 0; k < N; ++k) base += lqp[(size_t)k * L.rec + L.oc];
  }
  double best = 0.0;
  int best_i = -1;
  for (int e = 0; e < n_alpha; ++e) {  // candidates are sorted from the largest step length down: the first hit is the answer
    const double a = alphas[e];
    if (merit[(size_t)e * batch + prob] < base - armijo * a * is) {
      best = a;
      best_i = e;
      break;
    }
  }
  baseline_out[prob] = base;
  update_is[prob] = is;
  step[prob] = best;
  index[prob] = best_i;
}

}  // namespace

cudaError_t launch_merit(const Layout& L, const DeviceBuffers& buf, int out_nodes, int n_alpha, int batch, int begin, int count, double* merit,
                         cudaStream_t stream) {
  const long long tasks = (long long)count * n_alpha;
  if (tasks == 0) return cudaSuccess;
  const int grid = (int)((tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
  merit_kernel<<<grid, kWarpsPerBlock * 32, 0, stream>>>(L, buf, out_nodes, n_alpha, batch, begin, count, merit);
  return cudaGetLastError();
}

cudaError_t launch_select(const Layout& L, const DeviceBuffers& buf, const double* merit, const double* alphas_dev, int n_alpha, int batch,
                          int begin, int count, double armijo, const double* baseline_in_dev, double* baseline_out, double* update_is,
                          double* step, int* index, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  select_kernel<<<(count + 127) / 128, 128, 0, stream>>>(L, buf, merit, alphas_dev, n_alpha, batch, begin, count, armijo, baseline_in_dev,
                                                        baseline_out, update_is, step, index);
  return cudaGetLastError();
}

}  // namespace o2c
