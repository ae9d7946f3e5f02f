// rollout.cu — forward rollout of the LQ model under the LinearController, one warp per (step length, problem).
//
//   policy    u(t,x) = bias(t) + alpha * deltaBias(t) + K(t) x      incrementController ocs2_ddp/src/DDP_HelperFunctions.cpp:296-304,
//                                                                   LinearController::computeInput ocs2_core/src/control/LinearController.cpp:79-87
//   discrete  x_{k+1} = x_nom_{k+1} + A_k (x_k - x_nom_k) + B_k (u_k - u_nom_k) + Hv_k
//                                                                   (discrete-model semantics ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:48-52)
//   continuous xdot = A(t)(x - x_nom(t)) + B(t)(u - u_nom(t)) + Hv(t), lerped on the node grid, classic RK4 with the
//             constant-step schedule of boost::odeint integrate_adaptive for a plain stepper (host-precomputed RolloutStep list);
//             inputs re-evaluated at every output node (TimeTriggeredRollout.cpp:98-102).
#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) rollout_discrete_kernel(Layout L, DeviceBuffers buf, const double* __restrict__ alphas,
                                                                             int n_alpha, int batch, int begin, int count) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (task >= (long long)count * n_alpha) return;
  const int ia = (int)(task / count);
  const int prob = begin + (int)(task % count);
  const int n = L.n, m = L.m, N = L.N;
  const double alpha = alphas[ia];
  double* xs = smem + (size_t)warp * (2 * n + 2 * m);
  double* dx = xs + n;
  double* us = dx + n;
  double* du = us + m;
  const double* solp = buf.sol + (size_t)prob * (N + 1) * L.orec;
  const double* lqp = buf.lq + (size_t)prob * L.nodes * L.rec;
  double* xo = buf.xs + ((size_t)ia * batch + prob) * (size_t)(N + 1) * n;
  double* uo = buf.us + ((size_t)ia * batch + prob) * (size_t)(N + 1) * m;
  const double* xnom = buf.x_nom ? buf.x_nom + (size_t)prob * (N + 1) * n : nullptr;
  const double* unom = buf.u_nom ? buf.u_nom + (size_t)prob * (N + 1) * m : nullptr;
  for (int i = lane; i < n; i += 32) xs[i] = buf.x0[(size_t)prob * n + i];
  __syncwarp();
  bool finite = true;
  for (int k = 0; k <= N; ++k) {
    const double* rec = solp + (size_t)k * L.orec;
    const double* K = rec + L.oK;
    for (int i = lane; i < n; i += 32) {
      xo[(size_t)k * n + i] = xs[i];
      finite = finite && isfinite(xs[i]);
      dx[i] = xs[i] - (xnom ? xnom[(size_t)k * n + i] : 0.0);
    }
    for (int i = lane; i < m; i += 32) {
      double acc = rec[L.obias + i] + alpha * rec[L.odb + i];
      for (int j = 0; j < n; ++j) acc = fma(K[i + j * m], xs[j], acc);
      us[i] = acc;
      uo[(size_t)k * m + i] = acc;
      du[i] = acc - (unom ? unom[(size_t)k * m + i] : 0.0);
    }
    __syncwarp();
    if (k == N) break;
    const double* st = lqp + (size_t)k * L.rec;
    const double* A = st + L.oA;
    const double* B = st + L.oB;
    const bool jump = buf.event != nullptr && buf.event[(size_t)prob * L.nodes + k] != 0;  // pre-event node: x+ = x_nom+ + A_e dx + Hv_e
    double xn_[2];
    int cnt = 0;
    for (int i = lane; i < n; i += 32) {
      double acc = st[L.oHv + i] + (xnom ? xnom[(size_t)(k + 1) * n + i] : 0.0);
      for (int j = 0; j < n; ++j) acc = fma(A[i + j * n], dx[j], acc);
      if (!jump)
        for (int j = 0; j < m; ++j) acc = fma(B[i + j * n], du[j], acc);
      xn_[cnt++] = acc;  // n <= 64
    }
    __syncwarp();
    cnt = 0;
    for (int i = lane; i < n; i += 32) xs[i] = xn_[cnt++];
    __syncwarp();
  }
  if (!__all_sync(0xffffffffu, finite) && lane == 0) atomicOr(buf.status + prob, O2C_STATUS_NONFINITE);
}

struct ContWork {
  double *x, *xt, *k1, *k2, *k3, *k4, *u, *dx;
};

// u = lerp(bias + alpha dbias)(idx, a) + lerp(K)(idx, a) * xv
__device__ void policy_eval(const Layout& L, const double* solp, int idx, double a, double alpha, const double* xv, double* uout) {
  const int n = L.n, m = L.m;
  const double* l = solp + (size_t)idx * L.orec;
  const double* r = l + L.orec;
  const double b = 1.0 - a;
  for (int i = lane_id(); i < m; i += 32) {
    const double bl = l[L.obias + i] + alpha * l[L.odb + i];
    const double br = r[L.obias + i] + alpha * r[L.odb + i];
    double acc = a * bl + b * br;
    for (int j = 0; j < n; ++j) acc = fma(a * l[L.oK + i + j * m] + b * r[L.oK + i + j * m], xv[j], acc);
    uout[i] = acc;
  }
  __syncwarp();
}

__device__ void cont_flow(const Layout& L, const DeviceBuffers& buf, const double* lqp, const double* solp, const double* xnom,
                          const double* unom, int idx, double a, double alpha, const double* xv, double* dxdt, ContWork& W) {
  const int n = L.n, m = L.m, lane = lane_id();
  policy_eval(L, solp, idx, a, alpha, xv, W.u);
  const double* l = lqp + (size_t)idx * L.rec;
  const double* r = l + L.rec;
  const double b = 1.0 - a;
  for (int i = lane; i < n; i += 32)
    W.dx[i] = xv[i] - (xnom ? a * xnom[(size_t)idx * n + i] + b * xnom[(size_t)(idx + 1) * n + i] : 0.0);
  for (int i = lane; i < m; i += 32)
    if (unom) W.u[i] -= a * unom[(size_t)idx * m + i] + b * unom[(size_t)(idx + 1) * m + i];
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    double acc = a * l[L.oHv + i] + b * r[L.oHv + i];
    for (int j = 0; j < n; ++j) acc = fma(a * l[L.oA + i + j * n] + b * r[L.oA + i + j * n], W.dx[j], acc);
    for (int j = 0; j < m; ++j) acc = fma(a * l[L.oB + i + j * n] + b * r[L.oB + i + j * n], W.u[j], acc);
    dxdt[i] = acc;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    rollout_continuous_kernel(Layout L, DeviceBuffers buf, const RolloutStep* __restrict__ steps, int nsteps, int first_idx,
                              double first_alpha, int out_nodes, const double* __restrict__ alphas, int n_alpha, int batch, int begin,
                              int count) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (task >= (long long)count * n_alpha) return;
  const int ia = (int)(task / count);
  const int prob = begin + (int)(task % count);
  const int n = L.n, m = L.m, N = L.N;
  const double alpha = alphas[ia];
  double* base = smem + (size_t)warp * (7 * n + m);
  ContWork W{base, base + n, base + 2 * n, base + 3 * n, base + 4 * n, base + 5 * n, base + 7 * n, base + 6 * n};
  const double* solp = buf.sol + (size_t)prob * (N + 1) * L.orec;
  const double* lqp = buf.lq + (size_t)prob * L.nodes * L.rec;
  double* xo = buf.xs + ((size_t)ia * batch + prob) * (size_t)out_nodes * n;
  double* uo = buf.us + ((size_t)ia * batch + prob) * (size_t)out_nodes * m;
  const double* xnom = buf.x_nom ? buf.x_nom + (size_t)prob * (N + 1) * n : nullptr;
  const double* unom = buf.u_nom ? buf.u_nom + (size_t)prob * (N + 1) * m : nullptr;
  for (int i = lane; i < n; i += 32) W.x[i] = buf.x0[(size_t)prob * n + i];
  __syncwarp();
  bool finite = true;
  auto observe = [&](int o, int idx, double a) {
    policy_eval(L, solp, idx, a, alpha, W.x, W.u);
    for (int i = lane; i < n; i += 32) {
      xo[(size_t)o * n + i] = W.x[i];
      finite = finite && isfinite(W.x[i]);
    }
    for (int i = lane; i < m; i += 32) uo[(size_t)o * m + i] = W.u[i];
    __syncwarp();
  };
  observe(0, first_idx, first_alpha);
  for (int s = 0; s < nsteps; ++s) {
    const RolloutStep sp = steps[s];
    const double h = sp.h;
    if (sp.jump > 0) {
      // an event (TimeTriggeredRollout.cpp:104-108): the LQ model's jump map x+ = x_nom(post) + A_e (x - x_nom(pre)) + Hv_e
      const double* jr = buf.jump + ((size_t)prob * buf.jump_capacity + (sp.jump - 1)) * jump_rec(n);
      for (int i = lane; i < n; i += 32) W.dx[i] = W.x[i] - (xnom ? xnom[(size_t)sp.pre_node * n + i] : 0.0);
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        double acc = jr[jump_oHv(n) + i] + (xnom ? xnom[(size_t)(sp.pre_node + 1) * n + i] : 0.0);
        for (int k = 0; k < n; ++k) acc = fma(jr[i + (size_t)k * n], W.dx[k], acc);
        W.xt[i] = acc;
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) W.x[i] = W.xt[i];
      __syncwarp();
    }
    if (h == 0.0) {  // a jump or a degenerate interval: no integration
      observe(s + 1, sp.obs_idx, sp.obs_alpha);
      continue;
    }
    cont_flow(L, buf, lqp, solp, xnom, unom, sp.idx[0], sp.alpha[0], alpha, W.x, W.k1, W);
    for (int i = lane; i < n; i += 32) W.xt[i] = W.x[i] + (h * 0.5) * W.k1[i];
    __syncwarp();
    cont_flow(L, buf, lqp, solp, xnom, unom, sp.idx[1], sp.alpha[1], alpha, W.xt, W.k2, W);
    for (int i = lane; i < n; i += 32) W.xt[i] = W.x[i] + (h * 0.5) * W.k2[i];
    __syncwarp();
    cont_flow(L, buf, lqp, solp, xnom, unom, sp.idx[2], sp.alpha[2], alpha, W.xt, W.k3, W);
    for (int i = lane; i < n; i += 32) W.xt[i] = W.x[i] + h * W.k3[i];
    __syncwarp();
    cont_flow(L, buf, lqp, solp, xnom, unom, sp.idx[3], sp.alpha[3], alpha, W.xt, W.k4, W);
    const double b1 = h * (1.0 / 6.0), b2 = h * (1.0 / 3.0);
    for (int i = lane; i < n; i += 32) W.x[i] = W.x[i] + b1 * W.k1[i] + b2 * W.k2[i] + b2 * W.k3[i] + b1 * W.k4[i];
    __syncwarp();
    observe(s + 1, sp.obs_idx, sp.obs_alpha);
  }
  if (!__all_sync(0xffffffffu, finite) && lane == 0) atomicOr(buf.status + prob, O2C_STATUS_NONFINITE);
}

}  // namespace

cudaError_t launch_rollout_discrete(const Layout& L, const DeviceBuffers& buf, const double* alphas_dev, int n_alpha, int batch, int begin,
                                    int count, cudaStream_t stream) {
  if (L.n > 64) return cudaErrorInvalidValue;
  const long long tasks = (long long)count * n_alpha;
  const int grid = (int)((tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const size_t smem = (size_t)kWarpsPerBlock * (2 * L.n + 2 * L.m) * sizeof(double);
  rollout_discrete_kernel<<<grid, kWarpsPerBlock * 32, smem, stream>>>(L, buf, alphas_dev, n_alpha, batch, begin, count);
  return cudaGetLastError();
}

cudaError_t launch_rollout_continuous(const Layout& L, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps, int first_idx,
                                      double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch, int begin,
                                      int count, cudaStream_t stream) {
  const long long tasks = (long long)count * n_alpha;
  const int grid = (int)((tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const size_t smem = (size_t)kWarpsPerBlock * (7 * L.n + L.m) * sizeof(double);
  rollout_continuous_kernel<<<grid, kWarpsPerBlock * 32, smem, stream>>>(L, buf, steps, nsteps, first_idx, first_alpha, out_nodes,
                                                                          alphas_dev, n_alpha, batch, begin, count);
  return cudaGetLastError();
}

}  // namespace o2c
