// Parity test of the C++ host front-end (include/ocs2_ddp_cuda/BatchedRiccatiSolver.h) against the CPU oracle.
//
// The reference's Eigen-backed types are absent from this image, so the templates are driven with minimal stand-ins that
// offer the same members (ModelData.h:43-60, LinearController.h:109-112, Types.h ScalarFunctionQuadraticApproximation /
// VectorFunctionLinearApproximation; all blocks column-major like Eigen's default).
//
//   test_frontend --gpu       every case below on cuda:0, checked against oracle/liblq_oracle.so to 1e-9 relative
//   test_frontend --no-gpu    the constructor must throw (there is no CPU fallback)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ocs2_ddp_cuda/BatchedRiccatiSolver.h"
#include "lq_oracle.h"

namespace {

struct Dense {  // stand-in for vector_t / matrix_t
  std::vector<double> v;
  long r = 0, c = 0;
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  long size() const { return r * c; }
  long rows() const { return r; }
  long cols() const { return c; }
  void resize(long n) { r = n, c = 1, v.assign(n, 0.0); }
  void resize(long rr, long cc) { r = rr, c = cc, v.assign(rr * cc, 0.0); }
  void set(const double* src, long rr, long cc) { resize(rr, cc), std::copy(src, src + rr * cc, v.begin()); }
};
struct VectorFunctionLinearApproximation { Dense f, dfdx, dfdu; };
struct ScalarFunctionQuadraticApproximation { double f = 0.0; Dense dfdx, dfdu, dfdxx, dfdux, dfduu; };
struct ModelData {
  int stateDim = 0, inputDim = 0;
  double time = 0.0;
  Dense dynamicsBias;
  VectorFunctionLinearApproximation dynamics;
  ScalarFunctionQuadraticApproximation cost;
  VectorFunctionLinearApproximation stateInputEqConstraint;
};
struct LinearController {
  std::vector<double> timeStamp_;
  std::vector<Dense> biasArray_, deltaBiasArray_, gainArray_;
};

// one problem in the oracle's flat layout
struct Flat {
  int n, m, nc, N, nodes;
  std::vector<double> A, B, Hv, Q, P, R, q, r, c, C, D, e, Qf, qf, cf, x0, xnom, unom, time, jA, jHv, jQ, jq, jc;
  std::vector<int32_t> ncActive, event;
  orc_problem view(bool nominal, bool ragged, bool events) const {
    orc_problem p{};
    p.nx = n, p.nu = m, p.nc_max = nc, p.N = N;
    p.A = A.data(), p.B = B.data(), p.Hv = Hv.data(), p.Q = Q.data(), p.P = P.data(), p.R = R.data();
    p.q = q.data(), p.r = r.data(), p.c = c.data();
    if (nc > 0) p.C = C.data(), p.D = D.data(), p.e = e.data();
    if (ragged) p.nc = ncActive.data();
    p.Qf = Qf.data(), p.qf = qf.data(), p.cf = cf.data();
    if (nominal) p.x_nom = xnom.data(), p.u_nom = unom.data();
    p.time = time.data();
    if (events) p.event = event.data();
    if (events && !jA.empty()) p.jA = jA.data(), p.jHv = jHv.data(), p.jQ = jQ.data(), p.jq = jq.data(), p.jc = jc.data();
    return p;
  }
};

double relErr(const double* got, const double* want, size_t count) {
  double diff = 0.0, scale = 1.0;
  for (size_t i = 0; i < count; ++i) {
    if (!std::isfinite(got[i]) || !std::isfinite(want[i])) return INFINITY;
    diff = std::max(diff, std::fabs(got[i] - want[i]));
    scale = std::max(scale, std::fabs(want[i]));
  }
  return diff / scale;
}

uint64_t lcg(uint64_t& s) { return s = s * 6364136223846793005ull + 1442695040888963407ull; }
double uni(uint64_t& s) { return (double)(lcg(s) >> 11) / 9007199254740992.0 * 2.0 - 1.0; }

using LineSearchResultT = ocs2_ddp_cuda::BatchedRiccatiSolver::LineSearchResult;

struct Case {
  const char* name;
  int algorithm, n, m, nc, N, batch;
  bool nominal, ragged, events;
  const char* kernel;  // substring expected in kernelVariant()
};

int runCase(const Case& cs) {
  const double dt = 0.01, tol = 1e-9;
  const int n = cs.n, m = cs.m, nc = cs.nc, N = cs.N, nodes = cs.algorithm == ORC_ALG_ILQR ? N : N + 1;
  o2c_config cfg{};
  cfg.nx = n, cfg.nu = m, cfg.nc_max = nc, cfg.num_stages = N, cfg.batch = cs.batch, cfg.algorithm = cs.algorithm;
  cfg.riccati_form = O2C_FORM_REDUCED, cfg.strategy = O2C_STRATEGY_LINE_SEARCH, cfg.hessian_correction = O2C_HC_DIAGONAL_SHIFT;
  cfg.device = 0, cfg.max_alphas = 6, cfg.has_nominal = cs.nominal, cfg.hessian_multiple = 1e-5, cfg.time_step = dt;
  orc_settings ost{};
  ost.algorithm = cs.algorithm, ost.reduced_form = 1, ost.strategy = ORC_STRATEGY_LINE_SEARCH, ost.hessian_correction = ORC_HC_DIAGONAL_SHIFT;
  ost.hessian_multiple = 1e-5, ost.time_step = dt;

  const bool slqEvents = cs.events && cs.algorithm == ORC_ALG_SLQ;
  const int slqEventNodes[2] = {3, 8};  // pre-event nodes, the same for every instance (one time grid)
  ocs2_ddp_cuda::BatchedRiccatiSolver solver(cfg, slqEvents ? 2 : 0);
  std::vector<Flat> flats(cs.batch);
  uint64_t rng = 0x9e3779b97f4a7c15ull + n * 131 + m;
  for (int b = 0; b < cs.batch; ++b) {
    Flat& f = flats[b];
    f.n = n, f.m = m, f.nc = nc, f.N = N, f.nodes = nodes;
    f.A.resize((size_t)nodes * n * n), f.B.resize((size_t)nodes * n * m), f.Hv.resize((size_t)nodes * n);
    f.Q.resize((size_t)nodes * n * n), f.P.resize((size_t)nodes * m * n), f.R.resize((size_t)nodes * m * m);
    f.q.resize((size_t)nodes * n), f.r.resize((size_t)nodes * m), f.c.resize(nodes);
    f.C.resize((size_t)nodes * nc * n + 1), f.D.resize((size_t)nodes * nc * m + 1), f.e.resize((size_t)nodes * nc + 1);
    f.Qf.resize((size_t)n * n), f.qf.resize(n), f.cf.resize(1), f.x0.resize(n);
    orc_generate_problem(1234, b, cs.algorithm, n, m, nc, N, dt, f.A.data(), f.B.data(), f.Hv.data(), f.Q.data(), f.P.data(), f.R.data(),
                         f.q.data(), f.r.data(), f.c.data(), f.C.data(), f.D.data(), f.e.data(), f.Qf.data(), f.qf.data(), f.cf.data(), f.x0.data());
    f.time.resize(N + 1);
    for (int k = 0; k <= N; ++k) f.time[k] = dt * k;
    if (slqEvents)  // the post-event node is stamped weakEpsilon after the pre-event node (RolloutBase.cpp:62-64)
      for (int k = 1; k <= N; ++k) f.time[k] = f.time[k - 1] + ((k - 1 == slqEventNodes[0] || k - 1 == slqEventNodes[1]) ? 1e-9 : dt);
    f.xnom.resize((size_t)(N + 1) * n), f.unom.resize((size_t)(N + 1) * m);
    for (auto& x : f.xnom) x = 0.3 * uni(rng);
    for (auto& u : f.unom) u = 0.3 * uni(rng);
    f.ncActive.assign(nodes, nc);
    if (cs.ragged)
      for (int k = 0; k < nodes; ++k) f.ncActive[k] = nc - (k + b) % 2;
    f.event.assign(nodes, 0);

    // the instance's AoS data, as GaussNewtonDDP holds it
    std::vector<ModelData> traj(N + 1);
    for (int k = 0; k <= N; ++k) {
      ModelData& md = traj[k];
      md.stateDim = n, md.inputDim = m, md.time = f.time[k];
      const int kk = std::min(k, nodes - 1);  // ILQR: node N carries no stage data of its own
      md.dynamics.dfdx.set(&f.A[(size_t)kk * n * n], n, n), md.dynamics.dfdu.set(&f.B[(size_t)kk * n * m], n, m);
      md.dynamicsBias.set(&f.Hv[(size_t)kk * n], n, 1);
      md.cost.dfdxx.set(&f.Q[(size_t)kk * n * n], n, n), md.cost.dfdux.set(&f.P[(size_t)kk * m * n], m, n), md.cost.dfduu.set(&f.R[(size_t)kk * m * m], m, m);
      md.cost.dfdx.set(&f.q[(size_t)kk * n], n, 1), md.cost.dfdu.set(&f.r[(size_t)kk * m], m, 1), md.cost.f = f.c[kk];
      const int nck = nc > 0 ? f.ncActive[kk] : 0;  // nck x n with leading dimension nck, as Eigen stores it
      md.stateInputEqConstraint.f.resize(nck), md.stateInputEqConstraint.dfdx.resize(nck, n), md.stateInputEqConstraint.dfdu.resize(nck, m);
      for (int i = 0; i < nck; ++i) {
        md.stateInputEqConstraint.f.v[i] = f.e[(size_t)kk * nc + i];
        for (int j = 0; j < n; ++j) md.stateInputEqConstraint.dfdx.v[i + nck * j] = f.C[(size_t)kk * nc * n + i + nc * j];
        for (int j = 0; j < m; ++j) md.stateInputEqConstraint.dfdu.v[i + nck * j] = f.D[(size_t)kk * nc * m + i + nc * j];
      }
    }
    ScalarFunctionQuadraticApproximation fin;
    fin.dfdxx.set(f.Qf.data(), n, n), fin.dfdx.set(f.qf.data(), n, 1), fin.f = f.cf[0];
    solver.setModelData(b, traj, fin);
    if (cs.events) {
      for (int j = 0; j < (slqEvents ? 2 : b % 3); ++j) {  // ILQR: 0, 1 or 2 events per instance at random nodes
        const int k = slqEvents ? slqEventNodes[j] : (int)(lcg(rng) >> 33) % nodes;
        ModelData jump;
        jump.stateDim = n, jump.inputDim = m;
        jump.dynamics.dfdx.resize(n, n), jump.dynamicsBias.resize(n), jump.cost.dfdxx.resize(n, n), jump.cost.dfdx.resize(n);
        for (int a = 0; a < n; ++a) {
          for (int c2 = 0; c2 < n; ++c2) jump.dynamics.dfdx.v[a + n * c2] = (a == c2) + 0.3 * uni(rng);
          jump.dynamicsBias.v[a] = 0.1 * uni(rng), jump.cost.dfdx.v[a] = 0.2 * uni(rng);
          jump.cost.dfdxx.v[a + n * a] = 1.0 + 0.5 * uni(rng);
        }
        jump.cost.f = 0.4 * uni(rng);
        solver.setEvent(b, k, jump);
        f.event[k] = 1;
        if (slqEvents) {  // SLQ keeps the jump data per event
          f.jA.insert(f.jA.end(), jump.dynamics.dfdx.v.begin(), jump.dynamics.dfdx.v.end());
          f.jHv.insert(f.jHv.end(), jump.dynamicsBias.v.begin(), jump.dynamicsBias.v.end());
          f.jQ.insert(f.jQ.end(), jump.cost.dfdxx.v.begin(), jump.cost.dfdxx.v.end());
          f.jq.insert(f.jq.end(), jump.cost.dfdx.v.begin(), jump.cost.dfdx.v.end());
          f.jc.push_back(jump.cost.f);
          continue;
        }
        // ILQR: the oracle reads the jump data from the node's own A, Hv, Q, q, c
        std::copy(jump.dynamics.dfdx.v.begin(), jump.dynamics.dfdx.v.end(), &f.A[(size_t)k * n * n]);
        std::copy(jump.dynamicsBias.v.begin(), jump.dynamicsBias.v.end(), &f.Hv[(size_t)k * n]);
        std::copy(jump.cost.dfdxx.v.begin(), jump.cost.dfdxx.v.end(), &f.Q[(size_t)k * n * n]);
        std::copy(jump.cost.dfdx.v.begin(), jump.cost.dfdx.v.end(), &f.q[(size_t)k * n]);
        f.c[k] = jump.cost.f;
      }
    }
    if (cs.nominal) {
      std::vector<Dense> xs(N + 1), us(N + 1);
      for (int k = 0; k <= N; ++k) xs[k].set(&f.xnom[(size_t)k * n], n, 1), us[k].set(&f.unom[(size_t)k * m], m, 1);
      solver.setNominalTrajectories(b, xs, us);
    }
    Dense x0;
    x0.set(f.x0.data(), n, 1);
    solver.setInitState(b, x0);
  }
  solver.setTimeTrajectory(flats[0].time);
  solver.solveSequentialRiccatiEquations();
  if (solver.kernelVariant().find(cs.kernel) == std::string::npos) {
    std::printf("FAIL %s: kernel variant %s, expected %s\n", cs.name, solver.kernelVariant().c_str(), cs.kernel);
    return 1;
  }
  const std::vector<double> alphas = {1.0, 0.5};
  solver.rolloutTrajectory(alphas);
  const std::vector<double> rolloutTimes = solver.rolloutTimes();

  double worst = 0.0;
  for (int b = 0; b < cs.batch; ++b) {
    const Flat& f = flats[b];
    const orc_problem pb = f.view(cs.nominal, cs.ragged, cs.events);
    std::vector<double> K((size_t)(N + 1) * m * n), db((size_t)(N + 1) * m), bias((size_t)(N + 1) * m), Sm((size_t)(N + 1) * n * n), Sv((size_t)(N + 1) * n), s(N + 1);
    orc_solution ref{K.data(), db.data(), bias.data(), Sm.data(), Sv.data(), s.data(), 0};
    orc_backward(&ost, &pb, &ref);
    if ((solver.status(b) & 1) != (ref.status & 1)) {
      std::printf("FAIL %s: status of instance %d: %d vs oracle %d\n", cs.name, b, solver.status(b), ref.status);
      return 1;
    }
    std::vector<ScalarFunctionQuadraticApproximation> vf;
    solver.getValueFunctionTrajectory(b, vf);
    LinearController ctrl;
    solver.calculateController(b, ctrl);
    if ((int)vf.size() != N + 1 || (int)ctrl.gainArray_.size() != N + 1 || ctrl.timeStamp_ != f.time) {
      std::printf("FAIL %s: trajectory sizes / time stamps\n", cs.name);
      return 1;
    }
    for (int k = 0; k <= N; ++k) {
      worst = std::max(worst, relErr(vf[k].dfdxx.data(), &Sm[(size_t)k * n * n], (size_t)n * n));
      worst = std::max(worst, relErr(vf[k].dfdx.data(), &Sv[(size_t)k * n], n));
      worst = std::max(worst, relErr(&vf[k].f, &s[k], 1));
      worst = std::max(worst, relErr(ctrl.gainArray_[k].data(), &K[(size_t)k * m * n], (size_t)m * n));
      worst = std::max(worst, relErr(ctrl.deltaBiasArray_[k].data(), &db[(size_t)k * m], m));
      worst = std::max(worst, relErr(ctrl.biasArray_[k].data(), &bias[(size_t)k * m], m));
    }
    // the float wire format of the incremented controller: row i = [uff_i, K_i,:] (LinearController.cpp:107-140)
    std::vector<std::vector<float>> flat;
    solver.flatten(b, 0.5, flat);
    if ((int)flat.size() != N + 1) {
      std::printf("FAIL %s: flattened controller has %zu nodes\n", cs.name, flat.size());
      return 1;
    }
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < m; ++i) {
        const float uff = static_cast<float>(ctrl.biasArray_[k].v[i] + 0.5 * ctrl.deltaBiasArray_[k].v[i]);
        bool same = flat[k].size() == (size_t)m * (n + 1) && flat[k][(size_t)i * (n + 1)] == uff;
        for (int j = 0; same && j < n; ++j) same = flat[k][(size_t)i * (n + 1) + j + 1] == static_cast<float>(ctrl.gainArray_[k].v[i + (size_t)m * j]);
        if (!same) {
          std::printf("FAIL %s: flattened controller of instance %d, node %d, row %d\n", cs.name, b, k, i);
          return 1;
        }
      }
    for (size_t a = 0; a < alphas.size(); ++a) {
      const int cap = (int)rolloutTimes.size();
      std::vector<double> x((size_t)cap * n), u((size_t)cap * m), t(cap);
      int count = 0;
      orc_rollout(&ost, &pb, &ref, f.x0.data(), alphas[a], x.data(), u.data(), t.data(), cap, &count);
      std::vector<Dense> xs, us;
      solver.getRollout(b, (int)a, xs, us);
      if ((int)xs.size() != count) {
        std::printf("FAIL %s: rollout node count %zu vs oracle %d\n", cs.name, xs.size(), count);
        return 1;
      }
      for (int k = 0; k < count; ++k) {
        worst = std::max(worst, relErr(xs[k].data(), &x[(size_t)k * n], n));
        worst = std::max(worst, relErr(us[k].data(), &u[(size_t)k * m], m));
        worst = std::max(worst, relErr(&rolloutTimes[k], &t[k], 1));
      }
    }
  }
  if (cs.algorithm == ORC_ALG_ILQR) {
    // LineSearchStrategy on the LQ model: candidates 1, 1/2, ... >= 0.05; merit = LQ-model cost of the oracle's rollout; Armijo
    // against the baseline (default: the cost of the zero-deviation trajectory) with the trapezoidal IS of deltaBias
    const o2c_line_search_settings ls{0.05, 1.0, 0.5, 1e-4};
    const auto results = solver.lineSearch(ls);
    for (int b = 0; b < cs.batch; ++b) {
      const Flat& f = flats[b];
      const orc_problem pb = f.view(cs.nominal, cs.ragged, cs.events);
      std::vector<double> K((size_t)(N + 1) * m * n), db((size_t)(N + 1) * m), bias((size_t)(N + 1) * m), Sm((size_t)(N + 1) * n * n), Sv((size_t)(N + 1) * n), s(N + 1);
      orc_solution ref{K.data(), db.data(), bias.data(), Sm.data(), Sv.data(), s.data(), 0};
      orc_backward(&ost, &pb, &ref);
      double is = 0.0, base = f.cf[0];
      for (int k = 0; k < N; ++k) base += f.c[k];
      for (int k = 1; k <= N; ++k) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < m; ++i) s0 += db[(size_t)(k - 1) * m + i] * db[(size_t)(k - 1) * m + i], s1 += db[(size_t)k * m + i] * db[(size_t)k * m + i];
        is += (s0 + s1) * (0.5 * (f.time[k] - f.time[k - 1]));
      }
      const LineSearchResultT& got = results[b];
      if (got.merits.size() != 5) {
        std::printf("FAIL %s: %zu line-search candidates, expected 5\n", cs.name, got.merits.size());
        return 1;
      }
      int want = -1;
      double alpha = 1.0;
      for (int e = 0; e < 5; ++e, alpha *= 0.5) {
        std::vector<double> x((size_t)(N + 1) * n), u((size_t)(N + 1) * m);
        int count = 0;
        orc_rollout(&ost, &pb, &ref, f.x0.data(), alpha, x.data(), u.data(), nullptr, N + 1, &count);
        const double merit = orc_discrete_lq_cost(&pb, x.data(), u.data());
        worst = std::max(worst, relErr(&got.merits[e], &merit, 1));
        if (want < 0 && merit < base - 1e-4 * alpha * is) want = e;
        if (got.candidateIndex == e) {  // the winning rollout stays available
          std::vector<Dense> xs, us;
          solver.getRollout(b, e, xs, us);
          worst = std::max(worst, relErr(xs[N].data(), &x[(size_t)N * n], n));
        }
      }
      worst = std::max(worst, relErr(&got.baselineMerit, &base, 1));
      worst = std::max(worst, relErr(&got.controllerUpdateIS, &is, 1));
      if (got.candidateIndex != want) {
        std::printf("FAIL %s: line search picked candidate %d, oracle %d\n", cs.name, got.candidateIndex, want);
        return 1;
      }
    }
  }
  std::printf("%s %-28s kernel %-22s max rel err %.3e\n", worst <= tol ? "ok  " : "FAIL", cs.name, solver.kernelVariant().c_str(), worst);
  return worst <= tol ? 0 : 1;
}

// HpipmInterface::solve-style hand-over (HpipmInterface.h:85-87): x0, N dynamics {f, dfdx, dfdu}, N+1 costs. The QP solution must satisfy
// the optimality conditions of the equality-constrained QP: with the costates lambda_k = dV_k/dx (x_k) = Sm_k x_k + Sv_k,
//   r_k + P_k x_k + R_k u_k + B_k' lambda_{k+1} = 0   (stationarity in u_k),   x_{k+1} = A_k x_k + B_k u_k + b_k   (feasibility)
int qpCase() {
  const int n = 10, m = 3, N = 15, batch = 4;
  o2c_config cfg{};
  cfg.nx = n, cfg.nu = m, cfg.num_stages = N, cfg.batch = batch, cfg.algorithm = O2C_ALG_ILQR, cfg.riccati_form = O2C_FORM_REDUCED;
  cfg.max_alphas = 1, cfg.hessian_multiple = 0.0, cfg.time_step = 0.01;
  ocs2_ddp_cuda::BatchedRiccatiSolver solver(cfg);
  std::vector<std::vector<VectorFunctionLinearApproximation>> dyn(batch);
  std::vector<std::vector<ScalarFunctionQuadraticApproximation>> cost(batch);
  for (int b = 0; b < batch; ++b) {
    std::vector<double> A((size_t)N * n * n), B((size_t)N * n * m), Hv((size_t)N * n), Q((size_t)N * n * n), P((size_t)N * m * n), R((size_t)N * m * m),
        q((size_t)N * n), r((size_t)N * m), c(N), C(1), D(1), e(1), Qf((size_t)n * n), qf(n), cf(1), x0(n);
    orc_generate_problem(99, b, ORC_ALG_ILQR, n, m, 0, N, 0.01, A.data(), B.data(), Hv.data(), Q.data(), P.data(), R.data(), q.data(), r.data(),
                         c.data(), C.data(), D.data(), e.data(), Qf.data(), qf.data(), cf.data(), x0.data());
    dyn[b].resize(N), cost[b].resize(N + 1);
    for (int k = 0; k < N; ++k) {
      dyn[b][k].dfdx.set(&A[(size_t)k * n * n], n, n), dyn[b][k].dfdu.set(&B[(size_t)k * n * m], n, m), dyn[b][k].f.set(&Hv[(size_t)k * n], n, 1);
      cost[b][k].dfdxx.set(&Q[(size_t)k * n * n], n, n), cost[b][k].dfdux.set(&P[(size_t)k * m * n], m, n), cost[b][k].dfduu.set(&R[(size_t)k * m * m], m, m);
      cost[b][k].dfdx.set(&q[(size_t)k * n], n, 1), cost[b][k].dfdu.set(&r[(size_t)k * m], m, 1), cost[b][k].f = c[k];
    }
    cost[b][N].dfdxx.set(Qf.data(), n, n), cost[b][N].dfdx.set(qf.data(), n, 1), cost[b][N].f = cf[0];
    Dense x0v;
    x0v.set(x0.data(), n, 1);
    solver.setQp(b, x0v, dyn[b], cost[b]);
  }
  solver.solveQps();
  double worst = 0.0;
  for (int b = 0; b < batch; ++b) {
    std::vector<Dense> xs, us;
    solver.getQpSolution(b, xs, us);
    std::vector<ScalarFunctionQuadraticApproximation> vf;
    solver.getValueFunctionTrajectory(b, vf);
    if ((int)xs.size() != N + 1 || (int)us.size() != N) {
      std::printf("FAIL qp: trajectory sizes %zu, %zu\n", xs.size(), us.size());
      return 1;
    }
    for (int k = 0; k < N; ++k) {
      const auto& d = dyn[b][k];
      const auto& cs = cost[b][k];
      for (int i = 0; i < n; ++i) {  // feasibility
        double v = d.f.v[i];
        for (int j = 0; j < n; ++j) v += d.dfdx.v[i + (size_t)n * j] * xs[k].v[j];
        for (int j = 0; j < m; ++j) v += d.dfdu.v[i + (size_t)n * j] * us[k].v[j];
        worst = std::max(worst, std::fabs(v - xs[k + 1].v[i]));
      }
      std::vector<double> lam(n);  // costate at k+1
      for (int i = 0; i < n; ++i) {
        double v = vf[k + 1].dfdx.v[i];
        for (int j = 0; j < n; ++j) v += vf[k + 1].dfdxx.v[i + (size_t)n * j] * xs[k + 1].v[j];
        lam[i] = v;
      }
      for (int l = 0; l < m; ++l) {  // stationarity
        double v = cs.dfdu.v[l];
        for (int j = 0; j < n; ++j) v += cs.dfdux.v[l + (size_t)m * j] * xs[k].v[j];
        for (int j = 0; j < m; ++j) v += cs.dfduu.v[l + (size_t)m * j] * us[k].v[j];
        for (int i = 0; i < n; ++i) v += d.dfdu.v[i + (size_t)n * l] * lam[i];
        worst = std::max(worst, std::fabs(v));
      }
    }
  }
  std::printf("%s %-28s kernel %-22s max KKT residual %.3e\n", worst <= 1e-9 ? "ok  " : "FAIL", "hpipm-style qp hand-over", solver.kernelVariant().c_str(), worst);
  return worst <= 1e-9 ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "--gpu";
  if (mode == "--no-gpu") {
    o2c_config cfg{};
    cfg.nx = 4, cfg.nu = 1, cfg.num_stages = 10, cfg.batch = 2, cfg.max_alphas = 1, cfg.riccati_form = O2C_FORM_REDUCED, cfg.time_step = 0.01;
    try {
      ocs2_ddp_cuda::BatchedRiccatiSolver solver(cfg);
    } catch (const std::runtime_error& err) {
      std::printf("ok   constructor throws without a CUDA device: %s\n", err.what());
      return 0;
    }
    std::printf("FAIL the constructor succeeded without a CUDA device\n");
    return 1;
  }
  const Case cases[] = {
      {"legged ilqr", ORC_ALG_ILQR, 24, 24, 0, 20, 5, false, false, false, "ilqr_wpp"},
      {"legged ilqr nominal", ORC_ALG_ILQR, 24, 24, 0, 20, 5, true, false, false, "ilqr_wpp"},
      {"ballbot ilqr", ORC_ALG_ILQR, 10, 3, 0, 30, 7, true, false, false, "ilqr_rpl"},
      {"manipulator ilqr nc=3", ORC_ALG_ILQR, 9, 9, 3, 25, 7, false, false, false, "ilqr_rpl"},
      {"manipulator ragged nc", ORC_ALG_ILQR, 9, 9, 3, 25, 4, true, true, false, "ilqr_rpl"},
      {"quadrotor slq", ORC_ALG_SLQ, 12, 4, 0, 20, 5, true, false, false, "slq_rpl"},
      {"generic slq nc=2", ORC_ALG_SLQ, 6, 4, 2, 16, 3, false, false, false, "generic"},
      {"ilqr events", ORC_ALG_ILQR, 6, 4, 0, 14, 6, true, false, true, "generic"},
      {"legged ilqr events", ORC_ALG_ILQR, 24, 24, 0, 12, 4, false, false, true, "ilqr_wpp"},
      {"quadrotor slq events", ORC_ALG_SLQ, 12, 4, 0, 14, 4, true, false, true, "slq_rpl"},
  };
  int failures = 0;
  try {
    for (const Case& cs : cases) failures += runCase(cs);
    failures += qpCase();
  } catch (const std::exception& err) {
    std::printf("FAIL exception: %s\n", err.what());
    return 2;
  }
  return failures ? 1 : 0;
}
