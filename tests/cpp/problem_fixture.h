// One seeded LQ problem of the oracle's synthetic family, held twice: flat (what the oracle reads) and as the arrays-of-structs a
// GaussNewtonDDP instance keeps (std::vector<ModelData>, stand-in types of tests/cpp/stubs). TEST INFRASTRUCTURE.
#ifndef TESTS_CPP_PROBLEM_FIXTURE_H_
#define TESTS_CPP_PROBLEM_FIXTURE_H_

#include <cmath>
#include <cstdint>
#include <vector>

#include "lq_oracle.h"
#include "ocs2_standins.h"

namespace fixture {

inline uint64_t lcg(uint64_t& s) { return s = s * 6364136223846793005ull + 1442695040888963407ull; }
inline double uni(uint64_t& s) { return (double)(lcg(s) >> 11) / 9007199254740992.0 * 2.0 - 1.0; }

// max |got - want| / max |want| over a block (true relative error, no floor)
inline void accumulate(const double* got, const double* want, size_t count, double& diff, double& scale) {
  for (size_t i = 0; i < count; ++i) {
    if (!std::isfinite(got[i]) || !std::isfinite(want[i])) diff = INFINITY;
    diff = std::fmax(diff, std::fabs(got[i] - want[i]));
    scale = std::fmax(scale, std::fabs(want[i]));
  }
}

struct Problem {
  int algorithm = 0, n = 0, m = 0, nc = 0, N = 0, nodes = 0;
  std::vector<double> A, B, Hv, Q, P, R, q, r, c, C, D, e, Qf, qf, cf, x0, xnom, unom, time, jA, jHv, jQ, jq, jc;
  std::vector<int32_t> event;
  std::vector<ocs2::ModelData> modelDataTrajectory, modelDataEventTimes;
  std::vector<size_t> postEventIndices;
  ocs2::ScalarFunctionQuadraticApproximation finalValueFunction;
  ocs2::vector_array_t stateTrajectory, inputTrajectory;
  ocs2::vector_t initState;

  void generate(uint64_t seed, int64_t index, int algorithm_, int n_, int m_, int nc_, int N_, double dt) {
    algorithm = algorithm_, n = n_, m = m_, nc = nc_, N = N_, nodes = algorithm == ORC_ALG_ILQR ? N : N + 1;
    A.resize((size_t)nodes * n * n), B.resize((size_t)nodes * n * m), Hv.resize((size_t)nodes * n);
    Q.resize((size_t)nodes * n * n), P.resize((size_t)nodes * m * n), R.resize((size_t)nodes * m * m);
    q.resize((size_t)nodes * n), r.resize((size_t)nodes * m), c.resize(nodes);
    C.resize((size_t)nodes * nc * n + 1), D.resize((size_t)nodes * nc * m + 1), e.resize((size_t)nodes * nc + 1);
    Qf.resize((size_t)n * n), qf.resize(n), cf.resize(1), x0.resize(n);
    orc_generate_problem(seed, index, algorithm, n, m, nc, N, dt, A.data(), B.data(), Hv.data(), Q.data(), P.data(), R.data(), q.data(), r.data(),
                         c.data(), C.data(), D.data(), e.data(), Qf.data(), qf.data(), cf.data(), x0.data());
    time.resize(N + 1);
    for (int k = 0; k <= N; ++k) time[k] = dt * k;
    uint64_t rng = 0x9e3779b97f4a7c15ull + seed * 977 + (uint64_t)index * 131 + n;
    xnom.resize((size_t)(N + 1) * n), unom.resize((size_t)(N + 1) * m);
    for (auto& x : xnom) x = 0.3 * uni(rng);
    for (auto& u : unom) u = 0.3 * uni(rng);
    event.assign(nodes, 0);
    rebuild();
  }

  // ILQR pre-event node k: the jump ModelData replace the node's A, Hv, Q, q, c in the oracle's flat layout (lq_oracle.h)
  void addIlqrEvent(int k, uint64_t& rng) {
    ocs2::ModelData jump;
    jump.stateDim = n, jump.inputDim = m;
    jump.dynamics.dfdx.resize(n, n), jump.dynamicsBias.resize(n), jump.cost.dfdxx.resize(n, n), jump.cost.dfdx.resize(n);
    for (int a = 0; a < n; ++a) {
      for (int b = 0; b < n; ++b) jump.dynamics.dfdx.v[a + n * b] = (a == b) + 0.3 * uni(rng);
      jump.dynamicsBias.v[a] = 0.1 * uni(rng), jump.cost.dfdx.v[a] = 0.2 * uni(rng), jump.cost.dfdxx.v[a + n * a] = 1.0 + 0.5 * uni(rng);
    }
    jump.cost.f = 0.4 * uni(rng);
    modelDataEventTimes.push_back(jump);
    postEventIndices.push_back(k + 1);
    event[k] = 1;
    std::copy(jump.dynamics.dfdx.v.begin(), jump.dynamics.dfdx.v.end(), &A[(size_t)k * n * n]);
    std::copy(jump.dynamicsBias.v.begin(), jump.dynamicsBias.v.end(), &Hv[(size_t)k * n]);
    std::copy(jump.cost.dfdxx.v.begin(), jump.cost.dfdxx.v.end(), &Q[(size_t)k * n * n]);
    std::copy(jump.cost.dfdx.v.begin(), jump.cost.dfdx.v.end(), &q[(size_t)k * n]);
    c[k] = jump.cost.f;
  }

  // the instance's AoS data as GaussNewtonDDP holds them (the ModelData of an ILQR pre-event node keep the regular stage data: the
  // front-ends overwrite the jump blocks from modelDataEventTimes)
  void rebuild() {
    modelDataTrajectory.assign(N + 1, ocs2::ModelData());
    for (int k = 0; k <= N; ++k) {
      ocs2::ModelData& md = modelDataTrajectory[k];
      md.stateDim = n, md.inputDim = m, md.time = time[k];
      const int kk = k < nodes ? k : nodes - 1;  // ILQR: node N carries no stage data of its own
      md.dynamics.dfdx.set(&A[(size_t)kk * n * n], n, n), md.dynamics.dfdu.set(&B[(size_t)kk * n * m], n, m), md.dynamicsBias.set(&Hv[(size_t)kk * n], n, 1);
      md.cost.dfdxx.set(&Q[(size_t)kk * n * n], n, n), md.cost.dfdux.set(&P[(size_t)kk * m * n], m, n), md.cost.dfduu.set(&R[(size_t)kk * m * m], m, m);
      md.cost.dfdx.set(&q[(size_t)kk * n], n, 1), md.cost.dfdu.set(&r[(size_t)kk * m], m, 1), md.cost.f = c[kk];
      md.stateInputEqConstraint.f.resize(nc), md.stateInputEqConstraint.dfdx.resize(nc, n), md.stateInputEqConstraint.dfdu.resize(nc, m);
      for (int i = 0; i < nc; ++i) {
        md.stateInputEqConstraint.f.v[i] = e[(size_t)kk * nc + i];
        for (int j = 0; j < n; ++j) md.stateInputEqConstraint.dfdx.v[i + nc * j] = C[(size_t)kk * nc * n + i + nc * j];
        for (int j = 0; j < m; ++j) md.stateInputEqConstraint.dfdu.v[i + nc * j] = D[(size_t)kk * nc * m + i + nc * j];
      }
    }
    finalValueFunction.dfdxx.set(Qf.data(), n, n), finalValueFunction.dfdx.set(qf.data(), n, 1), finalValueFunction.f = cf[0];
    stateTrajectory.resize(N + 1), inputTrajectory.resize(N + 1);
    for (int k = 0; k <= N; ++k) stateTrajectory[k].set(&xnom[(size_t)k * n], n, 1), inputTrajectory[k].set(&unom[(size_t)k * m], m, 1);
    initState.set(x0.data(), n, 1);
  }

  orc_problem view(bool nominal) const {
    orc_problem p{};
    p.nx = n, p.nu = m, p.nc_max = nc, p.N = N;
    p.A = A.data(), p.B = B.data(), p.Hv = Hv.data(), p.Q = Q.data(), p.P = P.data(), p.R = R.data(), p.q = q.data(), p.r = r.data(), p.c = c.data();
    if (nc > 0) p.C = C.data(), p.D = D.data(), p.e = e.data();
    p.Qf = Qf.data(), p.qf = qf.data(), p.cf = cf.data();
    if (nominal) p.x_nom = xnom.data(), p.u_nom = unom.data();
    p.time = time.data();
    if (!postEventIndices.empty()) p.event = event.data();
    return p;
  }
};

struct OracleSolution {
  std::vector<double> K, dbias, bias, Sm, Sv, s;
  int status = 0;
  void solve(const orc_settings& st, const Problem& pb, bool nominal) {
    const size_t N1 = pb.N + 1, n = pb.n, m = pb.m;
    K.resize(N1 * m * n), dbias.resize(N1 * m), bias.resize(N1 * m), Sm.resize(N1 * n * n), Sv.resize(N1 * n), s.resize(N1);
    orc_solution ref{K.data(), dbias.data(), bias.data(), Sm.data(), Sv.data(), s.data(), 0};
    const orc_problem view = pb.view(nominal);
    orc_backward(&st, &view, &ref);
    status = ref.status;
  }
};

}  // namespace fixture

#endif  // TESTS_CPP_PROBLEM_FIXTURE_H_
