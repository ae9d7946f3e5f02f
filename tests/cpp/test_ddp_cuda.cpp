// ocs2::ILQR_CUDA / ocs2::SLQ_CUDA (include/ocs2_ddp_cuda/GaussNewtonDDP_CUDA.h) compiled against stand-ins of the reference's headers
// (tests/cpp/stubs: the seam of GaussNewtonDDP.h:149-193 with its protected data) and driven the way GaussNewtonDDP::runImpl drives the
// backward pass (GaussNewtonDDP.cpp:1043-1050): solveSequentialRiccatiEquations(finalValueFunction), then calculateController().
// valueFunctionTrajectory and unoptimizedController_ are compared with the CPU oracle (1e-9 relative, per field, no floor).
//
//   test_ddp_cuda --gpu | --no-gpu
#include <cstdio>
#include <cstring>
#include <string>

#include "ocs2_ddp_cuda/GaussNewtonDDP_CUDA.h"
#include "problem_fixture.h"

namespace {

using namespace ocs2;

// what a test (or GaussNewtonDDP::runImpl) does with the protected seam
template <class Solver>
class Harness final : public Solver {
 public:
  Harness(ddp::Settings settings, scalar_t riccatiMultiple) : Solver(std::move(settings)) {
    if (this->settings().strategy_ == search_strategy::Type::LEVENBERG_MARQUARDT) this->searchStrategyPtr_.reset(new LevenbergMarquardtStrategy(riccatiMultiple));
  }
  void setNominal(const fixture::Problem& pb) {
    auto& primal = this->nominalPrimalData_;
    primal.primalSolution.timeTrajectory_ = pb.time;
    primal.primalSolution.stateTrajectory_ = pb.stateTrajectory;
    primal.primalSolution.inputTrajectory_ = pb.inputTrajectory;
    primal.primalSolution.postEventIndices_ = pb.postEventIndices;
    primal.modelDataTrajectory = pb.modelDataTrajectory;
    primal.modelDataEventTimes = pb.modelDataEventTimes;
    this->initTime_ = pb.time.front(), this->finalTime_ = pb.time.back();
  }
  scalar_t backwardPass(const ScalarFunctionQuadraticApproximation& finalValueFunction) {
    const scalar_t avgTimeStep = this->solveSequentialRiccatiEquations(finalValueFunction);
    this->calculateController();
    return avgTimeStep;
  }
  const std::vector<ScalarFunctionQuadraticApproximation>& valueFunction() const { return this->nominalDualData_.valueFunctionTrajectory; }
  const LinearController& controller() const { return this->unoptimizedController_; }
};

struct Case {
  const char* name;
  int algorithm, n, m, nc, N;
  bool lm, gershgorin, events;
  const char* kernel;
};

template <class Solver>
int runCase(const Case& cs) {
  const double dt = 0.01, mu = 0.37;
  ddp::Settings settings;
  settings.algorithm_ = cs.algorithm == O2C_ALG_ILQR ? ddp::Algorithm::ILQR : ddp::Algorithm::SLQ;
  settings.timeStep_ = dt;
  settings.strategy_ = cs.lm ? search_strategy::Type::LEVENBERG_MARQUARDT : search_strategy::Type::LINE_SEARCH;
  settings.lineSearch_.hessianCorrectionStrategy = cs.gershgorin ? hessian_correction::Strategy::GERSHGORIN_MODIFICATION : hessian_correction::Strategy::DIAGONAL_SHIFT;
  settings.lineSearch_.hessianCorrectionMultiple = 1e-5;
  orc_settings ost{};
  ost.algorithm = cs.algorithm, ost.reduced_form = cs.lm ? 0 : 1, ost.strategy = cs.lm ? ORC_STRATEGY_LM : ORC_STRATEGY_LINE_SEARCH;
  ost.hessian_correction = cs.gershgorin ? ORC_HC_GERSHGORIN_MODIFICATION : ORC_HC_DIAGONAL_SHIFT;
  ost.hessian_multiple = 1e-5, ost.lm_riccati_multiple = cs.lm ? mu : 0.0, ost.time_step = dt;

  Harness<Solver> ddp(settings, mu);
  double worst = 0.0;
  for (int iteration = 0; iteration < 3; ++iteration) {  // the handle is created once and reused by the following iterations
    fixture::Problem pb;
    pb.generate(4321 + iteration, 5, cs.algorithm, cs.n, cs.m, cs.nc, cs.N, dt);
    uint64_t rng = 7 + iteration;
    if (cs.events) pb.addIlqrEvent(2 + iteration, rng), pb.addIlqrEvent(cs.N - 3, rng);
    ddp.setNominal(pb);
    const scalar_t avg = ddp.backwardPass(pb.finalValueFunction);
    fixture::OracleSolution ref;
    ref.solve(ost, pb, true);
    const size_t n = cs.n, m = cs.m, N1 = cs.N + 1;
    if (ddp.valueFunction().size() != N1 || ddp.controller().gainArray_.size() != N1 || ddp.controller().timeStamp_ != pb.time ||
        std::fabs(avg - dt) > 1e-12) {
      std::printf("FAIL %s: trajectory sizes / time stamps / average time step\n", cs.name);
      return 1;
    }
    double d[6] = {0, 0, 0, 0, 0, 0}, s[6] = {0, 0, 0, 0, 0, 0};
    for (size_t k = 0; k < N1; ++k) {
      fixture::accumulate(ddp.valueFunction()[k].dfdxx.data(), &ref.Sm[k * n * n], n * n, d[0], s[0]);
      fixture::accumulate(ddp.valueFunction()[k].dfdx.data(), &ref.Sv[k * n], n, d[1], s[1]);
      fixture::accumulate(&ddp.valueFunction()[k].f, &ref.s[k], 1, d[2], s[2]);
      fixture::accumulate(ddp.controller().gainArray_[k].data(), &ref.K[k * m * n], m * n, d[3], s[3]);
      fixture::accumulate(ddp.controller().deltaBiasArray_[k].data(), &ref.dbias[k * m], m, d[4], s[4]);
      fixture::accumulate(ddp.controller().biasArray_[k].data(), &ref.bias[k * m], m, d[5], s[5]);
    }
    for (int f = 0; f < 6; ++f) worst = std::fmax(worst, d[f] / s[f]);
  }
  const bool kernelOk = ddp.kernelVariant().find(cs.kernel) != std::string::npos;
  std::printf("%s %-34s kernel %-20s max rel err %.3e\n", worst <= 1e-9 && kernelOk ? "ok  " : "FAIL", cs.name, ddp.kernelVariant().c_str(), worst);
  return worst <= 1e-9 && kernelOk ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "--gpu";
  if (mode == "--no-gpu") {
    fixture::Problem pb;
    pb.generate(1, 0, O2C_ALG_ILQR, 4, 1, 0, 5, 0.01);
    ddp::Settings settings;
    Harness<ILQR_CUDA> ddp(settings, 0.0);
    ddp.setNominal(pb);
    try {
      ddp.backwardPass(pb.finalValueFunction);
    } catch (const std::runtime_error& err) {
      std::printf("refused: %s\n", err.what());
      return std::strstr(err.what(), "no CPU fallback") ? 0 : 1;
    }
    std::printf("FAIL: the backward pass ran without a device\n");
    return 1;
  }
  int failed = 0;
  const Case ilqr[] = {
      {"ILQR_CUDA legged 24x24", O2C_ALG_ILQR, 24, 24, 0, 20, false, false, false, "ilqr_wpp"},
      {"ILQR_CUDA legged LM (mu by probe)", O2C_ALG_ILQR, 24, 24, 0, 20, true, false, false, "ilqr_wpp"},
      {"ILQR_CUDA legged Gershgorin", O2C_ALG_ILQR, 24, 24, 0, 20, false, true, false, "ilqr_wpp"},
      {"ILQR_CUDA legged events", O2C_ALG_ILQR, 24, 24, 0, 20, false, false, true, "ilqr_wpp"},
      {"ILQR_CUDA manipulator nc=3", O2C_ALG_ILQR, 9, 9, 3, 20, false, false, false, "ilqr_rpl"},
      {"ILQR_CUDA ballbot LM", O2C_ALG_ILQR, 10, 3, 0, 20, true, false, false, "generic"},
      {"ILQR_CUDA cartpole events", O2C_ALG_ILQR, 4, 1, 0, 20, false, false, true, "ilqr_rpl"},
  };
  const Case slq[] = {
      {"SLQ_CUDA quadrotor 12x4", O2C_ALG_SLQ, 12, 4, 0, 20, false, false, false, "slq_rpl"},
      {"SLQ_CUDA legged 24x24", O2C_ALG_SLQ, 24, 24, 0, 12, false, false, false, "slq_wpp"},
      {"SLQ_CUDA generic 6x3 nc=1 LM", O2C_ALG_SLQ, 6, 3, 1, 12, true, false, false, "generic"},
  };
  for (const Case& cs : ilqr) {
    try {
      failed += runCase<ILQR_CUDA>(cs);
    } catch (const std::exception& err) {
      std::printf("FAIL %s: %s\n", cs.name, err.what());
      ++failed;
    }
  }
  for (const Case& cs : slq) {
    try {
      failed += runCase<SLQ_CUDA>(cs);
    } catch (const std::exception& err) {
      std::printf("FAIL %s: %s\n", cs.name, err.what());
      ++failed;
    }
  }
  std::printf("%d case(s) failed\n", failed);
  return failed ? 1 : 0;
}
