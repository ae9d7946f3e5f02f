/*
 * GaussNewtonDDP_CUDA.h — `ocs2::ILQR_CUDA` and `ocs2::SLQ_CUDA`: the subclasses a maintainer of RIVeR-Lab/ocs2 adds next to ocs2_ddp to
 * run the backward pass of one solver instance on the GPU.
 *
 * The seam is the pair of virtuals the backward pass of `GaussNewtonDDP::runImpl` goes through (GaussNewtonDDP.cpp:1043-1050):
 *
 *   solveSequentialRiccatiEquations(finalValueFunction)   GaussNewtonDDP.h:176, implemented by ILQR.cpp:186-212 / SLQ.cpp:174-201
 *       reads  nominalPrimalData_.{modelDataTrajectory, modelDataEventTimes, primalSolution}          (DDP_Data.h:52-75)
 *       writes nominalDualData_.valueFunctionTrajectory                                                (DDP_Data.h:96)
 *   calculateControllerWorker(k, primal, dual, dst)        GaussNewtonDDP.h:167, called per node by calculateController (:588-642)
 *       writes dst.gainArray_[k], biasArray_[k], deltaBiasArray_[k]                                    (LinearController.h:109-112)
 *
 * Everything else of the solver (LQ approximation, rollouts, search strategy, MPC) stays the reference's own code. The override packs
 * the instance's arrays-of-structs through BatchedRiccatiSolver (batch = 1 here; the lock-step multi-instance form is
 * BatchedRiccatiSolver / ShardedRiccatiSolver used directly), runs o2c_upload + o2c_backward + o2c_download, and keeps the controller
 * of the pass so that the per-node worker is a copy.
 *
 * What the GPU pass does not fill: nominalDualData_.projectedModelDataTrajectory and riccatiModificationTrajectory (intermediates of
 * the CPU workers that only getStateInputEqualityConstraintLagrangian reads afterwards).
 *
 * The code touches the reference's dense types only through data() / size() / resize(), so it compiles against Eigen-backed ocs2_core
 * and — in this repository, where Eigen is absent — against the stand-in headers of tests/cpp/stubs (tests/cpp/test_ddp_cuda.cpp builds
 * and runs it on the GPU against the CPU oracle).
 */
#ifndef OCS2_DDP_CUDA_GAUSS_NEWTON_DDP_CUDA_H_
#define OCS2_DDP_CUDA_GAUSS_NEWTON_DDP_CUDA_H_

#include <algorithm>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <utility>

#include <ocs2_ddp/ILQR.h>
#include <ocs2_ddp/SLQ.h>

#include "BatchedRiccatiSolver.h"

namespace ocs2 {

template <class Base, int Algorithm>
class GaussNewtonDDP_CUDA : public Base {
 public:
  using Base::Base;
  ~GaussNewtonDDP_CUDA() override = default;

  /** CUDA device of the handle (default 0); takes effect when the handle is (re)created */
  void setDevice(int device) {
    device_ = device;
    lq_.reset();
  }
  /** sweep kernel that served the last backward pass (diagnostics) */
  std::string kernelVariant() const { return lq_ ? lq_->kernelVariant() : std::string(); }

 protected:
  scalar_t solveSequentialRiccatiEquations(const ScalarFunctionQuadraticApproximation& finalValueFunction) override {
    const auto& primal = this->nominalPrimalData_.primalSolution;
    const auto& modelData = this->nominalPrimalData_.modelDataTrajectory;
    const std::size_t count = primal.timeTrajectory_.size();  // N + 1 time nodes
    if (count < 2 || modelData.size() != count) throw std::runtime_error("[GaussNewtonDDP_CUDA] the nominal trajectories hold fewer than two nodes");
    const int N = static_cast<int>(count) - 1;

    int ncMax = 0;
    for (const auto& md : modelData) ncMax = std::max<int>(ncMax, static_cast<int>(md.stateInputEqConstraint.f.size()));
    const int numEvents = static_cast<int>(primal.postEventIndices_.size());
    ensureSolver(static_cast<int>(modelData.front().stateDim), static_cast<int>(modelData.front().inputDim), N, ncMax, numEvents);

    lq_->setModelData(0, modelData, finalValueFunction);
    for (int i = 0; i < numEvents; ++i) {  // node postEventIndex - 1 is the pre-event node (ILQR.cpp:263, SLQ.cpp:256-302)
      const int post = static_cast<int>(primal.postEventIndices_[i]);
      if (post >= 1 && post <= N) lq_->setEvent(0, post - 1, this->nominalPrimalData_.modelDataEventTimes[i]);
    }
    lq_->setNominalTrajectories(0, primal.stateTrajectory_, primal.inputTrajectory_);
    lq_->setTimeTrajectory(primal.timeTrajectory_);
    if (this->settings().strategy_ == search_strategy::Type::LEVENBERG_MARQUARDT) lq_->setRiccatiMultiple(riccatiMultiple());

    lq_->solveSequentialRiccatiEquations();
    const int status = lq_->status(0);
    if (status & O2C_STATUS_NONFINITE) throw std::runtime_error("[GaussNewtonDDP_CUDA] the backward pass produced non-finite values");

    lq_->getValueFunctionTrajectory(0, this->nominalDualData_.valueFunctionTrajectory);
    lq_->calculateController(0, controller_, false);  // the stability check stays with GaussNewtonDDP::calculateController

    if (this->settings().checkNumericalStability_) {  // GaussNewtonDDP.cpp:555-579: checkBeingPSD of every value function
      check(o2c_check_numerical_stability(lq_->handle(), 0, 1), "o2c_check_numerical_stability");
      int32_t bits = 0;
      o2c_solution_view sv{};
      sv.status = &bits;
      check(o2c_download(lq_->handle(), &sv, 0, 1, 0), "o2c_download");
      check(o2c_sync(lq_->handle()), "o2c_sync");
      if (bits & O2C_STATUS_NOT_PSD) throw std::runtime_error("[GaussNewtonDDP_CUDA] ValueFunction is not PSD.");
    }
    return (this->finalTime_ - this->initTime_) / static_cast<scalar_t>(N);
  }

  void calculateControllerWorker(size_t timeIndex, const PrimalDataContainer& /*primalData*/, const DualDataContainer& /*dualData*/,
                                 LinearController& dstController) override {
    dstController.gainArray_[timeIndex] = controller_.gainArray_.at(timeIndex);
    dstController.biasArray_[timeIndex] = controller_.biasArray_[timeIndex];
    dstController.deltaBiasArray_[timeIndex] = controller_.deltaBiasArray_[timeIndex];
  }

 private:
  static void check(o2c_error e, const char* what) {
    if (e != O2C_OK) throw std::runtime_error(std::string("[ocs2_ddp_cuda] ") + what + ": " + o2c_last_error());
  }

  void ensureSolver(int n, int m, int N, int ncMax, int numEvents) {
    if (lq_) {
      const o2c_config& c = lq_->config();
      if (c.nx == n && c.nu == m && c.num_stages == N && c.nc_max >= ncMax && maxEvents_ >= numEvents) return;
    }
    const auto& st = this->settings();
    o2c_config cfg{};
    cfg.nx = n, cfg.nu = m, cfg.nc_max = ncMax, cfg.num_stages = N, cfg.batch = 1, cfg.algorithm = Algorithm;
    cfg.strategy = st.strategy_ == search_strategy::Type::LINE_SEARCH ? O2C_STRATEGY_LINE_SEARCH : O2C_STRATEGY_LEVENBERG_MARQUARDT;
    // ILQR.cpp:68 / SLQ.cpp:65: the reduced form is used only when the Riccati terms are pre-computed under LINE_SEARCH
    cfg.riccati_form = (st.preComputeRiccatiTerms_ && cfg.strategy == O2C_STRATEGY_LINE_SEARCH) ? O2C_FORM_REDUCED : O2C_FORM_FULL;
    cfg.hessian_correction = static_cast<int32_t>(st.lineSearch_.hessianCorrectionStrategy);  // same order as O2C_HC_*
    cfg.hessian_multiple = st.lineSearch_.hessianCorrectionMultiple;
    cfg.time_step = st.timeStep_;
    cfg.max_alphas = 1, cfg.has_nominal = 1, cfg.device = device_;
    maxEvents_ = std::max(numEvents, 4);
    lq_.reset(new ocs2_ddp_cuda::BatchedRiccatiSolver(cfg, maxEvents_));
  }

  // LevenbergMarquardtStrategy keeps riccatiMultiple private; augmentHamiltonianHessian (LevenbergMarquardtStrategy.cpp:244-249) returns
  // Hm + riccatiMultiple * B'B, so a 1x1 probe with B = [1], Hm = [0] reads it back exactly.
  scalar_t riccatiMultiple() const {
    ModelData probe;
    probe.stateDim = 1, probe.inputDim = 1;
    probe.dynamics.dfdu.resize(1, 1);
    probe.dynamics.dfdu.data()[0] = 1.0;
    matrix_t zero;
    zero.resize(1, 1);
    zero.data()[0] = 0.0;
    const matrix_t augmented = this->searchStrategyPtr_->augmentHamiltonianHessian(probe, zero);
    return augmented.data()[0];
  }

  std::unique_ptr<ocs2_ddp_cuda::BatchedRiccatiSolver> lq_;
  LinearController controller_;
  int device_ = 0, maxEvents_ = 0;
};

using ILQR_CUDA = GaussNewtonDDP_CUDA<ILQR, O2C_ALG_ILQR>;
using SLQ_CUDA = GaussNewtonDDP_CUDA<SLQ, O2C_ALG_SLQ>;

}  // namespace ocs2

#endif  // OCS2_DDP_CUDA_GAUSS_NEWTON_DDP_CUDA_H_
