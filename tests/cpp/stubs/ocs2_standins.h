// Stand-ins of the reference types the C++ front-ends touch, for an image without Eigen3 / Boost (the reference itself cannot be
// compiled here). Only the members that include/ocs2_ddp_cuda/*.h read or write exist; names, nesting and access levels follow
//   ocs2_core/include/ocs2_core/Types.h, model_data/ModelData.h:43-60, control/LinearController.h:109-112,
//   ocs2_oc/include/ocs2_oc/oc_data/PrimalSolution.h, ocs2_ddp/include/ocs2_ddp/{DDP_Settings.h, DDP_Data.h:52-110,
//   GaussNewtonDDP.h:149-193,346-362, ILQR.h, SLQ.h, search_strategy/SearchStrategyBase.h:124}.
// Dense blocks are column-major like Eigen's default. TEST INFRASTRUCTURE — not part of the product.
#ifndef OCS2_STANDINS_H_
#define OCS2_STANDINS_H_

#include <algorithm>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

namespace ocs2 {

using scalar_t = double;

struct Dense {  // vector_t / matrix_t
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
using vector_t = Dense;
using matrix_t = Dense;
using scalar_array_t = std::vector<scalar_t>;
using size_array_t = std::vector<size_t>;
using vector_array_t = std::vector<vector_t>;
using matrix_array_t = std::vector<matrix_t>;

struct VectorFunctionLinearApproximation { vector_t f; matrix_t dfdx, dfdu; };
struct ScalarFunctionQuadraticApproximation { scalar_t f = 0.0; vector_t dfdx, dfdu; matrix_t dfdxx, dfdux, dfduu; };

struct ModelData {
  int stateDim = 0, inputDim = 0;
  scalar_t time = 0.0;
  vector_t dynamicsBias;
  VectorFunctionLinearApproximation dynamics;
  ScalarFunctionQuadraticApproximation cost;
  VectorFunctionLinearApproximation stateInputEqConstraint;
};

struct LinearController {
  scalar_array_t timeStamp_;
  vector_array_t biasArray_, deltaBiasArray_;
  matrix_array_t gainArray_;
  size_t size() const { return timeStamp_.size(); }
  void clear() { timeStamp_.clear(), biasArray_.clear(), deltaBiasArray_.clear(), gainArray_.clear(); }
};

struct PrimalSolution {
  scalar_array_t timeTrajectory_;
  vector_array_t stateTrajectory_, inputTrajectory_;
  size_array_t postEventIndices_;
};
struct PrimalDataContainer {
  PrimalSolution primalSolution;
  ModelData modelDataFinalTime;
  std::vector<ModelData> modelDataEventTimes, modelDataTrajectory;
};
struct DualDataContainer {
  std::vector<ModelData> projectedModelDataTrajectory;
  std::vector<ScalarFunctionQuadraticApproximation> valueFunctionTrajectory;
};

namespace hessian_correction {
enum class Strategy { DIAGONAL_SHIFT, CHOLESKY_MODIFICATION, EIGENVALUE_MODIFICATION, GERSHGORIN_MODIFICATION };
}
namespace search_strategy {
enum class Type { LINE_SEARCH, LEVENBERG_MARQUARDT };
}
namespace line_search {
struct Settings {
  scalar_t minStepLength = 0.05, maxStepLength = 1.0, contractionRate = 0.5, armijoCoefficient = 1e-4;
  hessian_correction::Strategy hessianCorrectionStrategy = hessian_correction::Strategy::DIAGONAL_SHIFT;
  scalar_t hessianCorrectionMultiple = 1e-12;
};
}  // namespace line_search
namespace ddp {
enum class Algorithm { SLQ, ILQR };
struct Settings {
  Algorithm algorithm_ = Algorithm::SLQ;
  size_t nThreads_ = 1;
  bool checkNumericalStability_ = true;
  scalar_t timeStep_ = 1e-2;
  bool preComputeRiccatiTerms_ = true;
  search_strategy::Type strategy_ = search_strategy::Type::LINE_SEARCH;
  line_search::Settings lineSearch_;
};
}  // namespace ddp

class SearchStrategyBase {
 public:
  virtual ~SearchStrategyBase() = default;
  virtual matrix_t augmentHamiltonianHessian(const ModelData& modelData, const matrix_t& Hm) const = 0;
};
class LineSearchStrategy final : public SearchStrategyBase {
 public:
  matrix_t augmentHamiltonianHessian(const ModelData&, const matrix_t& Hm) const override { return Hm; }
};
class LevenbergMarquardtStrategy final : public SearchStrategyBase {
 public:
  explicit LevenbergMarquardtStrategy(scalar_t riccatiMultiple) { lmModule_.riccatiMultiple = riccatiMultiple; }
  matrix_t augmentHamiltonianHessian(const ModelData& modelData, const matrix_t& Hm) const override {  // Hm + riccatiMultiple * B'B
    matrix_t aug = Hm;
    const matrix_t& B = modelData.dynamics.dfdu;
    for (long i = 0; i < B.cols(); ++i)
      for (long j = 0; j < B.cols(); ++j)
        for (long k = 0; k < B.rows(); ++k) aug.v[i + aug.r * j] += lmModule_.riccatiMultiple * B.v[k + B.r * i] * B.v[k + B.r * j];
    return aug;
  }

 private:
  struct { scalar_t riccatiMultiple = 0.0; } lmModule_;
};

// The seam of GaussNewtonDDP: the two virtuals of the backward pass, the protected data they read and write, and the driver loop
// of calculateController (GaussNewtonDDP.cpp:588-642) around the per-node worker.
class GaussNewtonDDP {
 public:
  explicit GaussNewtonDDP(ddp::Settings ddpSettings) : ddpSettings_(std::move(ddpSettings)) {
    if (ddpSettings_.strategy_ == search_strategy::Type::LINE_SEARCH) searchStrategyPtr_.reset(new LineSearchStrategy);
  }
  virtual ~GaussNewtonDDP() = default;
  const ddp::Settings& settings() const { return ddpSettings_; }

 protected:
  virtual void calculateControllerWorker(size_t timeIndex, const PrimalDataContainer& primalData, const DualDataContainer& dualData,
                                         LinearController& dstController) = 0;
  virtual scalar_t solveSequentialRiccatiEquations(const ScalarFunctionQuadraticApproximation& finalValueFunction) = 0;

  void calculateController() {
    const size_t N = nominalPrimalData_.primalSolution.timeTrajectory_.size();
    unoptimizedController_.clear();
    unoptimizedController_.timeStamp_ = nominalPrimalData_.primalSolution.timeTrajectory_;
    unoptimizedController_.gainArray_.resize(N), unoptimizedController_.biasArray_.resize(N), unoptimizedController_.deltaBiasArray_.resize(N);
    for (size_t k = 0; k < N; ++k) calculateControllerWorker(k, nominalPrimalData_, nominalDualData_, unoptimizedController_);
    const auto& post = nominalPrimalData_.primalSolution.postEventIndices_;
    if ((post.empty() || post.back() != N - 1) && N >= 2) {
      unoptimizedController_.gainArray_.back() = unoptimizedController_.gainArray_[N - 2];
      unoptimizedController_.biasArray_.back() = unoptimizedController_.biasArray_[N - 2];
      unoptimizedController_.deltaBiasArray_.back() = unoptimizedController_.deltaBiasArray_[N - 2];
    }
  }

  DualDataContainer nominalDualData_;
  PrimalDataContainer nominalPrimalData_;
  LinearController unoptimizedController_;
  scalar_t initTime_ = 0.0, finalTime_ = 0.0;
  vector_t initState_;
  std::unique_ptr<SearchStrategyBase> searchStrategyPtr_;

 private:
  const ddp::Settings ddpSettings_;
};

// ILQR / SLQ: the CPU workers are the reference's own code and are not restated here
class ILQR : public GaussNewtonDDP {
 public:
  using GaussNewtonDDP::GaussNewtonDDP;

 protected:
  scalar_t solveSequentialRiccatiEquations(const ScalarFunctionQuadraticApproximation&) override { throw std::logic_error("stand-in: CPU ILQR backward pass"); }
  void calculateControllerWorker(size_t, const PrimalDataContainer&, const DualDataContainer&, LinearController&) override { throw std::logic_error("stand-in: CPU ILQR controller"); }
};
class SLQ : public GaussNewtonDDP {
 public:
  using GaussNewtonDDP::GaussNewtonDDP;

 protected:
  scalar_t solveSequentialRiccatiEquations(const ScalarFunctionQuadraticApproximation&) override { throw std::logic_error("stand-in: CPU SLQ backward pass"); }
  void calculateControllerWorker(size_t, const PrimalDataContainer&, const DualDataContainer&, LinearController&) override { throw std::logic_error("stand-in: CPU SLQ controller"); }
};

}  // namespace ocs2

#endif  // OCS2_STANDINS_H_
