/*
 * BatchedRiccatiSolver.h — C++ host side above the C ABI (ocs2_ddp_cuda.h): the batched front-end a maintainer of
 * RIVeR-Lab/ocs2 drops next to ocs2_ddp.
 *
 * B solver instances run their LQ sub-problem in lock-step (the pattern of
 * ocs2_mpcnet/ocs2_mpcnet_core/src/rollout/MpcnetRolloutManager.cpp:72-97): each instance hands over what
 * GaussNewtonDDP keeps in `nominalPrimalData_` (ocs2_ddp/include/ocs2_ddp/DDP_Data.h:52-110) — the
 * std::vector<ModelData> of its time nodes, the final value function, the nominal trajectories — the front-end packs these
 * arrays-of-structs into the pinned struct-of-arrays batch the ABI takes, ONE call solves all of them on the GPU, and every
 * instance reads back exactly the objects the reference's backward pass fills:
 *
 *   setModelData / setEvent / setNominalTrajectories   <- nominalPrimalData_.modelDataTrajectory, modelDataEventTimes,
 *                                                         primalSolution.{state,input}Trajectory_        (DDP_Data.h:52-75)
 *   solveSequentialRiccatiEquations()                  == GaussNewtonDDP.h:176 (ILQR.cpp:186-299, SLQ.cpp:174-302), all instances
 *   getValueFunctionTrajectory                         -> nominalDualData_.valueFunctionTrajectory        (DDP_Data.h:96)
 *   calculateController                                -> unoptimizedController_ (GaussNewtonDDP.cpp:588-642): timeStamp_,
 *                                                         gainArray_, biasArray_, deltaBiasArray_, last node = copy of N-1, and
 *                                                         the same std::runtime_error on non-finite gains / feedforward
 *   rolloutTrajectory / getRollout                     == incrementController + the LQ-model rollout (DDP_HelperFunctions.cpp:296-304)
 *   lineSearch                                         == LineSearchStrategy::run on the LQ model (LineSearchStrategy.cpp:125-258)
 *   flatten                                            == LinearController::flatten at its own time stamps (LinearController.cpp:87-140)
 *
 * The header is dependency-free and duck-typed: every template works with the reference's Eigen-backed types
 * (ocs2::ModelData, ocs2::ScalarFunctionQuadraticApproximation, ocs2::LinearController, vector_t / matrix_t, all
 * column-major) and with any type offering the same members and data() / rows() / cols() / size() / resize(). Eigen itself is
 * not available in this repository's image; tests/cpp/test_frontend.cpp drives the header with minimal stand-ins of those
 * types and checks the results against the CPU oracle.
 */
#ifndef OCS2_DDP_CUDA_BATCHED_RICCATI_SOLVER_H_
#define OCS2_DDP_CUDA_BATCHED_RICCATI_SOLVER_H_

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ocs2_ddp_cuda.h"

namespace ocs2_ddp_cuda {

/** Pinned host SoA batch [problem][node][block] of one ModelData field. */
struct HostField {
  double* ptr = nullptr;
  std::size_t block = 0, nodes = 0;
  double* at(int problem, int node) const { return ptr + (static_cast<std::size_t>(problem) * nodes + node) * block; }
  o2c_field field() const { return {ptr, static_cast<int64_t>(nodes * block), static_cast<int64_t>(block)}; }
};

class BatchedRiccatiSolver {
 public:
  /** Throws std::runtime_error (message of o2c_last_error) when the configuration is rejected or no CUDA device exists.
   * maxEvents: room for that many events per instance (only SLQ keeps separate jump model data). */
  explicit BatchedRiccatiSolver(const o2c_config& config, int maxEvents = 0) : cfg_(config), maxEvents_(maxEvents) {
    check(o2c_abi_version() == O2C_ABI_VERSION ? O2C_OK : O2C_ERR_INVALID_ARGUMENT, "ABI version mismatch");
    check(o2c_create(&cfg_, &h_), "o2c_create");
    const std::size_t n = cfg_.nx, m = cfg_.nu, nc = cfg_.nc_max, N = cfg_.num_stages;
    nodes_ = cfg_.algorithm == O2C_ALG_ILQR ? N : N + 1;
    try {
      alloc(A_, n * n, nodes_), alloc(B_, n * m, nodes_), alloc(Hv_, n, nodes_);
      alloc(Q_, n * n, nodes_), alloc(P_, m * n, nodes_), alloc(R_, m * m, nodes_);
      alloc(q_, n, nodes_), alloc(r_, m, nodes_), alloc(c_, 1, nodes_);
      if (nc > 0) {
        alloc(C_, nc * n, nodes_), alloc(D_, nc * m, nodes_), alloc(e_, nc, nodes_);
        ncActive_.assign(static_cast<std::size_t>(cfg_.batch) * nodes_, 0);
      }
      alloc(Qf_, n * n, 1), alloc(qf_, n, 1), alloc(cf_, 1, 1), alloc(x0_, n, 1);
      if (cfg_.has_nominal) alloc(xNom_, n, N + 1), alloc(uNom_, m, N + 1);
      alloc(K_, m * n, N + 1), alloc(dbias_, m, N + 1), alloc(bias_, m, N + 1);
      alloc(Sm_, n * n, N + 1), alloc(Sv_, n, N + 1), alloc(s_, 1, N + 1);
      status_.assign(cfg_.batch, 0);
      if (cfg_.algorithm == O2C_ALG_ILQR || maxEvents_ > 0) event_.assign(static_cast<std::size_t>(cfg_.batch) * nodes_, 0);
      if (cfg_.algorithm == O2C_ALG_SLQ && maxEvents_ > 0) {
        alloc(jA_, n * n, maxEvents_), alloc(jHv_, n, maxEvents_), alloc(jQ_, n * n, maxEvents_), alloc(jq_, n, maxEvents_), alloc(jc_, 1, maxEvents_);
        eventCount_.assign(cfg_.batch, 0);
        lastEventNode_.assign(cfg_.batch, -1);
      }
      time_.resize(N + 1);
      for (std::size_t k = 0; k <= N; ++k) time_[k] = cfg_.time_step * static_cast<double>(k);
    } catch (...) {
      release();
      throw;
    }
  }
  ~BatchedRiccatiSolver() { release(); }
  BatchedRiccatiSolver(const BatchedRiccatiSolver&) = delete;
  BatchedRiccatiSolver& operator=(const BatchedRiccatiSolver&) = delete;

  const o2c_config& config() const { return cfg_; }
  o2c_handle* handle() const { return h_; }
  /** sweep kernel serving the current data (diagnostics) */
  std::string kernelVariant() const { return o2c_kernel_variant(h_); }

  /**
   * Instance b's model data: `modelDataTrajectory` as in PrimalDataContainer (one ModelData per time node; ILQR reads the
   * discretised stage data of nodes 0..N-1, SLQ the continuous-time data of nodes 0..N) and the final value function
   * (heuristics, already Hessian-corrected: GaussNewtonDDP.cpp:724-727). Safe to call concurrently for different b.
   */
  template <class ModelDataArray, class ScalarQuadratic>
  void setModelData(int b, const ModelDataArray& modelDataTrajectory, const ScalarQuadratic& finalValueFunction) {
    checkInstance(b);
    if (modelDataTrajectory.size() < nodes_) throw std::runtime_error("[BatchedRiccatiSolver] modelDataTrajectory is shorter than the horizon");
    for (std::size_t k = 0; k < nodes_; ++k) {
      const auto& md = modelDataTrajectory[k];
      if (md.stateDim != cfg_.nx || md.inputDim != cfg_.nu) throw std::runtime_error(sizeError(b, k, "state/input dimension"));
      copyBlock(md.dynamics.dfdx, A_, b, k, "dynamics.dfdx");
      copyBlock(md.dynamics.dfdu, B_, b, k, "dynamics.dfdu");
      copyBlock(md.dynamicsBias, Hv_, b, k, "dynamicsBias");
      copyCost(md.cost, b, k);
      if (cfg_.nc_max > 0) copyConstraint(md.stateInputEqConstraint, b, k);
      if (!event_.empty()) event_[static_cast<std::size_t>(b) * nodes_ + k] = 0;
    }
    if (!eventCount_.empty()) eventCount_[b] = 0, lastEventNode_[b] = -1;
    copyRaw(finalValueFunction.dfdxx, Qf_.at(b, 0), Qf_.block, "finalValueFunction.dfdxx");
    copyRaw(finalValueFunction.dfdx, qf_.at(b, 0), qf_.block, "finalValueFunction.dfdx");
    *cf_.at(b, 0) = finalValueFunction.f;
  }

  /**
   * Marks node `preEventNode` of instance b (postEventIndex - 1) as a pre-event node and installs the jump model data
   * (nominalPrimalData_.modelDataEventTimes[i]: dynamics.dfdx, dynamicsBias, cost.{dfdxx, dfdx, f}). Call after setModelData.
   * ILQR: the node keeps the input-side blocks of setModelData, which shape its controller entry (ILQR.cpp:263-295).
   * SLQ: the node keeps all of its continuous-time model data; the jump data are stored per event (SLQ.cpp:256-302) — call in
   * increasing node order; the instances run in lock-step on one time grid, so every instance must mark the same nodes.
   */
  template <class ModelDataT>
  void setEvent(int b, int preEventNode, const ModelDataT& modelDataEventTime) {
    checkInstance(b);
    if (preEventNode < 0 || static_cast<std::size_t>(preEventNode) >= nodes_) throw std::runtime_error("[BatchedRiccatiSolver] event node out of range");
    if (cfg_.algorithm == O2C_ALG_SLQ) {
      if (maxEvents_ <= 0 || eventCount_[b] >= maxEvents_) throw std::runtime_error("[BatchedRiccatiSolver] more events than maxEvents");
      if (preEventNode <= lastEventNode_[b]) throw std::runtime_error("[BatchedRiccatiSolver] SLQ events must be set in increasing node order");
      const int e = eventCount_[b]++;
      lastEventNode_[b] = preEventNode;
      copyBlock(modelDataEventTime.dynamics.dfdx, jA_, b, e, "event dynamics.dfdx");
      copyBlock(modelDataEventTime.dynamicsBias, jHv_, b, e, "event dynamicsBias");
      copyBlock(modelDataEventTime.cost.dfdxx, jQ_, b, e, "event cost.dfdxx");
      copyBlock(modelDataEventTime.cost.dfdx, jq_, b, e, "event cost.dfdx");
      *jc_.at(b, e) = modelDataEventTime.cost.f;
      event_[static_cast<std::size_t>(b) * nodes_ + preEventNode] = 1;
      return;
    }
    const int k = preEventNode;
    copyBlock(modelDataEventTime.dynamics.dfdx, A_, b, k, "event dynamics.dfdx");
    copyBlock(modelDataEventTime.dynamicsBias, Hv_, b, k, "event dynamicsBias");
    copyBlock(modelDataEventTime.cost.dfdxx, Q_, b, k, "event cost.dfdxx");
    copyBlock(modelDataEventTime.cost.dfdx, q_, b, k, "event cost.dfdx");
    *c_.at(b, k) = modelDataEventTime.cost.f;
    event_[static_cast<std::size_t>(b) * nodes_ + k] = 1;
  }

  /**
   * The same LQ arrays as another in-tree consumer hands them over: HpipmInterface::solve (ocs2_sqp/hpipm_catkin/include/hpipm_catkin/
   * HpipmInterface.h:85-87, called from SqpSolver.cpp:287-293) takes x0, N dynamics {f, dfdx, dfdu} (x+ = dfdx x + dfdu u + f) and
   * N+1 costs {f, dfdx, dfdu, dfdxx, dfdux, dfduu}, the last one being the terminal cost. Without inequality / equality constraints
   * the QP solution is this class's backward pass + rollout with step length 1 (create the handle with hessian_multiple = 0 and
   * has_nominal = 0): solveQps() then getQpSolution(), getValueFunctionTrajectory() (= getRiccatiCostToGo, with f filled in) and
   * calculateController() (gainArray_ = getRiccatiFeedback, deltaBiasArray_ = getRiccatiFeedforward).
   */
  template <class Vector, class LinearArray, class QuadraticArray>
  void setQp(int b, const Vector& x0, const LinearArray& dynamics, const QuadraticArray& cost) {
    checkInstance(b);
    if (cfg_.algorithm != O2C_ALG_ILQR || cfg_.has_nominal) throw std::runtime_error("[BatchedRiccatiSolver] setQp needs a discrete (ILQR) handle without nominal trajectories");
    const std::size_t N = cfg_.num_stages;
    if (dynamics.size() < N || cost.size() < N + 1) throw std::runtime_error("[BatchedRiccatiSolver] setQp: N dynamics and N+1 costs are required");
    for (std::size_t k = 0; k < N; ++k) {
      copyBlock(dynamics[k].dfdx, A_, b, k, "dynamics.dfdx");
      copyBlock(dynamics[k].dfdu, B_, b, k, "dynamics.dfdu");
      copyBlock(dynamics[k].f, Hv_, b, k, "dynamics.f");
      copyCost(cost[k], b, k);
      if (cfg_.nc_max > 0) ncActive_[static_cast<std::size_t>(b) * nodes_ + k] = 0;
      if (!event_.empty()) event_[static_cast<std::size_t>(b) * nodes_ + k] = 0;
    }
    copyRaw(cost[N].dfdxx, Qf_.at(b, 0), Qf_.block, "terminal cost.dfdxx");
    copyRaw(cost[N].dfdx, qf_.at(b, 0), qf_.block, "terminal cost.dfdx");
    *cf_.at(b, 0) = cost[N].f;
    copyRaw(x0, x0_.at(b, 0), x0_.block, "x0");
  }
  /** solves every QP of the batch: backward pass + rollout with step length 1 */
  void solveQps() {
    solveSequentialRiccatiEquations();
    const double one[1] = {1.0};
    rolloutTrajectory(std::vector<double>(one, one + 1));
  }
  /** HpipmInterface::solve outputs: N+1 states and N inputs of the optimal trajectory of instance b */
  template <class VectorArray>
  void getQpSolution(int b, VectorArray& stateTrajectory, VectorArray& inputTrajectory) const {
    getRollout(b, 0, stateTrajectory, inputTrajectory);
    inputTrajectory.resize(cfg_.num_stages);
  }

  /** primalSolution.stateTrajectory_ / inputTrajectory_ of instance b (N+1 nodes); requires config.has_nominal. */
  template <class VectorArray>
  void setNominalTrajectories(int b, const VectorArray& stateTrajectory, const VectorArray& inputTrajectory) {
    checkInstance(b);
    if (!cfg_.has_nominal) throw std::runtime_error("[BatchedRiccatiSolver] the handle was created without nominal trajectories");
    const std::size_t count = static_cast<std::size_t>(cfg_.num_stages) + 1;
    if (stateTrajectory.size() < count || inputTrajectory.size() < count) throw std::runtime_error("[BatchedRiccatiSolver] nominal trajectories are shorter than the horizon");
    for (std::size_t k = 0; k < count; ++k) {
      copyRaw(stateTrajectory[k], xNom_.at(b, k), xNom_.block, "stateTrajectory");
      copyRaw(inputTrajectory[k], uNom_.at(b, k), uNom_.block, "inputTrajectory");
    }
  }

  /** initial state of the rollout of instance b (a deviation from x_nom[0] when the handle has no nominal trajectories); travels with
   * the model data, i.e. set it before solveSequentialRiccatiEquations */
  template <class Vector>
  void setInitState(int b, const Vector& initState) {
    checkInstance(b);
    copyRaw(initState, x0_.at(b, 0), x0_.block, "initState");
  }

  /** levenbergMarquardt riccatiMultiple of the next backward pass: the reference's strategy adapts it after every iteration
   * (LevenbergMarquardtStrategy.cpp:131-147), the handle otherwise keeps config.lm_riccati_multiple */
  void setRiccatiMultiple(double riccatiMultiple) {
    check(o2c_set_lm_riccati_multiple(h_, riccatiMultiple), "o2c_set_lm_riccati_multiple");
    cfg_.lm_riccati_multiple = riccatiMultiple;
  }

  /** primalSolution.timeTrajectory_ (N+1 node times, shared by the batch: the instances run in lock-step) */
  template <class ScalarArray>
  void setTimeTrajectory(const ScalarArray& timeTrajectory) {
    if (timeTrajectory.size() != time_.size()) throw std::runtime_error("[BatchedRiccatiSolver] timeTrajectory must hold N+1 nodes");
    for (std::size_t k = 0; k < time_.size(); ++k) time_[k] = timeTrajectory[k];
  }

  /**
   * GaussNewtonDDP.h:176 for all instances: uploads the batch, runs the backward pass and the controller computation on
   * the GPU, brings value function, controller and per-instance status back. Blocking.
   */
  void solveSequentialRiccatiEquations() {
    o2c_lq_view v = lqView();
    check(o2c_upload(h_, &v, 0, cfg_.batch), "o2c_upload");
    check(o2c_backward(h_, 0, cfg_.batch), "o2c_backward");
    o2c_solution_view sv{};
    sv.K = K_.field(), sv.dbias = dbias_.field(), sv.bias = bias_.field();
    sv.Sm = Sm_.field(), sv.Sv = Sv_.field(), sv.s = s_.field();
    sv.status = status_.data();
    check(o2c_download(h_, &sv, 0, cfg_.batch, 0), "o2c_download");
    check(o2c_sync(h_), "o2c_sync");
    solved_ = true;
  }

  /** O2C_STATUS_* bits of instance b after the last solve */
  int status(int b) const { return status_.at(b); }

  /** nominalDualData_.valueFunctionTrajectory of instance b: N+1 quadratic approximations {f, dfdx, dfdxx}. */
  template <class ScalarQuadraticArray>
  void getValueFunctionTrajectory(int b, ScalarQuadraticArray& valueFunctionTrajectory) const {
    requireSolved(b);
    const std::size_t count = static_cast<std::size_t>(cfg_.num_stages) + 1;
    valueFunctionTrajectory.resize(count);
    for (std::size_t k = 0; k < count; ++k) {
      auto& vf = valueFunctionTrajectory[k];
      vf.dfdxx.resize(cfg_.nx, cfg_.nx);
      vf.dfdx.resize(cfg_.nx);
      std::memcpy(vf.dfdxx.data(), Sm_.at(b, k), sizeof(double) * Sm_.block);
      std::memcpy(vf.dfdx.data(), Sv_.at(b, k), sizeof(double) * Sv_.block);
      vf.f = *s_.at(b, k);
    }
  }

  /**
   * GaussNewtonDDP::calculateController for instance b (GaussNewtonDDP.cpp:588-642): fills timeStamp_, gainArray_, biasArray_,
   * deltaBiasArray_ of a LinearController (the last node already is the copy of node N-1) and throws the reference's
   * std::runtime_error when a gain or feedforward entry is not finite (checkNumericalStability_).
   */
  template <class LinearControllerT>
  void calculateController(int b, LinearControllerT& controller, bool checkNumericalStability = true) const {
    requireSolved(b);
    const std::size_t count = static_cast<std::size_t>(cfg_.num_stages) + 1;
    controller.timeStamp_.assign(time_.begin(), time_.end());
    controller.gainArray_.resize(count);
    controller.biasArray_.resize(count);
    controller.deltaBiasArray_.resize(count);
    for (std::size_t k = 0; k < count; ++k) {
      controller.gainArray_[k].resize(cfg_.nu, cfg_.nx);
      controller.biasArray_[k].resize(cfg_.nu);
      controller.deltaBiasArray_[k].resize(cfg_.nu);
      std::memcpy(controller.gainArray_[k].data(), K_.at(b, k), sizeof(double) * K_.block);
      std::memcpy(controller.biasArray_[k].data(), bias_.at(b, k), sizeof(double) * bias_.block);
      std::memcpy(controller.deltaBiasArray_[k].data(), dbias_.at(b, k), sizeof(double) * dbias_.block);
    }
    if (!checkNumericalStability) return;
    for (std::size_t k = 0; k < count; ++k) {
      std::stringstream errorDescription;
      if (!allFinite(K_.at(b, k), K_.block)) errorDescription << "Feedback gains are unstable!\n";
      if (!allFinite(dbias_.at(b, k), dbias_.block)) errorDescription << "Feedforward control is unstable!\n";
      if (errorDescription.tellp() != 0) {
        std::stringstream errorMessage;
        errorMessage << "At time " << time_[k] << " [sec].\n" << errorDescription.str();
        throw std::runtime_error(errorMessage.str());
      }
    }
  }

  /**
   * LinearController::flatten of instance b at the controller's own time stamps (LinearController.cpp:87-140), after
   * incrementController(stepLength): flatArray2[k] = float rows [uff_i, K_i,:] of node k — the `data` entries of
   * ocs2_msgs/mpc_flattened_controller (MPC_ROS_Interface.cpp:175). The FP64 -> float conversion runs on the device.
   */
  template <class FloatArray2>
  void flatten(int b, double stepLength, FloatArray2& flatArray2) const {
    requireSolved(b);
    const std::size_t count = static_cast<std::size_t>(cfg_.num_stages) + 1, len = static_cast<std::size_t>(cfg_.nu) * (cfg_.nx + 1);
    std::vector<float> flat(count * len);
    check(o2c_download_flattened_controller(h_, flat.data(), stepLength, b, 1), "o2c_download_flattened_controller");
    flatArray2.resize(count);
    for (std::size_t k = 0; k < count; ++k) flatArray2[k].assign(flat.begin() + k * len, flat.begin() + (k + 1) * len);
  }

  /**
   * incrementController(alpha) + rollout of the LQ model for every step length of `stepLengths` and every instance
   * (one launch; LineSearchStrategy.cpp:169-170 runs them on worker threads). Results through getRollout. Blocking.
   */
  template <class ScalarArray>
  void rolloutTrajectory(const ScalarArray& stepLengths) {
    if (!solved_) throw std::runtime_error("[BatchedRiccatiSolver] rolloutTrajectory before solveSequentialRiccatiEquations");
    std::vector<double> alphas(stepLengths.begin(), stepLengths.end());
    check(o2c_rollout(h_, alphas.data(), static_cast<int32_t>(alphas.size()), 0, cfg_.batch), "o2c_rollout");
    fetchRollouts(static_cast<int>(alphas.size()));
  }

  /** number of nodes of a rollout and their times (N+1 node times for ILQR, the RK4 step schedule for SLQ) */
  std::vector<double> rolloutTimes() const {
    int32_t count = 0;
    check(o2c_rollout_num_nodes(h_, &count), "o2c_rollout_num_nodes");
    std::vector<double> t(count);
    check(o2c_rollout_times(h_, t.data()), "o2c_rollout_times");
    return t;
  }

  /** state / input trajectories of rollout `alphaIndex` of instance b; throws like rolloutTrajectory (DDP_HelperFunctions.cpp:132-134)
   * when the rollout is not finite. */
  template <class VectorArray>
  void getRollout(int b, int alphaIndex, VectorArray& stateTrajectory, VectorArray& inputTrajectory) const {
    checkInstance(b);
    if (alphaIndex < 0 || alphaIndex >= rolloutAlphas_) throw std::runtime_error("[BatchedRiccatiSolver] no such rollout");
    stateTrajectory.resize(rolloutNodes_);
    inputTrajectory.resize(rolloutNodes_);
    const double* x = xRoll_.data() + (static_cast<std::size_t>(alphaIndex) * cfg_.batch + b) * rolloutNodes_ * cfg_.nx;
    const double* u = uRoll_.data() + (static_cast<std::size_t>(alphaIndex) * cfg_.batch + b) * rolloutNodes_ * cfg_.nu;
    if (!allFinite(x, rolloutNodes_ * cfg_.nx) || !allFinite(u, rolloutNodes_ * cfg_.nu)) throw std::runtime_error("[rolloutTrajectory] System became unstable during the rollout!");
    for (std::size_t k = 0; k < rolloutNodes_; ++k) {
      stateTrajectory[k].resize(cfg_.nx);
      inputTrajectory[k].resize(cfg_.nu);
      std::memcpy(stateTrajectory[k].data(), x + k * cfg_.nx, sizeof(double) * cfg_.nx);
      std::memcpy(inputTrajectory[k].data(), u + k * cfg_.nu, sizeof(double) * cfg_.nu);
    }
  }

  /** result of lineSearch for one instance */
  struct LineSearchResult {
    double stepLength = 0.0;  //!< 0 when no candidate passed the Armijo test (LineSearchStrategy.cpp:150-152)
    int candidateIndex = -1;
    double baselineMerit = 0.0, controllerUpdateIS = 0.0;
    std::vector<double> merits;  //!< merit of every candidate step length
  };

  /**
   * LineSearchStrategy::run on the LQ model for all instances (ILQR). `baselineMerit`: the performance index of the nominal
   * trajectory per instance, or empty for the LQ cost of the zero-deviation trajectory. The winning rollouts stay available
   * through getRollout(b, result.candidateIndex, ...).
   */
  std::vector<LineSearchResult> lineSearch(const o2c_line_search_settings& settings, const std::vector<double>& baselineMerit = {}) {
    if (!solved_) throw std::runtime_error("[BatchedRiccatiSolver] lineSearch before solveSequentialRiccatiEquations");
    if (!baselineMerit.empty() && baselineMerit.size() != static_cast<std::size_t>(cfg_.batch)) throw std::runtime_error("[BatchedRiccatiSolver] one baseline merit per instance");
    check(o2c_line_search(h_, &settings, baselineMerit.empty() ? nullptr : baselineMerit.data(), 0, cfg_.batch), "o2c_line_search");
    const std::size_t B = cfg_.batch;
    std::vector<double> step(B), merits(static_cast<std::size_t>(cfg_.max_alphas) * B), base(B), is(B), cand(cfg_.max_alphas);
    std::vector<int32_t> index(B);
    int32_t count = 0;
    check(o2c_line_search_result(h_, step.data(), index.data(), merits.data(), base.data(), is.data(), cand.data(), &count, 0, cfg_.batch), "o2c_line_search_result");
    std::vector<LineSearchResult> out(B);
    for (std::size_t b = 0; b < B; ++b) {
      out[b].stepLength = step[b], out[b].candidateIndex = index[b], out[b].baselineMerit = base[b], out[b].controllerUpdateIS = is[b];
      out[b].merits.resize(count);
      for (int e = 0; e < count; ++e) out[b].merits[e] = merits[static_cast<std::size_t>(e) * B + b];
    }
    fetchRollouts(count);
    return out;
  }

  /** read access to the packed batch (tests) */
  const HostField& packedField(const char* name) const {
    const struct { const char* n; const HostField* f; } table[] = {{"A", &A_}, {"B", &B_}, {"Hv", &Hv_}, {"Q", &Q_}, {"P", &P_}, {"R", &R_},
        {"q", &q_}, {"r", &r_}, {"c", &c_}, {"C", &C_}, {"D", &D_}, {"e", &e_}, {"Qf", &Qf_}, {"qf", &qf_}, {"cf", &cf_}, {"x0", &x0_}};
    for (const auto& t : table)
      if (std::strcmp(t.n, name) == 0) return *t.f;
    throw std::runtime_error("[BatchedRiccatiSolver] unknown field");
  }

 private:
  static void check(o2c_error e, const char* what) {
    if (e != O2C_OK) throw std::runtime_error(std::string("[ocs2_ddp_cuda] ") + what + ": " + o2c_last_error());
  }
  void checkInstance(int b) const {
    if (b < 0 || b >= cfg_.batch) throw std::runtime_error("[BatchedRiccatiSolver] instance index out of range");
  }
  void requireSolved(int b) const {
    checkInstance(b);
    if (!solved_) throw std::runtime_error("[BatchedRiccatiSolver] solveSequentialRiccatiEquations has not run");
  }
  static bool allFinite(const double* p, std::size_t count) {
    for (std::size_t i = 0; i < count; ++i)
      if (!std::isfinite(p[i])) return false;
    return true;
  }
  static std::string sizeError(int b, std::size_t k, const char* what) {
    std::stringstream ss;
    ss << "[BatchedRiccatiSolver] instance " << b << ", node " << k << ": " << what << " does not match the configured dimensions";
    return ss.str();
  }
  void alloc(HostField& f, std::size_t block, std::size_t nodes) {
    f.block = block, f.nodes = nodes;
    void* p = nullptr;
    const std::size_t bytes = sizeof(double) * block * nodes * static_cast<std::size_t>(cfg_.batch);
    check(o2c_host_alloc(&p, bytes), "o2c_host_alloc");
    std::memset(p, 0, bytes);
    f.ptr = static_cast<double*>(p);
    owned_.push_back(p);
  }
  void release() {
    for (void* p : owned_) o2c_host_free(p);
    owned_.clear();
    if (h_) o2c_destroy(h_);
    h_ = nullptr;
  }
  template <class Dense>
  static void copyRaw(const Dense& src, double* dst, std::size_t count, const char* what) {
    if (static_cast<std::size_t>(src.size()) != count) throw std::runtime_error(std::string("[BatchedRiccatiSolver] size of ") + what + " does not match the configured dimensions");
    std::memcpy(dst, src.data(), sizeof(double) * count);
  }
  template <class Dense>
  void copyBlock(const Dense& src, const HostField& f, int b, std::size_t k, const char* what) const {
    if (static_cast<std::size_t>(src.size()) != f.block) throw std::runtime_error(sizeError(b, k, what));
    std::memcpy(f.at(b, static_cast<int>(k)), src.data(), sizeof(double) * f.block);
  }
  template <class ScalarQuadratic>
  void copyCost(const ScalarQuadratic& cost, int b, std::size_t k) const {
    copyBlock(cost.dfdxx, Q_, b, k, "cost.dfdxx");
    copyBlock(cost.dfdux, P_, b, k, "cost.dfdux");
    copyBlock(cost.dfduu, R_, b, k, "cost.dfduu");
    copyBlock(cost.dfdx, q_, b, k, "cost.dfdx");
    copyBlock(cost.dfdu, r_, b, k, "cost.dfdu");
    *c_.at(b, static_cast<int>(k)) = cost.f;
  }
  // nc x n / nc x m column-major blocks of the node -> leading dimension nc_max
  template <class VectorLinear>
  void copyConstraint(const VectorLinear& con, int b, std::size_t k) {
    const std::size_t nc = static_cast<std::size_t>(con.f.size()), ld = cfg_.nc_max;
    if (nc > ld) throw std::runtime_error(sizeError(b, k, "number of state-input equality constraints (> nc_max)"));
    if (nc > 0 && (static_cast<std::size_t>(con.dfdx.size()) != nc * cfg_.nx || static_cast<std::size_t>(con.dfdu.size()) != nc * cfg_.nu)) throw std::runtime_error(sizeError(b, k, "stateInputEqConstraint"));
    double* C = C_.at(b, static_cast<int>(k));
    double* D = D_.at(b, static_cast<int>(k));
    double* e = e_.at(b, static_cast<int>(k));
    std::memset(C, 0, sizeof(double) * C_.block), std::memset(D, 0, sizeof(double) * D_.block), std::memset(e, 0, sizeof(double) * e_.block);
    for (std::size_t j = 0; j < static_cast<std::size_t>(cfg_.nx); ++j)
      for (std::size_t i = 0; i < nc; ++i) C[i + ld * j] = con.dfdx.data()[i + nc * j];
    for (std::size_t j = 0; j < static_cast<std::size_t>(cfg_.nu); ++j)
      for (std::size_t i = 0; i < nc; ++i) D[i + ld * j] = con.dfdu.data()[i + nc * j];
    for (std::size_t i = 0; i < nc; ++i) e[i] = con.f.data()[i];
    ncActive_[static_cast<std::size_t>(b) * nodes_ + k] = static_cast<int32_t>(nc);
  }
  o2c_lq_view lqView() const {
    o2c_lq_view v{};
    v.A = A_.field(), v.B = B_.field(), v.Hv = Hv_.field();
    v.Q = Q_.field(), v.P = P_.field(), v.R = R_.field(), v.q = q_.field(), v.r = r_.field(), v.c = c_.field();
    if (cfg_.nc_max > 0) {
      v.C = C_.field(), v.D = D_.field(), v.e = e_.field();
      v.nc = ncActive_.data(), v.nc_problem_stride = static_cast<int64_t>(nodes_), v.nc_node_stride = 1;
    }
    v.Qf = Qf_.field(), v.qf = qf_.field(), v.cf = cf_.field(), v.x0 = x0_.field();
    if (cfg_.has_nominal) v.x_nom = xNom_.field(), v.u_nom = uNom_.field();
    v.time = time_.data();
    bool any = false;
    for (int32_t flag : event_) any = any || flag != 0;
    if (any) v.event = event_.data(), v.event_problem_stride = static_cast<int64_t>(nodes_), v.event_node_stride = 1;
    if (any && cfg_.algorithm == O2C_ALG_SLQ)
      v.jump_A = jA_.field(), v.jump_Hv = jHv_.field(), v.jump_Q = jQ_.field(), v.jump_q = jq_.field(), v.jump_c = jc_.field();
    return v;
  }
  void fetchRollouts(int nAlpha) {
    int32_t count = 0;
    check(o2c_rollout_num_nodes(h_, &count), "o2c_rollout_num_nodes");
    rolloutNodes_ = count, rolloutAlphas_ = nAlpha;
    const std::size_t B = cfg_.batch;
    xRoll_.resize(static_cast<std::size_t>(nAlpha) * B * count * cfg_.nx);
    uRoll_.resize(static_cast<std::size_t>(nAlpha) * B * count * cfg_.nu);
    o2c_solution_view sv{};
    sv.x = {xRoll_.data(), static_cast<int64_t>(count) * cfg_.nx, cfg_.nx};
    sv.u = {uRoll_.data(), static_cast<int64_t>(count) * cfg_.nu, cfg_.nu};
    sv.x_alpha_stride = static_cast<int64_t>(B) * count * cfg_.nx;
    sv.u_alpha_stride = static_cast<int64_t>(B) * count * cfg_.nu;
    check(o2c_download(h_, &sv, 0, cfg_.batch, nAlpha), "o2c_download");
    check(o2c_sync(h_), "o2c_sync");
  }

  o2c_config cfg_{};
  o2c_handle* h_ = nullptr;
  std::size_t nodes_ = 0;
  HostField A_, B_, Hv_, Q_, P_, R_, q_, r_, c_, C_, D_, e_, Qf_, qf_, cf_, x0_, xNom_, uNom_;
  HostField K_, dbias_, bias_, Sm_, Sv_, s_;
  HostField jA_, jHv_, jQ_, jq_, jc_;  // SLQ jump model data [instance][event]
  int maxEvents_ = 0;
  std::vector<int> eventCount_, lastEventNode_;
  std::vector<int32_t> ncActive_, event_, status_;
  std::vector<double> time_, xRoll_, uRoll_;
  std::vector<void*> owned_;
  std::size_t rolloutNodes_ = 0;
  int rolloutAlphas_ = 0;
  bool solved_ = false;
};

}  // namespace ocs2_ddp_cuda

#endif  // OCS2_DDP_CUDA_BATCHED_RICCATI_SOLVER_H_
