/*
 * ShardedRiccatiSolver.h — the multi-device form of BatchedRiccatiSolver (`device_ids[]` / `n_devices`, §8(e):
 * shard by problem index, no collective on the data path).
 *
 * One BatchedRiccatiSolver (= one o2c_handle, one set of CUDA streams, one set of pinned staging buffers) per entry of
 * `deviceIds`; instance b of the global batch lives on shard s with shardBegin(s) <= b < shardBegin(s + 1), contiguous blocks, the
 * first `batch % nShards` shards one instance larger — the same arithmetic as ocs2_b200/sharding.py: shard_bounds, which the
 * one-process-per-GPU launch (bench.py under torchrun) uses. Every blocking phase (solveSequentialRiccatiEquations, rolloutTrajectory,
 * lineSearch) runs on ONE HOST THREAD PER SHARD, the way the reference drives independent solvers from a pool of threads
 * (ocs2_mpcnet/ocs2_mpcnet_core/src/rollout/MpcnetRolloutManager.cpp:86-96); the per-instance setters / getters forward to the owning
 * shard and are safe to call concurrently for different b. Problems share no data, so the result of every instance is BIT-IDENTICAL
 * to the single-device solver's (tests/cpp/test_sharded.cpp checks that on the GPU, incl. two shards on one device).
 *
 * A device id may appear more than once (two handles on one GPU: their streams overlap upload, compute and download of the shards).
 */
#ifndef OCS2_DDP_CUDA_SHARDED_RICCATI_SOLVER_H_
#define OCS2_DDP_CUDA_SHARDED_RICCATI_SOLVER_H_

#include <exception>
#include <memory>
#include <thread>
#include <utility>

#include "BatchedRiccatiSolver.h"

namespace ocs2_ddp_cuda {

class ShardedRiccatiSolver {
 public:
  using LineSearchResult = BatchedRiccatiSolver::LineSearchResult;

  /** `config.batch` is the GLOBAL number of instances, `config.device` is ignored. An empty `deviceIds` means every visible device
   * (o2c_device_count). Throws std::runtime_error like BatchedRiccatiSolver (no device, rejected configuration). */
  explicit ShardedRiccatiSolver(const o2c_config& config, std::vector<int> deviceIds = {}, int maxEvents = 0) : cfg_(config) {
    if (deviceIds.empty()) {
      int32_t count = 0;
      if (o2c_device_count(&count) != O2C_OK) throw std::runtime_error(std::string("[ocs2_ddp_cuda] o2c_device_count: ") + o2c_last_error());
      for (int d = 0; d < count; ++d) deviceIds.push_back(d);
    }
    if (config.batch < static_cast<int>(deviceIds.size())) deviceIds.resize(config.batch > 0 ? config.batch : 1);
    const int S = static_cast<int>(deviceIds.size());
    begin_.resize(S + 1);
    for (int s = 0; s <= S; ++s) begin_[s] = shardBegin(config.batch, S, s);
    for (int s = 0; s < S; ++s) {
      o2c_config c = config;
      c.device = deviceIds[s];
      c.batch = begin_[s + 1] - begin_[s];
      shards_.emplace_back(new BatchedRiccatiSolver(c, maxEvents));
    }
    devices_ = std::move(deviceIds);
  }

  /** first global instance of shard s out of nShards (s == nShards: the batch size) */
  static int shardBegin(int batch, int nShards, int s) {
    const int base = batch / nShards, extra = batch % nShards;
    return s * base + (s < extra ? s : extra);
  }

  int numShards() const { return static_cast<int>(shards_.size()); }
  int shardBegin(int s) const { return begin_.at(s); }
  int shardOf(int b) const {
    if (b < 0 || b >= cfg_.batch) throw std::runtime_error("[ShardedRiccatiSolver] instance index out of range");
    int s = 0;
    while (b >= begin_[s + 1]) ++s;
    return s;
  }
  int deviceOf(int b) const { return devices_[shardOf(b)]; }
  BatchedRiccatiSolver& shard(int s) { return *shards_.at(s); }
  const BatchedRiccatiSolver& shard(int s) const { return *shards_.at(s); }
  const o2c_config& config() const { return cfg_; }

  // ---- per-instance hand-over: forwarded to the owning shard (see BatchedRiccatiSolver for the meaning of each call) ----
  template <class ModelDataArray, class ScalarQuadratic>
  void setModelData(int b, const ModelDataArray& modelDataTrajectory, const ScalarQuadratic& finalValueFunction) {
    const int s = shardOf(b);
    shards_[s]->setModelData(b - begin_[s], modelDataTrajectory, finalValueFunction);
  }
  template <class ModelDataT>
  void setEvent(int b, int preEventNode, const ModelDataT& modelDataEventTime) {
    const int s = shardOf(b);
    shards_[s]->setEvent(b - begin_[s], preEventNode, modelDataEventTime);
  }
  template <class VectorArray>
  void setNominalTrajectories(int b, const VectorArray& stateTrajectory, const VectorArray& inputTrajectory) {
    const int s = shardOf(b);
    shards_[s]->setNominalTrajectories(b - begin_[s], stateTrajectory, inputTrajectory);
  }
  template <class Vector>
  void setInitState(int b, const Vector& initState) {
    const int s = shardOf(b);
    shards_[s]->setInitState(b - begin_[s], initState);
  }
  template <class ScalarArray>
  void setTimeTrajectory(const ScalarArray& timeTrajectory) {
    for (auto& sh : shards_) sh->setTimeTrajectory(timeTrajectory);
  }
  /** levenbergMarquardt riccatiMultiple of the next backward pass, all shards */
  void setRiccatiMultiple(double riccatiMultiple) {
    for (auto& sh : shards_) sh->setRiccatiMultiple(riccatiMultiple);
  }

  // ---- blocking phases: one host thread per shard ----
  void solveSequentialRiccatiEquations() {
    parallel([](BatchedRiccatiSolver& sh) { sh.solveSequentialRiccatiEquations(); });
  }
  template <class ScalarArray>
  void rolloutTrajectory(const ScalarArray& stepLengths) {
    const std::vector<double> alphas(stepLengths.begin(), stepLengths.end());
    parallel([&alphas](BatchedRiccatiSolver& sh) { sh.rolloutTrajectory(alphas); });
  }
  /** `baselineMerit`: one entry per GLOBAL instance, or empty */
  std::vector<LineSearchResult> lineSearch(const o2c_line_search_settings& settings, const std::vector<double>& baselineMerit = {}) {
    if (!baselineMerit.empty() && baselineMerit.size() != static_cast<std::size_t>(cfg_.batch)) throw std::runtime_error("[ShardedRiccatiSolver] one baseline merit per instance");
    std::vector<std::vector<LineSearchResult>> parts(shards_.size());
    parallelIndexed([&](int s, BatchedRiccatiSolver& sh) {
      std::vector<double> base;
      if (!baselineMerit.empty()) base.assign(baselineMerit.begin() + begin_[s], baselineMerit.begin() + begin_[s + 1]);
      parts[s] = sh.lineSearch(settings, base);
    });
    std::vector<LineSearchResult> out;
    out.reserve(cfg_.batch);
    for (auto& p : parts) out.insert(out.end(), std::make_move_iterator(p.begin()), std::make_move_iterator(p.end()));
    return out;
  }

  // ---- per-instance results ----
  int status(int b) const {
    const int s = shardOf(b);
    return shards_[s]->status(b - begin_[s]);
  }
  template <class ScalarQuadraticArray>
  void getValueFunctionTrajectory(int b, ScalarQuadraticArray& valueFunctionTrajectory) const {
    const int s = shardOf(b);
    shards_[s]->getValueFunctionTrajectory(b - begin_[s], valueFunctionTrajectory);
  }
  template <class LinearControllerT>
  void calculateController(int b, LinearControllerT& controller, bool checkNumericalStability = true) const {
    const int s = shardOf(b);
    shards_[s]->calculateController(b - begin_[s], controller, checkNumericalStability);
  }
  template <class FloatArray2>
  void flatten(int b, double stepLength, FloatArray2& flatArray2) const {
    const int s = shardOf(b);
    shards_[s]->flatten(b - begin_[s], stepLength, flatArray2);
  }
  template <class VectorArray>
  void getRollout(int b, int alphaIndex, VectorArray& stateTrajectory, VectorArray& inputTrajectory) const {
    const int s = shardOf(b);
    shards_[s]->getRollout(b - begin_[s], alphaIndex, stateTrajectory, inputTrajectory);
  }
  std::vector<double> rolloutTimes() const { return shards_.front()->rolloutTimes(); }

 private:
  template <class F>
  void parallel(F&& f) {
    parallelIndexed([&f](int, BatchedRiccatiSolver& sh) { f(sh); });
  }
  // runs f(s, shard s) on one thread per shard (the calling thread takes shard 0); the first exception is rethrown after all joined
  template <class F>
  void parallelIndexed(F&& f) {
    const int S = numShards();
    std::vector<std::exception_ptr> errors(S);
    auto task = [&](int s) {
      try {
        f(s, *shards_[s]);
      } catch (...) {
        errors[s] = std::current_exception();
      }
    };
    std::vector<std::thread> workers;
    for (int s = 1; s < S; ++s) workers.emplace_back(task, s);
    task(0);
    for (auto& w : workers) w.join();
    for (auto& e : errors)
      if (e) std::rethrow_exception(e);
  }

  o2c_config cfg_{};
  std::vector<int> devices_, begin_;
  std::vector<std::unique_ptr<BatchedRiccatiSolver>> shards_;
};

}  // namespace ocs2_ddp_cuda

#endif  // OCS2_DDP_CUDA_SHARDED_RICCATI_SOLVER_H_
