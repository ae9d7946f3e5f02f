// ShardedRiccatiSolver (include/ocs2_ddp_cuda/ShardedRiccatiSolver.h): N handles on the devices of one process, one host thread per
// shard. Problems share no data, so every instance must come out BIT-IDENTICAL to the single-handle solve (and that one is held to
// the CPU oracle at 1e-9 relative). With one visible GPU the shards are several handles on device 0 driven concurrently from their
// own threads; with two or more they spread over the devices (device ids cycle through the visible ones).
//
//   test_sharded --gpu     all cases
//   test_sharded --no-gpu  the constructor must throw (there is no CPU fallback)
#include <cstdio>
#include <cstring>
#include <string>

#include "ocs2_ddp_cuda/ShardedRiccatiSolver.h"
#include "problem_fixture.h"

namespace {

using namespace ocs2;

struct Case {
  const char* name;
  int algorithm, n, m, nc, N, batch, shards;
  bool events;
};

template <class Solver>
void handOver(Solver& solver, const std::vector<fixture::Problem>& problems) {
  for (int b = 0; b < (int)problems.size(); ++b) {
    const fixture::Problem& pb = problems[b];
    solver.setModelData(b, pb.modelDataTrajectory, pb.finalValueFunction);
    for (size_t i = 0; i < pb.postEventIndices.size(); ++i) solver.setEvent(b, (int)pb.postEventIndices[i] - 1, pb.modelDataEventTimes[i]);
    solver.setNominalTrajectories(b, pb.stateTrajectory, pb.inputTrajectory);
    solver.setInitState(b, pb.initState);
  }
  solver.setTimeTrajectory(problems[0].time);
}

bool same(const Dense& a, const Dense& b) { return a.size() == b.size() && std::memcmp(a.data(), b.data(), sizeof(double) * a.size()) == 0; }

int runCase(const Case& cs, int deviceCount) {
  const double dt = 0.01;
  o2c_config cfg{};
  cfg.nx = cs.n, cfg.nu = cs.m, cfg.nc_max = cs.nc, cfg.num_stages = cs.N, cfg.batch = cs.batch, cfg.algorithm = cs.algorithm;
  cfg.riccati_form = O2C_FORM_REDUCED, cfg.strategy = O2C_STRATEGY_LINE_SEARCH, cfg.hessian_correction = O2C_HC_DIAGONAL_SHIFT;
  cfg.device = 0, cfg.max_alphas = 6, cfg.has_nominal = 1, cfg.hessian_multiple = 1e-5, cfg.time_step = dt;
  orc_settings ost{};
  ost.algorithm = cs.algorithm, ost.reduced_form = 1, ost.strategy = ORC_STRATEGY_LINE_SEARCH, ost.hessian_correction = ORC_HC_DIAGONAL_SHIFT;
  ost.hessian_multiple = 1e-5, ost.time_step = dt;

  std::vector<fixture::Problem> problems(cs.batch);
  uint64_t rng = 42;
  for (int b = 0; b < cs.batch; ++b) {
    problems[b].generate(777, b, cs.algorithm, cs.n, cs.m, cs.nc, cs.N, dt);
    if (cs.events)
      for (int j = 0; j < b % 3; ++j) {
        const int k = (int)(fixture::lcg(rng) >> 33) % (cs.N - 1);  // not the last stage
        if (!problems[b].event[k]) problems[b].addIlqrEvent(k, rng);
      }
    // events must be handed over in node order
    fixture::Problem& pb = problems[b];
    for (size_t i = 0; i + 1 < pb.postEventIndices.size(); ++i)
      for (size_t j = i + 1; j < pb.postEventIndices.size(); ++j)
        if (pb.postEventIndices[j] < pb.postEventIndices[i]) std::swap(pb.postEventIndices[i], pb.postEventIndices[j]), std::swap(pb.modelDataEventTimes[i], pb.modelDataEventTimes[j]);
  }

  std::vector<int> devices;
  for (int s = 0; s < cs.shards; ++s) devices.push_back(s % deviceCount);
  ocs2_ddp_cuda::BatchedRiccatiSolver single(cfg);
  ocs2_ddp_cuda::ShardedRiccatiSolver sharded(cfg, devices);
  if (sharded.numShards() != cs.shards || sharded.shardBegin(cs.shards) != cs.batch) {
    std::printf("FAIL %s: shard bookkeeping\n", cs.name);
    return 1;
  }
  handOver(single, problems);
  handOver(sharded, problems);
  single.solveSequentialRiccatiEquations();
  sharded.solveSequentialRiccatiEquations();
  const std::vector<double> alphas = {1.0, 0.5};
  single.rolloutTrajectory(alphas);
  sharded.rolloutTrajectory(alphas);

  double diff = 0.0, scale = 0.0;
  for (int b = 0; b < cs.batch; ++b) {
    std::vector<ScalarFunctionQuadraticApproximation> vfA, vfB;
    LinearController ctrlA, ctrlB;
    single.getValueFunctionTrajectory(b, vfA), sharded.getValueFunctionTrajectory(b, vfB);
    single.calculateController(b, ctrlA), sharded.calculateController(b, ctrlB);
    bool equal = single.status(b) == sharded.status(b) && vfA.size() == vfB.size() && ctrlA.timeStamp_ == ctrlB.timeStamp_;
    for (size_t k = 0; equal && k < vfA.size(); ++k)
      equal = same(vfA[k].dfdxx, vfB[k].dfdxx) && same(vfA[k].dfdx, vfB[k].dfdx) && std::memcmp(&vfA[k].f, &vfB[k].f, sizeof(double)) == 0 &&
              same(ctrlA.gainArray_[k], ctrlB.gainArray_[k]) && same(ctrlA.biasArray_[k], ctrlB.biasArray_[k]) && same(ctrlA.deltaBiasArray_[k], ctrlB.deltaBiasArray_[k]);
    for (int a = 0; equal && a < (int)alphas.size(); ++a) {
      vector_array_t xA, uA, xB, uB;
      single.getRollout(b, a, xA, uA), sharded.getRollout(b, a, xB, uB);
      equal = xA.size() == xB.size();
      for (size_t k = 0; equal && k < xA.size(); ++k) equal = same(xA[k], xB[k]) && same(uA[k], uB[k]);
    }
    if (!equal) {
      std::printf("FAIL %s: instance %d (shard %d, device %d) differs from the single-handle solve\n", cs.name, b, sharded.shardOf(b), sharded.deviceOf(b));
      return 1;
    }
    fixture::OracleSolution ref;  // and the single-handle solve against the oracle
    ref.solve(ost, problems[b], true);
    const size_t n = cs.n, m = cs.m;
    for (int k = 0; k <= cs.N; ++k) {
      fixture::accumulate(vfB[k].dfdxx.data(), &ref.Sm[k * n * n], n * n, diff, scale);
      fixture::accumulate(ctrlB.gainArray_[k].data(), &ref.K[k * m * n], m * n, diff, scale);
    }
  }
  if (cs.algorithm == O2C_ALG_ILQR) {
    const o2c_line_search_settings ls{0.05, 1.0, 0.5, 1e-4};
    const auto a = single.lineSearch(ls);
    const auto b = sharded.lineSearch(ls);
    bool equal = a.size() == b.size();
    for (size_t i = 0; equal && i < a.size(); ++i)
      equal = a[i].candidateIndex == b[i].candidateIndex && a[i].stepLength == b[i].stepLength && a[i].merits == b[i].merits &&
              a[i].baselineMerit == b[i].baselineMerit && a[i].controllerUpdateIS == b[i].controllerUpdateIS;
    if (!equal) {
      std::printf("FAIL %s: line search of the sharded solver differs\n", cs.name);
      return 1;
    }
  }
  const double rel = diff / scale;
  std::printf("%s %-30s %d shards on %d device(s), kernel %-18s bit-identical to one handle; vs oracle %.2e\n", rel <= 1e-9 ? "ok  " : "FAIL", cs.name,
              cs.shards, std::min(cs.shards, deviceCount), sharded.shard(0).kernelVariant().c_str(), rel);
  return rel <= 1e-9 ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "--gpu";
  if (mode == "--no-gpu") {
    o2c_config cfg{};
    cfg.nx = 4, cfg.nu = 1, cfg.num_stages = 5, cfg.batch = 3, cfg.max_alphas = 1, cfg.time_step = 0.01;
    try {
      ocs2_ddp_cuda::ShardedRiccatiSolver solver(cfg);
    } catch (const std::runtime_error& err) {
      std::printf("refused: %s\n", err.what());
      return std::strstr(err.what(), "no CPU fallback") ? 0 : 1;
    }
    std::printf("FAIL: constructed without a device\n");
    return 1;
  }
  int32_t deviceCount = 0;
  if (o2c_device_count(&deviceCount) != O2C_OK) {
    std::printf("FAIL: %s\n", o2c_last_error());
    return 1;
  }
  const Case cases[] = {
      {"legged ilqr 24x24", O2C_ALG_ILQR, 24, 24, 0, 12, 11, 2, false},
      {"legged ilqr 24x24 events", O2C_ALG_ILQR, 24, 24, 0, 9, 7, 3, true},
      {"manipulator ilqr 9x9 nc=3", O2C_ALG_ILQR, 9, 9, 3, 15, 10, 3, false},
      {"ballbot ilqr 10x3", O2C_ALG_ILQR, 10, 3, 0, 20, 64, 4, false},
      {"quadrotor slq 12x4", O2C_ALG_SLQ, 12, 4, 0, 10, 9, 2, false},
      {"legged slq 24x24", O2C_ALG_SLQ, 24, 24, 0, 6, 5, 2, false},
      {"generic ilqr 7x5 nc=2", O2C_ALG_ILQR, 7, 5, 2, 8, 6, 8, false},  // more shards than instances: clipped to 6
  };
  int failed = 0;
  for (const Case& cs : cases) {
    Case c = cs;
    if (c.shards > c.batch) c.shards = c.batch;
    try {
      failed += runCase(c, deviceCount);
    } catch (const std::exception& err) {
      std::printf("FAIL %s: %s\n", cs.name, err.what());
      ++failed;
    }
  }
  std::printf("%d device(s) visible, %d case(s) failed\n", deviceCount, failed);
  return failed ? 1 : 0;
}
