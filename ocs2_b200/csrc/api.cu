// api.cu — C ABI of libocs2_ddp_cuda (see include/ocs2_ddp_cuda.h). Host-side plumbing only: handle life cycle, device buffers,
// strided-SoA <-> record conversion, RK4 step schedules (boost::odeint integrate_times / integrate_adaptive semantics), kernel dispatch,
// and the chunked H2D / compute / D2H pipeline of o2c_solve_host. No CPU arithmetic fallback exists: every compute call launches CUDA kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "o2c_common.cuh"

using namespace o2c;

namespace {

thread_local std::string g_last_error;

o2c_error fail(o2c_error code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define O2C_CUDA(expr)                                                                                             \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess) {                                                                                       \
      return fail(_e == cudaErrorMemoryAllocation ? O2C_ERR_OUT_OF_MEMORY : O2C_ERR_CUDA,                          \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                             \
    }                                                                                                              \
  } while (0)

constexpr int kLanes = 3;

struct Lane {
  cudaStream_t stream = nullptr;
  double* stage_in = nullptr;
  double* stage_out = nullptr;
  size_t stage_in_doubles = 0, stage_out_doubles = 0;
  std::vector<double> bounce;  // host bounce buffer for irregular strides
  int* h_nc = nullptr;         // pinned staging of the per-node constraint counts of one chunk (asynchronous H2D without a lane sync)
  size_t h_nc_count = 0;
  cudaEvent_t nc_copied = nullptr;  // recorded after the last H2D out of h_nc: the next chunk on this lane waits for it before refilling
};

}  // namespace

struct o2c_handle {
  o2c_config cfg{};
  Layout L{};
  SolverSettings st{};
  Lane lanes[kLanes];
  double *d_lq = nullptr, *d_term = nullptr, *d_xnom = nullptr, *d_unom = nullptr, *d_x0 = nullptr, *d_time = nullptr;
  double *d_sol = nullptr, *d_xs = nullptr, *d_us = nullptr, *d_alphas = nullptr;
  int *d_nc = nullptr, *d_status = nullptr;
  double *d_ls_merit = nullptr, *d_ls_base = nullptr, *d_ls_is = nullptr, *d_ls_step = nullptr, *d_ls_basein = nullptr;  // line search
  int* d_ls_index = nullptr;
  double* d_dt = nullptr;  // step lengths of o2c_discretize [N], allocated on first use
  float* d_flat = nullptr;  // flattened controllers of the whole batch (o2c_download_flattened_controller), allocated on first use
  int* d_event = nullptr;       // [batch][nodes] pre-event flags, allocated by the first upload that carries events
  bool events_present = false;
  std::vector<int> slq_events;  // SLQ: pre-event node flags shared by the batch (empty = none)
  double* d_jump = nullptr;     // SLQ: [batch][jump_capacity] jump records
  int jump_capacity = 0;
  std::vector<double> ls_candidates;
  bool backward_done = false;
  int* h_status = nullptr;  // pinned bounce buffer [batch]: a caller's status array may be pageable, and an asynchronous copy into
                            // pageable memory blocks the host thread (it serialised the H2D / compute / D2H pipeline of o2c_solve_host)
  struct PendingStatus {
    int32_t* dst;
    int begin, count;
  };
  std::vector<PendingStatus> pending_status;
  SlqStep* d_slq_steps = nullptr;
  int n_slq_steps = 0;
  RolloutStep* d_ro_steps = nullptr;
  int n_ro_steps = 0, ro_first_idx = 0;
  double ro_first_alpha = 1.0;
  int out_nodes = 0;
  std::vector<double> time, ro_times;
  bool time_set = false;
  bool use_fast = false;  // legged shape: warp-per-problem DMMA kernel
  bool use_rpl = false;   // small shapes: row-per-lane kernel
  bool nc_ragged = false;  // a caller supplied per-node constraint counts (otherwise every node has nc_max: the kernels skip the lookup)
  int64_t launches = 0;
  int stage_chunk = 0;
  double* d_slq_ws = nullptr;  // projected node data of the legged SLQ kernels (slq_wpp.cu), allocated on first use
  int* d_counter = nullptr;  // [kLanes] work counters of the persistent kernels' dynamic problem fetch, one per stream lane

  int lane_of(cudaStream_t s) const {
    for (int i = 0; i < kLanes; ++i)
      if (lanes[i].stream == s) return i;
    return 0;
  }
  DeviceBuffers buffers(cudaStream_t s) const {
    DeviceBuffers b = buffers();
    b.work_counter = d_counter ? d_counter + lane_of(s) : nullptr;
    return b;
  }
  DeviceBuffers buffers() const {
    DeviceBuffers b{};
    b.lq = d_lq;
    b.term = d_term;
    b.x_nom = d_xnom;
    b.u_nom = d_unom;
    b.nc = nc_ragged ? d_nc : nullptr;
    b.event = events_present ? d_event : nullptr;
    b.jump = d_jump;
    b.jump_capacity = jump_capacity;
    b.x0 = d_x0;
    b.time = d_time;
    b.sol = d_sol;
    b.xs = d_xs;
    b.us = d_us;
    b.status = d_status;
    b.work_counter = nullptr;
    return b;
  }
};

namespace {

__global__ void fill_int_kernel(int* p, int v, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}

// LinearInterpolation::timeSegment (ocs2_core/include/ocs2_core/misc/implementation/LinearInterpolation.h:69-107)
void time_segment(double t, const std::vector<double>& time, int& index, double& alpha) {
  const int count = (int)time.size();
  if (count <= 1) {
    index = 0;
    alpha = 1.0;
    return;
  }
  const int idx = (int)(std::lower_bound(time.begin(), time.end(), t) - time.begin()) - 1;
  const int lastInterval = count - 1;
  if (idx >= 0) {
    if (idx < lastInterval) {
      const double len = time[idx + 1] - time[idx];
      const double till = time[idx + 1] - t;
      index = idx;
      if (len > 2.0 * 1e-9) {
        alpha = till / len;
      } else {
        alpha = (till < 0.5 * len) ? 0.0 : 1.0;
      }
      return;
    }
    index = std::max(lastInterval - 1, 0);
    alpha = 0.0;
    return;
  }
  index = 0;
  alpha = 1.0;
}
inline bool less_with_sign(double t1, double t2) { return (t2 - t1) > std::numeric_limits<double>::epsilon(); }
inline bool less_eq_with_sign(double t1, double t2) { return (t1 - t2) <= std::numeric_limits<double>::epsilon(); }

// SLQ backward schedule: boost::numeric::odeint::integrate_times with a plain RK4 stepper over z = -t reversed
// (SLQ.cpp:256-302, DDP_HelperFunctions.cpp:309-328, implementation/Integrator.h:298-311)
o2c_error build_slq_schedule(o2c_handle* h, std::vector<SlqStep>& steps) {
  const int N = h->L.N;
  const double dt = h->st.time_step;
  if (!(dt > 0.0)) return fail(O2C_ERR_INVALID_ARGUMENT, "time_step must be positive for SLQ");
  std::vector<double> z(N + 1);
  for (int j = 0; j <= N; ++j) z[j] = -h->time[N - j];
  steps.clear();
  const double cs[4] = {0.0, 0.5, 0.5, 1.0};
  int n_events = 0;
  for (int flag : h->slq_events) n_events += flag != 0;
  int events_below = n_events;  // events at nodes < the current interval's lower node, counted while walking backwards
  for (int j = 0; j < N; ++j) {
    const int i0 = N - 1 - j;
    if (!h->slq_events.empty() && h->slq_events[i0]) {
      // pre-event node i0 / post-event node i0 + 1: the segments are integrated separately (SLQ.cpp:269-296); the value function
      // crosses the event through computeJumpMap
      events_below -= 1;
      SlqStep s{};
      s.interval = i0;
      s.observe_node = i0;
      s.jump = events_below + 1;
      steps.push_back(s);
      continue;
    }
    double current_time = z[j];
    double current_dt = dt;
    const size_t first = steps.size();
    while (less_with_sign(current_time, z[j + 1])) {
      current_dt = std::min(dt, z[j + 1] - current_time);
      SlqStep s{};
      s.interval = i0;
      s.observe_node = -1;
      s.h = current_dt;
      for (int c = 0; c < 4; ++c) {
        const double zt = current_time + current_dt * cs[c];
        int idx;
        double a;
        time_segment(-zt, h->time, idx, a);
        if (idx == i0) {
          s.alpha[c] = a;
        } else if (idx == i0 - 1 && a == 0.0) {
          s.alpha[c] = 1.0;
        } else if (idx == i0 + 1 && a == 1.0) {
          s.alpha[c] = 0.0;
        } else {
          return fail(O2C_ERR_UNSUPPORTED, "SLQ step leaves its time interval (non-monotone or degenerate time grid)");
        }
      }
      steps.push_back(s);
      current_time += current_dt;
      current_dt = std::max(dt, current_dt);
      if (steps.size() > 50000000) return fail(O2C_ERR_INVALID_ARGUMENT, "SLQ schedule too long (time_step too small)");
    }
    if (steps.size() == first) {
      // zero-length interval: the observer still records the unchanged state; emit a zero step so the kernel writes node i0
      SlqStep s{};
      s.interval = i0;
      s.h = 0.0;
      for (int c = 0; c < 4; ++c) s.alpha[c] = 0.0;
      steps.push_back(s);
    }
    steps.back().observe_node = i0;
  }
  return O2C_OK;
}

// continuous rollout schedule: TimeTriggeredRollout::run -> integrateAdaptive with a plain stepper = integrate_const steps of
// timeStep + one truncated last step; start nudged by weakEpsilon (RolloutBase.cpp:62-64)
o2c_error build_rollout_schedule(o2c_handle* h, std::vector<RolloutStep>& steps) {
  const int N = h->L.N;
  const double dt = h->st.time_step;
  if (!(dt > 0.0)) return fail(O2C_ERR_INVALID_ARGUMENT, "time_step must be positive for the continuous rollout");
  const double t0 = h->time[0], tf = h->time[N];
  steps.clear();
  h->ro_times.clear();
  // RolloutBase::findActiveModesTimeInterval (RolloutBase.cpp:43-67): the event times (stamps of the pre-event nodes) split [t0, tf]
  // into intervals whose start is nudged by weakEpsilon
  std::vector<double> switching{t0};
  std::vector<int> event_node;
  for (int k = 0; k < N && !h->slq_events.empty(); ++k)
    if (h->slq_events[k]) {
      switching.push_back(h->time[k]);
      event_node.push_back(k);
    }
  switching.push_back(tf);
  auto add = [&](double t, double hh, double tnext) {
    RolloutStep s{};
    s.h = hh;
    const double cs[4] = {0.0, 0.5, 0.5, 1.0};
    for (int c = 0; c < 4; ++c) time_segment(t + hh * cs[c], h->time, s.idx[c], s.alpha[c]);
    time_segment(tnext, h->time, s.obs_idx, s.obs_alpha);
    steps.push_back(s);
    h->ro_times.push_back(tnext);
  };
  const int intervals = (int)switching.size() - 1;
  for (int iv = 0; iv < intervals; ++iv) {
    const double tEnd = switching[iv + 1];
    const double tBegin = std::min(switching[iv] + 1e-9, tEnd);
    if (iv == 0) {
      time_segment(tBegin, h->time, h->ro_first_idx, h->ro_first_alpha);
      h->ro_times.push_back(tBegin);
    } else {  // the jump at the end of the previous interval, observed at the start of this one (TimeTriggeredRollout.cpp:104-108)
      RolloutStep s{};
      s.jump = iv;
      s.pre_node = event_node[iv - 1];
      time_segment(tBegin, h->time, s.obs_idx, s.obs_alpha);
      for (int c = 0; c < 4; ++c) s.idx[c] = s.obs_idx, s.alpha[c] = s.obs_alpha;
      steps.push_back(s);
      h->ro_times.push_back(tBegin);
    }
    if (tBegin < tEnd) {
      double t = tBegin;
      int step = 0;
      while (less_eq_with_sign(t + dt, tEnd)) {
        ++step;
        const double tn = tBegin + (double)step * dt;
        add(t, dt, tn);
        t = tn;
        if (steps.size() > 50000000) return fail(O2C_ERR_INVALID_ARGUMENT, "rollout schedule too long (time_step too small)");
      }
      const double end = tBegin + dt * (double)step;
      if (less_with_sign(end, tEnd)) add(end, tEnd - end, tEnd);
    }
  }
  return O2C_OK;
}

// (re)builds the SLQ backward and rollout step schedules from the node times and the event nodes of the handle
o2c_error rebuild_schedules(o2c_handle* h) {
  if (h->st.algorithm == O2C_ALG_SLQ) {
    for (auto& lane : h->lanes) O2C_CUDA(cudaStreamSynchronize(lane.stream));  // kernels in flight still read the old schedules
    std::vector<SlqStep> steps;
    o2c_error e = build_slq_schedule(h, steps);
    if (e != O2C_OK) return e;
    if (h->d_slq_steps) cudaFree(h->d_slq_steps);
    h->d_slq_steps = nullptr;
    h->n_slq_steps = (int)steps.size();
    if (!steps.empty()) {
      O2C_CUDA(cudaMalloc(&h->d_slq_steps, sizeof(SlqStep) * steps.size()));
      O2C_CUDA(cudaMemcpy(h->d_slq_steps, steps.data(), sizeof(SlqStep) * steps.size(), cudaMemcpyHostToDevice));
    }
    std::vector<RolloutStep> rs;
    e = build_rollout_schedule(h, rs);
    if (e != O2C_OK) return e;
    const int out_nodes = (int)rs.size() + 1;
    if (h->d_ro_steps) cudaFree(h->d_ro_steps);
    h->d_ro_steps = nullptr;
    h->n_ro_steps = (int)rs.size();
    if (!rs.empty()) {
      O2C_CUDA(cudaMalloc(&h->d_ro_steps, sizeof(RolloutStep) * rs.size()));
      O2C_CUDA(cudaMemcpy(h->d_ro_steps, rs.data(), sizeof(RolloutStep) * rs.size(), cudaMemcpyHostToDevice));
    }
    if (out_nodes != h->out_nodes) {
      if (h->d_xs) cudaFree(h->d_xs);
      if (h->d_us) cudaFree(h->d_us);
      h->d_xs = h->d_us = nullptr;
      h->out_nodes = out_nodes;
      const size_t tasks = (size_t)h->cfg.max_alphas * h->cfg.batch * out_nodes;
      O2C_CUDA(cudaMalloc(&h->d_xs, sizeof(double) * tasks * h->L.n));
      O2C_CUDA(cudaMalloc(&h->d_us, sizeof(double) * tasks * h->L.m));
    }
  } else {
    h->ro_times = h->time;
  }
  return O2C_OK;
}

o2c_error install_time(o2c_handle* h, const double* host_time) {
  const int N = h->L.N;
  h->time.assign(host_time, host_time + N + 1);
  for (int k = 0; k < N; ++k)
    if (!(h->time[k + 1] >= h->time[k])) return fail(O2C_ERR_INVALID_ARGUMENT, "time nodes must be non-decreasing");
  O2C_CUDA(cudaMemcpyAsync(h->d_time, h->time.data(), sizeof(double) * (N + 1), cudaMemcpyHostToDevice, h->lanes[0].stream));
  O2C_CUDA(cudaStreamSynchronize(h->lanes[0].stream));
  o2c_error e = rebuild_schedules(h);
  if (e != O2C_OK) return e;
  h->time_set = true;
  return O2C_OK;
}

struct FieldSpec {
  const o2c_field* f;
  int block;
  int nodes;  // 1 for per-problem fields
};

// size of the dense staging area (doubles) one problem needs on the way in / out
size_t stage_in_per_problem(const Layout& L, bool nominal) {
  size_t t = (size_t)L.nodes * L.rec + L.trec + L.n;
  if (nominal) t += (size_t)(L.N + 1) * (L.n + L.m);
  t += (size_t)(L.nodes + 1) / 2 + 2;  // int32 nc per node
  return t + 64;
}
size_t stage_out_per_problem(const Layout& L, int out_nodes, int n_alpha) {
  return (size_t)(L.N + 1) * L.orec + (size_t)n_alpha * out_nodes * (L.n + L.m) + 64;
}

o2c_error ensure_stage(Lane& lane, size_t in_doubles, size_t out_doubles) {
  if (in_doubles > lane.stage_in_doubles) {
    if (lane.stage_in) cudaFree(lane.stage_in);
    lane.stage_in = nullptr;
    lane.stage_in_doubles = 0;
    O2C_CUDA(cudaMalloc(&lane.stage_in, in_doubles * sizeof(double)));
    lane.stage_in_doubles = in_doubles;
  }
  if (out_doubles > lane.stage_out_doubles) {
    if (lane.stage_out) cudaFree(lane.stage_out);
    lane.stage_out = nullptr;
    lane.stage_out_doubles = 0;
    O2C_CUDA(cudaMalloc(&lane.stage_out, out_doubles * sizeof(double)));
    lane.stage_out_doubles = out_doubles;
  }
  return O2C_OK;
}

// copies one host field (count problems x nodes blocks) into dense device staging and returns the device-side strided view
o2c_error field_h2d(Lane& lane, const o2c_field& f, int block, int nodes, int count, double*& cursor, FieldDev& out) {
  out = FieldDev{nullptr, 0, 0};
  if (f.ptr == nullptr || count == 0) return O2C_OK;
  double* dst = cursor;
  const size_t total = (size_t)count * nodes * block;
  cursor += (total + 1) & ~(size_t)1;
  const long long ps = f.problem_stride, ns = (nodes > 1) ? f.node_stride : block;
  if (ns == block && (count == 1 || ps == (long long)nodes * block)) {
    O2C_CUDA(cudaMemcpyAsync(dst, f.ptr, total * sizeof(double), cudaMemcpyHostToDevice, lane.stream));
    out = FieldDev{dst, (long long)nodes * block, block};
  } else if (ns == block && ps >= (long long)nodes * block) {
    O2C_CUDA(cudaMemcpy2DAsync(dst, (size_t)nodes * block * sizeof(double), f.ptr, (size_t)ps * sizeof(double),
                               (size_t)nodes * block * sizeof(double), count, cudaMemcpyHostToDevice, lane.stream));
    out = FieldDev{dst, (long long)nodes * block, block};
  } else if (ps == block && ns >= (long long)count * block) {
    // [node][problem] order: rows are nodes
    O2C_CUDA(cudaMemcpy2DAsync(dst, (size_t)count * block * sizeof(double), f.ptr, (size_t)ns * sizeof(double),
                               (size_t)count * block * sizeof(double), nodes, cudaMemcpyHostToDevice, lane.stream));
    out = FieldDev{dst, block, (long long)count * block};
  } else {
    // irregular strides: gather on the host (synchronous)
    lane.bounce.resize(total);
    for (int p = 0; p < count; ++p)
      for (int k = 0; k < nodes; ++k)
        std::memcpy(lane.bounce.data() + ((size_t)p * nodes + k) * block, f.ptr + p * ps + k * ns, sizeof(double) * block);
    O2C_CUDA(cudaMemcpyAsync(dst, lane.bounce.data(), total * sizeof(double), cudaMemcpyHostToDevice, lane.stream));
    O2C_CUDA(cudaStreamSynchronize(lane.stream));
    out = FieldDev{dst, (long long)nodes * block, block};
  }
  return O2C_OK;
}

o2c_error field_d2h(Lane& lane, const o2c_field& f, int block, int nodes, int count, const double* src) {
  if (f.ptr == nullptr || count == 0) return O2C_OK;
  const size_t total = (size_t)count * nodes * block;
  const long long ps = f.problem_stride, ns = (nodes > 1) ? f.node_stride : block;
  if (ns == block && (count == 1 || ps == (long long)nodes * block)) {
    O2C_CUDA(cudaMemcpyAsync(f.ptr, src, total * sizeof(double), cudaMemcpyDeviceToHost, lane.stream));
  } else if (ns == block && ps >= (long long)nodes * block) {
    O2C_CUDA(cudaMemcpy2DAsync(f.ptr, (size_t)ps * sizeof(double), src, (size_t)nodes * block * sizeof(double),
                               (size_t)nodes * block * sizeof(double), count, cudaMemcpyDeviceToHost, lane.stream));
  } else {
    lane.bounce.resize(total);
    O2C_CUDA(cudaMemcpyAsync(lane.bounce.data(), src, total * sizeof(double), cudaMemcpyDeviceToHost, lane.stream));
    O2C_CUDA(cudaStreamSynchronize(lane.stream));
    for (int p = 0; p < count; ++p)
      for (int k = 0; k < nodes; ++k)
        std::memcpy(f.ptr + p * ps + k * ns, lane.bounce.data() + ((size_t)p * nodes + k) * block, sizeof(double) * block);
  }
  return O2C_OK;
}

o2c_field offset_field(const o2c_field& f, long long problems) {
  o2c_field g = f;
  if (g.ptr) g.ptr += problems * f.problem_stride;
  return g;
}

o2c_error check_range(const o2c_handle* h, int begin, int count) {
  if (!h) return fail(O2C_ERR_INVALID_ARGUMENT, "null handle");
  if (begin < 0 || count < 0 || (long long)begin + count > h->cfg.batch) return fail(O2C_ERR_INVALID_ARGUMENT, "problem range outside the batch");
  return O2C_OK;
}

// pre-event flags of problems [begin, begin+count): host (or device) array -> d_event; a view without events clears the range.
// SLQ: the flags are shared by the batch (they shape the step schedules) and every event brings its jump record.
o2c_error install_events(o2c_handle* h, cudaStream_t stream, const o2c_lq_view* v, bool device_memory, int begin, int count) {
  const int nodes = h->L.nodes, n = h->L.n;
  const bool slq = h->st.algorithm == O2C_ALG_SLQ;
  const bool whole = begin == 0 && count == h->cfg.batch;
  const int32_t* ev = v ? v->event : nullptr;
  if (ev == nullptr || count == 0) {
    if (count == 0) return O2C_OK;
    if (h->events_present) O2C_CUDA(cudaMemsetAsync(h->d_event + (size_t)begin * nodes, 0, sizeof(int) * (size_t)count * nodes, stream));
    if (whole) {  // whole batch replaced: the specialised kernels serve it again
      h->events_present = false;
      if (slq && !h->slq_events.empty()) {
        h->slq_events.clear();
        if (h->time_set) return rebuild_schedules(h);
      }
    } else if (slq && !h->slq_events.empty()) {
      return fail(O2C_ERR_INVALID_ARGUMENT, "SLQ event nodes are shared by the batch: a partial upload must carry the same events");
    }
    return O2C_OK;
  }
  const long long ps = v->event_problem_stride, ns = v->event_node_stride;
  if (slq && device_memory) return fail(O2C_ERR_UNSUPPORTED, "o2c_import_device does not take SLQ events (upload them from host memory)");
  if (!h->d_event) {
    O2C_CUDA(cudaMalloc(&h->d_event, sizeof(int) * (size_t)h->cfg.batch * nodes));
    O2C_CUDA(cudaMemsetAsync(h->d_event, 0, sizeof(int) * (size_t)h->cfg.batch * nodes, stream));
  }
  int* dst = h->d_event + (size_t)begin * nodes;
  if (device_memory) {
    if (ns != 1) return fail(O2C_ERR_UNSUPPORTED, "device event flags must be contiguous over the nodes (event_node_stride == 1)");
    h->events_present = true;
    O2C_CUDA(cudaMemcpy2DAsync(dst, sizeof(int) * nodes, ev, sizeof(int) * (size_t)ps, sizeof(int) * nodes, count, cudaMemcpyDeviceToDevice, stream));
    return O2C_OK;
  }
  std::vector<int> tmp((size_t)count * nodes);
  for (int p = 0; p < count; ++p)
    for (int k = 0; k < nodes; ++k) tmp[(size_t)p * nodes + k] = ev[p * ps + k * ns] != 0;
  if (std::find(tmp.begin(), tmp.end(), 1) == tmp.end())  // an all-zero flag array is "no events": keep the event-free kernels
    return install_events(h, stream, nullptr, false, begin, count);
  if (slq) {
    std::vector<int> flags(tmp.begin(), tmp.begin() + nodes);
    for (int p = 1; p < count; ++p)
      if (!std::equal(flags.begin(), flags.end(), tmp.begin() + (size_t)p * nodes))
        return fail(O2C_ERR_INVALID_ARGUMENT, "SLQ event nodes must be the same for every problem (the time grid is shared by the batch)");
    if (flags[nodes - 1]) return fail(O2C_ERR_INVALID_ARGUMENT, "the last node cannot be a pre-event node");
    int n_ev = 0;
    for (int f : flags) n_ev += f;
    if (!whole && flags != h->slq_events)
      return fail(O2C_ERR_INVALID_ARGUMENT, "SLQ event nodes are shared by the batch: a partial upload must carry the same events");
    if (n_ev > 0 && (!v->jump_A.ptr || !v->jump_Q.ptr)) return fail(O2C_ERR_INVALID_ARGUMENT, "jump_A and jump_Q are required with SLQ events");
    const size_t jrec = (size_t)jump_rec(n);
    if (n_ev > h->jump_capacity) {
      for (auto& lane : h->lanes) O2C_CUDA(cudaStreamSynchronize(lane.stream));
      if (h->d_jump) cudaFree(h->d_jump);
      h->d_jump = nullptr;
      h->jump_capacity = 0;
      O2C_CUDA(cudaMalloc(&h->d_jump, sizeof(double) * jrec * n_ev * h->cfg.batch));
      h->jump_capacity = n_ev;
    }
    if (n_ev > 0) {
      std::vector<double> jr((size_t)count * n_ev * jrec, 0.0);
      auto gather = [&](const o2c_field& f, int block, int offset) {
        if (!f.ptr) return;
        for (int p = 0; p < count; ++p)
          for (int e = 0; e < n_ev; ++e) {
            const double* src = f.ptr + (long long)p * f.problem_stride + (long long)e * f.node_stride;
            std::copy(src, src + block, jr.begin() + ((size_t)p * n_ev + e) * jrec + offset);
          }
      };
      gather(v->jump_A, n * n, 0);
      gather(v->jump_Hv, n, jump_oHv(n));
      gather(v->jump_Q, n * n, jump_oQ(n));
      gather(v->jump_q, n, jump_oq(n));
      gather(v->jump_c, 1, jump_oc(n));
      O2C_CUDA(cudaMemcpy2DAsync(h->d_jump + (size_t)begin * h->jump_capacity * jrec, sizeof(double) * h->jump_capacity * jrec, jr.data(),
                                 sizeof(double) * n_ev * jrec, sizeof(double) * n_ev * jrec, count, cudaMemcpyHostToDevice, stream));
      O2C_CUDA(cudaStreamSynchronize(stream));
    }
    if (flags != h->slq_events) {
      h->slq_events = flags;
      if (h->time_set) {
        o2c_error e = rebuild_schedules(h);
        if (e != O2C_OK) return e;
      }
    }
    if (n_ev == 0) h->slq_events.clear();
  }
  h->events_present = true;
  O2C_CUDA(cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  O2C_CUDA(cudaStreamSynchronize(stream));  // tmp is pageable and about to go out of scope
  return O2C_OK;
}

// Bookkeeping of an upload that replaces the WHOLE batch, done once before its chunks (a chunk alone cannot tell that the batch is
// being replaced): stale event flags and the "ragged constraint counts" mark of the previous contents are dropped, so that the
// shape-specialised kernels serve the new data again, and SLQ event nodes (they shape the step schedules) and their jump records are
// installed up front. The chunks then only add what they carry.
o2c_error begin_whole_upload(o2c_handle* h, const o2c_lq_view* lq) {
  cudaStream_t stream = h->lanes[0].stream;
  h->nc_ragged = false;
  const bool slq = h->st.algorithm == O2C_ALG_SLQ;
  if (!lq->event || !slq) {  // no events, or ILQR flags that the chunks install range by range: start from a clean slate
    o2c_error e = install_events(h, stream, nullptr, false, 0, h->cfg.batch);
    if (e != O2C_OK) return e;
  }
  if (lq->event && slq) return install_events(h, stream, lq, false, 0, h->cfg.batch);
  return O2C_OK;
}

// host view (already offset so that index 0 is the first problem of the chunk) -> records of problems [begin, begin+count)
o2c_error upload_chunk(o2c_handle* h, Lane& lane, const o2c_lq_view& v, int begin, int count) {
  const Layout& L = h->L;
  const bool nominal = h->cfg.has_nominal != 0;
  o2c_error e = ensure_stage(lane, stage_in_per_problem(L, nominal) * (size_t)count, 0);
  if (e != O2C_OK) return e;
  double* cur = lane.stage_in;
  LqViewDev d{};
  const int n = L.n, m = L.m, ncm = L.ncmax, nodes = L.nodes;
#define H2D(field, block, nn)                                                   \
  if ((e = field_h2d(lane, v.field, block, nn, count, cur, d.field)) != O2C_OK) return e;
  const bool packed = (v.flags & O2C_LQ_SYMMETRIC_PACKED) != 0;
  const int qblock = packed ? n * (n + 1) / 2 : n * n, rblock = packed ? m * (m + 1) / 2 : m * m;
  d.sym_packed = packed ? 1 : 0;
  H2D(A, n * n, nodes) H2D(B, n * m, nodes) H2D(Hv, n, nodes) H2D(Q, qblock, nodes) H2D(P, m * n, nodes) H2D(R, rblock, nodes)
  H2D(q, n, nodes) H2D(r, m, nodes) H2D(c, 1, nodes)
  if (ncm > 0) {
    H2D(C, ncm * n, nodes) H2D(D, ncm * m, nodes) H2D(e, ncm, nodes)
  }
  H2D(Qf, qblock, 1) H2D(qf, n, 1) H2D(cf, 1, 1) H2D(x0, n, 1)
  if (nominal) {
    H2D(x_nom, n, L.N + 1) H2D(u_nom, m, L.N + 1)
  }
#undef H2D
  int* nc_stage = nullptr;
  if (ncm > 0 && v.nc != nullptr) {
    // int32 per (problem, node): gathered densely into the lane's pinned staging, copied after the doubles without blocking the host
    const size_t cnt = (size_t)count * nodes;
    if (lane.nc_copied) O2C_CUDA(cudaEventSynchronize(lane.nc_copied));  // the previous chunk of this lane has left the staging
    if (cnt > lane.h_nc_count) {
      if (lane.h_nc) cudaFreeHost(lane.h_nc);
      lane.h_nc = nullptr;
      lane.h_nc_count = 0;
      O2C_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&lane.h_nc), cnt * sizeof(int), cudaHostAllocDefault));
      lane.h_nc_count = cnt;
    }
    if (!lane.nc_copied) O2C_CUDA(cudaEventCreateWithFlags(&lane.nc_copied, cudaEventDisableTiming));
    bool ragged = false;
    for (int p = 0; p < count; ++p)
      for (int k = 0; k < nodes; ++k) {
        const int c = v.nc[p * v.nc_problem_stride + k * v.nc_node_stride];
        if (c < 0 || c > ncm) return fail(O2C_ERR_INVALID_ARGUMENT, "nc[problem][node] must lie in [0, nc_max]");
        ragged = ragged || c != ncm;
        lane.h_nc[(size_t)p * nodes + k] = c;
      }
    nc_stage = reinterpret_cast<int*>(cur);
    O2C_CUDA(cudaMemcpyAsync(nc_stage, lane.h_nc, cnt * sizeof(int), cudaMemcpyHostToDevice, lane.stream));
    O2C_CUDA(cudaEventRecord(lane.nc_copied, lane.stream));
    d.nc = nc_stage;
    // counts that all equal nc_max keep the kernels that skip the lookup; a whole-batch upload resets the flag first (begin_whole_upload)
    h->nc_ragged = (begin == 0 && count == h->cfg.batch) ? ragged : (h->nc_ragged || ragged);
    d.nc_ps = nodes;
    d.nc_ns = 1;
  } else if (ncm > 0 && begin == 0 && count == h->cfg.batch) {
    h->nc_ragged = false;  // pack_kernel rewrites every count to nc_max
  }
  if ((e = install_events(h, lane.stream, &v, false, begin, count)) != O2C_OK) return e;
  O2C_CUDA(launch_pack(L, d, h->d_lq, h->d_term, h->d_xnom, h->d_unom, h->d_nc, h->d_x0, begin, count, lane.stream));
  h->launches += 1;
  return O2C_OK;
}

o2c_error download_chunk(o2c_handle* h, Lane& lane, const o2c_solution_view& v, int begin, int count, int n_alpha) {
  const Layout& L = h->L;
  const int n = L.n, m = L.m, N = L.N, on = h->out_nodes;
  o2c_error e = ensure_stage(lane, 0, stage_out_per_problem(L, on, std::max(n_alpha, 1)) * (size_t)count);
  if (e != O2C_OK) return e;
  double* cur = lane.stage_out;
  SolViewDev d{};
  auto take = [&](const o2c_field& f, int block, int nodes, FieldDev& out) {
    out = FieldDev{nullptr, 0, 0};
    if (!f.ptr) return;
    out = FieldDev{cur, (long long)nodes * block, block};
    cur += ((size_t)count * nodes * block + 1) & ~(size_t)1;
  };
  take(v.K, m * n, N + 1, d.K);
  take(v.dbias, m, N + 1, d.dbias);
  take(v.bias, m, N + 1, d.bias);
  take(v.Sm, n * n, N + 1, d.Sm);
  take(v.Sv, n, N + 1, d.Sv);
  take(v.s, 1, N + 1, d.s);
  double* xbase = nullptr;
  double* ubase = nullptr;
  if (v.x.ptr && n_alpha > 0) {
    xbase = cur;
    d.x = FieldDev{cur, (long long)on * n, n};
    d.x_as = (long long)count * on * n;
    cur += (size_t)n_alpha * count * on * n;
  }
  if (v.u.ptr && n_alpha > 0) {
    ubase = cur;
    d.u = FieldDev{cur, (long long)on * m, m};
    d.u_as = (long long)count * on * m;
    cur += (size_t)n_alpha * count * on * m;
  }
  cur = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(cur) + 15) & ~(uintptr_t)15);
  int* status_stage = nullptr;
  if (v.status) {
    status_stage = reinterpret_cast<int*>(cur);
    d.status = status_stage;
  }
  O2C_CUDA(launch_unpack(L, d, h->d_sol, h->d_xs, h->d_us, h->d_status, on, (xbase || ubase) ? n_alpha : 0, h->cfg.batch, begin, count,
                         lane.stream));
  h->launches += 1;
#define D2H(field, block, nn) \
  if ((e = field_d2h(lane, v.field, block, nn, count, d.field.ptr)) != O2C_OK) return e;
  D2H(K, m * n, N + 1) D2H(dbias, m, N + 1) D2H(bias, m, N + 1) D2H(Sm, n * n, N + 1) D2H(Sv, n, N + 1) D2H(s, 1, N + 1)
#undef D2H
  for (int a = 0; a < n_alpha; ++a) {
    if (xbase) {
      o2c_field fx = v.x;
      fx.ptr += (long long)a * v.x_alpha_stride;
      if ((e = field_d2h(lane, fx, n, on, count, xbase + (size_t)a * count * on * n)) != O2C_OK) return e;
    }
    if (ubase) {
      o2c_field fu = v.u;
      fu.ptr += (long long)a * v.u_alpha_stride;
      if ((e = field_d2h(lane, fu, m, on, count, ubase + (size_t)a * count * on * m)) != O2C_OK) return e;
    }
  }
  if (v.status) {
    if (!h->h_status) O2C_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_status), sizeof(int) * (size_t)h->cfg.batch, cudaHostAllocDefault));
    O2C_CUDA(cudaMemcpyAsync(h->h_status + begin, status_stage, sizeof(int) * count, cudaMemcpyDeviceToHost, lane.stream));
    h->pending_status.push_back({v.status, begin, count});  // handed to the caller by flush_status once the lanes are synchronised
  }
  return O2C_OK;
}

void flush_status(o2c_handle* h) {
  for (const auto& p : h->pending_status) std::memcpy(p.dst, h->h_status + p.begin, sizeof(int) * (size_t)p.count);
  h->pending_status.clear();
}

o2c_error backward_on(o2c_handle* h, cudaStream_t stream, int begin, int count) {
  if (count == 0) return O2C_OK;
  if (h->st.algorithm == O2C_ALG_ILQR && h->events_present && h->st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT)
    return fail(O2C_ERR_UNSUPPORTED,
                "ILQR events under LEVENBERG_MARQUARDT: deltaGm / deltaGv of a pre-event node need the node's regular dynamics "
                "(ILQR.cpp:263-295), which the record layout replaces by the jump map");
  h->backward_done = true;
  const DeviceBuffers buf = h->buffers(stream);
  if (h->st.algorithm == O2C_ALG_ILQR) {
    if (h->use_fast && wpp_ilqr_supported(h->L, h->st, buf)) {  // (handles with events fall through to the generic kernel)
      int l = 0;
      O2C_CUDA(launch_ilqr_wpp(h->L, h->st, buf, false, 1.0, h->cfg.batch, begin, count, stream, &l));
      h->launches += l;
    } else if (h->use_rpl && rpl_ilqr_supported(h->L, h->st, buf)) {  // (ragged constraint counts fall through to the generic kernel)
      int l = 0;
      O2C_CUDA(launch_ilqr_rpl(h->L, h->st, buf, false, 1.0, begin, count, stream, &l));
      h->launches += l;
    } else {
      O2C_CUDA(launch_ilqr_generic(h->L, h->st, buf, begin, count, stream));
      h->launches += 1;
    }
  } else {
    if (!h->time_set) return fail(O2C_ERR_NOT_READY, "SLQ needs the node times (o2c_set_time or lq_view.time) before o2c_backward");
    if (slq_wpp_supported(h->L, h->st, buf)) {
      if (!h->d_slq_ws) O2C_CUDA(cudaMalloc(&h->d_slq_ws, sizeof(double) * slq_wpp_workspace_doubles(h->L, h->cfg.batch)));
      int l = 0;
      O2C_CUDA(launch_slq_wpp(h->L, h->st, buf, h->d_slq_ws, h->d_slq_steps, h->n_slq_steps, begin, count, stream, &l));
      h->launches += l;
      return O2C_OK;
    }
    if (rpl_slq_supported(h->L, h->st, buf))
      O2C_CUDA(launch_slq_rpl(h->L, h->st, buf, h->d_slq_steps, h->n_slq_steps, begin, count, stream));
    else
      O2C_CUDA(launch_slq_generic(h->L, h->st, buf, h->d_slq_steps, h->n_slq_steps, begin, count, stream));
    h->launches += 1;
  }
  return O2C_OK;
}

o2c_error rollout_on(o2c_handle* h, cudaStream_t stream, const double* alphas_dev, int n_alpha, int begin, int count) {
  if (count == 0 || n_alpha == 0) return O2C_OK;
  const DeviceBuffers buf = h->buffers();
  if (h->st.algorithm == O2C_ALG_ILQR) {
    O2C_CUDA(launch_rollout_discrete(h->L, buf, alphas_dev, n_alpha, h->cfg.batch, begin, count, stream));
  } else {
    if (!h->time_set) return fail(O2C_ERR_NOT_READY, "continuous rollout needs the node times");
    if (rollout_cont24_supported(h->L, h->st, buf))
      O2C_CUDA(launch_rollout_cont24(h->L, h->st, buf, h->d_ro_steps, h->n_ro_steps, h->ro_first_idx, h->ro_first_alpha, h->out_nodes, alphas_dev,
                                     n_alpha, h->cfg.batch, begin, count, stream));
    else if (rpl_rollout_cont_supported(h->L, h->st, buf))
      O2C_CUDA(launch_rollout_cont_rpl(h->L, h->st, buf, h->d_ro_steps, h->n_ro_steps, h->ro_first_idx, h->ro_first_alpha, h->out_nodes,
                                       alphas_dev, n_alpha, h->cfg.batch, begin, count, stream));
    else
      O2C_CUDA(launch_rollout_continuous(h->L, buf, h->d_ro_steps, h->n_ro_steps, h->ro_first_idx, h->ro_first_alpha, h->out_nodes, alphas_dev,
                                         n_alpha, h->cfg.batch, begin, count, stream));
  }
  h->launches += 1;
  return O2C_OK;
}

o2c_error solve_on(o2c_handle* h, cudaStream_t stream, double* alpha_slot_dev, double alpha, int begin, int count) {
  if (count == 0) return O2C_OK;
  if (h->st.algorithm == O2C_ALG_ILQR && ((h->use_fast && wpp_ilqr_supported(h->L, h->st, h->buffers())) ||
                                          (h->use_rpl && rpl_ilqr_supported(h->L, h->st, h->buffers())))) {
    int l = 0;
    h->backward_done = true;
    if (h->use_fast)
      O2C_CUDA(launch_ilqr_wpp(h->L, h->st, h->buffers(stream), true, alpha, h->cfg.batch, begin, count, stream, &l));
    else
      O2C_CUDA(launch_ilqr_rpl(h->L, h->st, h->buffers(stream), true, alpha, begin, count, stream, &l));
    h->launches += l;
    return O2C_OK;
  }
  o2c_error e = backward_on(h, stream, begin, count);
  if (e != O2C_OK) return e;
  O2C_CUDA(cudaMemcpyAsync(alpha_slot_dev, &alpha, sizeof(double), cudaMemcpyHostToDevice, stream));
  return rollout_on(h, stream, alpha_slot_dev, 1, begin, count);
}

void release(o2c_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (auto& lane : h->lanes) {
    if (lane.stream) cudaStreamSynchronize(lane.stream);
  }
  void* ptrs[] = {h->d_lq,  h->d_term, h->d_xnom, h->d_unom,   h->d_x0,        h->d_time,    h->d_sol,
                  h->d_xs,  h->d_us,   h->d_alphas, h->d_nc,   h->d_status,    h->d_slq_steps, h->d_ro_steps,
                  h->d_ls_merit, h->d_ls_base, h->d_ls_is, h->d_ls_step, h->d_ls_basein, h->d_ls_index, h->d_event, h->d_flat, h->d_jump, h->d_dt, h->d_counter, h->d_slq_ws};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_status) cudaFreeHost(h->h_status);
  for (auto& lane : h->lanes) {
    if (lane.stage_in) cudaFree(lane.stage_in);
    if (lane.stage_out) cudaFree(lane.stage_out);
    if (lane.h_nc) cudaFreeHost(lane.h_nc);
    if (lane.nc_copied) cudaEventDestroy(lane.nc_copied);
    if (lane.stream) cudaStreamDestroy(lane.stream);
  }
  delete h;
}

}  // namespace

extern "C" {

int o2c_abi_version(void) { return O2C_ABI_VERSION; }
const char* o2c_last_error(void) { return g_last_error.c_str(); }

o2c_error o2c_create(const o2c_config* cfg, o2c_handle** out) {
  if (!cfg || !out) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (cfg->nx < 1 || cfg->nu < 1 || cfg->nx > 64 || cfg->nu > 64) return fail(O2C_ERR_INVALID_ARGUMENT, "nx, nu must be in [1, 64]");
  if (cfg->nc_max < 0 || cfg->nc_max > cfg->nu) return fail(O2C_ERR_INVALID_ARGUMENT, "nc_max must be in [0, nu]");
  if (cfg->nc_max > 32) return fail(O2C_ERR_UNSUPPORTED, "nc_max > 32 is not supported");
  if (cfg->num_stages < 1 || cfg->num_stages > 1022) return fail(O2C_ERR_INVALID_ARGUMENT, "num_stages must be in [1, 1022]");
  if (cfg->batch < 1) return fail(O2C_ERR_INVALID_ARGUMENT, "batch must be positive");
  if (cfg->algorithm != O2C_ALG_ILQR && cfg->algorithm != O2C_ALG_SLQ) return fail(O2C_ERR_INVALID_ARGUMENT, "unknown algorithm");
  if (cfg->strategy != O2C_STRATEGY_LINE_SEARCH && cfg->strategy != O2C_STRATEGY_LEVENBERG_MARQUARDT)
    return fail(O2C_ERR_INVALID_ARGUMENT, "unknown strategy");
  if (cfg->riccati_form != O2C_FORM_FULL && cfg->riccati_form != O2C_FORM_REDUCED) return fail(O2C_ERR_INVALID_ARGUMENT, "unknown riccati_form");
  if (cfg->strategy == O2C_STRATEGY_LINE_SEARCH && cfg->hessian_correction != O2C_HC_DIAGONAL_SHIFT &&
      cfg->hessian_correction != O2C_HC_GERSHGORIN_MODIFICATION && cfg->hessian_correction != O2C_HC_EIGENVALUE_MODIFICATION)
    return fail(O2C_ERR_UNSUPPORTED, "hessian_correction must be DIAGONAL_SHIFT, GERSHGORIN_MODIFICATION or EIGENVALUE_MODIFICATION "
                                     "(CHOLESKY_MODIFICATION is Eigen::IncompleteCholesky in the reference and is not provided)");
  if (cfg->riccati_form == O2C_FORM_REDUCED && cfg->strategy != O2C_STRATEGY_LINE_SEARCH)
    return fail(O2C_ERR_INVALID_ARGUMENT, "the reduced Riccati form is only valid with LINE_SEARCH (ILQR.cpp:68, SLQ.cpp:65)");
  if (cfg->max_alphas < 1) return fail(O2C_ERR_INVALID_ARGUMENT, "max_alphas must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(O2C_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(O2C_ERR_INVALID_ARGUMENT, "device ordinal out of range");
  O2C_CUDA(cudaSetDevice(cfg->device));
  o2c_handle* h = new o2c_handle();
  h->cfg = *cfg;
  h->L = make_layout(cfg->nx, cfg->nu, cfg->nc_max, cfg->num_stages, cfg->algorithm);
  h->st = SolverSettings{cfg->algorithm, cfg->riccati_form == O2C_FORM_REDUCED ? 1 : 0, cfg->strategy, cfg->hessian_correction,
                         cfg->hessian_multiple, cfg->lm_riccati_multiple, cfg->time_step};
  const Layout& L = h->L;
  const size_t B = (size_t)cfg->batch;
  auto cleanup_fail = [&](o2c_error code) {
    release(h);
    return code;
  };
#define ALLOC(ptr, count)                                                                                      \
  do {                                                                                                         \
    cudaError_t _e = cudaMalloc(&(ptr), (count));                                                              \
    if (_e != cudaSuccess) {                                                                                   \
      cudaGetLastError();                                                                                      \
      fail(O2C_ERR_OUT_OF_MEMORY, std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(_e));            \
      return cleanup_fail(O2C_ERR_OUT_OF_MEMORY);                                                              \
    }                                                                                                          \
  } while (0)
  for (auto& lane : h->lanes) {
    if (cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking) != cudaSuccess) {
      fail(O2C_ERR_CUDA, "cudaStreamCreate failed");
      return cleanup_fail(O2C_ERR_CUDA);
    }
  }
  h->out_nodes = L.N + 1;
  ALLOC(h->d_lq, sizeof(double) * B * L.nodes * L.rec);
  ALLOC(h->d_term, sizeof(double) * B * L.trec);
  ALLOC(h->d_x0, sizeof(double) * B * L.n);
  ALLOC(h->d_time, sizeof(double) * (L.N + 1));
  ALLOC(h->d_sol, sizeof(double) * B * (L.N + 1) * L.orec);
  ALLOC(h->d_xs, sizeof(double) * (size_t)cfg->max_alphas * B * h->out_nodes * L.n);
  ALLOC(h->d_us, sizeof(double) * (size_t)cfg->max_alphas * B * h->out_nodes * L.m);
  ALLOC(h->d_alphas, sizeof(double) * (cfg->max_alphas + kLanes));
  ALLOC(h->d_status, sizeof(int) * B);
  ALLOC(h->d_counter, sizeof(int) * kLanes);
  if (cfg->has_nominal) {
    ALLOC(h->d_xnom, sizeof(double) * B * (L.N + 1) * L.n);
    ALLOC(h->d_unom, sizeof(double) * B * (L.N + 1) * L.m);
    cudaMemsetAsync(h->d_xnom, 0, sizeof(double) * B * (L.N + 1) * L.n, h->lanes[0].stream);
    cudaMemsetAsync(h->d_unom, 0, sizeof(double) * B * (L.N + 1) * L.m, h->lanes[0].stream);
  }
  if (cfg->nc_max > 0) {
    ALLOC(h->d_nc, sizeof(int) * B * L.nodes);
    const size_t cnt = B * L.nodes;
    fill_int_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->lanes[0].stream>>>(h->d_nc, cfg->nc_max, cnt);
  }
#undef ALLOC
  cudaMemsetAsync(h->d_status, 0, sizeof(int) * B, h->lanes[0].stream);
  cudaMemsetAsync(h->d_x0, 0, sizeof(double) * B * L.n, h->lanes[0].stream);
  h->use_fast = wpp_ilqr_supported(h->L, h->st, h->buffers());
  h->use_rpl = !h->use_fast && rpl_ilqr_supported(h->L, h->st, h->buffers());
  // default uniform time grid t_k = k * time_step (ILQR does not need it; SLQ callers normally override it)
  std::vector<double> t(L.N + 1);
  const double dt = cfg->time_step > 0.0 ? cfg->time_step : 1.0;
  for (int k = 0; k <= L.N; ++k) t[k] = dt * (double)k;
  o2c_error e = install_time(h, t.data());
  if (e != O2C_OK) return cleanup_fail(e);
  if (cudaStreamSynchronize(h->lanes[0].stream) != cudaSuccess) {
    fail(O2C_ERR_CUDA, "initialisation failed");
    return cleanup_fail(O2C_ERR_CUDA);
  }
  *out = h;
  return O2C_OK;
}

o2c_error o2c_destroy(o2c_handle* h) {
  release(h);
  return O2C_OK;
}

o2c_error o2c_get_config(const o2c_handle* h, o2c_config* cfg) {
  if (!h || !cfg) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  *cfg = h->cfg;
  return O2C_OK;
}

o2c_error o2c_device_count(int32_t* count) {
  if (!count) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    *count = 0;
    return fail(O2C_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  *count = n;
  return O2C_OK;
}

o2c_error o2c_set_lm_riccati_multiple(o2c_handle* h, double mu) {
  if (!h) return fail(O2C_ERR_INVALID_ARGUMENT, "null handle");
  if (!(mu >= 0.0) || !std::isfinite(mu)) return fail(O2C_ERR_INVALID_ARGUMENT, "riccati_multiple must be finite and non-negative");
  h->cfg.lm_riccati_multiple = mu;
  h->st.mu = mu;  // kernel parameter blocks are built per launch: in-flight launches keep the value they were enqueued with
  return O2C_OK;
}

o2c_error o2c_sync(o2c_handle* h) {
  if (!h) return fail(O2C_ERR_INVALID_ARGUMENT, "null handle");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  for (auto& lane : h->lanes) O2C_CUDA(cudaStreamSynchronize(lane.stream));
  return O2C_OK;
}

o2c_error o2c_compute_stream(o2c_handle* h, void** stream) {
  if (!h || !stream) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  *stream = (void*)h->lanes[0].stream;
  return O2C_OK;
}

o2c_error o2c_device_lq_view(o2c_handle* h, o2c_lq_view* v) {
  if (!h || !v) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  const Layout& L = h->L;
  std::memset(v, 0, sizeof(*v));
  const int64_t ps = (int64_t)L.nodes * L.rec, ns = L.rec;
  auto rec = [&](int off) { return o2c_field{h->d_lq + off, ps, ns}; };
  v->A = rec(L.oA);
  v->B = rec(L.oB);
  v->Hv = rec(L.oHv);
  v->Q = rec(L.oQ);
  v->P = rec(L.oP);
  v->R = rec(L.oR);
  v->q = rec(L.oq);
  v->r = rec(L.or_);
  v->c = rec(L.oc);
  if (L.ncmax > 0) {
    v->C = rec(L.oC);
    v->D = rec(L.oD);
    v->e = rec(L.oe);
    v->nc = h->d_nc;
    h->nc_ragged = true;  // a device-side producer may write per-node counts through this view
    v->nc_problem_stride = L.nodes;
    v->nc_node_stride = 1;
  }
  v->Qf = o2c_field{h->d_term + L.oQf, L.trec, 0};
  v->qf = o2c_field{h->d_term + L.oqf, L.trec, 0};
  v->cf = o2c_field{h->d_term + L.ocf, L.trec, 0};
  if (h->d_xnom) v->x_nom = o2c_field{h->d_xnom, (int64_t)(L.N + 1) * L.n, L.n};
  if (h->d_unom) v->u_nom = o2c_field{h->d_unom, (int64_t)(L.N + 1) * L.m, L.m};
  v->x0 = o2c_field{h->d_x0, L.n, 0};
  v->time = h->d_time;
  return O2C_OK;
}

o2c_error o2c_device_solution_view(o2c_handle* h, o2c_solution_view* v) {
  if (!h || !v) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  const Layout& L = h->L;
  std::memset(v, 0, sizeof(*v));
  const int64_t ps = (int64_t)(L.N + 1) * L.orec, ns = L.orec;
  auto rec = [&](int off) { return o2c_field{h->d_sol + off, ps, ns}; };
  v->K = rec(L.oK);
  v->dbias = rec(L.odb);
  v->bias = rec(L.obias);
  v->Sm = rec(L.oSm);
  v->Sv = rec(L.oSv);
  v->s = rec(L.os);
  v->x = o2c_field{h->d_xs, (int64_t)h->out_nodes * L.n, L.n};
  v->u = o2c_field{h->d_us, (int64_t)h->out_nodes * L.m, L.m};
  v->x_alpha_stride = (int64_t)h->cfg.batch * h->out_nodes * L.n;
  v->u_alpha_stride = (int64_t)h->cfg.batch * h->out_nodes * L.m;
  v->status = h->d_status;
  return O2C_OK;
}

o2c_error o2c_rollout_num_nodes(o2c_handle* h, int32_t* out_nodes) {
  if (!h || !out_nodes) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  *out_nodes = h->out_nodes;
  return O2C_OK;
}

o2c_error o2c_rollout_times(o2c_handle* h, double* times) {
  if (!h || !times) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  std::copy(h->ro_times.begin(), h->ro_times.end(), times);
  return O2C_OK;
}

o2c_error o2c_set_time(o2c_handle* h, const double* host_time) {
  if (!h || !host_time) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  return install_time(h, host_time);
}

o2c_error o2c_upload(o2c_handle* h, const o2c_lq_view* v, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!v) return fail(O2C_ERR_INVALID_ARGUMENT, "null view");
  if (!v->A.ptr || !v->B.ptr || !v->Q.ptr || !v->R.ptr || !v->Qf.ptr)
    return fail(O2C_ERR_INVALID_ARGUMENT, "A, B, Q, R and Qf are required (absent Hv, P, q, r, c, qf, cf are taken as zero)");
  if (h->L.ncmax > 0 && (!v->C.ptr || !v->D.ptr || !v->e.ptr)) return fail(O2C_ERR_INVALID_ARGUMENT, "C, D, e are required when nc_max > 0");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  if (v->time) {
    e = install_time(h, v->time);
    if (e != O2C_OK) return e;
  }
  Lane& lane = h->lanes[0];
  const size_t per = stage_in_per_problem(h->L, h->cfg.has_nominal != 0) * sizeof(double);
  size_t chunk_bytes = (size_t)1 << 30;
  if (const char* env = getenv("O2C_UPLOAD_CHUNK_BYTES")) {  // test knob: split small uploads like the 38.8 GB legged batch is split
    const long long v = atoll(env);
    if (v > 0) chunk_bytes = (size_t)v;
  }
  int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)count, chunk_bytes / per));
  if (begin == 0 && count == h->cfg.batch && chunk < count)  // a whole-batch upload that is split: do the whole-range bookkeeping once
    if ((e = begin_whole_upload(h, v)) != O2C_OK) return e;
  for (int off = 0; off < count; off += chunk) {
    const int c = std::min(chunk, count - off);
    o2c_lq_view sub = *v;
#define OFF(field) sub.field = offset_field(v->field, off);
    OFF(A) OFF(B) OFF(Hv) OFF(Q) OFF(P) OFF(R) OFF(q) OFF(r) OFF(c) OFF(C) OFF(D) OFF(e) OFF(Qf) OFF(qf) OFF(cf) OFF(x_nom) OFF(u_nom) OFF(x0)
    OFF(jump_A) OFF(jump_Hv) OFF(jump_Q) OFF(jump_q) OFF(jump_c)
#undef OFF
    if (v->nc) sub.nc = v->nc + (long long)off * v->nc_problem_stride;
    if (v->event) sub.event = v->event + (long long)off * v->event_problem_stride;
    e = upload_chunk(h, lane, sub, begin + off, c);
    if (e != O2C_OK) return e;
  }
  return O2C_OK;
}

o2c_error o2c_import_device(o2c_handle* h, const o2c_lq_view* v, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!v) return fail(O2C_ERR_INVALID_ARGUMENT, "null view");
  if (!v->A.ptr || !v->B.ptr || !v->Q.ptr || !v->R.ptr || !v->Qf.ptr)
    return fail(O2C_ERR_INVALID_ARGUMENT, "A, B, Q, R and Qf are required (absent Hv, P, q, r, c, qf, cf are taken as zero)");
  if (h->L.ncmax > 0 && (!v->C.ptr || !v->D.ptr || !v->e.ptr)) return fail(O2C_ERR_INVALID_ARGUMENT, "C, D, e are required when nc_max > 0");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  if (v->time) {
    std::vector<double> t(h->L.N + 1);
    O2C_CUDA(cudaMemcpy(t.data(), v->time, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
    e = install_time(h, t.data());
    if (e != O2C_OK) return e;
  }
  LqViewDev d{};
  auto cv = [](const o2c_field& f) { return FieldDev{f.ptr, (long long)f.problem_stride, (long long)f.node_stride}; };
  d.A = cv(v->A);
  d.B = cv(v->B);
  d.Hv = cv(v->Hv);
  d.Q = cv(v->Q);
  d.P = cv(v->P);
  d.R = cv(v->R);
  d.q = cv(v->q);
  d.r = cv(v->r);
  d.c = cv(v->c);
  d.C = cv(v->C);
  d.D = cv(v->D);
  d.e = cv(v->e);
  d.Qf = cv(v->Qf);
  d.qf = cv(v->qf);
  d.cf = cv(v->cf);
  d.x_nom = cv(v->x_nom);
  d.u_nom = cv(v->u_nom);
  d.x0 = cv(v->x0);
  d.nc = v->nc;
  if (h->L.ncmax > 0) {  // device-side counts cannot be inspected here: any supplied array may be ragged
    if (v->nc) h->nc_ragged = true;
    else if (begin == 0 && count == h->cfg.batch) h->nc_ragged = false;  // pack_kernel rewrites every count to nc_max
  }
  d.nc_ps = v->nc_problem_stride;
  d.nc_ns = v->nc_node_stride;
  d.sym_packed = (v->flags & O2C_LQ_SYMMETRIC_PACKED) ? 1 : 0;
  if ((e = install_events(h, h->lanes[0].stream, v, true, begin, count)) != O2C_OK) return e;
  O2C_CUDA(launch_pack(h->L, d, h->d_lq, h->d_term, h->d_xnom, h->d_unom, h->d_nc, h->d_x0, begin, count, h->lanes[0].stream));
  h->launches += 1;
  return O2C_OK;
}

o2c_error o2c_download(o2c_handle* h, const o2c_solution_view* v, int32_t begin, int32_t count, int32_t n_alpha) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!v) return fail(O2C_ERR_INVALID_ARGUMENT, "null view");
  if (n_alpha < 0 || n_alpha > h->cfg.max_alphas) return fail(O2C_ERR_INVALID_ARGUMENT, "n_alpha outside [0, max_alphas]");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  Lane& lane = h->lanes[0];
  const size_t per = stage_out_per_problem(h->L, h->out_nodes, std::max(n_alpha, 1)) * sizeof(double);
  int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)count, ((size_t)1 << 30) / per));
  for (int off = 0; off < count; off += chunk) {
    const int c = std::min(chunk, count - off);
    o2c_solution_view sub = *v;
#define OFF(field) sub.field = offset_field(v->field, off);
    OFF(K) OFF(dbias) OFF(bias) OFF(Sm) OFF(Sv) OFF(s) OFF(x) OFF(u)
#undef OFF
    if (v->status) sub.status = v->status + off;
    e = download_chunk(h, lane, sub, begin + off, c, n_alpha);
    if (e != O2C_OK) return e;
  }
  O2C_CUDA(cudaStreamSynchronize(lane.stream));
  flush_status(h);
  return O2C_OK;
}

o2c_error o2c_backward(o2c_handle* h, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  return backward_on(h, h->lanes[0].stream, begin, count);
}

o2c_error o2c_rollout(o2c_handle* h, const double* alphas, int32_t n_alpha, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!alphas || n_alpha < 1 || n_alpha > h->cfg.max_alphas) return fail(O2C_ERR_INVALID_ARGUMENT, "n_alpha outside [1, max_alphas]");
  if (!h->backward_done) return fail(O2C_ERR_NOT_READY, "o2c_rollout needs the controller of o2c_backward");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  O2C_CUDA(cudaMemcpyAsync(h->d_alphas, alphas, sizeof(double) * n_alpha, cudaMemcpyHostToDevice, h->lanes[0].stream));
  return rollout_on(h, h->lanes[0].stream, h->d_alphas, n_alpha, begin, count);
}

o2c_error o2c_solve(o2c_handle* h, double alpha, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  return solve_on(h, h->lanes[0].stream, h->d_alphas, alpha, begin, count);
}

o2c_error o2c_line_search(o2c_handle* h, const o2c_line_search_settings* ls, const double* baseline, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!ls) return fail(O2C_ERR_INVALID_ARGUMENT, "null settings");
  if (h->st.algorithm != O2C_ALG_ILQR) return fail(O2C_ERR_UNSUPPORTED, "o2c_line_search evaluates the discrete (ILQR) LQ model only");
  if (!h->backward_done) return fail(O2C_ERR_NOT_READY, "o2c_line_search needs the controller of o2c_backward");
  if (!(ls->max_step_length > 0.0) || !(ls->min_step_length > 0.0) || !(ls->contraction_rate > 0.0 && ls->contraction_rate < 1.0) ||
      !(ls->armijo_coefficient >= 0.0))
    return fail(O2C_ERR_INVALID_ARGUMENT, "line search settings out of range");
  // candidates alpha_e = max * rate^e while almost_ge(alpha_e, min) (LineSearchStrategy.cpp:189-199, Numerics.h almost_eq)
  std::vector<double> cand;
  for (int ex = 0; ex < 64; ++ex) {
    const double a = ls->max_step_length * std::pow(ls->contraction_rate, (double)ex);
    const double diff = std::fabs(a - ls->min_step_length), mag = std::min(std::fabs(a), std::fabs(ls->min_step_length));
    const bool almost_eq = diff <= std::numeric_limits<double>::epsilon() * mag || diff < std::numeric_limits<double>::min();
    if (!(a > ls->min_step_length || almost_eq)) break;
    cand.push_back(a);
  }
  if (cand.empty()) return fail(O2C_ERR_INVALID_ARGUMENT, "max_step_length is below min_step_length: no candidate step length");
  if ((int)cand.size() > h->cfg.max_alphas)
    return fail(O2C_ERR_INVALID_ARGUMENT, "the settings give " + std::to_string(cand.size()) + " candidate step lengths but the handle was created with max_alphas = " +
                                              std::to_string(h->cfg.max_alphas));
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t stream = h->lanes[0].stream;
  const size_t B = (size_t)h->cfg.batch;
  if (!h->d_ls_merit) {
    O2C_CUDA(cudaMalloc(&h->d_ls_merit, sizeof(double) * B * h->cfg.max_alphas));
    O2C_CUDA(cudaMalloc(&h->d_ls_base, sizeof(double) * B));
    O2C_CUDA(cudaMalloc(&h->d_ls_is, sizeof(double) * B));
    O2C_CUDA(cudaMalloc(&h->d_ls_step, sizeof(double) * B));
    O2C_CUDA(cudaMalloc(&h->d_ls_basein, sizeof(double) * B));
    O2C_CUDA(cudaMalloc(&h->d_ls_index, sizeof(int) * B));
  }
  h->ls_candidates = cand;
  const int na = (int)cand.size();
  O2C_CUDA(cudaMemcpyAsync(h->d_alphas, cand.data(), sizeof(double) * na, cudaMemcpyHostToDevice, stream));
  if (baseline) O2C_CUDA(cudaMemcpyAsync(h->d_ls_basein, baseline, sizeof(double) * count, cudaMemcpyHostToDevice, stream));
  if ((e = rollout_on(h, stream, h->d_alphas, na, begin, count)) != O2C_OK) return e;
  const DeviceBuffers buf = h->buffers();
  O2C_CUDA(launch_merit(h->L, buf, h->out_nodes, na, h->cfg.batch, begin, count, h->d_ls_merit, stream));
  O2C_CUDA(launch_select(h->L, buf, h->d_ls_merit, h->d_alphas, na, h->cfg.batch, begin, count, ls->armijo_coefficient,
                         baseline ? h->d_ls_basein : nullptr, h->d_ls_base, h->d_ls_is, h->d_ls_step, h->d_ls_index, stream));
  h->launches += 2;
  if (baseline) O2C_CUDA(cudaStreamSynchronize(stream));  // the caller's baseline array may be pageable memory
  return O2C_OK;
}

o2c_error o2c_line_search_result(o2c_handle* h, double* step, int32_t* index, double* merits, double* baseline, double* update_is,
                                 double* candidates, int32_t* n_candidates, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!h->d_ls_merit || h->ls_candidates.empty()) return fail(O2C_ERR_NOT_READY, "o2c_line_search has not run on this handle");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t stream = h->lanes[0].stream;
  const int na = (int)h->ls_candidates.size();
  if (step) O2C_CUDA(cudaMemcpyAsync(step, h->d_ls_step + begin, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
  if (index) O2C_CUDA(cudaMemcpyAsync(index, h->d_ls_index + begin, sizeof(int) * count, cudaMemcpyDeviceToHost, stream));
  if (baseline) O2C_CUDA(cudaMemcpyAsync(baseline, h->d_ls_base + begin, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
  if (update_is) O2C_CUDA(cudaMemcpyAsync(update_is, h->d_ls_is + begin, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
  if (merits)
    for (int ex = 0; ex < na; ++ex)
      O2C_CUDA(cudaMemcpyAsync(merits + (size_t)ex * count, h->d_ls_merit + (size_t)ex * h->cfg.batch + begin, sizeof(double) * count,
                               cudaMemcpyDeviceToHost, stream));
  O2C_CUDA(cudaStreamSynchronize(stream));
  if (candidates) std::copy(h->ls_candidates.begin(), h->ls_candidates.end(), candidates);
  if (n_candidates) *n_candidates = na;
  return O2C_OK;
}

o2c_error o2c_download_flattened_controller(o2c_handle* h, float* host_out, double step_length, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!host_out) return fail(O2C_ERR_INVALID_ARGUMENT, "null output");
  if (!h->backward_done) return fail(O2C_ERR_NOT_READY, "o2c_backward has not run on this handle");
  if (!std::isfinite(step_length)) return fail(O2C_ERR_INVALID_ARGUMENT, "step_length must be finite");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  const size_t per = (size_t)(h->L.N + 1) * h->L.m * (h->L.n + 1);
  // converted in slices of at most 256 MiB of device scratch (the whole legged batch would be 4 GB of floats)
  const int slice = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->cfg.batch, ((size_t)256 << 20) / (per * sizeof(float))));
  if (!h->d_flat) O2C_CUDA(cudaMalloc(&h->d_flat, sizeof(float) * per * slice));
  cudaStream_t stream = h->lanes[0].stream;
  for (int off = 0; off < count; off += slice) {
    const int c = std::min(slice, count - off);
    O2C_CUDA(launch_flatten(h->L, h->d_sol, h->d_flat, step_length, begin + off, c, stream));
    h->launches += 1;
    O2C_CUDA(cudaMemcpyAsync(host_out + per * off, h->d_flat, sizeof(float) * per * c, cudaMemcpyDeviceToHost, stream));
    O2C_CUDA(cudaStreamSynchronize(stream));  // the scratch is reused by the next slice
  }
  return O2C_OK;
}

o2c_error o2c_discretize(o2c_handle* h, const o2c_discretization_view* v, int32_t scale_cost, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!v) return fail(O2C_ERR_INVALID_ARGUMENT, "null view");
  if (h->st.algorithm != O2C_ALG_ILQR) return fail(O2C_ERR_UNSUPPORTED, "discretisation belongs to the ILQR path (SLQ consumes continuous-time data)");
  if (v->stages != 1 && v->stages != 4) return fail(O2C_ERR_INVALID_ARGUMENT, "stages must be 1 or 4");
  for (int s = 0; s < v->stages; ++s)
    if (!v->dfdx[s].ptr || !v->dfdu[s].ptr) return fail(O2C_ERR_INVALID_ARGUMENT, "dfdx / dfdu of every stage are required");
  const int N = h->L.N;
  std::vector<double> dt(N);
  if (v->dt) {
    std::copy(v->dt, v->dt + N, dt.begin());
  } else {
    if (!h->time_set) return fail(O2C_ERR_NOT_READY, "no step lengths: pass dt or set the node times first");
    for (int k = 0; k < N; ++k) dt[k] = h->time[k + 1] - h->time[k];
  }
  for (double d : dt)
    if (!(d >= 0.0) || !std::isfinite(d)) return fail(O2C_ERR_INVALID_ARGUMENT, "step lengths must be finite and non-negative");
  if (5 * (size_t)h->L.n * h->L.n + 4 * (size_t)h->L.n * h->L.m > 227 * 1024 / sizeof(double))
    return fail(O2C_ERR_UNSUPPORTED, "state / input dimensions too large for the discretisation kernel");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t stream = h->lanes[0].stream;
  if (!h->d_dt) O2C_CUDA(cudaMalloc(&h->d_dt, sizeof(double) * (size_t)N));
  O2C_CUDA(cudaMemcpyAsync(h->d_dt, dt.data(), sizeof(double) * N, cudaMemcpyHostToDevice, stream));
  O2C_CUDA(cudaStreamSynchronize(stream));  // dt is pageable and about to go out of scope
  DiscretizeArgs a{};
  for (int s = 0; s < 4; ++s) {
    const int src = v->stages == 4 ? s : 0;
    a.dfdx[s] = FieldDev{v->dfdx[src].ptr, (long long)v->dfdx[src].problem_stride, (long long)v->dfdx[src].node_stride};
    a.dfdu[s] = FieldDev{v->dfdu[src].ptr, (long long)v->dfdu[src].problem_stride, (long long)v->dfdu[src].node_stride};
  }
  a.dt = h->d_dt;
  a.stages = v->stages;
  a.scale_cost = scale_cost != 0;
  O2C_CUDA(launch_discretize(h->L, a, h->d_lq, begin, count, stream));
  h->launches += 1;
  return O2C_OK;
}

o2c_error o2c_check_numerical_stability(o2c_handle* h, int32_t begin, int32_t count) {
  o2c_error e = check_range(h, begin, count);
  if (e != O2C_OK) return e;
  if (!h->backward_done) return fail(O2C_ERR_NOT_READY, "o2c_check_numerical_stability needs the value function of o2c_backward");
  if (count == 0) return O2C_OK;
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  O2C_CUDA(launch_check_psd(h->L, h->d_sol, h->d_status, begin, count, h->lanes[0].stream));
  h->launches += 1;
  return O2C_OK;
}

o2c_error o2c_launch_count(const o2c_handle* h, int64_t* launches) {
  if (!h || !launches) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  *launches = h->launches;
  return O2C_OK;
}

const char* o2c_kernel_variant(const o2c_handle* h) {
  if (!h) return "";
  if (h->st.algorithm == O2C_ALG_SLQ && slq_wpp_supported(h->L, h->st, h->buffers())) return "slq_wpp_kernel";
  if (h->st.algorithm == O2C_ALG_SLQ && rpl_slq_supported(h->L, h->st, h->buffers())) return "slq_rpl_kernel";
  if (h->st.algorithm == O2C_ALG_ILQR && h->use_fast && wpp_ilqr_supported(h->L, h->st, h->buffers())) return "ilqr_wpp_kernel";
  if (h->st.algorithm == O2C_ALG_ILQR && h->use_rpl && rpl_ilqr_supported(h->L, h->st, h->buffers())) return "ilqr_rpl_kernel";
  return generic_variant_name(h->L, h->st);
}

o2c_error o2c_solve_host(o2c_handle* h, const o2c_lq_view* lq, const o2c_solution_view* sol, double alpha, int32_t count, int32_t chunk) {
  o2c_error e = check_range(h, 0, count);
  if (e != O2C_OK) return e;
  if (!lq || !sol) return fail(O2C_ERR_INVALID_ARGUMENT, "null view");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  if (lq->time) {
    e = install_time(h, lq->time);
    if (e != O2C_OK) return e;
  }
  if (chunk <= 0) {
    const size_t per = stage_in_per_problem(h->L, h->cfg.has_nominal != 0) * sizeof(double);
    chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)(count + kLanes - 1) / kLanes, ((size_t)1 << 29) / per));
  }
  // lanes run [H2D -> pack -> sweep+rollout -> unpack -> D2H] for alternating chunks; copies of one lane overlap the kernels of another
  for (auto& lane : h->lanes) O2C_CUDA(cudaStreamSynchronize(lane.stream));
  if (count == h->cfg.batch)  // whole batch replaced: whole-range bookkeeping once, before the chunks (a chunk alone cannot tell)
    if ((e = begin_whole_upload(h, lq)) != O2C_OK) return e;
  int li = 0;
  o2c_error failed = O2C_OK;
  for (int off = 0; off < count && failed == O2C_OK; off += chunk, li = (li + 1) % kLanes) {
    const int c = std::min(chunk, count - off);
    Lane& lane = h->lanes[li];
    o2c_lq_view sub = *lq;
#define OFF(field) sub.field = offset_field(lq->field, off);
    OFF(A) OFF(B) OFF(Hv) OFF(Q) OFF(P) OFF(R) OFF(q) OFF(r) OFF(c) OFF(C) OFF(D) OFF(e) OFF(Qf) OFF(qf) OFF(cf) OFF(x_nom) OFF(u_nom) OFF(x0)
    OFF(jump_A) OFF(jump_Hv) OFF(jump_Q) OFF(jump_q) OFF(jump_c)
#undef OFF
    if (lq->nc) sub.nc = lq->nc + (long long)off * lq->nc_problem_stride;
    if (lq->event) sub.event = lq->event + (long long)off * lq->event_problem_stride;
    o2c_solution_view ss = *sol;
#define OFF(field) ss.field = offset_field(sol->field, off);
    OFF(K) OFF(dbias) OFF(bias) OFF(Sm) OFF(Sv) OFF(s) OFF(x) OFF(u)
#undef OFF
    if (sol->status) ss.status = sol->status + off;
    if ((failed = upload_chunk(h, lane, sub, off, c)) != O2C_OK) break;
    if ((failed = solve_on(h, lane.stream, h->d_alphas + h->cfg.max_alphas + li, alpha, off, c)) != O2C_OK) break;
    if ((failed = download_chunk(h, lane, ss, off, c, 1)) != O2C_OK) break;
  }
  // success or not, every lane is drained before returning: the caller's host buffers are no longer touched by copies in flight
  const std::string first_error = failed != O2C_OK ? g_last_error : std::string();
  cudaError_t drain = cudaSuccess;
  for (auto& lane : h->lanes) {
    const cudaError_t se = cudaStreamSynchronize(lane.stream);
    if (se != cudaSuccess && drain == cudaSuccess) drain = se;
  }
  if (failed != O2C_OK) {
    h->pending_status.clear();
    return fail(failed, first_error);
  }
  if (drain != cudaSuccess) {
    h->pending_status.clear();
    return fail(O2C_ERR_CUDA, std::string("o2c_solve_host: ") + cudaGetErrorString(drain));
  }
  flush_status(h);
  return O2C_OK;
}

o2c_error o2c_generate_synthetic(o2c_handle* h, uint64_t seed, int64_t first_problem_index, double dt) {
  if (!h) return fail(O2C_ERR_INVALID_ARGUMENT, "null handle");
  O2C_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->lanes[0].stream;
  h->events_present = false;
  if (h->d_event)  // stale flags of an earlier upload must not come back to life with a later partial upload that carries events
    O2C_CUDA(cudaMemsetAsync(h->d_event, 0, sizeof(int) * (size_t)h->cfg.batch * h->L.nodes, s));
  h->nc_ragged = false;
  h->slq_events.clear();  // the generated family has no events; the schedules are rebuilt with the generated time grid below
  O2C_CUDA(launch_generate(h->L, h->st.algorithm, h->d_lq, h->d_term, h->d_x0, seed, first_problem_index, dt, h->cfg.batch, s));
  h->launches += 1;
  if (h->d_nc) {
    const size_t cnt = (size_t)h->cfg.batch * h->L.nodes;
    fill_int_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(h->d_nc, h->L.ncmax, cnt);
    O2C_CUDA(cudaGetLastError());
  }
  if (h->d_xnom) {
    O2C_CUDA(cudaMemsetAsync(h->d_xnom, 0, sizeof(double) * (size_t)h->cfg.batch * (h->L.N + 1) * h->L.n, s));
    O2C_CUDA(cudaMemsetAsync(h->d_unom, 0, sizeof(double) * (size_t)h->cfg.batch * (h->L.N + 1) * h->L.m, s));
  }
  std::vector<double> t(h->L.N + 1);
  for (int k = 0; k <= h->L.N; ++k) t[k] = dt * (double)k;
  return install_time(h, t.data());
}

o2c_error o2c_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return fail(O2C_ERR_INVALID_ARGUMENT, "null argument");
  O2C_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return O2C_OK;
}

o2c_error o2c_host_free(void* ptr) {
  if (ptr) O2C_CUDA(cudaFreeHost(ptr));
  return O2C_OK;
}

}  // extern "C"
