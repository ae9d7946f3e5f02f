// o2c_common.cuh — shared declarations of libocs2_ddp_cuda (device layout, kernel parameter blocks, warp helpers).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ocs2_ddp_cuda.h"

namespace o2c {

// ---------------------------------------------------------------------------------------------------------------------
// Device-resident layout (library-owned). One interleaved record per (problem, node):
//   lq   [batch][nodes][rec]   rec = { A | B | Hv | q | r | c | Q | P | R | C | D | e }, every block padded to an even number of doubles
//   term [batch][trec]         trec = { Qf | qf | cf } padded
//   sol  [batch][N+1][orec]    orec = { K | dbias | bias | Sm | Sv | s } padded
//   xs   [alpha][batch][out_nodes][n],  us [alpha][batch][out_nodes][m]
// Every record starts 16-byte aligned so a whole record is one cp.async.bulk (TMA) transfer.
// ---------------------------------------------------------------------------------------------------------------------
struct Layout {
  int n, m, ncmax, N, nodes;
  int rec, oA, oB, oQ, oP, oR, oHv, oq, or_, oc, oC, oD, oe;
  int trec, oQf, oqf, ocf;
  int orec, oK, odb, obias, oSm, oSv, os;
};

inline int pad2(int v) { return (v + 1) & ~1; }

inline Layout make_layout(int n, int m, int ncmax, int N, int algorithm) {
  Layout L{};
  L.n = n;
  L.m = m;
  L.ncmax = ncmax;
  L.N = N;
  L.nodes = (algorithm == O2C_ALG_SLQ) ? N + 1 : N;
  int o = 0;
  // every matrix block starts on a 16-byte boundary (even double offset) so that vector loads / bulk copies stay aligned
  auto take = [&](int count) {
    int at = o;
    o = pad2(o + count);
    return at;
  };
  // big shapes (a matrix block of >= 4 KB): the cost-Hessian part and every record start on a 128-byte boundary, so that the
  // per-column L2 prefetches of the shape-specialised kernel touch whole 64-byte DRAM bursts. Small shapes stay densely packed
  // (they are HBM-bound; padding would be pure extra traffic).
  const int align = (n * n >= 512 || n * m >= 512) ? 16 : 2;
  auto align_up = [&]() { o = (o + align - 1) / align * align; };
  // operand part first (what the specialised sweep kernel stages into shared memory with one bulk copy), then the cost Hessians
  L.oA = take(n * n);
  L.oB = take(n * m);
  L.oHv = take(n);
  L.oq = take(n);
  L.or_ = take(m);
  L.oc = take(1);
  align_up();
  L.oQ = take(n * n);
  L.oP = take(m * n);
  L.oR = take(m * m);
  L.oC = take(ncmax * n);
  L.oD = take(ncmax * m);
  L.oe = take(ncmax);
  align_up();
  L.rec = o;
  o = 0;
  L.oQf = take(n * n);
  L.oqf = take(n);
  L.ocf = take(1);
  align_up();
  L.trec = o;
  o = 0;
  L.oK = take(m * n);
  L.odb = take(m);
  L.obias = take(m);
  L.oSm = take(n * n);
  L.oSv = take(n);
  L.os = take(1);
  align_up();
  L.orec = o;
  return L;
}

struct SolverSettings {
  int algorithm, reduced, strategy, hc;
  double eps, mu, time_step;
};

struct DeviceBuffers {
  const double* lq;
  const double* term;
  const double* x_nom;  // [batch][N+1][n] or nullptr
  const double* u_nom;  // [batch][N+1][m] or nullptr
  const int* nc;        // [batch][nodes] or nullptr
  const int* event;     // [batch][nodes] pre-event node flags or nullptr: no events anywhere
  const double* jump;   // SLQ: [batch][jump_capacity] jump records { A_e | Hv_e | Q_e | q_e | c_e } (offsets jump_offsets), or nullptr
  int jump_capacity;    // events per problem the jump array has room for
  const double* x0;     // [batch][n]
  const double* time;   // [N+1]
  double* sol;
  double* xs;
  double* us;
  int* status;
  int* work_counter;  // one int owned by the launching stream lane (dynamic problem fetch of the persistent kernels), or nullptr
};

// SM count of the CURRENT device, queried per call: nothing is cached in function-local statics, so handles on different devices
// and launches from different host threads never share launcher state
inline int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

// one RK4 step of the SLQ backward integration (host-precomputed, mirrors boost::odeint integrate_times)
struct SlqStep {
  int interval;      // data interval i: lerp between nodes i and i+1
  int observe_node;  // >= 0: after this step the state is the value function of this node
  int jump;          // > 0: node `interval` is a pre-event node; instead of integrating, apply the jump record jump - 1
                     // (ContinuousTimeRiccatiEquations::computeJumpMap, riccatiTransversalityConditions)
  double h;
  double alpha[4];   // weight of node i at the four RK4 stage times (1 - alpha on node i+1)
};
// one RK4 step of the continuous rollout (mirrors integrate_const + truncated last step)
struct RolloutStep {
  double h;
  int idx[4];       // timeSegment index at the four stage times
  double alpha[4];
  int obs_idx;      // timeSegment of the observation time after the step
  double obs_alpha;
  int jump;         // > 0: before observing, the state jumps through record jump - 1 at pre-event node pre_node (h == 0: no integration)
  int pre_node;
};
// jump record layout for state dimension n (doubles): A_e n*n | Hv_e n | Q_e n*n | q_e n | c_e 1, padded to an even count
__host__ __device__ inline int jump_oHv(int n) { return n * n; }
__host__ __device__ inline int jump_oQ(int n) { return n * n + n; }
__host__ __device__ inline int jump_oq(int n) { return 2 * n * n + n; }
__host__ __device__ inline int jump_oc(int n) { return 2 * n * n + 2 * n; }
__host__ __device__ inline int jump_rec(int n) { return (2 * n * n + 2 * n + 2) & ~1; }

// ---------------------------------------------------------------------------------------------------------------------
// launchers implemented in the .cu files
// ---------------------------------------------------------------------------------------------------------------------
cudaError_t launch_ilqr_generic(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count,
                                cudaStream_t stream);
const char* generic_variant_name(const Layout& L, const SolverSettings& st);
cudaError_t launch_slq_generic(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps,
                               int begin, int count, cudaStream_t stream);
cudaError_t launch_rollout_discrete(const Layout& L, const DeviceBuffers& buf, const double* alphas_dev, int n_alpha, int batch,
                                    int begin, int count, cudaStream_t stream);
cudaError_t launch_rollout_continuous(const Layout& L, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps, int first_idx,
                                      double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch, int begin,
                                      int count, cudaStream_t stream);
cudaError_t launch_generate(const Layout& L, int algorithm, double* lq, double* term, double* x0, uint64_t seed, int64_t first_index,
                            double dt, int batch, cudaStream_t stream);
struct FieldDev {
  double* ptr;
  long long ps, ns;
};
struct LqViewDev {
  FieldDev A, B, Hv, Q, P, R, q, r, c, C, D, e, Qf, qf, cf, x_nom, u_nom, x0;
  const int* nc;
  long long nc_ps, nc_ns;
  int sym_packed;  // O2C_LQ_SYMMETRIC_PACKED: Q, R, Qf are packed upper triangles
};
struct SolViewDev {
  FieldDev K, dbias, bias, Sm, Sv, s, x, u;
  long long x_as, u_as;
  int* status;
};
// strided SoA (device) -> records, for problems [begin, begin+count); view indexed from 0
cudaError_t launch_pack(const Layout& L, const LqViewDev& v, double* lq, double* term, double* x_nom, double* u_nom, int* nc, double* x0,
                        int begin, int count, cudaStream_t stream);
// records -> strided SoA (device)
cudaError_t launch_unpack(const Layout& L, const SolViewDev& v, const double* sol, const double* xs, const double* us, const int* status,
                          int out_nodes, int n_alpha, int batch, int begin, int count, cudaStream_t stream);

// ILQR::discreteLQWorker on caller-supplied stage linearisations (device, strided): A, B, Hv of the records, cost blocks scaled by dt
struct DiscretizeArgs {
  FieldDev dfdx[4], dfdu[4];
  const double* dt;  // device, [N]
  int stages, scale_cost;
};
cudaError_t launch_discretize(const Layout& L, const DiscretizeArgs& a, double* lq, int begin, int count, cudaStream_t stream);

// controller records -> float wire format of LinearController::flatten, out[count][N+1][m*(n+1)] (device)
cudaError_t launch_flatten(const Layout& L, const double* sol, float* out, double alpha, int begin, int count, cudaStream_t stream);

// shape-specialised fast path (riccati_wpp.cu): warp-per-problem DMMA kernel for nx = nu = 24; cudaErrorNotSupported otherwise
bool wpp_ilqr_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
cudaError_t launch_ilqr_wpp(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, bool with_rollout, double alpha, int batch,
                            int begin, int count, cudaStream_t stream, int* launches);

// shape-specialised fast path for small shapes (riccati_rpl.cu): row-per-lane kernel, several problems per warp
bool rpl_ilqr_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
cudaError_t launch_ilqr_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, bool with_rollout, double alpha, int begin,
                            int count, cudaStream_t stream, int* launches);

bool rpl_slq_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
cudaError_t launch_slq_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin,
                           int count, cudaStream_t stream);

// SLQ backward pass of the legged shape on the FP64 tensor pipe (slq_wpp.cu): projection, RK4 flow map, controller = 3 launches; needs
// a workspace of slq_wpp_workspace_doubles(L, batch) doubles
bool slq_wpp_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
size_t slq_wpp_workspace_doubles(const Layout& L, int batch);
cudaError_t launch_slq_wpp(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, double* workspace, const SlqStep* steps, int nsteps,
                           int begin, int count, cudaStream_t stream, int* launches);

// continuous LQ rollout of the legged shape (slq_wpp.cu)
bool rollout_cont24_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
cudaError_t launch_rollout_cont24(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps,
                                  int first_idx, double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch, int begin,
                                  int count, cudaStream_t stream);

bool rpl_rollout_cont_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf);
cudaError_t launch_rollout_cont_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps,
                                    int first_idx, double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch,
                                    int begin, int count, cudaStream_t stream);

// checkBeingPSD of every S_k (stability.cu): ors O2C_STATUS_NOT_PSD / O2C_STATUS_NONFINITE into status[problem]
cudaError_t launch_check_psd(const Layout& L, const double* sol, int* status, int begin, int count, cudaStream_t stream);

// batched Armijo line search on the LQ model (line_search.cu)
cudaError_t launch_merit(const Layout& L, const DeviceBuffers& buf, int out_nodes, int n_alpha, int batch, int begin, int count, double* merit,
                         cudaStream_t stream);
cudaError_t launch_select(const Layout& L, const DeviceBuffers& buf, const double* merit, const double* alphas_dev, int n_alpha, int batch,
                          int begin, int count, double armijo, const double* baseline_in_dev, double* baseline_out, double* update_is,
                          double* step, int* index, cudaStream_t stream);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy (cp.async.bulk), L2 prefetch
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "O2C_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra O2C_MBAR_DONE;\n"
      "bra O2C_MBAR_WAIT;\n"
      "O2C_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form); completion is signalled on the mbarrier as transferred bytes
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// per-thread L2 prefetch of the line holding p (CCTL.PF2): one instruction for the whole warp, every lane its own address
__device__ __forceinline__ void l2_touch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// orders this thread's ordinary (generic-proxy) global stores before later async-proxy (TMA) reads of the same addresses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }


// ---------------------------------------------------------------------------------------------------------------------
// warp-cooperative dense helpers: one warp owns one problem; matrices live in shared memory, column-major.
// Every helper ends with __syncwarp() so results are visible to all lanes.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor pipe: lane (r = lane / 4, c = lane % 4) supplies a = A[r][c], b = B[c][r] and owns
// d.x = D[r][2c], d.y = D[r][2c + 1]
__device__ __forceinline__ void dmma_m8n8k4(double2& d, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}

// C(MxN) = beta*C + alpha * op(A)*op(B); op(A) is MxK, op(B) is KxN. TA: A stored KxM; TB: B stored NxK.
// Shapes made of whole 8x8x4 tiles (the legged shape in every configuration the specialised kernel does not serve: LM, Gershgorin,
// eigenvalue correction, SLQ) go through DMMA, three column tiles per A fragment; everything else one output element per lane.
template <bool TA, bool TB>
__device__ __forceinline__ void wgemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
                                      const double* __restrict__ B, int ldb, double beta, double* C, int ldc) {
  const int lane = lane_id();
  if (((M | N) & 7) == 0 && (K & 3) == 0) {
    const int r = lane >> 2, c = lane & 3;
    for (int ti = 0; ti < M; ti += 8) {
      for (int tj = 0; tj < N; tj += 24) {
        double2 acc[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        const int ntile = (N - tj) >= 24 ? 3 : (N - tj) / 8;
        for (int k0 = 0; k0 < K; k0 += 4) {
          const double a = TA ? A[(k0 + c) + (ti + r) * lda] : A[(ti + r) + (k0 + c) * lda];
#pragma unroll
          for (int t = 0; t < 3; ++t)
            if (t < ntile) {
              const int j = tj + 8 * t + r;
              const double b = TB ? B[j + (k0 + c) * ldb] : B[(k0 + c) + j * ldb];
              dmma_m8n8k4(acc[t], a, b);
            }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (t < ntile) {
            double* c0 = C + (ti + r) + (tj + 8 * t + 2 * c) * ldc;
            double* c1 = c0 + ldc;
            *c0 = (beta == 0.0 ? 0.0 : beta * (*c0)) + alpha * acc[t].x;
            *c1 = (beta == 0.0 ? 0.0 : beta * (*c1)) + alpha * acc[t].y;
          }
      }
    }
    __syncwarp();
    return;
  }
  const int total = M * N;
  for (int idx = lane; idx < total; idx += 32) {
    const int i = idx % M, j = idx / M;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
      const double a = TA ? A[k + i * lda] : A[i + k * lda];
      const double b = TB ? B[j + k * ldb] : B[k + j * ldb];
      acc = fma(a, b, acc);
    }
    double* c = C + i + j * ldc;
    *c = (beta == 0.0 ? 0.0 : beta * (*c)) + alpha * acc;
  }
  __syncwarp();
}

__device__ __forceinline__ void wcopy(int count, const double* __restrict__ src, double* __restrict__ dst) {
  for (int i = lane_id(); i < count; i += 32) dst[i] = src[i];
  __syncwarp();
}
__device__ __forceinline__ void wfill(int count, double v, double* dst) {
  for (int i = lane_id(); i < count; i += 32) dst[i] = v;
  __syncwarp();
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wdot(int count, const double* a, const double* b) {
  double acc = 0.0;
  for (int i = lane_id(); i < count; i += 32) acc = fma(a[i], b[i], acc);
  return warp_sum(acc);
}
#endif

}  // namespace o2c
