// riccati_wpp.cu — warp-per-problem ILQR sweep + LQ rollout for nx = nu = 24 (the legged-robot shape), FP64, sm_100a.
//
// One WARP owns one problem's whole time-sequential sweep; one persistent CTA per SM, 12 warps (168 registers per thread fill the
// register file), no CTA barrier after the prologue. The warps of a CTA have roles: SWEEPERS run backward pass after backward pass,
// ROLLERS run the forward rollouts of finished problems, which reach them through a ticketed queue in shared memory (see the kernel
// for why: fused behind each sweep the rollout took 21 % of the time for 1.6 % of the flops). While one sweeper sits in the latency
// chain of its Cholesky the other warps of the same scheduler keep the FP64 pipe busy with their contractions. Template variants:
// nominal trajectories, events, the Riccati modification (line search / Levenberg-Marquardt / Gershgorin), up to 16 state-input
// equality constraints as 8-row tiles, and a 14-warp / 128-register instantiation for batches of one round.
//
//   * The per-node operand block {A | B | Hv | q | r | c} (9.8 KB) is staged into the warp's shared-memory slot by ONE TMA bulk copy
//     (cp.async.bulk + a per-warp mbarrier), issued half a stage ahead (the second half of a stage does not touch A, B). The cost
//     Hessians Q, P, R are only ever added to accumulators: they are touched into L2 per lane at the top of the stage and read
//     straight from L2 into the accumulator registers, one contraction ahead of their first use. {C | D | e} of a constrained node
//     arrive by their own bulk copy. A rollout stage reads {A | B | Hv} and {K | dbias} out of a TMA ring of stage sets.
//   * Every 24x24x24 contraction runs on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no FP64 kind) in the form
//     Z = X'Y, for which both operand fragments have the same register layout ("op": lane (r,c) holds M[8kb+2c..2c+1][8jb+r]) and the
//     accumulator fragment of Z ("acc": lane holds Z[8ib+r][8jb+2c..2c+1]) IS the operand fragment of Z'. The sweep is arranged so
//     that every product is consumed in exactly that transposed role, so the chain SA -> {G, Hm, T} -> Y -> {S, K} never leaves the
//     register file:  ZA = A'S (= (SA)'), ZB = B'S, G' = P' + SA'B, Hm = R + SB'B, T = Q + eps I + SA'A, Y' = G' L^-T,
//     S = T - Y'Y, K' = -Y' L^-1.
//   * Only Hm takes a detour through a 5 KB shared scratch (leading dimension 26: conflict-free 8-byte transposed access): it is
//     factorised in 8x8 blocks (block columns one row per lane, trailing updates and the off-diagonal blocks of L^-1 on the tensor
//     pipe, see factor_hm), and the value function S is parked in the same scratch between stages so that it can be re-read in
//     operand layout.
//   * The matrix-vector terms ride on the fragments of the contractions (S Hv, A'Sv, B'Sv in the ZA/ZB loop; (SA)'Hv and (SB)'Hv
//     from the accumulators), so w = Sv + S Hv never exists in memory.
//
// Math (unconstrained, LINE_SEARCH, reduced Riccati form, DIAGONAL_SHIFT; same quantities as the reference, re-associated):
//   Hm = R + B'(S B)                         ILQR::computeHamiltonianHessian              ocs2_ddp/src/ILQR.cpp:217-222
//   Hm = L L',  Pu = U^-1 = L^-T             LinearAlgebra::computeInverseMatrixUUT       ocs2_core/src/misc/LinearAlgebra.cpp:119-124
//   projected G~m = Pu'(P + B'SA) = L^-1 G =: Y,  G~v = L^-1 (r + B'w) =: Yv,  w = Sv + S Hv
//                                            DiscreteTimeRiccatiEquations::computeMapILQR  .../DiscreteTimeRiccatiEquations.cpp:65-154
//   S  = Q + eps I + A'(SA) - Y'Y ;  Sv = q + A'w - Y'Yv ;  s = s+ + c + Hv.w - 1/2 Hv.(S Hv) - 1/2 Yv.Yv
//   K  = Pu K~ = -L^-T Y ;  dbias = -L^-T Yv ;  bias = 0 (deviation coordinates)        ILQR::calculateControllerWorker ILQR.cpp:162-181
//   dQ = eps I: the reference forms (M + eps I) - M with M = Q~ - P~'P~ (LineSearchStrategy.cpp:294-312), which equals eps I up
//   to one rounding of M_ii + eps (<= 1e-16 |M_ii|).
// Rollout (roller warps): du_k = K_k dx_k + alpha dbias_k ; dx_{k+1} = A_k dx_k + B_k du_k + Hv_k
//                                            DDP_HelperFunctions.cpp:125-138, 296-304; LinearController.cpp:79-87
#include <cstdio>
#include <cstdlib>

#include "o2c_common.cuh"
#include "wpp_tiles.cuh"

namespace o2c {
namespace {

// NCB: tiles of 8 state-input equality constraints the kernel instantiation carries (0: unconstrained; nc_max <= 8 NCB)
template <int NCB>
struct __align__(16) WarpSmemT {
  static constexpr int kC8 = NCB > 0 ? 8 * NCB : 2;
  double in[kOperand];   // TMA destination: {A | B | Hv | q | r | c,pad}
  double W[kN * kLd];    // S of node k+1 (both triangles) -> Hm (lower) -> L -> L^-T (upper)
  double Sv[kN], Gv[kN], Yv[kN], xb[kN], ub[kN];
  double cde[NCB > 0 ? 2 * kC8 * kN + kC8 : 2];  // TMA destination: {C | D | e} as in the record (leading dimension nc_max)
  double Wm[NCB > 0 ? kC8 * kLd : 2];            // M = Z'Z (lower) -> L_M -> L_M^-T (upper), leading rows of a ld-26 scratch
  double vw[kC8], vv[kC8];                       // Z'Yv - e -> L_M^-T vv ; vv = L_M^-1 (Z'Yv - e)
  unsigned long long full;   // operand block landed (sweep)
  unsigned long long cfull;  // {C | D | e} landed
};
using WarpSmem = WarpSmemT<0>;
static_assert(sizeof(WarpSmemT<0>) % 16 == 0 && sizeof(WarpSmemT<1>) % 16 == 0 && sizeof(WarpSmemT<2>) % 16 == 0, "warp slots must keep 16-byte alignment");

struct Args {
  const double* lq;
  const double* term;
  const double* x0;
  const double* x_nom;  // [batch][N+1][24] nominal trajectories, or nullptr (deviation coordinates): kernel instantiation NOM
  const double* u_nom;
  double* sol;
  double* xs;
  double* us;
  int* status;
  const int* event;  // [batch][N] pre-event flags (ILQR.cpp:263-295), or nullptr: kernel instantiation EV
  const int* nc;     // [batch][N] active constraints per node, or nullptr (nc_max everywhere); constrained instantiations only
  int rec, oC, cdeD, cdeE, ncmax;  // constrained instantiations: record stride, offset of {C | D | e}, D and e inside that block
  uint32_t cde_bytes;
  int N;
  int oQf, oqf, ocf, trec;
  int begin, count, with_rollout;
  int sweep_count;  // problems [0, sweep_count) belong to the sweepers; [sweep_count, count) are swept by the rollers before their first rollout
  int ring_depth;  // stage sets in a roller's ring (the rollers' rings follow the sweepers' slots in dynamic shared memory)
  int nsweep;    // warps [0, nsweep) of a CTA sweep; the others only roll out (0 < nsweep <= warps per CTA)
  int* counter;  // dynamic problem fetch: zeroed before the launch; nullptr = static stride
  int dyn_limit;  // problems [resident slots, dyn_limit) are fetched dynamically, [dyn_limit, count) = the last partial wave, static
  double eps, alpha, mu;
};

// L2 prefetch of the cost Hessians a stage reads: all of P and Q (they are copied to shared memory whole by TMA), and the tiles of R on
// or above the block diagonal (R is symmetric; the sweep reads R[8jb+2c..][8ib+r] for jb <= ib). One 16-byte touch per lane and tile.
__device__ __forceinline__ void prefetch_hessians(const double* rec, const Args& a, int lo24) {
#pragma unroll
  for (int ib = 0; ib < 3; ++ib)
#pragma unroll
    for (int jb = 0; jb < 3; ++jb) {
      l2_touch(rec + kOP + lo24 + t24(jb, ib));
      if (jb <= ib) l2_touch(rec + kOQ + lo24 + t24(jb, ib));  // Q is read like R: only the tiles on or above the block diagonal
      if (jb <= ib) l2_touch(rec + kOR + lo24 + t24(jb, ib));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// next problem of this warp. The first problem is static (warp-major over the CTAs: CTA b gets b, b + grid, b + 2 grid, ... so that a
// batch smaller than the resident slots still spreads evenly over the SMs). The problems of the following FULL waves come from a global
// counter, so a warp that finishes early picks up the next problem instead of idling until the slowest warp of a static schedule is
// done. The last, partial wave is static again (warp-major): handed out dynamically it would be grabbed by the SMs that finished
// first, twelve problems each, and run at full-machine contention on a few SMs while the others idle (measured: 3.4 ms instead of 2.5 ms
// for 2048 problems, profiles/r02_wpp_residency.jsonl).
__device__ __forceinline__ int next_problem(const Args& a, int lane, int total_warps, int pi, int slot, bool& tail_taken) {
  if (a.counter == nullptr) return pi + total_warps;
  if (tail_taken) return a.sweep_count;
  int nx = 0;
  if (lane == 0) nx = atomicAdd(a.counter, 1) + total_warps;
  nx = __shfl_sync(kFull, nx, 0);
  if (nx < a.dyn_limit) return nx;
  tail_taken = true;
  return a.dyn_limit + slot;  // >= sweep_count when this warp has no problem in the last wave
}

// MODE: the Riccati modification (search strategy) the stage carries, everything else is shared:
//   kModeLS    LINE_SEARCH + DIAGONAL_SHIFT: dQ = eps I                                   LineSearchStrategy.cpp:294-312, HessianCorrection.cpp:53-74
//   kModeGersh LINE_SEARCH + GERSHGORIN_MODIFICATION: dQ = diag(max(0, R_i + eps - M_ii)), M = Q - P'Hm^-1 P, R_i = sum_{j != i} |M_ji|
//                                                                                        LinearAlgebra.cpp:77-85 (the off-diagonal part of dQ,
//              (M' - M) / 2, is rounding noise of a symmetric M and is not formed)
//   kModeLM    LEVENBERG_MARQUARDT (full Riccati form): Hm += mu B'B, dGm = mu B~'A~, dGv = mu B~'Hv~, dQ = 0
//                                                                                        LevenbergMarquardtStrategy.cpp:230-249
//              With S' = S + mu I this is the LINE_SEARCH stage evaluated on S' (which yields Hm, K, dbias exactly) minus mu times the Gram
//              terms of the closed loop: S = T(S') - Y'Y - mu Acl'Acl, Sv = tv(S') - Y'Yv - mu Acl'hcl, s = s(S') - mu/2 |hcl|^2 with
//              Acl = A + B K, hcl = Hv + B dbias (expand the full-form map of DiscreteTimeRiccatiEquations.cpp:91-153 with
//              H~m = I - mu B~'B~: the cross terms cancel).
constexpr int kModeLS = 0, kModeLM = 1, kModeGersh = 2;

// Profiling builds only (tools/wpp_ablate.py): -DO2C_WPP_ABLATE=<bits> removes one cost at a time from the kernel to measure what it is
// worth at full residency. The results of such a build are WRONG by construction; the product is always built with 0.
//   1 no factorisation   2 every stage reads the record of node N-1 (no DRAM reads in the sweep)   4 every stage writes node 0's record
//   8 no K / S global stores   16 the rollout reads node 0's records at every stage (no DRAM reads in the rollout)
#ifndef O2C_WPP_ABLATE
#define O2C_WPP_ABLATE 0
#endif
constexpr int kAblate = O2C_WPP_ABLATE;
constexpr int kQueue = 64;  // finished problems waiting for a roller, per CTA
constexpr int kDefaultRollers = 1;
#ifndef O2C_WPP_CTA_WARPS
#define O2C_WPP_CTA_WARPS 12
#endif
constexpr int kMaxCtaWarps = O2C_WPP_CTA_WARPS;  // sweepers + rollers of a CTA (the register file holds 12 warps at 168 registers)
constexpr int kWideWarps = 14;  // the WIDE instantiations: 14 sweepers at 128 registers (register allocation is per 4 warps: 16 x 32 x 128)
constexpr int kRingK = 2 * kMat + kN;        // stage set of the rollout ring: { A | B | Hv } then { K | dbias }
constexpr int kRing = kRingK + kMat + kN;    // 1776 doubles = 14.2 KB
constexpr int kMaxRingDepth = 4;
#ifndef O2C_WPP_ROLL_AHEAD
#define O2C_WPP_ROLL_AHEAD 3
#endif
constexpr int kRollAhead = O2C_WPP_ROLL_AHEAD;  // rollout: L2 prefetch distance in stages

// WIDE: 14 warps per SM at 128 registers (a few spills) instead of 12 at 168. One sweep is ~10 % slower, but a batch of 12 x SMs <
// count <= 14 x SMs problems — 2048 on 148 SMs, the 8-GPU share of BASELINE config 5 — runs as ONE round of sweeps instead of two:
// 2.12 ms instead of 2.21 ms (profiles/r02_wpp_wide.jsonl). Every sweeper rolls its own problem out (no room for a roller's ring).
template <bool NOM, bool EV, int MODE, int NCB, bool WIDE = false>
// (the constrained instantiations have room for 10 / 8 warps in shared memory: their register budget is that of 320 / 256 threads)
__global__ void __launch_bounds__(WIDE ? 32 * kWideWarps : (NCB == 0 ? 32 * kMaxCtaWarps : (NCB == 1 ? 320 : 256)), 1) ilqr_wpp_kernel(const Args a) {
  static_assert(!(EV && MODE == kModeLM), "ILQR events under LEVENBERG_MARQUARDT are refused by the API");
  static_assert(NCB == 0 || MODE == kModeLS, "the constrained instantiations serve LINE_SEARCH + DIAGONAL_SHIFT");
  using WarpSmem = WarpSmemT<NCB>;
  constexpr int NCB_ = NCB > 0 ? NCB : 1;
  const int rec_stride = NCB > 0 ? a.rec : kRec;  // (a compile-time constant for the unconstrained layout)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a roller's ring doubles as the WarpSmem of the one sweep it does before its first rollout
  WarpSmem& ws = warp < a.nsweep ? reinterpret_cast<WarpSmem*>(smem_raw)[warp]
                                 : *reinterpret_cast<WarpSmem*>(smem_raw + (size_t)a.nsweep * sizeof(WarpSmem) +
                                                                (size_t)(warp - a.nsweep) * a.ring_depth * kRing * sizeof(double));
  const int r = lane >> 2, c = lane & 3;
  const int lo24 = 2 * c + kN * r, lo26 = 2 * c + kLd * r;
  const int li = lane < kN ? lane : kN - 1;
  const int N = a.N;
  const uint32_t opBytes = kOperand * sizeof(double);

  const int nwarps = blockDim.x >> 5, nsweep = a.nsweep, nroll = nwarps - nsweep;
  constexpr int kCtaWarps = WIDE ? kWideWarps : kMaxCtaWarps;
  __shared__ unsigned long long ring_full[kCtaWarps][kMaxRingDepth];  // rollout ring: stage set landed
  __shared__ __align__(16) double rvec[kCtaWarps][2][kN];               // rollout: x, u of the current node
  if (lane == 0) {
    if (warp < nsweep || a.sweep_count < a.count) {
      mbar_init(&ws.full, 1);
      mbar_init(&ws.cfull, 1);
    }
    for (int d = 0; d < kMaxRingDepth; ++d) mbar_init(&ring_full[warp][d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0, cphase = 0, rphase = 0;  // rphase: bit d = parity to wait for on ring_full[warp][d]

  // Warp roles. Warps [0, nsweep) are SWEEPERS: backward pass after backward pass, the FP64 pipe never waits for a rollout. The
  // remaining warps are ROLLERS: they take finished problems from a queue in shared memory and run the forward rollout, which is
  // 1.6 % of the flops but streams {A|B|Hv} and {K|dbias} once more (1.4 MB per problem) and is bound by DRAM latency and bandwidth:
  // fused behind each sweep in the same warp it took 21 % of the kernel time at full residency (profiles/r02_wpp_ablation.jsonl). A
  // sweeper that runs out of problems becomes a roller, so the tail of a launch drains on every warp of the CTA. With no rollers
  // (nsweep == warps per CTA) every sweeper rolls its own problem out, as before.
  // queue entry = (ticket + 1) << 32 | problem, 0 = empty. A producer writes its slot once the ticket one lap earlier has been claimed
  // AND its entry consumed (slot back to 0); a consumer claims a ticket by CAS on the head and takes exactly that ticket's entry.
  __shared__ unsigned long long queue[kQueue];
  __shared__ unsigned q_tail, q_head, sweepers_done;
  if (threadIdx.x < kQueue) queue[threadIdx.x] = 0ull;
  if (threadIdx.x == 0) q_tail = q_head = sweepers_done = 0;
  __syncthreads();
  const int total_warps = gridDim.x * nsweep;
  const int slot = warp * gridDim.x + blockIdx.x;
  bool tail_taken = false;
  // The rollers have nothing to roll out until the first sweeps finish (1.4 ms): each of them first sweeps one problem from the end of
  // the batch itself (its ring is the scratch), which is what keeps a launch from ending in a nearly empty extra round when the batch
  // is a little more than a whole number of rounds (16384 problems = 10.06 rounds of 148 x 11 sweepers).
  const int init_pi = a.sweep_count + (warp - nsweep) * (int)gridDim.x + (int)blockIdx.x;
  bool sweeping = warp < nsweep || init_pi < a.count;
  int pi = warp < nsweep ? slot : init_pi;
  int done_target = nsweep;
  for (int q = 0; q < nroll; ++q) done_target += a.sweep_count + q * (int)gridDim.x + (int)blockIdx.x < a.count ? 1 : 0;
#ifdef O2C_WPP_STATS
  long long st_t0 = clock64(), st_sweep = 0, st_roll = 0, st_wait = 0, st_mark = 0;
  int st_nsweep = 0, st_nroll = 0;
#endif
  for (;;) {
    int job = -1;
#ifdef O2C_WPP_STATS
    st_mark = clock64();
#endif
    if (sweeping) {
      if (pi >= (warp < nsweep ? a.sweep_count : a.count)) {
        sweeping = false;
        if (nroll == 0) break;
        __syncwarp();
        if (lane == 0 && warp >= nsweep) {  // the scratch of the roller's own sweep becomes its ring
          mbar_inval(&ws.full);
          mbar_inval(&ws.cfull);
        }
        if (lane == 0) {
          __threadfence_block();
          atomicAdd(&sweepers_done, 1u);
        }
      }
    }
    if (sweeping) {
    const int prob = a.begin + pi;
    const double* lqp = a.lq + (size_t)prob * N * rec_stride;
    const double* term = a.term + (size_t)prob * a.trec;
    double* solp = a.sol + (size_t)prob * (N + 1) * kORec;
    const int* evp = EV ? a.event + (size_t)prob * N : nullptr;
    const int* ncp = (NCB > 0 && a.nc != nullptr) ? a.nc + (size_t)prob * N : nullptr;

    // operand block of node N-1 (TMA) and the L2 prefetch of its cost Hessians
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&ws.full, opBytes);
      tma_load(ws.in, lqp + (size_t)(N - 1) * rec_stride, opBytes, &ws.full);
      if (NCB > 0) {
        mbar_expect_tx(&ws.cfull, a.cde_bytes);
        tma_load(ws.cde, lqp + (size_t)(N - 1) * rec_stride + a.oC, a.cde_bytes, &ws.cfull);
      }
    }
    prefetch_hessians(lqp + (size_t)(N - 1) * rec_stride, a, lo24);
    // terminal condition: valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526)
    {
      double* outN = solp + (size_t)N * kORec;
#pragma unroll 1
      for (int i = lane; i < kMat; i += 32) {
        const double v = term[a.oQf + i];
        ws.W[(i % kN) + kLd * (i / kN)] = v;
        outN[kOSm + i] = v;
      }
      if (lane < kN) {
        const double v = term[a.oqf + lane];
        ws.Sv[lane] = v;
        outN[kOSv + lane] = v;
      }
      if (lane == 0) outN[kOs] = term[a.ocf];
    }
    double sval = term[a.ocf];  // s of node k+1
    bool pd = true, rank_ok = true;
    double2 t[6];               // T -> S (lower tiles); after the loop: S of node 0 for the finiteness test
    double svn = 0.0;
    __syncwarp();

#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
      const double* A = ws.in;
      const double* B = ws.in + kMat;
      const double* Hv = ws.in + 2 * kMat;
      const double* qv = Hv + kN;
      const double* rv = qv + kN;
      const double* rec = lqp + (size_t)((kAblate & 2) ? N - 1 : k) * rec_stride;
      double* out = solp + (size_t)((kAblate & 4) ? 0 : k) * kORec;
      // pre-event node (ILQR.cpp:263-295): the staged A, Hv, q, c and Q are the jump map and the pre-jump cost. The value function goes
      // through riccatiTransversalityConditions (S- = Q_e + A_e'S A_e, Sv- = q_e + A_e'(Sv + S Hv), s- = s + c_e + Hv.(Sv + S Hv / 2));
      // the controller entry comes from the node's B, P, R, r against a zero next value function (Hm = R) and S-, Sv-:
      // G = P + B'S-, Gv = r + B'Sv-. Warp-uniform.
      const bool ev = EV && __ldg(evp + k) != 0;
      mbar_wait(&ws.full, phase);
      phase ^= 1u;
      if (k < N - 1) prefetch_hessians(rec, a, lo24);  // this node's Q, P, R into L2 now: first needed a third of a stage from here
                                                        // (node N-1's were touched in the prologue); prefetching any earlier only
                                                        // loses lines to L2 capacity misses with 1776 problems in flight

      // Hm accumulators start from R (in flight during the ZA / ZB contractions)
      double2 h[6];
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) h[lt(ib, jb)] = ldg2(rec + kOR + lo24 + t24(jb, ib));  // R[8ib+r][8jb+2c..] = R[8jb+2c..][8ib+r]

      // ---- ZA = A'S (= op fragments of SA), ZB = B'S (= op fragments of SB); the matrix-vector terms ride on the same fragments:
      //      S Hv, A'Sv, B'Sv here, and (SA)'Hv = ZA Hv, (SB)'Hv = ZB Hv from the accumulators, so that
      //      Gv = r + B'(Sv + S Hv), tv = q + A'(Sv + S Hv) never need w = Sv + S Hv in shared memory ----
      double2 zA[3][3], zB[3][3];
      double2 hvf[3];
      double pSH[3] = {0.0, 0.0, 0.0}, pA[3] = {0.0, 0.0, 0.0}, pB[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) zA[i][j] = zB[i][j] = zero2();
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        double2 s[3];
        hvf[kb] = ld2(Hv + 8 * kb + 2 * c);
        const double2 svf = ld2(ws.Sv + 8 * kb + 2 * c);
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) {
          s[jb] = ld2(ws.W + lo26 + t26(kb, jb));
          if (MODE == kModeLM && kb == jb) {  // S' = S + mu I: rows 8kb + 2c, 2c + 1 of column 8jb + r
            s[jb].x += (2 * c == r) ? a.mu : 0.0;
            s[jb].y += (2 * c + 1 == r) ? a.mu : 0.0;
          }
          pSH[jb] = fma(s[jb].x, hvf[kb].x, pSH[jb]);
          pSH[jb] = fma(s[jb].y, hvf[kb].y, pSH[jb]);
        }
#pragma unroll
        for (int ib = 0; ib < 3; ++ib) {
          const double2 af = ld2(A + lo24 + t24(kb, ib)), bf = ld2(B + lo24 + t24(kb, ib));
          pA[ib] = fma(af.x, svf.x, pA[ib]);
          pA[ib] = fma(af.y, svf.y, pA[ib]);
          pB[ib] = fma(bf.x, svf.x, pB[ib]);
          pB[ib] = fma(bf.y, svf.y, pB[ib]);
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            dmma2(zA[ib][jb], af, s[jb]);
            dmma2(zB[ib][jb], bf, s[jb]);
          }
        }
      }
      double spart = 0.0, tvj = 0.0;
      {
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            pA[ib] = fma(zA[ib][jb].x, hvf[jb].x, pA[ib]);
            pA[ib] = fma(zA[ib][jb].y, hvf[jb].y, pA[ib]);
            pB[ib] = fma(zB[ib][jb].x, hvf[jb].x, pB[ib]);
            pB[ib] = fma(zB[ib][jb].y, hvf[jb].y, pB[ib]);
          }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          pSH[i] = quad_sum(pSH[i]);
          pA[i] = quad_sum(pA[i]);
          pB[i] = quad_sum(pB[i]);
        }
        if (c < 3) {
          const int j = 8 * c + r;
          spart = Hv[j] * (ws.Sv[j] + 0.5 * pick3(pSH, c));  // Hv.w - 1/2 Hv.(S Hv)
          ws.Gv[j] = rv[j] + pick3(pB, c);
          tvj = qv[j] + pick3(pA, c);
        }
      }
      const double cval = ws.in[2 * kMat + 3 * kN];
      const double epsk = (ev || MODE != kModeLS) ? 0.0 : a.eps;  // the transversality condition carries no Hessian correction; LM: dQ = 0;
                                                                   // Gershgorin: dQ is added to S at the end of the stage
      __syncwarp();  // S (scratch) is dead from here on

      // Q (its six tiles on or below the block diagonal) and P go from L2 straight into the accumulators of T and G', each issued one
      // contraction (576 pipe cycles) ahead of its first use, like R at the top of the stage
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) t[lt(ib, jb)] = ldg2(rec + kOQ + lo24 + t24(jb, ib));  // Q[8ib+r][8jb+2c..] = Q[8jb+2c..][8ib+r]

      // ---- Hm = R + SB'B (lower tiles) ----
      if (!ev) {
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          double2 bf[3];
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) bf[jb] = ld2(B + lo24 + t24(kb, jb));
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(h[lt(ib, jb)], zB[ib][kb], bf[jb]);
        }
      }

      // ---- T = Q + eps I + SA'A (lower tiles) ----
      double2 g[3][3];
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) g[ib][jb] = ldg2(rec + kOP + lo24 + t24(jb, ib));  // P'[8ib+r][8jb+2c..] = P[8jb+2c..][8ib+r]
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        double2 af[3];
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) af[jb] = ld2(A + lo24 + t24(kb, jb));
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) dmma2(t[lt(ib, jb)], zA[ib][kb], af[jb]);
      }
#pragma unroll
      for (int ib = 0; ib < 3; ++ib) {
        t[lt(ib, ib)].x += (2 * c == r) ? epsk : 0.0;
        t[lt(ib, ib)].y += (2 * c + 1 == r) ? epsk : 0.0;
      }
      if (ev) {
        // S- = T and Sv- = tv are final: through the scratch once more for G' = P' + S-B and Gv = r + B'Sv-
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            st2(ws.W + lo26 + t26(jb, ib), t[lt(ib, jb)]);
            if (ib != jb) tput(ws.W, ib, jb, r, c, t[lt(ib, jb)]);
          }
        if (c < 3) ws.Sv[8 * c + r] = tvj;
        __syncwarp();
        double pG[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          double2 sf[3], bf[3];
          const double2 svf = ld2(ws.Sv + 8 * kb + 2 * c);
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            sf[jb] = ld2(ws.W + lo26 + t26(kb, jb));
            bf[jb] = ld2(B + lo24 + t24(kb, jb));
            pG[jb] = fma(bf[jb].x, svf.x, pG[jb]);
            pG[jb] = fma(bf[jb].y, svf.y, pG[jb]);
          }
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) dmma2(g[ib][jb], sf[ib], bf[jb]);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) pG[i] = quad_sum(pG[i]);
        if (c < 3) ws.Gv[8 * c + r] = rv[8 * c + r] + pick3(pG, c);
        __syncwarp();  // every lane is done reading S- from the scratch
      }
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) tput(ws.W, ib, jb, r, c, h[lt(ib, jb)]);

      // ---- G' = P' + SA'B (op fragments of G) ----
      if (!ev) {
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          double2 bf[3];
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) bf[jb] = ld2(B + lo24 + t24(kb, jb));
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) dmma2(g[ib][jb], zA[ib][kb], bf[jb]);
        }
      }
      __syncwarp();  // all lanes are done with the staged operand block: refill it for node k-1 while this stage finishes
      if (MODE != kModeLM && lane == 0 && k >= 1) {  // (LM needs A, B, Hv once more after the gains: its refill follows Acl)
        mbar_expect_tx(&ws.full, opBytes);
        tma_load(ws.in, lqp + (size_t)((kAblate & 2) ? N - 1 : k - 1) * rec_stride, opBytes, &ws.full);
      }

      // ---- blocked Cholesky of Hm and L^-T into the scratch ----
      if (!(kAblate & 1)) pd = factor_hm(ws.W, lane, li, r, c) && pd;

      // ---- Yv = L^-1 Gv ----
      {
        double z[3];
        matvec_cols<kLd, true>(ws.W, ws.Gv, r, c, z);  // (L^-1 Gv)[j] = sum_k (L^-T)[k][j] Gv[k]
        if (c < 3) {
          const double yv = pick3(z, c);
          ws.Yv[8 * c + r] = yv;
          if (!ev) spart = fma(-0.5 * yv, yv, spart);
        }
        __syncwarp();
      }
      // ---- dbias = -L^-T Yv, s (with constraints: after the projection, on the constrained Yv) ----
      auto finish_vectors = [&]() {
        double z[3];
        // (L^-T Yv)[8jb + r] = sum_{kb >= jb} (L^-T)[8jb+r][8kb+2c..] Yv[8kb+2c..]
        double2 vf[3];
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) vf[kb] = ld2(ws.Yv + 8 * kb + 2 * c);
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) {
          double p = 0.0;
#pragma unroll
          for (int kb = jb; kb < 3; ++kb) {
            const double2 mv = tfrag(ws.W, jb, kb, r, c);
            p = fma(mv.x, vf[kb].x, p);
            p = fma(mv.y, vf[kb].y, p);
          }
          z[jb] = quad_sum(p);
        }
        if (c < 3) {
          const int j = 8 * c + r;
          __stcg(out + kOdb + j, -pick3(z, c));
          if (!NOM) __stcg(out + kObias + j, 0.0);
          if (MODE == kModeLM) ws.ub[j] = -pick3(z, c);
        }
        if (MODE == kModeLM) {  // hcl = Hv + B dbias (one row per lane), kept in xb for Sv; s -= mu/2 |hcl|^2
          __syncwarp();
          const double hcl = Hv[li] + matvec_rows(B, ws.ub, li);
          if (lane < kN) {
            ws.xb[lane] = hcl;
            spart = fma(-0.5 * a.mu * hcl, hcl, spart);
          }
          __syncwarp();
        }
        sval = sval + cval + warp_sum_all(spart);
        if (lane == 0) __stcg(out + kOs, sval);
      };
      if (NCB == 0) finish_vectors();

      // ---- Y' = G' L^-T (op fragments of Y); L^-T is block upper triangular ----
      double2 y[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) y[i][j] = zero2();
#pragma unroll
      for (int jb = 0; jb < 3; ++jb)
#pragma unroll
        for (int kb = 0; kb <= jb; ++kb) {
          const double2 lf = ld2(ws.W + lo26 + t26(kb, jb));
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) dmma2(y[ib][jb], g[ib][kb], lf);
        }

      // ---- state-input equality constraints C x + D u + e = 0: the range-space form of the reference's projection
      //      (LinearAlgebra::computeConstraintProjection, LinearAlgebra.cpp:129-155; GaussNewtonDDP.cpp:751-773), as in riccati_rpl.cu:
      //      Z = L^-1 D', M = Z'Z = L_M L_M' (L_M' is the R factor of the reference's QR of U^-T D'), Vx = L_M^-1 (Z'Y - C),
      //      vv = L_M^-1 (Z'Yv - e); the constrained minimiser is Y^ = Y - Z L_M^-T Vx, Yv^ = Yv - Z L_M^-T vv and the value function
      //      gains + Vx'Vx, + Vx'vv, + 1/2 vv'vv. Everything on 8-row constraint tiles; rows beyond the node's active count are zero rows
      //      of D, C, e with a unit diagonal in M, which leaves every result untouched (ragged counts cost nothing). ----
      double zsv[3] = {0.0, 0.0, 0.0};  // Y'Yv - Vx'vv
      if (NCB > 0) {
        mbar_wait(&ws.cfull, cphase);
        cphase ^= 1u;
        const int ncm = a.ncmax;
        const int nca = ncp ? __ldg(ncp + k) : ncm;
        const double* Cc = ws.cde;
        const double* Dd = ws.cde + a.cdeD;
        const double* ee = ws.cde + a.cdeE;
        // op fragment of D' tile (kb, ib): lane (r,c) holds D'[8kb+2c..2c+1][8ib+r] = D[8ib+r][8kb+2c..2c+1]
        auto dfrag = [&](int ib, int kb) -> double2 {
          const int row = 8 * ib + r;
          if (row >= nca) return zero2();
          const double* q = Dd + row + ncm * (8 * kb + 2 * c);
          return make_double2(q[0], q[ncm]);
        };
        // Z' = D L^-T (accumulators = op fragments of Z)
        double2 zp[NCB_][3];
#pragma unroll
        for (int ib = 0; ib < NCB; ++ib) {
          double2 df[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) df[kb] = dfrag(ib, kb);
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            zp[ib][jb] = zero2();
#pragma unroll
            for (int kb = 0; kb <= jb; ++kb) dmma2(zp[ib][jb], df[kb], ld2(ws.W + lo26 + t26(kb, jb)));
          }
        }
        // M = Z'Z (lower tiles) into its scratch, unit diagonal on the inactive rows; L_M^-T by the blocked factorisation
#pragma unroll
        for (int ib = 0; ib < NCB; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            double2 mm = zero2();
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) dmma2(mm, zp[ib][kb], zp[jb][kb]);
            if (ib == jb && 8 * ib + r >= nca) {
              mm.x += (2 * c == r) ? 1.0 : 0.0;
              mm.y += (2 * c + 1 == r) ? 1.0 : 0.0;
            }
            tput(ws.Wm, ib, jb, r, c, mm);
          }
        __syncwarp();
        rank_ok = factor_blocks<NCB_, true>(ws.Wm, lane, r, c) && rank_ok;
        // R' = Y'Z - C' (n x 8 NCB; accumulators = op fragments of R = Z'Y - C)
        double2 rp[3][NCB_];
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb < NCB; ++jb) {
            const int row = 8 * jb + 2 * c;
            const double* q = Cc + row + ncm * (8 * ib + r);
            rp[ib][jb] = make_double2(row < nca ? -q[0] : 0.0, row + 1 < nca ? -q[1] : 0.0);
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) dmma2(rp[ib][jb], y[ib][kb], zp[jb][kb]);
          }
        // w = Z'Yv - e, vv = L_M^-1 w
        {
          double2 yvf[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) yvf[kb] = ld2(ws.Yv + 8 * kb + 2 * c);
#pragma unroll
          for (int ib = 0; ib < NCB; ++ib) {
            double q = 0.0;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              q = fma(zp[ib][kb].x, yvf[kb].x, q);
              q = fma(zp[ib][kb].y, yvf[kb].y, q);
            }
            q = quad_sum(q);
            const int row = 8 * ib + r;
            if (c == 0) ws.vw[row] = row < nca ? q - ee[row] : 0.0;
          }
          __syncwarp();
          double zz[NCB_];
          matvec_cols<kLd, true, NCB_>(ws.Wm, ws.vw, r, c, zz);
          if (c < NCB) {
            const double v = (NCB_ == 1 || c == 0) ? zz[0] : zz[NCB_ - 1];
            ws.vv[8 * c + r] = v;
            if (!ev) spart = fma(0.5 * v, v, spart);  // (a pre-event node's value function comes from the transversality condition alone)
          }
          __syncwarp();
        }
        // Vx' = R' L_M^-T (accumulators = op fragments of Vx)
        double2 vx[3][NCB_];
#pragma unroll
        for (int jb = 0; jb < NCB; ++jb)
#pragma unroll
          for (int kb = 0; kb <= jb; ++kb) {
            const double2 lf = ld2(ws.Wm + lo26 + t26(kb, jb));
#pragma unroll
            for (int ib = 0; ib < 3; ++ib) {
              if (kb == 0) vx[ib][jb] = zero2();
              dmma2(vx[ib][jb], rp[ib][kb], lf);
            }
          }
        // the value-function terms take the UNPROJECTED Y: zsv = Y'Yv - Vx'vv ; t = Y'Y - T - Vx'Vx
        {
          double2 yvf[3], vvf[NCB_];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) yvf[kb] = ld2(ws.Yv + 8 * kb + 2 * c);
#pragma unroll
          for (int kb = 0; kb < NCB; ++kb) vvf[kb] = ld2(ws.vv + 8 * kb + 2 * c);
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            double q = 0.0;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              q = fma(y[cb][kb].x, yvf[kb].x, q);
              q = fma(y[cb][kb].y, yvf[kb].y, q);
            }
#pragma unroll
            for (int kb = 0; kb < NCB; ++kb) {
              q = fma(-vx[cb][kb].x, vvf[kb].x, q);
              q = fma(-vx[cb][kb].y, vvf[kb].y, q);
            }
            zsv[cb] = quad_sum(q);
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) t[i] = neg2(t[i]);
        if (!ev) {
#pragma unroll
          for (int kb = 0; kb < 3; ++kb)
#pragma unroll
            for (int ib = 0; ib < 3; ++ib)
#pragma unroll
              for (int jb = 0; jb <= ib; ++jb) dmma2(t[lt(ib, jb)], y[ib][kb], y[jb][kb]);
#pragma unroll
          for (int kb = 0; kb < NCB; ++kb)
#pragma unroll
            for (int ib = 0; ib < 3; ++ib) {
              const double2 nv = neg2(vx[ib][kb]);
#pragma unroll
              for (int jb = 0; jb <= ib; ++jb) dmma2(t[lt(ib, jb)], nv, vx[jb][kb]);
            }
        }
        // U' = Vx' L_M^-1 (accumulators = op fragments of U = L_M^-T Vx)
        double2 up[3][NCB_];
#pragma unroll
        for (int jb = 0; jb < NCB; ++jb)
#pragma unroll
          for (int kb = jb; kb < NCB; ++kb) {
            const double2 lf = tfrag(ws.Wm, jb, kb, r, c);  // L_M^-1[8kb+2c..][8jb+r] = L_M^-T[8jb+r][8kb+2c..]
#pragma unroll
            for (int ib = 0; ib < 3; ++ib) {
              if (kb == jb) up[ib][jb] = zero2();
              dmma2(up[ib][jb], vx[ib][kb], lf);
            }
          }
        // Z = L^-1 D' (accumulators = op fragments of Z')
        double2 zq[3][NCB_];
#pragma unroll
        for (int jb = 0; jb < NCB; ++jb) {
          double2 df[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) df[kb] = dfrag(jb, kb);
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            zq[ib][jb] = zero2();
#pragma unroll
            for (int kb = 0; kb <= ib; ++kb) dmma2(zq[ib][jb], ld2(ws.W + lo26 + t26(kb, ib)), df[kb]);
          }
        }
        // Y^' = Y' - U'Z'
#pragma unroll
        for (int kb = 0; kb < NCB; ++kb)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double2 nu = neg2(up[ib][kb]);
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) dmma2(y[ib][jb], nu, zq[jb][kb]);
          }
        // Yv^ = Yv - Z (L_M^-T vv)
        {
          double2 vvf[NCB_];
#pragma unroll
          for (int kb = 0; kb < NCB; ++kb) vvf[kb] = ld2(ws.vv + 8 * kb + 2 * c);
#pragma unroll
          for (int jb = 0; jb < NCB; ++jb) {
            double q = 0.0;
#pragma unroll
            for (int kb = jb; kb < NCB; ++kb) {
              const double2 mv = tfrag(ws.Wm, jb, kb, r, c);
              q = fma(mv.x, vvf[kb].x, q);
              q = fma(mv.y, vvf[kb].y, q);
            }
            q = quad_sum(q);
            if (c == 0) ws.vw[8 * jb + r] = q;
          }
          __syncwarp();
          double2 tf[NCB_];
#pragma unroll
          for (int jb = 0; jb < NCB; ++jb) tf[jb] = ld2(ws.vw + 8 * jb + 2 * c);
          double zt[3];
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            double q = 0.0;
#pragma unroll
            for (int jb = 0; jb < NCB; ++jb) {
              q = fma(zq[ib][jb].x, tf[jb].x, q);
              q = fma(zq[ib][jb].y, tf[jb].y, q);
            }
            zt[ib] = quad_sum(q);
          }
          if (c < 3) ws.Yv[8 * c + r] -= pick3(zt, c);
          __syncwarp();
        }
        finish_vectors();
      }

      // ---- K' = -Y' L^-1 (op fragments of K -> 16-byte global stores); L^-1 is block lower triangular ----
      double2 zc[3][3];  // LM: -Acl' = -(A + B K)' as accumulators = operand fragments of -Acl
      {
        double2 kk[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) kk[i][j] = zero2();
#pragma unroll
        for (int jb = 0; jb < 3; ++jb)
#pragma unroll
          for (int kb = jb; kb < 3; ++kb) {
            const double2 lf = tfrag(ws.W, jb, kb, r, c);  // L^-1[8kb+2c..][8jb+r] = L^-T[8jb+r][8kb+2c..]
#pragma unroll
            for (int ib = 0; ib < 3; ++ib) dmma2(kk[ib][jb], y[ib][kb], lf);
          }
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb < 3; ++jb)
            if (!(kAblate & 8)) stg2(out + kOK + lo24 + t24(jb, ib), neg2(kk[ib][jb]));
        if (NOM) {
          // bias = u_nom - K x_nom (GaussNewtonDDP.cpp:604-606). The accumulators hold -K' = Y'L^-1: lane (r,c) owns the terms of
          // state 8ib + r for inputs 8jb + 2c, 2c+1; the sum over the states is a butterfly over r
          const double* xn = a.x_nom + ((size_t)prob * (N + 1) + k) * kN;
          const double* un = a.u_nom + ((size_t)prob * (N + 1) + k) * kN;
          double xr[3];
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) xr[ib] = __ldg(xn + 8 * ib + r);
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            double px = 0.0, py = 0.0;
#pragma unroll
            for (int ib = 0; ib < 3; ++ib) {
              px = fma(kk[ib][jb].x, xr[ib], px);
              py = fma(kk[ib][jb].y, xr[ib], py);
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              px += __shfl_xor_sync(kFull, px, o);
              py += __shfl_xor_sync(kFull, py, o);
            }
            if (r == 0) {
              const double2 u2 = *reinterpret_cast<const double2*>(un + 8 * jb + 2 * c);
              stg2(out + kObias + 8 * jb + 2 * c, make_double2(u2.x + px, u2.y + py));
            }
          }
        }
        if (MODE == kModeLM) {
          // -Acl' = -A' - K'B' = -A' + (-K)'B': X = -K (the accumulators kk are its operand fragments), Y = B' (transposed reads of B)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) zc[ib][jb] = neg2(ld2(A + lo24 + t24(jb, ib)));  // A'[8ib+r][8jb+2c..] = A[8jb+2c..][8ib+r]
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            double2 bt[3];
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) {  // B'[8kb+2c..2c+1][8jb+r] = B[8jb+r][8kb+2c..2c+1]
              const double* bp = B + (8 * jb + r) + kN * (8 * kb + 2 * c);
              bt[jb] = make_double2(bp[0], bp[kN]);
            }
#pragma unroll
            for (int ib = 0; ib < 3; ++ib)
#pragma unroll
              for (int jb = 0; jb < 3; ++jb) dmma2(zc[ib][jb], kk[ib][kb], bt[jb]);
          }
          __syncwarp();  // A, B, Hv are consumed: the operand slot takes node k-1
          if (lane == 0 && k >= 1) {
            mbar_expect_tx(&ws.full, opBytes);
            tma_load(ws.in, lqp + (size_t)(k - 1) * rec_stride, opBytes, &ws.full);
          }
        }
      }

      // ---- Sv = tv - Y'Yv (+ Vx'vv) ----
      {
        double z[3];
        if (NCB == 0) {
          double2 vf[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) vf[kb] = ld2(ws.Yv + 8 * kb + 2 * c);
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            double p = 0.0;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              p = fma(y[cb][kb].x, vf[kb].x, p);
              p = fma(y[cb][kb].y, vf[kb].y, p);
            }
            z[cb] = quad_sum(p);
          }
        } else {
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) z[cb] = zsv[cb];
        }
        if (MODE == kModeLM) {  // - mu Acl'hcl: the fragments hold -Acl
          double2 hf[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) hf[kb] = ld2(ws.xb + 8 * kb + 2 * c);
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            double p = 0.0;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              p = fma(zc[cb][kb].x, hf[kb].x, p);
              p = fma(zc[cb][kb].y, hf[kb].y, p);
            }
            z[cb] = fma(-a.mu, quad_sum(p), z[cb]);
          }
        }
        if (c < 3) {
          const int j = 8 * c + r;
          svn = ev ? tvj : tvj - pick3(z, c);
          ws.Sv[j] = svn;
          __stcg(out + kOSv + j, svn);
        }
      }

      // ---- S = T - Y'Y (+ Vx'Vx) (lower tiles): accumulate Y'Y - T, flip the sign (constrained: done with the unprojected Y above) ----
      if (NCB == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) t[i] = neg2(t[i]);
      }
      if (NCB == 0 && !ev) {
#pragma unroll
        for (int kb = 0; kb < 3; ++kb)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(t[lt(ib, jb)], y[ib][kb], y[jb][kb]);
      }
      if (MODE == kModeLM) {  // + mu Acl'Acl
#pragma unroll
        for (int kb = 0; kb < 3; ++kb)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double2 zs = make_double2(a.mu * zc[ib][kb].x, a.mu * zc[ib][kb].y);
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(t[lt(ib, jb)], zs, zc[jb][kb]);
          }
      }
      if (MODE == kModeGersh && !ev) {
        // dQ = makePsdGershgorin(M) - M on M = Q~ - P~'P~ = Q - Yp'Yp, Yp = L^-1 P (LineSearchStrategy.cpp:294-312): only its diagonal
        // max(0, R_i + eps - M_ii) differs from rounding noise. Yp' = P'L^-T like Y', M like S; P and Q come from L2 once more.
        double2 yp[3][3];
#pragma unroll
        for (int ib = 0; ib < 3; ++ib) {
          double2 pf[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) pf[kb] = ldg2(rec + kOP + lo24 + t24(kb, ib));
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            yp[ib][jb] = zero2();
#pragma unroll
            for (int kb = 0; kb <= jb; ++kb) dmma2(yp[ib][jb], pf[kb], ld2(ws.W + lo26 + t26(kb, jb)));
          }
        }
        double2 mq[6];
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) mq[lt(ib, jb)] = ldg2(rec + kOQ + lo24 + t24(jb, ib));
#pragma unroll
        for (int kb = 0; kb < 3; ++kb)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double2 ny = neg2(yp[ib][kb]);
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(mq[lt(ib, jb)], ny, yp[jb][kb]);
          }
        // Gershgorin radii: lane (r,c) holds M[8ib+r][8jb+2c..2c+1] of the lower tiles; a row's sum takes the row sums of its own
        // tiles (over c) and, by symmetry, the column sums of the tiles below it (over r)
        double rowp[3] = {0.0, 0.0, 0.0};
        double2 colp[2] = {zero2(), zero2()};
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            const double ax = fabs(mq[lt(ib, jb)].x), ay = fabs(mq[lt(ib, jb)].y);
            rowp[ib] += ax + ay;
            if (ib != jb) {
              colp[jb].x += ax;
              colp[jb].y += ay;
            }
          }
#pragma unroll
        for (int i = 0; i < 3; ++i) rowp[i] = quad_sum(rowp[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            colp[j].x += __shfl_xor_sync(kFull, colp[j].x, o);
            colp[j].y += __shfl_xor_sync(kFull, colp[j].y, o);
          }
        if (c == 0) {
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) ws.Gv[8 * ib + r] = rowp[ib];  // Gv is free between the gain computation and the next stage
        }
        __syncwarp();
        if (r == 0) {
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            ws.Gv[8 * jb + 2 * c] += colp[jb].x;
            ws.Gv[8 * jb + 2 * c + 1] += colp[jb].y;
          }
        }
        __syncwarp();
        if (2 * c == r || 2 * c + 1 == r) {  // this lane holds the diagonal element of row 8ib + r in its diagonal tiles
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double mii = (2 * c == r) ? mq[lt(ib, ib)].x : mq[lt(ib, ib)].y;
            const double dq = fmax(0.0, ws.Gv[8 * ib + r] - fabs(mii) + a.eps - mii);
            if (2 * c == r) t[lt(ib, ib)].x -= dq; else t[lt(ib, ib)].y -= dq;  // t holds Y'Y - T here: S = -t gains + dq
          }
        }
      }
      __syncwarp();  // every lane is done reading L^-T from the scratch
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) {
          const double2 sv = neg2(t[lt(ib, jb)]);  // S[8ib+r][8jb+2c..] = S[8jb+2c..][8ib+r]
          t[lt(ib, jb)] = sv;
          st2(ws.W + lo26 + t26(jb, ib), sv);
          if (!(kAblate & 8)) stg2(out + kOSm + lo24 + t24(jb, ib), sv);
          if (ib != jb) {
            tput(ws.W, ib, jb, r, c, sv);
            double* gq = out + kOSm + (8 * ib + r) + kN * (8 * jb + 2 * c);
            if (!(kAblate & 8)) {
              __stcg(gq, sv.x);
              __stcg(gq + kN, sv.y);
            }
          }
        }
      __syncwarp();
      if (NCB > 0 && lane == 0 && k >= 1) {  // every lane is done with this node's {C | D | e}
        mbar_expect_tx(&ws.cfull, a.cde_bytes);
        tma_load(ws.cde, lqp + (size_t)(k - 1) * rec_stride + a.oC, a.cde_bytes, &ws.cfull);
      }
    }

    // node N of the controller := node N-1 (GaussNewtonDDP.cpp:609-618)
    {
      const double* src = solp + (size_t)(N - 1) * kORec;
      double* dst = solp + (size_t)N * kORec;
#pragma unroll 1
      for (int i = lane; i < (kMat + 2 * kN) / 2; i += 32) stg2(dst + kOK + 2 * i, ldg2(src + kOK + 2 * i));  // K | dbias | bias are contiguous
    }

    // ---- status: a non-finite value anywhere in the sweep propagates into S, Sv, s of node 0 ----
    int bits = 0;
    {
      bool finite = finite_bits(sval) && finite_bits(svn);
#pragma unroll
      for (int i = 0; i < 6; ++i) finite = finite && finite2(t[i]);
      if (!__all_sync(kFull, pd)) bits |= O2C_STATUS_CHOL_NOT_PD;
      if (!__all_sync(kFull, finite)) bits |= O2C_STATUS_NONFINITE;
      if (NCB > 0 && !__all_sync(kFull, rank_ok)) bits |= O2C_STATUS_CONSTRAINT_RANK;
    }
    if (lane == 0) a.status[prob] = bits;
#ifdef O2C_WPP_STATS
    st_sweep += clock64() - st_mark, ++st_nsweep, st_mark = clock64();
#endif
    const int done_pi = pi;
    pi = warp < nsweep ? next_problem(a, lane, total_warps, pi, slot, tail_taken) : a.count;  // (a roller sweeps one problem only)
    if (!a.with_rollout) continue;
    if (nroll == 0) {
      job = done_pi;
    } else {
      // hand the problem to the rollers: every lane's K / dbias / bias stores are ordered before the queue entry becomes visible
      __threadfence();
      fence_proxy_async_global();  // ... and before the rollers' TMA reads of them
      __syncwarp();
      if (lane == 0) {
        const unsigned ticket = atomicAdd(&q_tail, 1u);
        volatile unsigned long long* entry = &queue[ticket % kQueue];
        while ((int)(ticket - *(volatile unsigned*)&q_head) >= kQueue || *entry != 0ull) {  // ring full: the rollers are behind (the head may
                                                                                           // already be past this ticket: a roller claims it first)
          __nanosleep(100);
        }
        *entry = ((unsigned long long)(ticket + 1u) << 32) | (unsigned)done_pi;
      }
      continue;
    }
    }  // sweeping

    if (job < 0) {
      // roller: next finished problem of this CTA, or -1 once every sweeper is done and the queue is empty
      if (lane == 0) {
        for (;;) {
          const unsigned done = *(volatile unsigned*)&sweepers_done;  // read before the tail: every push precedes its sweeper's done
          const unsigned h = *(volatile unsigned*)&q_head, t = *(volatile unsigned*)&q_tail;
          if (h != t) {
            if (atomicCAS(&q_head, h, h + 1u) != h) continue;
            volatile unsigned long long* entry = &queue[h % kQueue];
            unsigned long long v;
            while ((unsigned)((v = *entry) >> 32) != h + 1u) {
            }
            *entry = 0ull;
            job = (int)(unsigned)v;
            break;
          }
          if (done == (unsigned)done_target) break;
          __nanosleep(200);
        }
      }
      job = __shfl_sync(kFull, job, 0);
#ifdef O2C_WPP_STATS
      st_wait += clock64() - st_mark, st_mark = clock64();
#endif
      if (job < 0) break;
      __threadfence();
    }
    const int prob = a.begin + job;
    const double* lqp = a.lq + (size_t)prob * N * rec_stride;
    double* solp = a.sol + (size_t)prob * (N + 1) * kORec;
    const int* evp = EV ? a.event + (size_t)prob * N : nullptr;

    // ---- forward rollout of the LQ model: du_k = K_k dx_k + alpha dbias_k, dx_{k+1} = A_k dx_k + B_k du_k + Hv_k. One row per lane.
    // Everything a stage reads — {A | B | Hv} of the LQ record and {K | dbias} of the solution record, 14.2 KB — arrives by two TMA bulk
    // copies in a ring of `depth` stage sets in shared memory, issued `depth` stages ahead. The stage itself is a short dependent chain
    // (two 24-term dot products per lane, ~0.3 us), so a rollout runs at DRAM latency / depth per stage: nothing waits in L2 (with 1600
    // sweeps in flight, lines prefetched into L2 ahead of a rollout were evicted before their use and every stage paid a full DRAM round
    // trip: 2.3 - 2.7 us per stage, profiles/r02_wpp_stats.log). Rollers own a deep ring behind the sweepers' slots; a sweeper rolling
    // out (no rollers, or the tail of a launch) uses its own slot as a ring of one.
    const bool own_slot = warp < nsweep;
    const int depth = own_slot ? 1 : a.ring_depth;
    double* ring = own_slot ? ws.in : reinterpret_cast<double*>(smem_raw + (size_t)nsweep * sizeof(WarpSmem)) + (size_t)(warp - nsweep) * a.ring_depth * kRing;
    unsigned long long* rfull = ring_full[warp];
    double* xb = rvec[warp][0];
    double* ub = rvec[warp][1];
    const uint32_t lqBytes = (2 * kMat + kN) * sizeof(double), solBytes = (kMat + kN) * sizeof(double);
    if (lane == 0) {
      fence_proxy_async();        // the slot was last touched through the generic proxy
      fence_proxy_async_global();  // K, dbias were written by ordinary stores (this warp or a sweeper of this CTA) and are read by TMA
      for (int d = 0; d < depth && d < N; ++d) {
        mbar_expect_tx(&rfull[d], lqBytes + solBytes);
        tma_load(ring + (size_t)d * kRing, lqp + (size_t)d * rec_stride, lqBytes, &rfull[d]);
        tma_load(ring + (size_t)d * kRing + kRingK, solp + (size_t)d * kORec + kOK, solBytes, &rfull[d]);
      }
    }
    double* xo = a.xs + (size_t)prob * (N + 1) * kN;
    double* uo = a.us + (size_t)prob * (N + 1) * kN;
    // with nominal trajectories the rollout runs in deviation coordinates dx = x - x_nom, du = u - u_nom (the LQ model is
    // x_{k+1} = x_nom_{k+1} + A dx + B du + Hv, u = u_nom + K dx + alpha dbias) and the outputs are shifted back
    const double* xnp = NOM ? a.x_nom + (size_t)prob * (N + 1) * kN : nullptr;
    const double* unp = NOM ? a.u_nom + (size_t)prob * (N + 1) * kN : nullptr;
    double xnk = NOM ? __ldg(xnp + li) : 0.0, unk = NOM ? __ldg(unp + li) : 0.0;  // nominal state / input of the current node
    double x = a.x0[(size_t)prob * kN + li] - xnk;
    if (lane < kN) xb[lane] = x;
    bool xfinite = true;
    bool jump = EV && __ldg(evp) != 0;  // pre-event node: x+ = A_e x + Hv_e, the input does not enter the jump map
    int rs = 0;                         // ring slot of stage k
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      const double* S = ring + (size_t)rs * kRing;
      const double* Kk = S + kRingK;
      mbar_wait(&rfull[rs], (rphase >> rs) & 1u);
      rphase ^= 1u << rs;
      // u = alpha dbias + K x ; ax = Hv + A x (independent of u)
      double u0 = a.alpha * Kk[kMat + li], u1 = 0.0, u2 = 0.0, u3 = 0.0;
      double a0 = S[2 * kMat + li], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int j = 0; j < kN; j += 4) {
        const double2 x01 = ld2(xb + j), x23 = ld2(xb + j + 2);
        u0 = fma(Kk[li + kN * j], x01.x, u0);
        u1 = fma(Kk[li + kN * (j + 1)], x01.y, u1);
        u2 = fma(Kk[li + kN * (j + 2)], x23.x, u2);
        u3 = fma(Kk[li + kN * (j + 3)], x23.y, u3);
        a0 = fma(S[li + kN * j], x01.x, a0);
        a1 = fma(S[li + kN * (j + 1)], x01.y, a1);
        a2 = fma(S[li + kN * (j + 2)], x23.x, a2);
        a3 = fma(S[li + kN * (j + 3)], x23.y, a3);
      }
      const double u = (u0 + u1) + (u2 + u3);
      if (lane < kN) {
        ub[lane] = u;
        __stcg(xo + (size_t)k * kN + lane, x + xnk);
        __stcg(uo + (size_t)k * kN + lane, u + unk);
      }
      xfinite = xfinite && finite_bits(x);
      if (NOM) {  // next node's nominal values: in flight during the second half of the stage
        xnk = __ldg(xnp + (size_t)(k + 1) * kN + li);
        unk = __ldg(unp + (size_t)(k + 1) * kN + li);
      }
      const bool jump_next = EV && k + 1 < N && __ldg(evp + k + 1) != 0;
      __syncwarp();  // u of every lane is in shared memory
      const double xn = ((a0 + a1) + (a2 + a3)) + (jump ? 0.0 : matvec_rows(S + kMat, ub, li));
      jump = jump_next;
      __syncwarp();  // every lane is done with x, u and the stage set
      x = xn;
      if (lane < kN) xb[lane] = x;
      if (lane == 0 && k + depth < N) {
        mbar_expect_tx(&rfull[rs], lqBytes + solBytes);
        tma_load(ring + (size_t)rs * kRing, lqp + (size_t)((kAblate & 16) ? 0 : k + depth) * rec_stride, lqBytes, &rfull[rs]);
        tma_load(ring + (size_t)rs * kRing + kRingK, solp + (size_t)((kAblate & 16) ? 0 : k + depth) * kORec + kOK, solBytes, &rfull[rs]);
      }
      __syncwarp();
      rs = rs + 1 == depth ? 0 : rs + 1;
    }
    // node N: state, and the input of the copied last policy (K, dbias of node N-1, still in their ring slot) re-evaluated at x_N
    // (TimeTriggeredRollout.cpp:98-102)
    {
      const double* Kl = ring + (size_t)((N - 1) % depth) * kRing + kRingK;
      const double xabs = x + xnk;
      double u0 = a.alpha * Kl[kMat + li], u1 = 0.0;
      if (NOM) {  // the copied policy of node N-1 is evaluated at the absolute state: u = bias + alpha dbias + K x
        __syncwarp();
        if (lane < kN) xb[lane] = xabs;
        __syncwarp();
        u0 += __ldcg(solp + (size_t)N * kORec + kObias + li);
      }
#pragma unroll
      for (int j = 0; j < kN; j += 2) {
        const double2 x01 = ld2(xb + j);
        u0 = fma(Kl[li + kN * j], x01.x, u0);
        u1 = fma(Kl[li + kN * (j + 1)], x01.y, u1);
      }
      if (lane < kN) {
        __stcg(xo + (size_t)N * kN + lane, xabs);
        __stcg(uo + (size_t)N * kN + lane, u0 + u1);
      }
      xfinite = xfinite && finite_bits(x);
    }
    if (!__all_sync(kFull, xfinite) && lane == 0) atomicOr(a.status + prob, O2C_STATUS_NONFINITE);
    __syncwarp();
#ifdef O2C_WPP_STATS
    st_roll += clock64() - st_mark, ++st_nroll;
#endif
  }
#ifdef O2C_WPP_STATS
  if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
    printf("cta %3d warp %2d: %3d sweeps %8.1f us, %3d rollouts %8.1f us, queue wait %8.1f us, total %8.1f us\n", blockIdx.x, warp, st_nsweep,
           st_sweep / 1965.0, st_nroll, st_roll / 1965.0, st_wait / 1965.0, (clock64() - st_t0) / 1965.0);
#endif
}

}  // namespace

bool wpp_ilqr_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  // Both Riccati forms are served: under LINE_SEARCH the full form (preComputeRiccatiTerms = false) is the same map written with
  // K~'G~ + G~'K~ + K~'H~K~ in place of -G~'G~ (H~ = Pu'Hm Pu = I); the reference's own RiccatiTest.cpp:87-105 holds them equal to 1e-9.
  const bool ls = st.strategy == O2C_STRATEGY_LINE_SEARCH && (st.hc == O2C_HC_DIAGONAL_SHIFT || st.hc == O2C_HC_GERSHGORIN_MODIFICATION);
  const bool lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT && buf.event == nullptr;  // (events under LM: refused by the API)
  // state-input equality constraints (up to 16, any per-node count, with or without events): LINE_SEARCH + DIAGONAL_SHIFT
  const bool constrained = L.ncmax > 0;
  const bool cons_ok = !constrained || (L.ncmax <= 16 && st.strategy == O2C_STRATEGY_LINE_SEARCH && st.hc == O2C_HC_DIAGONAL_SHIFT && L.oC == kRec);
  return L.n == kN && L.m == kN && cons_ok && st.algorithm == O2C_ALG_ILQR && (ls || lm) &&
         (buf.x_nom == nullptr) == (buf.u_nom == nullptr) && L.N >= 1 && (constrained || L.rec == kRec) && L.oQ == kOQ &&
         L.oP == kOP && L.oR == kOR && L.orec == kORec && L.oK == kOK && L.odb == kOdb && L.obias == kObias && L.oSm == kOSm &&
         L.oSv == kOSv && L.os == kOs && L.oA == 0 && L.oB == kMat && L.oHv == 2 * kMat;
}

// Sweeping warps per SM for a launch of `count` problems on `sms` SMs with `rollers` rollout warps per SM. The time a sweep takes
// depends on how many sweeps share its SM's FP64 pipe: tb(w) in ms for ONE round of w resident sweeps per SM (N = 100; B200,
// profiles/r02_wpp_residency2.jsonl; the steps at 4 -> 5 and 8 -> 9 are a scheduler taking its second / third warp). A full machine
// has the best steady-state throughput, but a batch of a few rounds pays for a ragged last round: 2048 problems on 148 SMs
// (BASELINE.json config 5: 16384 over 8 GPUs) are 13.8 per SM; with 11 sweepers + 1 roller that is a round of 12 and a round of 2,
// tb(12) + tb(2), while 7 sweepers + 1 roller run a round of 8 and a round of 6: tb(8) + tb(6), 2.2 ms instead of 2.5 ms. The choice
// minimises the modelled makespan (first round: w sweepers plus one sweep per roller; full rounds of w; the partial last round); only
// the ratios of tb matter. Batches of many rounds keep the full machine: the dynamic fetch de-synchronises the rounds and the
// steady-state throughput is what counts.
int choose_sweepers(int count, int sms, int rollers, int slots) {  // slots: sweepers the shared memory has room for
  static const double tb[kMaxWarps + 1] = {0.0, 0.80, 0.81, 0.83, 0.86, 1.00, 1.03, 1.08, 1.12, 1.30, 1.32, 1.36, 1.40};
  int wmax = kMaxCtaWarps - rollers < kMaxWarps ? kMaxCtaWarps - rollers : kMaxWarps;
  if (wmax > slots) wmax = slots;
  if (count >= 4L * wmax * sms) return wmax;
  int best = wmax;
  double best_time = 1e300;
  for (int w = wmax; w >= 1; --w) {
    long left = count;
    const long first = left < (long)(w + rollers) * sms ? left : (long)(w + rollers) * sms;
    double time = tb[(first + sms - 1) / sms];
    left -= first;
    const long full = left / ((long)w * sms);
    time += (double)full * tb[w];
    left -= full * w * sms;
    if (left > 0) time += tb[(left + sms - 1) / sms];
    if (time < best_time - 1e-9) {
      best_time = time;
      best = w;
    }
  }
  return best;
}

cudaError_t launch_ilqr_wpp(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, bool with_rollout, double alpha, int batch,
                            int begin, int count, cudaStream_t stream, int* launches) {
  if (!wpp_ilqr_supported(L, st, buf)) return cudaErrorNotSupported;
  const bool nom = buf.x_nom != nullptr, ev = buf.event != nullptr;
  using Kernel = void (*)(const Args);
  const int ncb = (L.ncmax + 7) / 8;  // constraint tiles: 0, 1 or 2
  const size_t slot_bytes = ncb == 0 ? sizeof(WarpSmemT<0>) : (ncb == 1 ? sizeof(WarpSmemT<1>) : sizeof(WarpSmemT<2>));
  Kernel kernel = nullptr;
#ifdef O2C_WPP_ONLY_BASE  // profiling builds: one instantiation, seconds to compile
  kernel = ilqr_wpp_kernel<false, false, kModeLS, 0>;
  (void)nom, (void)ev;
#else
  if (ncb == 0) {
    const Kernel kernels[3][4] = {
        {ilqr_wpp_kernel<false, false, kModeLS, 0>, ilqr_wpp_kernel<true, false, kModeLS, 0>, ilqr_wpp_kernel<false, true, kModeLS, 0>,
         ilqr_wpp_kernel<true, true, kModeLS, 0>},
        {ilqr_wpp_kernel<false, false, kModeLM, 0>, ilqr_wpp_kernel<true, false, kModeLM, 0>, nullptr, nullptr},
        {ilqr_wpp_kernel<false, false, kModeGersh, 0>, ilqr_wpp_kernel<true, false, kModeGersh, 0>, ilqr_wpp_kernel<false, true, kModeGersh, 0>,
         ilqr_wpp_kernel<true, true, kModeGersh, 0>}};
    const int mode = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT ? kModeLM : (st.hc == O2C_HC_GERSHGORIN_MODIFICATION ? kModeGersh : kModeLS);
    kernel = kernels[mode][(nom ? 1 : 0) + (ev ? 2 : 0)];
  } else if (ncb == 1) {
    const Kernel kernels[4] = {ilqr_wpp_kernel<false, false, kModeLS, 1>, ilqr_wpp_kernel<true, false, kModeLS, 1>, ilqr_wpp_kernel<false, true, kModeLS, 1>,
                               ilqr_wpp_kernel<true, true, kModeLS, 1>};
    kernel = kernels[(nom ? 1 : 0) + (ev ? 2 : 0)];
  } else {
    const Kernel kernels[4] = {ilqr_wpp_kernel<false, false, kModeLS, 2>, ilqr_wpp_kernel<true, false, kModeLS, 2>, ilqr_wpp_kernel<false, true, kModeLS, 2>,
                               ilqr_wpp_kernel<true, true, kModeLS, 2>};
    kernel = kernels[(nom ? 1 : 0) + (ev ? 2 : 0)];
  }
#endif
  if (kernel == nullptr) return cudaErrorNotSupported;
  const int num_sms = device_sm_count();
  if (num_sms <= 0) return cudaErrorInvalidDevice;
  // one round of 13 or 14 sweeps per SM instead of a round of 12 and a nearly empty one: the WIDE instantiation
  bool wide = false;
#ifndef O2C_WPP_ONLY_BASE
  if (ncb == 0 && !ev && st.strategy == O2C_STRATEGY_LINE_SEARCH && st.hc == O2C_HC_DIAGONAL_SHIFT && count > kMaxWarps * num_sms &&
      count <= kWideWarps * num_sms) {
    wide = true;
    if (const char* e = getenv("O2C_WPP_WIDE")) wide = atoi(e) != 0;
    if (wide) kernel = nom ? ilqr_wpp_kernel<true, false, kModeLS, 0, true> : ilqr_wpp_kernel<false, false, kModeLS, 0, true>;
  }
#endif
  // profiling knobs, read per launch (no state is cached in statics: launches from several host threads / on several devices are
  // independent): O2C_WPP_RESIDENT = sweeping warps per SM, O2C_WPP_ROLLERS = rollout-only warps per SM (0 = every sweeper rolls its own
  // problem out), O2C_WPP_DYNAMIC = 0 switches the dynamic problem fetch off
  // (two constraint tiles: shared memory has room for 8 sweepers or for 7 and a roller — 8 sweepers rolling their own problems out
  //  are 3 % faster, profiles/r02_legged_constraints.jsonl)
  int rollers = (with_rollout && !wide && ncb < 2) ? kDefaultRollers : 0;
  if (const char* e = getenv("O2C_WPP_ROLLERS")) {
    const int v = atoi(e);
    if (with_rollout && !wide && v >= 0 && v < kMaxCtaWarps) rollers = v;
  }
  cudaFuncAttributes fattr{};
  cudaError_t e = cudaFuncGetAttributes(&fattr, kernel);
  if (e != cudaSuccess) return e;
  const size_t smem_cap = (227 * 1024 - fattr.sharedSizeBytes) & ~(size_t)127;  // minus the static part (queue, ring barriers, x / u vectors)
  // sweeper slots the shared memory has room for next to the rollers' rings (two stage sets each at least, and room for the one sweep
  // a roller does first)
  const size_t ring_min = 2 * sizeof(double) * kRing > slot_bytes ? 2 * sizeof(double) * kRing : slot_bytes;
  const int slots = (int)((smem_cap - rollers * ring_min) / slot_bytes);
  int sweepers = choose_sweepers(count, num_sms, rollers, slots);
  if (const char* e = getenv("O2C_WPP_RESIDENT")) {
    const int v = atoi(e);
    if (v >= 1 && v <= kMaxCtaWarps) sweepers = v;
  }
  if (wide) sweepers = (count + num_sms - 1) / num_sms;  // 13 or 14: one round
  if (!wide && sweepers + rollers > kMaxCtaWarps) sweepers = kMaxCtaWarps - rollers;
  if (sweepers > slots) sweepers = slots;
  if (sweepers < 1) return cudaErrorInvalidConfiguration;
  const int warps = sweepers + rollers;
  // shared memory: the sweepers' slots, then the rollers' rings (as deep as fits, at most kMaxRingDepth stage sets)
  int ring_depth = 0;
  if (rollers > 0) {
    ring_depth = (int)((smem_cap - slot_bytes * sweepers) / (sizeof(double) * kRing * rollers));
    if (ring_depth > kMaxRingDepth) ring_depth = kMaxRingDepth;
    if (const char* e = getenv("O2C_WPP_RING")) {
      const int v = atoi(e);
      if (v >= 1 && v <= ring_depth) ring_depth = v;
    }
    if (ring_depth < 1) return cudaErrorInvalidConfiguration;
  }
  bool dynamic = true;
  if (const char* e = getenv("O2C_WPP_DYNAMIC")) dynamic = atoi(e) != 0;
  const size_t smem = slot_bytes * sweepers + sizeof(double) * kRing * rollers * ring_depth;
  // function attributes are per device: set on every launch (sub-microsecond) instead of caching "configured" in a static
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  Args a{};
  a.lq = buf.lq;
  a.term = buf.term;
  a.x0 = buf.x0;
  a.x_nom = buf.x_nom;
  a.u_nom = buf.u_nom;
  a.sol = buf.sol;
  a.xs = buf.xs;
  a.us = buf.us;
  a.status = buf.status;
  a.event = buf.event;
  a.nc = buf.nc;
  a.rec = L.rec;
  a.oC = L.oC;
  a.cdeD = L.oD - L.oC;
  a.cdeE = L.oe - L.oC;
  a.ncmax = L.ncmax;
  a.cde_bytes = (uint32_t)((L.oe + ((L.ncmax + 1) & ~1) - L.oC) * sizeof(double));
  a.N = L.N;
  a.oQf = L.oQf;
  a.oqf = L.oqf;
  a.ocf = L.ocf;
  a.trec = L.trec;
  a.begin = begin;
  a.count = count;
  a.with_rollout = with_rollout ? 1 : 0;
  a.nsweep = sweepers;
  a.sweep_count = count;
  a.ring_depth = ring_depth;
  a.eps = st.eps;
  a.alpha = alpha;
  a.mu = st.mu;
  (void)batch;
  // warp-major problem order (CTA b starts with problems b, b + grid, ...): the grid is as wide as the machine even when the batch
  // does not fill every warp slot, so a small batch spreads evenly over the SMs
  const int grid = count < num_sms ? count : num_sms;
  // the rollers' initial sweeps: whatever exceeds the sweepers' first static round, at most one problem per roller
  if (rollers > 0 && sizeof(double) * kRing * ring_depth >= slot_bytes && count > grid * sweepers) {
    const int extra = count - grid * sweepers;
    a.sweep_count = count - (extra < grid * rollers ? extra : grid * rollers);
  }
  if (const char* e = getenv("O2C_WPP_ROLLER_SWEEPS"))
    if (atoi(e) == 0) a.sweep_count = count;
  const int scount = a.sweep_count;
  a.counter = nullptr;
  if (dynamic && buf.work_counter != nullptr && scount > grid * sweepers) {
    e = cudaMemsetAsync(buf.work_counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    a.counter = buf.work_counter;
    a.dyn_limit = scount - scount % (grid * sweepers);
  }
  if (getenv("O2C_WPP_VERBOSE"))
    fprintf(stderr, "ilqr_wpp launch: sweepers/SM %d rollers/SM %d (ring %d) grid %d count %d dynamic %d\n", sweepers, rollers, ring_depth, grid, count, a.counter != nullptr);
  kernel<<<grid, 32 * warps, smem, stream>>>(a);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace o2c
