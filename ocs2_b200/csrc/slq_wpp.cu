// slq_wpp.cu — SLQ backward pass for nx = nu = 24 (the legged-robot shape) on the FP64 tensor pipe, warp per problem, sm_100a.
//
// The reference's legged example runs SLQ (ocs2_robotic_examples/ocs2_legged_robot/config/mpc/task.info:85); this is its backward pass for
// the unconstrained LINE_SEARCH / DIAGONAL_SHIFT configuration, in the three phases the reference itself has:
//
//   slq_project_kernel      SLQ::solveSequentialRiccatiEquations, the per-node projection loop (SLQ.cpp:174-208; node-parallel there too):
//                           Hm = R = L L', Pu = L^-T, B~ = B Pu, P~ = Pu'P, r~ = Pu'r (GaussNewtonDDP.cpp:734-782 with nc = 0,
//                           DDP_HelperFunctions.cpp:143-201) into a workspace record { B~ | P~ | L^-T | r~ } per (problem, node)
//   slq_flow_kernel         SLQ::riccatiEquationsWorker / integrateRiccatiEquationNominalTime (SLQ.cpp:213-302): classic RK4 over the
//                           host-built step schedule (boost::odeint integrate_times semantics, api.cu build_slq_schedule), four evaluations
//                           of ContinuousTimeRiccatiEquations::computeFlowMapSLQ (ContinuousTimeRiccatiEquations.cpp:152-292) per step on
//                           the projected data lerped between the two nodes of the interval. One warp per problem (time-sequential):
//                               dS  = Q~ + eps I + S'A~ + (S'A~)' - G'G,  G = P~ + B~'S
//                               dSv = q~ + S Hv + A~'Sv - G'Gv,           Gv = r~ + B~'Sv
//                               ds  = c~ + Hv.Sv - Gv.Gv / 2
//                           (reduced form, K~ = -G, L~ = -Gv under LINE_SEARCH; the full form is the same map). 162 DMMA per evaluation:
//                           the lower tiles of S'A + A'S (72), G' = P~' + S B~ as operand fragments of G (54), G'G lower (36).
//   slq_controller_kernel   SLQ::calculateControllerWorker (SLQ.cpp:127-169) + GaussNewtonDDP::calculateController (GaussNewtonDDP.cpp:588-618),
//                           node-parallel: K = -Pu (P~ + B~'S_k), dbias = -Pu (r~ + B~'Sv_k), bias = u_nom - K x_nom, node N := node N-1
//
// Fragment conventions are those of riccati_wpp.cu (wpp_tiles.cuh): every product is Z = X'Y on "op" fragments, the accumulator fragment
// of Z is the operand fragment of Z'. The value function lives in the lower tiles of accumulator layout; the stage argument of every
// evaluation goes through the 5 KB scratch (both triangles) to be re-read as operand fragments. The reference integrates the upper
// triangle of S (convert2Vector); this kernel integrates the lower tiles and mirrors them — dS is symmetric up to rounding.
#include <cstdio>
#include <cstdlib>

#include "o2c_common.cuh"
#include "wpp_tiles.cuh"

namespace o2c {
namespace {

// workspace record per (problem, node), 128-byte multiple
constexpr int kWB = 0, kWP = kMat, kWL = 2 * kMat, kWr = 3 * kMat, kWRec = 3 * kMat + 32;
constexpr int kTail = kOperand - 2 * kMat;  // { Hv | q | r | c, pad } = 80 doubles
constexpr int kSlot = kOperand + kN;        // { A | B~ | Hv | q | r | c, pad | r~ }
constexpr int kFlowWarps = 8;               // per SM: two 10 KB node slots + the scratch per warp = 25.5 KB
constexpr int kNodeWarps = 4;
constexpr int kStepCache = 256;             // RK4 steps of the schedule kept in shared memory (14 KB; longer schedules read the rest from global memory)

struct SlqWppArgs {
  const double* lq;
  const double* term;
  const double* x_nom;
  const double* u_nom;
  double* ws;
  double* sol;
  int* status;
  const SlqStep* steps;
  const double* jump;  // SLQ jump records [batch][jump_capacity] (events), or nullptr
  int jump_capacity;
  int nsteps, N, begin, count;
  int oQf, oqf, ocf, trec;
  double eps;
};

// ---------------------------------------------------------------------------------------------------------------------
// projection of every node (node-parallel)
// ---------------------------------------------------------------------------------------------------------------------
struct __align__(16) NodeSmem {
  double W[kN * kLd];
  double v[kN], y[kN];
};

__global__ void __launch_bounds__(32 * kNodeWarps) slq_project_kernel(const SlqWppArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  NodeSmem& ws = reinterpret_cast<NodeSmem*>(smem_raw)[warp];
  const int r = lane >> 2, c = lane & 3;
  const int lo24 = 2 * c + kN * r, lo26 = 2 * c + kLd * r;
  const int li = lane < kN ? lane : kN - 1;
  const int nodes = a.N + 1;
  const long long items = (long long)a.count * nodes;
  for (long long item = (long long)blockIdx.x * kNodeWarps + warp; item < items; item += (long long)gridDim.x * kNodeWarps) {
    const int prob = a.begin + (int)(item / nodes), node = (int)(item % nodes);
    const double* rec = a.lq + ((size_t)prob * nodes + node) * kRec;
    double* out = a.ws + ((size_t)prob * nodes + node) * kWRec;
    // Hm = R (SLQ.cpp:206-208): lower tiles into the scratch, r into shared memory
#pragma unroll
    for (int ib = 0; ib < 3; ++ib)
#pragma unroll
      for (int jb = 0; jb <= ib; ++jb) tput(ws.W, ib, jb, r, c, ldg2(rec + kOR + lo24 + t24(jb, ib)));  // R[8ib+r][8jb+2c..] = R[8jb+2c..][8ib+r]
    if (lane < kN) ws.v[lane] = rec[2 * kMat + 2 * kN + lane];
    __syncwarp();
    const bool pd = factor_hm(ws.W, lane, li, r, c);  // the scratch now holds L^-T in its block upper triangle
    // B~' = L^-1 B' = (L^-T)'B': X = L^-T, Y = B' (transposed reads of B); acc(B~')(ib, jb) = B~[8jb+2c..][8ib+r] -> column-major B~
    // P~  = L^-1 P  = (L^-T)'P : X = L^-T, Y = P;                          acc(P~)(ib, jb)  = P~[8ib+r][8jb+2c..]
    double2 bt[3][3], pt[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) bt[i][j] = pt[i][j] = zero2();
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) {
      double2 bf[3], pf[3];
#pragma unroll
      for (int jb = 0; jb < 3; ++jb) {
        const double* bp = rec + kMat + (8 * jb + r) + kN * (8 * kb + 2 * c);  // B'[8kb+2c..2c+1][8jb+r] = B[8jb+r][8kb+2c..2c+1]
        bf[jb] = make_double2(__ldcg(bp), __ldcg(bp + kN));
        pf[jb] = ldg2(rec + kOP + lo24 + t24(kb, jb));
      }
#pragma unroll
      for (int ib = kb; ib < 3; ++ib) {  // L^-T is block upper triangular: tiles (kb, ib) with kb <= ib
        const double2 lf = ld2(ws.W + lo26 + t26(kb, ib));
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) {
          dmma2(bt[ib][jb], lf, bf[jb]);
          dmma2(pt[ib][jb], lf, pf[jb]);
        }
      }
    }
#pragma unroll
    for (int ib = 0; ib < 3; ++ib)
#pragma unroll
      for (int jb = 0; jb < 3; ++jb) {
        stg2(out + kWB + lo24 + t24(jb, ib), bt[ib][jb]);
        double* pq = out + kWP + (8 * ib + r) + kN * (8 * jb + 2 * c);
        __stcg(pq, pt[ib][jb].x);
        __stcg(pq + kN, pt[ib][jb].y);
        if (ib <= jb) stg2(out + kWL + lo24 + t24(ib, jb), ld2(ws.W + lo26 + t26(ib, jb)));  // L^-T, tiles on and above the block diagonal
      }
    // r~ = L^-1 r
    {
      double z[3];
      matvec_cols<kLd, true>(ws.W, ws.v, r, c, z);
      if (c < 3) __stcg(out + kWr + 8 * c + r, pick3(z, c));
    }
    if (!__all_sync(kFull, pd) && lane == 0) atomicOr(a.status + prob, O2C_STATUS_CHOL_NOT_PD);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the flow map under RK4 (one warp per problem)
// ---------------------------------------------------------------------------------------------------------------------
struct __align__(16) FlowSmem {
  double slot[2][kSlot];  // node k lives in slot k & 1: { A | B~ | Hv | q | r | c, pad | r~ }
  double W[kN * kLd];     // stage argument of the flow map, both triangles
  double Sv[kN], Gv[kN];
  unsigned long long full[2];
};
static_assert(sizeof(FlowSmem) % 16 == 0, "warp slots must keep 16-byte alignment");

__device__ __forceinline__ double2 lerp2(double al, double be, const double2& u, const double2& v) {
  return make_double2(fma(al, u.x, be * v.x), fma(al, u.y, be * v.y));
}

// EV: the schedule carries jump steps (events); a template variant because the event-free kernel sits at its register limit
template <bool EV>
__global__ void __launch_bounds__(32 * kFlowWarps, 1) slq_flow_kernel(const SlqWppArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FlowSmem& ws = reinterpret_cast<FlowSmem*>(smem_raw)[warp];
  const int r = lane >> 2, c = lane & 3;
  const int lo24 = 2 * c + kN * r, lo26 = 2 * c + kLd * r;
  const int N = a.N, nodes = N + 1;
  const int jv = c < 3 ? 8 * c + r : 0;  // the vector element this lane owns (lanes with c < 3)
  const bool own = c < 3;
  if (lane == 0) {
    mbar_init(&ws.full[0], 1);
    mbar_init(&ws.full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;  // bit q: parity to wait for on full[q]
  const int total_warps = gridDim.x * (blockDim.x >> 5);
  // the step schedule is read once per step by every warp: from shared memory (a dependent global load at the top of every step
  // was 5 % of all stall samples, profiles/r02_ncu_slq_flow_summary.txt)
  __shared__ SlqStep step_cache[kStepCache];
  for (int i = threadIdx.x; i < a.nsteps && i < kStepCache; i += blockDim.x) step_cache[i] = a.steps[i];
  __syncthreads();

  for (int pi = warp * gridDim.x + blockIdx.x; pi < a.count; pi += total_warps) {
    const int prob = a.begin + pi;
    const double* lqp = a.lq + (size_t)prob * nodes * kRec;
    const double* wsp = a.ws + (size_t)prob * nodes * kWRec;
    const double* term = a.term + (size_t)prob * a.trec;
    double* solp = a.sol + (size_t)prob * nodes * kORec;

    auto issue_node = [&](int k) {  // four bulk copies, one mbarrier
      if (lane == 0) {
        double* dst = ws.slot[k & 1];
        const double* rec = lqp + (size_t)k * kRec;
        const double* wr = wsp + (size_t)k * kWRec;
        fence_proxy_async();
        mbar_expect_tx(&ws.full[k & 1], (uint32_t)(sizeof(double) * kSlot));
        tma_load(dst, rec, sizeof(double) * kMat, &ws.full[k & 1]);
        tma_load(dst + kMat, wr + kWB, sizeof(double) * kMat, &ws.full[k & 1]);
        tma_load(dst + 2 * kMat, rec + 2 * kMat, sizeof(double) * kTail, &ws.full[k & 1]);
        tma_load(dst + kOperand, wr + kWr, sizeof(double) * kN, &ws.full[k & 1]);
      }
    };
    auto wait_node = [&](int k) {
      const int q = k & 1;
      mbar_wait(&ws.full[q], (phase >> q) & 1u);
      phase ^= 1u << q;
    };
    auto touch_node = [&](int k) {  // pull the pieces of node k towards L2 (operands by bulk prefetch, the accumulator inits per lane)
      const double* rec = lqp + (size_t)k * kRec;
      const double* wr = wsp + (size_t)k * kWRec;
      if (lane == 0) {
        l2_prefetch(rec, sizeof(double) * kMat);
        l2_prefetch(wr + kWB, sizeof(double) * kMat);
      }
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) {
          l2_touch(wr + kWP + lo24 + t24(jb, ib));
          if (jb <= ib) l2_touch(rec + kOQ + lo24 + t24(jb, ib));
        }
    };

    // ---- terminal condition: valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526) ----
    double2 S[6];  // lower tiles, accumulator layout: S[8ib+r][8jb+2c..2c+1]
#pragma unroll
    for (int ib = 0; ib < 3; ++ib)
#pragma unroll
      for (int jb = 0; jb <= ib; ++jb) S[lt(ib, jb)] = ld2(term + a.oQf + lo24 + t24(jb, ib));  // Qf[8jb+2c..][8ib+r] (symmetric)
    double Svj = own ? term[a.oqf + jv] : 0.0;
    double sval = term[a.ocf];

    auto stage_to_scratch = [&](const double2 (&Y)[6], double svj) {  // value function -> operand layout (both triangles), Sv -> shared
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) {
          st2(ws.W + lo26 + t26(jb, ib), Y[lt(ib, jb)]);
          if (ib != jb) tput(ws.W, ib, jb, r, c, Y[lt(ib, jb)]);
        }
      if (own) ws.Sv[jv] = svj;
      __syncwarp();
    };
    auto write_value = [&](int k) {
      double* out = solp + (size_t)k * kORec;
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) {
          const double2 sv = S[lt(ib, jb)];
          stg2(out + kOSm + lo24 + t24(jb, ib), sv);
          if (ib != jb) {
            double* gq = out + kOSm + (8 * ib + r) + kN * (8 * jb + 2 * c);
            __stcg(gq, sv.x);
            __stcg(gq + kN, sv.y);
          }
        }
      if (own) __stcg(out + kOSv + jv, Svj);
      if (lane == 0) __stcg(out + kOs, sval);
    };

    write_value(N);
    issue_node(N);
    if (N >= 1) issue_node(N - 1);
    wait_node(N);
    if (N >= 1) wait_node(N - 1);
    if (N >= 2) touch_node(N - 2);
    int loaded_lo = N >= 1 ? N - 1 : N;
    stage_to_scratch(S, Svj);

#pragma unroll 1
    for (int sidx = 0; sidx < a.nsteps; ++sidx) {
      const SlqStep sp = sidx < kStepCache ? step_cache[sidx] : a.steps[sidx];
      const int i0 = sp.interval;
      if (i0 < loaded_lo) {  // the interval moved one node down: node i0 takes the slot of node i0 + 2 (dead)
        __syncwarp();
        issue_node(i0);
        wait_node(i0);
        if (i0 >= 1) touch_node(i0 - 1);
        loaded_lo = i0;
      }
      if (EV && sp.jump > 0) {  // (after the node bookkeeping: the step after the event lerps between nodes i0 - 1 and i0)
        // node i0 is a pre-event node: instead of integrating, the value function crosses the event through
        // ContinuousTimeRiccatiEquations::computeJumpMap = riccatiTransversalityConditions on the event's jump model data
        // (SLQ.cpp:286-296, RiccatiTransversalityConditions.h:40-56): S- = Q_e + (S A_e)' A_e, Sv- = q_e + A_e'(Sv + S Hv_e),
        // s- = s + c_e + Hv_e.(Sv + S Hv_e / 2). S and Sv of the post-event node are in the scratch (stage argument of the last step).
        const double* jr = a.jump + ((size_t)prob * a.jump_capacity + (sp.jump - 1)) * jump_rec(kN);
        const double* Ae = jr;
        const double* Hve = jr + jump_oHv(kN);
        double2 zA[3][3];
        double pSH[3] = {0.0, 0.0, 0.0}, pA[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) zA[i][j] = zero2();
        double2 hvf[3];
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          hvf[kb] = ldg2(Hve + 8 * kb + 2 * c);
          const double2 svf = ld2(ws.Sv + 8 * kb + 2 * c);
          double2 s[3];
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            s[jb] = ld2(ws.W + lo26 + t26(kb, jb));
            pSH[jb] = fma(s[jb].x, hvf[kb].x, pSH[jb]);
            pSH[jb] = fma(s[jb].y, hvf[kb].y, pSH[jb]);
          }
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double2 af = ldg2(Ae + lo24 + t24(kb, ib));
            pA[ib] = fma(af.x, svf.x, pA[ib]);
            pA[ib] = fma(af.y, svf.y, pA[ib]);
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) dmma2(zA[ib][jb], af, s[jb]);  // A_e'S = op fragments of S A_e
          }
        }
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {  // + A_e'(S Hv_e) from the accumulators
            pA[ib] = fma(zA[ib][jb].x, hvf[jb].x, pA[ib]);
            pA[ib] = fma(zA[ib][jb].y, hvf[jb].y, pA[ib]);
          }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          pSH[i] = quad_sum(pSH[i]);
          pA[i] = quad_sum(pA[i]);
        }
        double spart = 0.0;
        if (own) {
          spart = __ldg(Hve + jv) * (ws.Sv[jv] + 0.5 * pick3(pSH, c));
          Svj = __ldg(jr + jump_oq(kN) + jv) + pick3(pA, c);
        }
        sval = sval + __ldg(jr + jump_oc(kN)) + warp_sum_all(spart);
        // S- = Q_e + (S A_e)' A_e (lower tiles)
#pragma unroll
        for (int ib = 0; ib < 3; ++ib)
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) S[lt(ib, jb)] = ldg2(jr + jump_oQ(kN) + lo24 + t24(jb, ib));
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          double2 af[3];
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) af[jb] = ldg2(Ae + lo24 + t24(kb, jb));
#pragma unroll
          for (int ib = 0; ib < 3; ++ib)
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(S[lt(ib, jb)], zA[ib][kb], af[jb]);
        }
        __syncwarp();  // every lane is done reading the post-event value function from the scratch
        stage_to_scratch(S, Svj);
        if (sp.observe_node >= 0) write_value(sp.observe_node);
        continue;
      }
      const double* s0 = ws.slot[i0 & 1];        // node i0
      const double* s1 = ws.slot[(i0 + 1) & 1];  // node i0 + 1
      const double* q0 = lqp + (size_t)i0 * kRec + kOQ;
      const double* q1 = q0 + kRec;
      const double* p0 = wsp + (size_t)i0 * kWRec + kWP;
      const double* p1 = p0 + kWRec;
      const double h = sp.h;
      double2 acc[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[i] = S[i];
      double accv = Svj, accs = sval;

#pragma unroll 1
      for (int stg = 0; stg < 4; ++stg) {
        // (selects on constant indices keep the step in registers: a run-time index would put it into local memory)
        const double al = (stg == 0 ? sp.alpha[0] : (stg == 1 ? sp.alpha[1] : (stg == 2 ? sp.alpha[2] : sp.alpha[3]))), be = 1.0 - al;
        // Q (lower tiles, + eps I) and P~' (all tiles) come from L2, lerped; they are issued here and added to the accumulators after the
        // contractions, so that their latency hides behind the first 126 DMMAs
        double2 d[6], g[3][3], qd[6], pg[3][3];
#pragma unroll
        for (int ib = 0; ib < 3; ++ib) {
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            qd[lt(ib, jb)] = lerp2(al, be, ldg2(q0 + lo24 + t24(jb, ib)), ldg2(q1 + lo24 + t24(jb, ib)));
            d[lt(ib, jb)] = zero2();
          }
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            pg[ib][jb] = lerp2(al, be, ldg2(p0 + lo24 + t24(jb, ib)), ldg2(p1 + lo24 + t24(jb, ib)));
            g[ib][jb] = zero2();
          }
        }
        double pSH[3] = {0.0, 0.0, 0.0}, pA[3] = {0.0, 0.0, 0.0}, pB[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const double2 hvf = lerp2(al, be, ld2(s0 + 2 * kMat + 8 * kb + 2 * c), ld2(s1 + 2 * kMat + 8 * kb + 2 * c));
          const double2 svf = ld2(ws.Sv + 8 * kb + 2 * c);
          double2 s[3], af[3], bf[3];
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            s[jb] = ld2(ws.W + lo26 + t26(kb, jb));
            pSH[jb] = fma(s[jb].x, hvf.x, pSH[jb]);
            pSH[jb] = fma(s[jb].y, hvf.y, pSH[jb]);
            af[jb] = lerp2(al, be, ld2(s0 + lo24 + t24(kb, jb)), ld2(s1 + lo24 + t24(kb, jb)));
            bf[jb] = lerp2(al, be, ld2(s0 + kMat + lo24 + t24(kb, jb)), ld2(s1 + kMat + lo24 + t24(kb, jb)));
            pA[jb] = fma(af[jb].x, svf.x, pA[jb]);
            pA[jb] = fma(af[jb].y, svf.y, pA[jb]);
            pB[jb] = fma(bf[jb].x, svf.x, pB[jb]);
            pB[jb] = fma(bf[jb].y, svf.y, pB[jb]);
          }
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) {
              dmma2(d[lt(ib, jb)], s[ib], af[jb]);  // (S'A)(ib, jb)
              dmma2(d[lt(ib, jb)], af[ib], s[jb]);  // (A'S)(ib, jb)
            }
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) dmma2(g[ib][jb], s[ib], bf[jb]);  // G' = P~' + S B~: operand fragments of G
          }
        }
#pragma unroll
        for (int ib = 0; ib < 3; ++ib) {
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            d[lt(ib, jb)].x += qd[lt(ib, jb)].x;
            d[lt(ib, jb)].y += qd[lt(ib, jb)].y;
          }
          d[lt(ib, ib)].x += (2 * c == r) ? a.eps : 0.0;
          d[lt(ib, ib)].y += (2 * c + 1 == r) ? a.eps : 0.0;
#pragma unroll
          for (int jb = 0; jb < 3; ++jb) {
            g[ib][jb].x += pg[ib][jb].x;
            g[ib][jb].y += pg[ib][jb].y;
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          pSH[i] = quad_sum(pSH[i]);
          pA[i] = quad_sum(pA[i]);
          pB[i] = quad_sum(pB[i]);
        }
        double kv = 0.0, spart = 0.0;
        if (own) {
          const double hvj = fma(al, s0[2 * kMat + jv], be * s1[2 * kMat + jv]);
          const double gvj = fma(al, s0[kOperand + jv], be * s1[kOperand + jv]) + pick3(pB, c);  // Gv = r~ + B~'Sv
          ws.Gv[jv] = gvj;
          kv = fma(al, s0[2 * kMat + kN + jv], be * s1[2 * kMat + kN + jv]) + pick3(pSH, c) + pick3(pA, c);  // q~ + S Hv + A'Sv
          spart = fma(hvj, ws.Sv[jv], -0.5 * gvj * gvj);                                                     // Hv.Sv - Gv.Gv / 2
        }
        const double cl = fma(al, s0[2 * kMat + 3 * kN], be * s1[2 * kMat + 3 * kN]);
        __syncwarp();  // Gv is in shared memory; every lane is done reading the stage argument from the scratch
        // dS -= G'G (lower tiles); dSv -= G'Gv
#pragma unroll
        for (int kb = 0; kb < 3; ++kb)
#pragma unroll
          for (int ib = 0; ib < 3; ++ib) {
            const double2 ng = neg2(g[ib][kb]);
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma2(d[lt(ib, jb)], ng, g[jb][kb]);
          }
        {
          double2 gvf[3];
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) gvf[kb] = ld2(ws.Gv + 8 * kb + 2 * c);
          double z[3];
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            double p = 0.0;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              p = fma(g[cb][kb].x, gvf[kb].x, p);
              p = fma(g[cb][kb].y, gvf[kb].y, p);
            }
            z[cb] = quad_sum(p);
          }
          if (own) kv -= pick3(z, c);
        }
        const double ks = cl + warp_sum_all(spart);
        // classic RK4 (boost::odeint runge_kutta4): y += h (k1 + 2 k2 + 2 k3 + k4) / 6, next argument y + c_s h k_s
        const double bw = h * ((stg == 0 || stg == 3) ? (1.0 / 6.0) : (1.0 / 3.0));
        const double cw = h * ((stg == 2) ? 1.0 : 0.5);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          acc[i].x = fma(bw, d[i].x, acc[i].x);
          acc[i].y = fma(bw, d[i].y, acc[i].y);
        }
        accv = fma(bw, kv, accv);
        accs = fma(bw, ks, accs);
        if (stg < 3) {
          double2 ys[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) ys[i] = make_double2(fma(cw, d[i].x, S[i].x), fma(cw, d[i].y, S[i].y));
          stage_to_scratch(ys, fma(cw, kv, Svj));
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) S[i] = acc[i];
      Svj = accv;
      sval = accs;
      stage_to_scratch(S, Svj);
      if (sp.observe_node >= 0) write_value(sp.observe_node);
    }

    bool finite = finite_bits(sval) && finite_bits(Svj);
#pragma unroll
    for (int i = 0; i < 6; ++i) finite = finite && finite2(S[i]);
    if (!__all_sync(kFull, finite) && lane == 0) atomicOr(a.status + prob, O2C_STATUS_NONFINITE);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// controller of every node (node-parallel)
// ---------------------------------------------------------------------------------------------------------------------
template <bool NOM>
__global__ void __launch_bounds__(32 * kNodeWarps) slq_controller_kernel(const SlqWppArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  NodeSmem& ws = reinterpret_cast<NodeSmem*>(smem_raw)[warp];
  const int r = lane >> 2, c = lane & 3;
  const int lo24 = 2 * c + kN * r, lo26 = 2 * c + kLd * r;
  const int N = a.N, nodes = N + 1;
  const long long items = (long long)a.count * N;  // nodes 0 .. N-1; node N is the copy of node N-1 (GaussNewtonDDP.cpp:609-618)
  for (long long item = (long long)blockIdx.x * kNodeWarps + warp; item < items; item += (long long)gridDim.x * kNodeWarps) {
    const int prob = a.begin + (int)(item / N), node = (int)(item % N);
    const double* wr = a.ws + ((size_t)prob * nodes + node) * kWRec;
    double* out = a.sol + ((size_t)prob * nodes + node) * kORec;
    double* outN = (node == N - 1) ? a.sol + ((size_t)prob * nodes + N) * kORec : nullptr;
    // L^-T into the scratch (tiles on and above the block diagonal), Sv into shared memory
#pragma unroll
    for (int ib = 0; ib < 3; ++ib)
#pragma unroll
      for (int jb = ib; jb < 3; ++jb) st2(ws.W + lo26 + t26(ib, jb), ldg2(wr + kWL + lo24 + t24(ib, jb)));
    if (lane < kN) ws.v[lane] = out[kOSv + lane];
    __syncwarp();
    // Yc' = P~' + S B~ (operand fragments of Yc = P~ + B~'S_k), Gv = r~ + B~'Sv_k
    double2 y[3][3];
#pragma unroll
    for (int ib = 0; ib < 3; ++ib)
#pragma unroll
      for (int jb = 0; jb < 3; ++jb) y[ib][jb] = ldg2(wr + kWP + lo24 + t24(jb, ib));
    double pB[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) {
      const double2 svf = ld2(ws.v + 8 * kb + 2 * c);
      double2 s[3], bf[3];
#pragma unroll
      for (int jb = 0; jb < 3; ++jb) {
        s[jb] = ldg2(out + kOSm + lo24 + t24(kb, jb));
        bf[jb] = ldg2(wr + kWB + lo24 + t24(kb, jb));
        pB[jb] = fma(bf[jb].x, svf.x, pB[jb]);
        pB[jb] = fma(bf[jb].y, svf.y, pB[jb]);
      }
#pragma unroll
      for (int ib = 0; ib < 3; ++ib)
#pragma unroll
        for (int jb = 0; jb < 3; ++jb) dmma2(y[ib][jb], s[ib], bf[jb]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) pB[i] = quad_sum(pB[i]);
    if (c < 3) ws.y[8 * c + r] = __ldcg(wr + kWr + 8 * c + r) + pick3(pB, c);
    __syncwarp();
    // dbias = -L^-T Gv
    {
      double2 vf[3];
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) vf[kb] = ld2(ws.y + 8 * kb + 2 * c);
      double z[3];
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
        if (outN) __stcg(outN + kOdb + j, -pick3(z, c));
        if (!NOM) {
          __stcg(out + kObias + j, 0.0);
          if (outN) __stcg(outN + kObias + j, 0.0);
        }
      }
    }
    // K' = -Yc' L^-1 (operand fragments of K -> 16-byte stores)
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
      for (int jb = 0; jb < 3; ++jb) {
        stg2(out + kOK + lo24 + t24(jb, ib), neg2(kk[ib][jb]));
        if (outN) stg2(outN + kOK + lo24 + t24(jb, ib), neg2(kk[ib][jb]));
      }
    if (NOM) {  // bias = u_nom - K x_nom (GaussNewtonDDP.cpp:604-606): the accumulators hold -K'
      const double* xn = a.x_nom + ((size_t)prob * nodes + node) * kN;
      const double* un = a.u_nom + ((size_t)prob * nodes + node) * kN;
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
          const double2 b2 = make_double2(u2.x + px, u2.y + py);
          stg2(out + kObias + 8 * jb + 2 * c, b2);
          if (outN) stg2(outN + kObias + 8 * jb + 2 * c, b2);
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// continuous rollout of the LQ model under the LinearController for nx = nu = 24 (TimeTriggeredRollout::run on the linearised model,
// ocs2_oc/src/rollout/TimeTriggeredRollout.cpp:46-115; LinearController::computeInput, ocs2_core/src/control/LinearController.cpp:79-87):
// xdot = A(t) (x - x_nom(t)) + B(t) (u - u_nom(t)) + Hv(t), u(t, x) = bias_alpha(t) + K(t) x, every quantity lerped between the two nodes
// of its time segment, classic RK4 over the host-built step schedule (api.cu build_rollout_schedule). One warp per (problem, step
// length), one state row per lane; {A | B | Hv} and {K | dbias | bias} of the nodes stream through a three-slot TMA ring and are lerped
// on the fly from shared memory (the row-per-lane kernels of the small shapes keep them in registers: 6 x 24 doubles do not fit here).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDyn = 2 * kMat + kN;  // { A | B | Hv }
constexpr int kPol = kMat + 2 * kN;  // { K | dbias | bias }
constexpr int kRoWarps = 4;

struct __align__(16) RoSmem {
  double dyn[3][kDyn];
  double pol[3][kPol];
  double x[kN], u[kN], dx[kN];
  unsigned long long bar;
  unsigned long long pad;
};
static_assert(sizeof(RoSmem) % 16 == 0, "warp slots must keep 16-byte alignment");

struct Ro24Args {
  const double* lq;
  const double* sol;
  const double* x0;
  const double* x_nom;
  const double* u_nom;
  double* xs;
  double* us;
  int* status;
  const RolloutStep* steps;
  const double* alphas;
  const double* jump;
  int jump_capacity;
  int nsteps, first_idx, out_nodes, N, batch, begin, count;
  double first_alpha;
};

__global__ void __launch_bounds__(32 * kRoWarps, 1) rollout_cont24_kernel(const Ro24Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RoSmem& ws = reinterpret_cast<RoSmem*>(smem_raw)[warp];
  const bool valid = lane < kN;
  const int i = valid ? lane : kN - 1;
  const int N = a.N, nodes = N + 1;
  const int ia = blockIdx.y;
  const double alpha = a.alphas[ia];
  if (lane == 0) {
    mbar_init(&ws.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
  const int warps_total = gridDim.x * kRoWarps;
  __shared__ RolloutStep step_cache[kStepCache];  // the schedule out of shared memory instead of a dependent global load per step
  for (int i = threadIdx.x; i < a.nsteps && i < kStepCache; i += blockDim.x) step_cache[i] = a.steps[i];
  __syncthreads();
  for (int pi = blockIdx.x * kRoWarps + warp; pi < a.count; pi += warps_total) {
    const int prob = a.begin + pi;
    double* xo = a.xs + ((size_t)ia * a.batch + prob) * (size_t)a.out_nodes * kN;
    double* uo = a.us + ((size_t)ia * a.batch + prob) * (size_t)a.out_nodes * kN;
    auto issue_node = [&](int k) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&ws.bar, (uint32_t)(sizeof(double) * (kDyn + kPol)));
        tma_load(ws.dyn[k % 3], a.lq + ((size_t)prob * nodes + k) * kRec, sizeof(double) * kDyn, &ws.bar);
        tma_load(ws.pol[k % 3], a.sol + ((size_t)prob * nodes + k) * kORec, sizeof(double) * kPol, &ws.bar);
      }
    };
    int loaded_hi = -1, issued_hi = -1;
    // make nodes <= q resident (the segment index never decreases along the schedule: when node q has landed node q - 2 is dead and its
    // slot takes the prefetch of node q + 1)
    auto ensure = [&](int q) {
      q = q < N ? q : N;
      while (loaded_hi < q) {
        if (issued_hi == loaded_hi) {
          issue_node(loaded_hi + 1);
          issued_hi = loaded_hi + 1;
        }
        mbar_wait(&ws.bar, parity);
        parity ^= 1u;
        loaded_hi += 1;
        if (loaded_hi >= q && loaded_hi + 1 <= N) {
          __syncwarp();
          issue_node(loaded_hi + 1);
          issued_hi = loaded_hi + 1;
        }
      }
    };
    // u(t, x) for the x in shared memory (row i of K per lane), broadcast through shared memory
    // (a time that sits on a node — w0 = 1 or w0 = 0, i.e. the first and the last stage of every RK4 step on the nodes' own grid — reads
    //  that node's rows alone: no lerp, half the shared-memory reads)
    auto policy = [&](int idx, double w0) {
      ensure(idx + 1);
      const int hi = idx + 1 < N ? idx + 1 : N;
      const double w1 = 1.0 - w0;
      double u0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
      if (w1 == 0.0 || w0 == 0.0) {
        const double* pn = ws.pol[(w1 == 0.0 ? idx : hi) % 3];
        u0 = pn[kMat + kN + i] + alpha * pn[kMat + i];
#pragma unroll
        for (int j = 0; j < kN; j += 4) {
          const double2 xa = ld2(ws.x + j), xb = ld2(ws.x + j + 2);
          u0 = fma(pn[i + kN * j], xa.x, u0);
          u1 = fma(pn[i + kN * (j + 1)], xa.y, u1);
          u2 = fma(pn[i + kN * (j + 2)], xb.x, u2);
          u3 = fma(pn[i + kN * (j + 3)], xb.y, u3);
        }
      } else {
        const double* plo = ws.pol[idx % 3];
        const double* phi = ws.pol[hi % 3];
        u0 = w0 * (plo[kMat + kN + i] + alpha * plo[kMat + i]) + w1 * (phi[kMat + kN + i] + alpha * phi[kMat + i]);
#pragma unroll
        for (int j = 0; j < kN; j += 4) {
          const double2 xa = ld2(ws.x + j), xb = ld2(ws.x + j + 2);
          u0 = fma(fma(w0, plo[i + kN * j], w1 * phi[i + kN * j]), xa.x, u0);
          u1 = fma(fma(w0, plo[i + kN * (j + 1)], w1 * phi[i + kN * (j + 1)]), xa.y, u1);
          u2 = fma(fma(w0, plo[i + kN * (j + 2)], w1 * phi[i + kN * (j + 2)]), xb.x, u2);
          u3 = fma(fma(w0, plo[i + kN * (j + 3)], w1 * phi[i + kN * (j + 3)]), xb.y, u3);
        }
      }
      if (valid) ws.u[lane] = (u0 + u1) + (u2 + u3);
      __syncwarp();
    };
    auto flow = [&](int idx, double w0) -> double {
      policy(idx, w0);
      const int hi = idx + 1 < N ? idx + 1 : N;
      const double* dlo = ws.dyn[idx % 3];
      const double* dhi = ws.dyn[hi % 3];
      const double w1 = 1.0 - w0;
      const double* xv = ws.x;
      if (a.x_nom) {  // deviations from the lerped nominal trajectories
        const double* xn = a.x_nom + ((size_t)prob * nodes + idx) * kN;
        const double* un = a.u_nom + ((size_t)prob * nodes + idx) * kN;
        const int step = hi > idx ? kN : 0;
        const double dxi = ws.x[i] - fma(w0, __ldg(xn + i), w1 * __ldg(xn + step + i));
        const double dui = ws.u[i] - fma(w0, __ldg(un + i), w1 * __ldg(un + step + i));
        __syncwarp();
        if (valid) {
          ws.dx[lane] = dxi;
          ws.u[lane] = dui;
        }
        __syncwarp();
        xv = ws.dx;
      }
      double a0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      if (w1 == 0.0 || w0 == 0.0) {
        const double* dn = (w1 == 0.0) ? dlo : dhi;
        a0 = dn[2 * kMat + i];
#pragma unroll
        for (int j = 0; j < kN; j += 2) {
          const double2 x2 = ld2(xv + j), u2 = ld2(ws.u + j);
          a0 = fma(dn[i + kN * j], x2.x, a0);
          a1 = fma(dn[i + kN * (j + 1)], x2.y, a1);
          b0 = fma(dn[kMat + i + kN * j], u2.x, b0);
          b1 = fma(dn[kMat + i + kN * (j + 1)], u2.y, b1);
        }
      } else {
        a0 = fma(w0, dlo[2 * kMat + i], w1 * dhi[2 * kMat + i]);
#pragma unroll
        for (int j = 0; j < kN; j += 2) {
          const double2 x2 = ld2(xv + j), u2 = ld2(ws.u + j);
          a0 = fma(fma(w0, dlo[i + kN * j], w1 * dhi[i + kN * j]), x2.x, a0);
          a1 = fma(fma(w0, dlo[i + kN * (j + 1)], w1 * dhi[i + kN * (j + 1)]), x2.y, a1);
          b0 = fma(fma(w0, dlo[kMat + i + kN * j], w1 * dhi[kMat + i + kN * j]), u2.x, b0);
          b1 = fma(fma(w0, dlo[kMat + i + kN * (j + 1)], w1 * dhi[kMat + i + kN * (j + 1)]), u2.y, b1);
        }
      }
      __syncwarp();  // x, u in shared memory are dead
      return (a0 + a1) + (b0 + b1);
    };
    double x = a.x0[(size_t)prob * kN + i];
    bool finite = true;
    auto observe = [&](int o, int idx, double w0) {
      if (valid) ws.x[lane] = x;
      __syncwarp();
      policy(idx, w0);
      if (valid) {
        __stcg(xo + (size_t)o * kN + lane, x);
        __stcg(uo + (size_t)o * kN + lane, ws.u[lane]);
      }
      finite = finite && finite_bits(x);
      __syncwarp();
    };
    observe(0, a.first_idx, a.first_alpha);
#pragma unroll 1
    for (int sidx = 0; sidx < a.nsteps; ++sidx) {
      const RolloutStep sp = sidx < kStepCache ? step_cache[sidx] : a.steps[sidx];
      const double h = sp.h;
      if (sp.jump > 0) {  // an event (TimeTriggeredRollout.cpp:104-108): x+ = x_nom(post) + A_e (x - x_nom(pre)) + Hv_e
        const double* jr = a.jump + ((size_t)prob * a.jump_capacity + (sp.jump - 1)) * jump_rec(kN);
        if (valid) ws.x[lane] = x - (a.x_nom ? __ldg(a.x_nom + ((size_t)prob * nodes + sp.pre_node) * kN + i) : 0.0);
        __syncwarp();
        double xn = __ldg(jr + jump_oHv(kN) + i) + (a.x_nom ? __ldg(a.x_nom + ((size_t)prob * nodes + sp.pre_node + 1) * kN + i) : 0.0);
#pragma unroll
        for (int kk = 0; kk < kN; ++kk) xn = fma(__ldg(jr + i + kN * kk), ws.x[kk], xn);
        __syncwarp();
        x = xn;
      }
      if (h == 0.0) {  // a jump or a degenerate interval: no integration
        observe(sidx + 1, sp.obs_idx, sp.obs_alpha);
        continue;
      }
      double acc = x, xs = x;
#pragma unroll 1
      for (int stg = 0; stg < 4; ++stg) {
        if (valid) ws.x[lane] = xs;
        __syncwarp();
        const double kx = flow((stg == 0 ? sp.idx[0] : (stg == 1 ? sp.idx[1] : (stg == 2 ? sp.idx[2] : sp.idx[3]))), (stg == 0 ? sp.alpha[0] : (stg == 1 ? sp.alpha[1] : (stg == 2 ? sp.alpha[2] : sp.alpha[3]))));  // (constant indices: the step stays in registers)
        acc = fma(h * ((stg == 0 || stg == 3) ? (1.0 / 6.0) : (1.0 / 3.0)), kx, acc);
        xs = fma(h * ((stg == 2) ? 1.0 : 0.5), kx, x);
      }
      x = acc;
      observe(sidx + 1, sp.obs_idx, sp.obs_alpha);
    }
    if (issued_hi > loaded_hi) {  // drain a prefetch that is still in flight before the ring is reused
      mbar_wait(&ws.bar, parity);
      parity ^= 1u;
    }
    if (!__all_sync(kFull, finite) && lane == 0) atomicOr(a.status + prob, O2C_STATUS_NONFINITE);
    __syncwarp();
  }
}

}  // namespace

size_t slq_wpp_workspace_doubles(const Layout& L, int batch) { return (size_t)batch * (L.N + 1) * kWRec; }

bool slq_wpp_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  return L.n == kN && L.m == kN && L.ncmax == 0 && st.algorithm == O2C_ALG_SLQ && st.strategy == O2C_STRATEGY_LINE_SEARCH &&
         st.hc == O2C_HC_DIAGONAL_SHIFT && (buf.event == nullptr || buf.jump != nullptr) && (buf.x_nom == nullptr) == (buf.u_nom == nullptr) && L.N >= 1 &&
         L.rec == kRec && L.oQ == kOQ && L.oP == kOP && L.oR == kOR && L.orec == kORec && L.oK == kOK && L.odb == kOdb &&
         L.obias == kObias && L.oSm == kOSm && L.oSv == kOSv && L.os == kOs && L.oA == 0 && L.oB == kMat && L.oHv == 2 * kMat &&
         L.oq == 2 * kMat + kN && L.or_ == 2 * kMat + 2 * kN && L.oc == 2 * kMat + 3 * kN;
}

cudaError_t launch_slq_wpp(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, double* workspace, const SlqStep* steps, int nsteps,
                           int begin, int count, cudaStream_t stream, int* launches) {
  if (!slq_wpp_supported(L, st, buf) || workspace == nullptr) return cudaErrorNotSupported;
  const int sms = device_sm_count();
  if (sms <= 0) return cudaErrorInvalidDevice;
  SlqWppArgs a{};
  a.lq = buf.lq;
  a.term = buf.term;
  a.x_nom = buf.x_nom;
  a.u_nom = buf.u_nom;
  a.ws = workspace;
  a.sol = buf.sol;
  a.status = buf.status;
  a.steps = steps;
  a.jump = buf.jump;
  a.jump_capacity = buf.jump_capacity;
  a.nsteps = nsteps;
  a.N = L.N;
  a.begin = begin;
  a.count = count;
  a.oQf = L.oQf;
  a.oqf = L.oqf;
  a.ocf = L.ocf;
  a.trec = L.trec;
  a.eps = st.eps;
  cudaError_t e = cudaMemsetAsync(buf.status + begin, 0, sizeof(int) * (size_t)count, stream);  // the three kernels or their bits in
  if (e != cudaSuccess) return e;
  const size_t node_smem = sizeof(NodeSmem) * kNodeWarps;
  {
    const long long items = (long long)count * (L.N + 1);
    const long long ctas = (items + kNodeWarps - 1) / kNodeWarps;
    const int grid = (int)(ctas < (long long)sms * 16 ? ctas : (long long)sms * 16);
    slq_project_kernel<<<grid, 32 * kNodeWarps, node_smem, stream>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    int warps = kFlowWarps;
    if (const char* env = getenv("O2C_SLQ_WPP_RESIDENT")) {  // profiling knob
      const int v = atoi(env);
      if (v >= 1 && v <= kFlowWarps) warps = v;
    }
    const size_t smem = sizeof(FlowSmem) * warps;
    void (*flow)(const SlqWppArgs) = (buf.event != nullptr && buf.jump != nullptr) ? slq_flow_kernel<true> : slq_flow_kernel<false>;
    e = cudaFuncSetAttribute(flow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FlowSmem) * kFlowWarps));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(flow, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int grid = count < sms ? count : sms;
    flow<<<grid, 32 * warps, smem, stream>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    const long long items = (long long)count * L.N;
    const long long ctas = (items + kNodeWarps - 1) / kNodeWarps;
    const int grid = (int)(ctas < (long long)sms * 16 ? ctas : (long long)sms * 16);
    if (buf.x_nom)
      slq_controller_kernel<true><<<grid, 32 * kNodeWarps, node_smem, stream>>>(a);
    else
      slq_controller_kernel<false><<<grid, 32 * kNodeWarps, node_smem, stream>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (launches) *launches = 3;
  return cudaSuccess;
}

bool rollout_cont24_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  return L.n == kN && L.m == kN && L.ncmax == 0 && st.algorithm == O2C_ALG_SLQ && (buf.x_nom == nullptr) == (buf.u_nom == nullptr) &&
         L.N >= 1 && L.nodes == L.N + 1 && L.rec == kRec && L.orec == kORec && L.oA == 0 && L.oB == kMat && L.oHv == 2 * kMat && L.oK == kOK &&
         L.odb == kOdb && L.obias == kObias;
}

cudaError_t launch_rollout_cont24(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps,
                                  int first_idx, double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch, int begin,
                                  int count, cudaStream_t stream) {
  if (!rollout_cont24_supported(L, st, buf)) return cudaErrorNotSupported;
  Ro24Args a{};
  a.lq = buf.lq;
  a.sol = buf.sol;
  a.x0 = buf.x0;
  a.x_nom = buf.x_nom;
  a.u_nom = buf.u_nom;
  a.xs = buf.xs;
  a.us = buf.us;
  a.status = buf.status;
  a.steps = steps;
  a.alphas = alphas_dev;
  a.jump = buf.jump;
  a.jump_capacity = buf.jump_capacity;
  a.nsteps = nsteps;
  a.first_idx = first_idx;
  a.first_alpha = first_alpha;
  a.out_nodes = out_nodes;
  a.N = L.N;
  a.batch = batch;
  a.begin = begin;
  a.count = count;
  const size_t smem = sizeof(RoSmem) * kRoWarps;
  cudaError_t e = cudaFuncSetAttribute(rollout_cont24_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int sms = device_sm_count();
  if (sms <= 0) return cudaErrorInvalidDevice;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, rollout_cont24_kernel, 32 * kRoWarps, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int needed = (count + kRoWarps - 1) / kRoWarps;
  const int cap = sms * ctas_per_sm;
  dim3 grid(needed < cap ? needed : cap, n_alpha);
  rollout_cont24_kernel<<<grid, 32 * kRoWarps, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace o2c
