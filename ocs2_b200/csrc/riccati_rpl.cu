// riccati_rpl.cu — "row per lane" ILQR sweep + fused LQ rollout for SMALL shapes (nx <= 16), FP64, sm_100a.
//
// Small problems (ballbot nx = 10, nu = 3; cartpole nx = 4, nu = 1) are HBM-bound: 2.3 KB in / 1.3 KB out
// and only 3.5 k FMA per stage. A warp per problem wastes most lanes on 10x10 matrices and a tensor-core tile (8x8x4) would be
// three quarters padding, so here a warp carries P = 32 / nx problems at once and inside a problem LANE i OWNS ROW i of every
// nx-row matrix (S, SA, T, G', Y', K'), kept in registers; the operand that every row needs in full (A, B, SA, SB, Y', L) is
// broadcast from shared memory with 16-byte loads (two FMAs per load, no bank conflicts: all lanes of a problem read one address).
// Every product is arranged to be nx-row oriented (the nu-row ones are computed transposed), so all nx lanes work all the time:
//
//   SA_i  = S_i A            SB_i = S_i B            w_i = Sv_i + S_i Hv                       (S_i: row i of S, registers)
//   T_i   = Q_i + eps e_i + A(:,i)' SA            tv_i = q_i + A(:,i)' w
//   G'_i  = P(:,i)' + SA(:,i)' B                  (row i of G' = column i of G = P + B'SA)
//   Hm_l  = R_l + B(:,l)' SB   (lanes l < nu)     Cholesky Hm = L L' by shuffles inside the lane group, 1/L_jj kept on the diagonal
//   Y'_i  = G'_i L^-T  (forward substitution)     K'_i = -Y'_i L^-1 (back substitution)       -> K(:,i), coalesced
//   S_i   = T_i - Y'_i Y''   Sv_i = tv_i - Y'_i Yv   s = s+ + c + Hv.w - 1/2 Hv.(S Hv) - 1/2 Yv.Yv
//
// which are the quantities of ILQR::riccatiEquationsWorker (ocs2_ddp/src/ILQR.cpp:227-299), computeMapILQR
// (riccati_equations/DiscreteTimeRiccatiEquations.cpp:65-154), computeInverseMatrixUUT (ocs2_core/src/misc/LinearAlgebra.cpp:119-124)
// and calculateControllerWorker (ILQR.cpp:162-181) for nc = 0, LINE_SEARCH, reduced form, DIAGONAL_SHIFT (same re-association as
// riccati_wpp.cu: Y = L^-1 G, S = T - Y'Y, K = -L^-T Y). The stage records of the P problems arrive by TMA bulk copies
// (cp.async.bulk + one mbarrier per warp), double buffered, node k-1 in flight while node k is processed; outputs are written once.
// The forward rollout of the LQ model (DDP_HelperFunctions.cpp:125-138, 296-304; LinearController.cpp:79-87) is fused behind the
// sweep: { A | B | Hv } of the records and { K | dbias } of the solution stream forwards through a TMA ring of 3-6 stage sets laid over
// the part of the slot the sweep no longer needs.
#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int cpad2(int v) { return (v + 1) & ~1; }

template <int NX, int NU, int NC>
struct Shape {
  static constexpr int G = NX;        // lanes per problem
  static constexpr int XP = NX + (NX & 1);  // leading dimension of the kernel's own shared matrices: columns stay 16-byte aligned
  static constexpr int P = 32 / NX;   // problems per warp
  // record offsets of make_layout(NX, NU, NC, N, ILQR) for shapes below the 128-byte alignment threshold
  static constexpr int oA = 0, oB = cpad2(NX * NX), oHv = oB + cpad2(NX * NU), oq = oHv + cpad2(NX), or_ = oq + cpad2(NX),
                       oc = or_ + cpad2(NU), oQ = oc + 2, oP = oQ + cpad2(NX * NX), oR = oP + cpad2(NU * NX), oC = oR + cpad2(NU * NU),
                       oD = oC + cpad2(NC * NX), oe = oD + cpad2(NC * NU), rec = oe + cpad2(NC);
  static constexpr int oK = 0, odb = cpad2(NU * NX), obias = odb + cpad2(NU), oSm = obias + cpad2(NU), oSv = oSm + cpad2(NX * NX),
                       os = oSv + cpad2(NX), orec = os + 2;
  static constexpr int oQf = 0, oqf = cpad2(NX * NX), ocf = oqf + cpad2(NX), trec = ocf + 2;
  // shared memory of one problem slot (doubles)
  static constexpr int sRec = 0, sSA = 2 * rec, sSB = sSA + XP * NX, sYt = sSB + XP * NU, sL = sYt + XP * NU,
                       sW = sL + cpad2(NU * NU), sGv = sW + cpad2(NX), sYv = sGv + cpad2(NU), sX = sYv + cpad2(NU), sU = sX + cpad2(NX),
                       sK = sU + cpad2(NU), sZ = sK + cpad2(NU * NX + NU), sVx = sZ + cpad2(NU * NC), slot = sVx + NC * XP;
  // fused rollout: ring of stage sets { A | B | Hv } + { K | dbias } over the part of the slot the sweep no longer needs
  static constexpr int dyn = oq, pol = odb + cpad2(NU), set = dyn + pol;
  static constexpr int ring = (sX / set) < 6 ? (sX / set) : 6;
  static_assert(ring >= 2, "the rollout ring needs at least two stage sets");
  static constexpr int warp_doubles = P * slot + 2 + 6;  // + the warp's mbarriers: sweep, and one per ring slot
};

struct Args {
  const double* lq;
  const double* term;
  const double* x0;
  const double* x_nom;  // [batch][N+1][nx] nominal trajectories, or nullptr (deviation coordinates): kernel instantiation NOM
  const double* u_nom;
  double* sol;
  double* xs;
  double* us;
  int* status;
  const int* event;  // [batch][N] pre-event flags (ILQR.cpp:263-295), or nullptr: kernel instantiation EV
  const int* nc;     // [batch][N] active constraints per node (<= NC), or nullptr: NC everywhere
  int N, begin, count, with_rollout;
  double eps, alpha;
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ bool finite_bits(double v) { return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000; }
__device__ __forceinline__ double rsqrt_pivot(double d) {  // MUFU.RSQ64H seed + one third-order step (see riccati_wpp.cu)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}
// acc += sum_k a[k] * M[k] for a contiguous shared-memory column M, a in registers. Two FMAs per 16-byte load; par = 1 says that M sits
// an odd number of doubles after a 16-byte boundary (odd leading dimensions: the pairs then start at element 1). par is a
// compile-time constant at every call site once the loops are unrolled.
template <int LEN>
__device__ __forceinline__ double dot_col(const double (&a)[LEN], const double* M, double acc, int par = 0) {
  if (par) {
    acc = fma(a[0], M[0], acc);
#pragma unroll
    for (int k = 1; k + 1 < LEN; k += 2) {
      const double2 v = ld2(M + k);
      acc = fma(a[k], v.x, acc);
      acc = fma(a[k + 1], v.y, acc);
    }
    if (LEN % 2 == 0) acc = fma(a[LEN - 1], M[LEN - 1], acc);
  } else {
#pragma unroll
    for (int k = 0; k + 1 < LEN; k += 2) {
      const double2 v = ld2(M + k);
      acc = fma(a[k], v.x, acc);
      acc = fma(a[k + 1], v.y, acc);
    }
    if (LEN % 2 == 1) acc = fma(a[LEN - 1], M[LEN - 1], acc);
  }
  return acc;
}

// Two warps per CTA. Ballbot: capping the registers at 170 (6 CTAs = 12 warps per SM instead of 10) is worth +7 % (8.4 -> 9.0 M solves/s);
// the manipulator kernel (252 registers, nc = 3) loses 17 % to spills under the same cap and keeps the default.
template <int NX, int NU, int NC, bool NOM, bool EV>
__global__ void __launch_bounds__(64, (NX == 10 ? 6 : 1)) ilqr_rpl_kernel(const Args a) {
  using S = Shape<NX, NU, NC>;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)warp * S::warp_doubles;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(wbase + S::P * S::slot);
  unsigned long long* rbar = bar + 1;  // rollout ring: stage set landed
  const int graw = lane / NX;
  const bool in_group = graw < S::P;
  const int gi = in_group ? graw : S::P - 1;   // left-over lanes shadow the last slot and never store
  const int i = in_group ? lane - graw * NX : NX - 1;
  const int gbase = gi * NX;                   // first lane of the group
  double* sm = wbase + gi * S::slot;
  const int N = a.N;
  const uint32_t recBytes = S::rec * sizeof(double);
  const uint32_t dynBytes = S::dyn * sizeof(double);  // { A | B | Hv }: all the rollout reads of a record
  const uint32_t polBytes = S::pol * sizeof(double);  // { K | dbias }

  if (lane == 0) {
    mbar_init(bar, 1);
    for (int d = 0; d < S::ring; ++d) mbar_init(&rbar[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0, rphase = 0;  // rphase: bit d = parity to wait for on rbar[d]
  const int warps_total = gridDim.x * (blockDim.x >> 5);

  for (int base = (blockIdx.x * (blockDim.x >> 5) + warp) * S::P; base < a.count; base += warps_total * S::P) {
    const int nprob = (a.count - base) < S::P ? (a.count - base) : S::P;  // problems this warp carries in this round
    const bool valid = in_group && gi < nprob;
    const int prob = a.begin + base + (gi < nprob ? gi : nprob - 1);      // inactive groups shadow a live problem (never store)
    const double* lqp = a.lq + (size_t)prob * N * S::rec;
    const double* term = a.term + (size_t)prob * S::trec;
    double* solp = a.sol + (size_t)prob * (N + 1) * S::orec;
    const int* evp = EV ? a.event + (size_t)prob * N : nullptr;

    // records of node N-1 (one TMA copy per carried problem)
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, recBytes * nprob);
      for (int g = 0; g < nprob; ++g)
        tma_load(wbase + g * S::slot + S::sRec + ((N - 1) & 1) * S::rec, a.lq + ((size_t)(a.begin + base + g) * N + (N - 1)) * S::rec, recBytes, bar);
    }
    // terminal condition: valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526)
    double Srow[NX];
    double* outN = solp + (size_t)N * S::orec;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      Srow[j] = term[S::oQf + i + NX * j];
      if (valid) outN[S::oSm + i + NX * j] = Srow[j];
    }
    double Svi = term[S::oqf + i];
    if (valid) outN[S::oSv + i] = Svi;
    double sval = term[S::ocf];
    if (valid && i == 0) outN[S::os] = sval;
    bool pd = true, rank_ok = true;

#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
      mbar_wait(bar, parity);
      parity ^= 1u;
      const double* rec = sm + S::sRec + (k & 1) * S::rec;
      if (lane == 0 && k >= 1) {  // the other buffer was released by the __syncwarp that ended node k+1
        fence_proxy_async();
        mbar_expect_tx(bar, recBytes * nprob);
        for (int g = 0; g < nprob; ++g)
          tma_load(wbase + g * S::slot + S::sRec + ((k - 1) & 1) * S::rec, a.lq + ((size_t)(a.begin + base + g) * N + (k - 1)) * S::rec, recBytes, bar);
      }
      const double* A = rec + S::oA;
      const double* B = rec + S::oB;
      const double* Hv = rec + S::oHv;
      double* out = solp + (size_t)k * S::orec;
      // pre-event node (ILQR.cpp:263-295): A, Hv, Q, q, c of the record are the jump map and the pre-jump cost. The value function goes
      // through riccatiTransversalityConditions (S- = Q_e + A_e'S A_e, Sv- = q_e + A_e'w, s- = s + c_e + Hv.(w - S Hv / 2)), the controller
      // entry comes from B, P, R, r (and C, D, e) against a zero next value function and S-, Sv-: Hm = R, G = P + B'S-, Gv = r + B'Sv-.
      // The flag differs between the problems of a warp: selects, no branches, around the warp-wide shuffles and barriers.
      const bool ev = EV && __ldg(evp + k) != 0;

      // ---- w_i = Sv_i + S_i Hv ; SA_i = S_i A ; SB_i = S_i B ----
      const double shv = dot_col<NX>(Srow, Hv, 0.0);
      const double wi = Svi + shv;
      sm[S::sW + i] = wi;
      double spart = Hv[i] * (wi - 0.5 * shv);  // this row's share of Hv.w - 1/2 Hv.(S Hv)
#pragma unroll
      for (int j = 0; j < NX; ++j) sm[S::sSA + i + S::XP * j] = dot_col<NX>(Srow, A + NX * j, 0.0, (NX * j) & 1);
#pragma unroll
      for (int l = 0; l < NU; ++l) sm[S::sSB + i + S::XP * l] = dot_col<NX>(Srow, B + NX * l, 0.0, (NX * l) & 1);
      __syncwarp();

      // ---- T_i = Q_i + eps e_i + A(:,i)' SA ; tv_i = q_i + A(:,i)' w ----
      double acol[NX];
#pragma unroll
      for (int kk = 0; kk < NX; ++kk) acol[kk] = A[kk + NX * i];
      double Trow[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) Trow[j] = dot_col<NX>(acol, sm + S::sSA + S::XP * j, rec[S::oQ + i + NX * j] + ((j == i && !ev) ? a.eps : 0.0));
      const double tvi = dot_col<NX>(acol, sm + S::sW, rec[S::oq + i]);

      // ---- G'_i = P(:,i)' + SA(:,i)' B (row i of G') ----
      double gt[NU];
      {
        double sacol[NX];
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) sacol[kk] = ev ? Trow[kk] : sm[S::sSA + kk + S::XP * i];  // event: row i of S- (symmetric)
#pragma unroll
        for (int l = 0; l < NU; ++l) gt[l] = dot_col<NX>(sacol, B + NX * l, rec[S::oP + l + NU * i], (NX * l) & 1);
      }

      // ---- Hm_l = R_l + B(:,l)' SB and Gv_l = r_l + B(:,l)' w on lanes l < nu; Cholesky by shuffles inside the group ----
      double h[NU];
      if (EV) {  // Sv- replaces w for the event groups (every lane has read w by now)
        __syncwarp();
        if (ev) sm[S::sW + i] = tvi;
        __syncwarp();
      }
      {
        const int l = i < NU ? i : NU - 1;
        double bcol[NX];
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) bcol[kk] = B[kk + NX * l];
#pragma unroll
        for (int l2 = 0; l2 < NU; ++l2) {
          const double rl = rec[S::oR + l + NU * l2];
          const double hl = dot_col<NX>(bcol, sm + S::sSB + S::XP * l2, rl);
          h[l2] = ev ? rl : hl;
        }
        const double gv = dot_col<NX>(bcol, sm + S::sW, rec[S::or_ + l]);
        if (i < NU) sm[S::sGv + i] = gv;
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const double d = __shfl_sync(kFull, h[j], gbase + j);
        pd = pd && (__double2hiint(d) > 0);
        const double rs = rsqrt_pivot(d);  // a non-positive pivot gives NaN / Inf, which propagates like the reference's LLT
        const double lj = h[j] * rs;       // L[i][j] on lane i (i >= j); lane j holds sqrt(d)
        if (i < NU) sm[S::sL + i + NU * j] = (i == j) ? rs : lj;  // 1/L_jj on the diagonal
#pragma unroll
        for (int k2 = j + 1; k2 < NU; ++k2) {
          const double lk = __shfl_sync(kFull, lj, gbase + k2);
          h[k2] = fma(-lj, lk, h[k2]);
        }
      }
      __syncwarp();

      // ---- Y'_i = G'_i L^-T (forward substitution), K'_i = -Y'_i L^-1 (back substitution), Yv = L^-1 Gv, dbias = -L^-T Yv ----
      const double* Lm = sm + S::sL;
      double yt[NU], kt[NU], yv[NU], db[NU];
#pragma unroll
      for (int l = 0; l < NU; ++l) {
        double v = gt[l], vv = sm[S::sGv + l];
#pragma unroll
        for (int l2 = 0; l2 < l; ++l2) {
          v = fma(-Lm[l + NU * l2], yt[l2], v);
          vv = fma(-Lm[l + NU * l2], yv[l2], vv);
        }
        yt[l] = v * Lm[l + NU * l];
        yv[l] = vv * Lm[l + NU * l];
      }
      // ---- state-input equality constraints C x + D u + e = 0 (range-space form of the reference's projection, see the file header):
      //      Z = L^-1 D', M = Z'Z = L_M L_M' (L_M' is the R factor of the reference's QR of U^-T D'), Vx = L_M^-1 (Z'Y - C),
      //      vv = L_M^-1 (Z'Yv - e);  the constrained minimiser is Y^ = Y - Z L_M^-T Vx, Yv^ = Yv - Z L_M^-T vv and the value function
      //      gains + Vx'Vx, + Vx'vv, + 1/2 vv'vv ----
      double yh[NU], yvh[NU], vx[NC > 0 ? NC : 1], vv[NC > 0 ? NC : 1];
#pragma unroll
      for (int l = 0; l < NU; ++l) {
        yh[l] = yt[l];
        yvh[l] = yv[l];
      }
      if (NC > 0) {
        // rows beyond the node's active count are zero rows of D, C, e with a unit diagonal in M: they leave every result untouched, so
        // per-node (ragged) counts run the same code
        const int nca = a.nc ? __ldg(a.nc + (size_t)prob * N + k) : NC;
        {  // lane c < nc: column c of Z by forward substitution
          const int cc = i < NC ? i : NC - 1;
          const double keep = cc < nca ? 1.0 : 0.0;
          double z[NU];
#pragma unroll
          for (int l = 0; l < NU; ++l) {
            double v = keep * rec[S::oD + cc + NC * l];
#pragma unroll
            for (int l2 = 0; l2 < l; ++l2) v = fma(-Lm[l + NU * l2], z[l2], v);
            z[l] = v * Lm[l + NU * l];
          }
          if (i < NC) {
#pragma unroll
            for (int l = 0; l < NU; ++l) sm[S::sZ + l + NU * i] = z[l];
          }
        }
        __syncwarp();
        const double* Z = sm + S::sZ;
        // M = Z'Z and its Cholesky, redundantly in every lane (nc is tiny); 1/L_M,cc on the diagonal
        double lm[NC > 0 ? NC : 1][NC > 0 ? NC : 1];
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1)
#pragma unroll
          for (int c2 = 0; c2 <= c1; ++c2) {
            double v = 0.0;
#pragma unroll
            for (int l = 0; l < NU; ++l) v = fma(Z[l + NU * c1], Z[l + NU * c2], v);
            lm[c1][c2] = (c1 == c2 && c1 >= nca) ? 1.0 : v;
          }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          double d = lm[j][j];
          // the reference clamps |Rc_jj| = L_M,jj to 1e-9 (LinearAlgebra.cpp:38-47) and we flag it
          if (!(d >= 1e-18)) {
            rank_ok = false;
            d = 1e-18;
          }
          const double rs = rsqrt_pivot(d);
          lm[j][j] = rs;
#pragma unroll
          for (int c1 = j + 1; c1 < NC; ++c1) lm[c1][j] *= rs;
#pragma unroll
          for (int c1 = j + 1; c1 < NC; ++c1)
#pragma unroll
            for (int c2 = j + 1; c2 <= c1; ++c2) lm[c1][c2] = fma(-lm[c1][j], lm[c2][j], lm[c1][c2]);
        }
        // Vx(:,i) = L_M^-1 (Z'Y(:,i) - C(:,i)),  vv = L_M^-1 (Z'Yv - e)
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1) {
          double v = c1 < nca ? -rec[S::oC + c1 + NC * i] : 0.0, w2 = c1 < nca ? -rec[S::oe + c1] : 0.0;
#pragma unroll
          for (int l = 0; l < NU; ++l) {
            v = fma(Z[l + NU * c1], yt[l], v);
            w2 = fma(Z[l + NU * c1], yv[l], w2);
          }
#pragma unroll
          for (int c2 = 0; c2 < c1; ++c2) {
            v = fma(-lm[c1][c2], vx[c2], v);
            w2 = fma(-lm[c1][c2], vv[c2], w2);
          }
          vx[c1] = v * lm[c1][c1];
          vv[c1] = w2 * lm[c1][c1];
        }
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1) sm[S::sVx + i + S::XP * c1] = vx[c1];
        // t = L_M^-T Vx(:,i), tv = L_M^-T vv;  Y^ = Y - Z t, Yv^ = Yv - Z tv
        double tx[NC > 0 ? NC : 1], tvv[NC > 0 ? NC : 1];
#pragma unroll
        for (int c1 = NC - 1; c1 >= 0; --c1) {
          double v = vx[c1], w2 = vv[c1];
#pragma unroll
          for (int c2 = c1 + 1; c2 < NC; ++c2) {
            v = fma(-lm[c2][c1], tx[c2], v);
            w2 = fma(-lm[c2][c1], tvv[c2], w2);
          }
          tx[c1] = v * lm[c1][c1];
          tvv[c1] = w2 * lm[c1][c1];
        }
#pragma unroll
        for (int l = 0; l < NU; ++l)
#pragma unroll
          for (int c1 = 0; c1 < NC; ++c1) {
            yh[l] = fma(-Z[l + NU * c1], tx[c1], yh[l]);
            yvh[l] = fma(-Z[l + NU * c1], tvv[c1], yvh[l]);
          }
      }
#pragma unroll
      for (int l = NU - 1; l >= 0; --l) {
        double v = -yh[l], vv2 = -yvh[l];
#pragma unroll
        for (int l2 = l + 1; l2 < NU; ++l2) {
          v = fma(-Lm[l2 + NU * l], kt[l2], v);
          vv2 = fma(-Lm[l2 + NU * l], db[l2], vv2);
        }
        kt[l] = v * Lm[l + NU * l];
        db[l] = vv2 * Lm[l + NU * l];
      }
#pragma unroll
      for (int l = 0; l < NU; ++l) sm[S::sYt + i + S::XP * l] = yt[l];
      if (valid) {
#pragma unroll
        for (int l = 0; l < NU; ++l) out[S::oK + l + NU * i] = kt[l];  // column i of K: the group writes nu*nx contiguous doubles
        if (i < NU) {
          double dbi = db[0];
#pragma unroll
          for (int l = 1; l < NU; ++l) dbi = (i == l) ? db[l] : dbi;
          out[S::odb + i] = dbi;
          if (!NOM) out[S::obias + i] = 0.0;
        }
      }
      if (NOM) {  // bias = u_nom - K x_nom (GaussNewtonDDP.cpp:604-606): this lane's column of K times x_nom_i, summed over the group
        const double xni = __ldg(a.x_nom + ((size_t)prob * (N + 1) + k) * NX + i);
#pragma unroll
        for (int l = 0; l < NU; ++l) sm[S::sK + l + NU * i] = kt[l] * xni;
        __syncwarp();
        const int l = i < NU ? i : NU - 1;
        double kx = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) kx += sm[S::sK + l + NU * j];
        if (valid && i < NU) out[S::obias + i] = __ldg(a.u_nom + ((size_t)prob * (N + 1) + k) * NU + i) - kx;
      }
      __syncwarp();

      // ---- S_i = T_i - Y'_i Y'' ; Sv_i = tv_i - Y'_i Yv ; s ----
#pragma unroll
      for (int j = 0; j < NX; ++j) Srow[j] = Trow[j];
#pragma unroll
      for (int l = 0; l < NU; ++l) {  // column l of Y' (contiguous, 16-byte aligned: leading dimension XP) updates the whole row
        const double* col = sm + S::sYt + S::XP * l;
#pragma unroll
        for (int j = 0; j + 1 < NX; j += 2) {
          const double2 v = ld2(col + j);
          Srow[j] = fma(-yt[l], v.x, Srow[j]);
          Srow[j + 1] = fma(-yt[l], v.y, Srow[j + 1]);
        }
        if (NX % 2 == 1) Srow[NX - 1] = fma(-yt[l], col[NX - 1], Srow[NX - 1]);
      }
#pragma unroll
      for (int c1 = 0; c1 < NC; ++c1) {
        const double* col = sm + S::sVx + S::XP * c1;
#pragma unroll
        for (int j = 0; j + 1 < NX; j += 2) {
          const double2 v = ld2(col + j);
          Srow[j] = fma(vx[c1], v.x, Srow[j]);
          Srow[j + 1] = fma(vx[c1], v.y, Srow[j + 1]);
        }
        if (NX % 2 == 1) Srow[NX - 1] = fma(vx[c1], col[NX - 1], Srow[NX - 1]);
      }
      Svi = tvi;
#pragma unroll
      for (int l = 0; l < NU; ++l) Svi = fma(-yt[l], yv[l], Svi);
#pragma unroll
      for (int c1 = 0; c1 < NC; ++c1) Svi = fma(vx[c1], vv[c1], Svi);
      if (ev) {
        Svi = tvi;
#pragma unroll
        for (int j = 0; j < NX; ++j) Srow[j] = Trow[j];
      }
      // s: rows sum their shares inside the group (xor butterflies stay inside aligned power-of-two blocks only when nx is one,
      // so the group sum goes through shared memory)
      sm[S::sX + i] = spart;
      __syncwarp();
      {
        double sh = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) sh += sm[S::sX + j];
        double yy = 0.0;
#pragma unroll
        for (int l = 0; l < NU; ++l) yy = fma(yv[l], yv[l], yy);
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1) yy = fma(-vv[c1], vv[c1], yy);
        sval = sval + rec[S::oc] + sh - (ev ? 0.0 : 0.5 * yy);
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < NX; ++j) out[S::oSm + i + NX * j] = Srow[j];
        out[S::oSv + i] = Svi;
        if (i == 0) out[S::os] = sval;
      }
      __syncwarp();  // every lane is done with this node's record and scratch
    }

    // node N of the controller := node N-1 (GaussNewtonDDP.cpp:609-618)
    if (valid) {
      const double* src = solp + (size_t)(N - 1) * S::orec;
      double* dst = solp + (size_t)N * S::orec;
      for (int e = i; e < S::oSm; e += NX) dst[e] = __ldcg(src + e);  // K | dbias | bias precede Sm in the record
    }
    // status: a non-finite value anywhere in the sweep propagates into S, Sv, s of node 0
    bool finite = finite_bits(Svi) && finite_bits(sval);
#pragma unroll
    for (int j = 0; j < NX; ++j) finite = finite && finite_bits(Srow[j]);
    const unsigned gmask = (NX == 32) ? kFull : (((1u << NX) - 1u) << gbase);
    const unsigned bad_pd = __ballot_sync(kFull, !pd), bad_fin = __ballot_sync(kFull, !finite), bad_rank = __ballot_sync(kFull, !rank_ok);
    int bits = ((bad_pd & gmask) ? O2C_STATUS_CHOL_NOT_PD : 0) | ((bad_fin & gmask) ? O2C_STATUS_NONFINITE : 0) |
               ((bad_rank & gmask) ? O2C_STATUS_CONSTRAINT_RANK : 0);
    if (!a.with_rollout) {
      if (valid && i == 0) a.status[prob] = bits;
      __syncwarp();
      continue;
    }

    // ---- fused forward rollout: du_k = K_k dx_k + alpha dbias_k ; dx_{k+1} = A_k dx_k + B_k du_k + Hv_k ----
    // Everything a stage reads — { A | B | Hv } of the record and { K | dbias } of the solution — comes by TMA through a ring of
    // S::ring stage sets per problem, issued S::ring stages ahead, so that the short dependent chain of a stage never waits for a DRAM
    // round trip (one stage ahead for the record and plain loads for the gains cost 2.3 us per stage: profiles/r02_prof_modes.jsonl).
    fence_proxy_async_global();  // this warp's K / dbias stores of the sweep are read back through the async proxy
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      for (int d = 0; d < S::ring && d < N; ++d) {
        mbar_expect_tx(&rbar[d], (dynBytes + polBytes) * nprob);
        for (int g = 0; g < nprob; ++g) {
          double* dst = wbase + g * S::slot + d * S::set;
          tma_load(dst, a.lq + ((size_t)(a.begin + base + g) * N + d) * S::rec, dynBytes, &rbar[d]);
          tma_load(dst + S::dyn, a.sol + ((size_t)(a.begin + base + g) * (N + 1) + d) * S::orec + S::oK, polBytes, &rbar[d]);
        }
      }
    }
    double* xo = a.xs + (size_t)prob * (N + 1) * NX;
    double* uo = a.us + (size_t)prob * (N + 1) * NU;
    // with nominal trajectories the rollout runs in deviation coordinates (dx = x - x_nom, du = u - u_nom) and shifts the outputs back
    const double* xnp = NOM ? a.x_nom + (size_t)prob * (N + 1) * NX : nullptr;
    const double* unp = NOM ? a.u_nom + (size_t)prob * (N + 1) * NU : nullptr;
    double xnk = NOM ? __ldg(xnp + i) : 0.0, unk = NOM ? __ldg(unp + (i < NU ? i : 0)) : 0.0;  // nominal state / input of the current node
    double x = a.x0[(size_t)prob * NX + i] - xnk;
    bool xfinite = true;
    bool jump = EV && __ldg(evp) != 0;  // pre-event node: x+ = A_e x + Hv_e, the input does not enter the jump map
    int rs = 0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      sm[S::sX + i] = x;
      mbar_wait(&rbar[rs], (rphase >> rs) & 1u);
      rphase ^= 1u << rs;
      __syncwarp();
      const double* rec = sm + rs * S::set;
      const double* pol = rec + S::dyn;
      // u_l = alpha dbias_l + K(l,:) x on lanes l < nu
      {
        const int l = i < NU ? i : NU - 1;
        double u = a.alpha * pol[S::odb + l];
#pragma unroll
        for (int j = 0; j < NX; ++j) u = fma(pol[l + NU * j], sm[S::sX + j], u);
        if (i < NU) {
          sm[S::sU + i] = u;
          if (valid) __stcg(uo + (size_t)k * NU + i, u + unk);
        }
      }
      if (valid) __stcg(xo + (size_t)k * NX + i, x + xnk);
      xfinite = xfinite && finite_bits(x);
      if (NOM) {  // next node's nominal values: in flight during the second half of the stage
        xnk = __ldg(xnp + (size_t)(k + 1) * NX + i);
        unk = __ldg(unp + (size_t)(k + 1) * NU + (i < NU ? i : 0));
      }
      const bool jump_next = EV && k + 1 < N && __ldg(evp + k + 1) != 0;
      double xn = rec[S::oHv + i];
#pragma unroll
      for (int j = 0; j < NX; ++j) xn = fma(rec[S::oA + i + NX * j], sm[S::sX + j], xn);
      __syncwarp();  // u of every lane is in shared memory
      if (!jump) {
#pragma unroll
        for (int l = 0; l < NU; ++l) xn = fma(rec[S::oB + i + NX * l], sm[S::sU + l], xn);
      }
      jump = jump_next;
      __syncwarp();  // every lane is done with x, u and the stage set
      if (lane == 0 && k + S::ring < N) {
        mbar_expect_tx(&rbar[rs], (dynBytes + polBytes) * nprob);
        for (int g = 0; g < nprob; ++g) {
          double* dst = wbase + g * S::slot + rs * S::set;
          tma_load(dst, a.lq + ((size_t)(a.begin + base + g) * N + (k + S::ring)) * S::rec, dynBytes, &rbar[rs]);
          tma_load(dst + S::dyn, a.sol + ((size_t)(a.begin + base + g) * (N + 1) + (k + S::ring)) * S::orec + S::oK, polBytes, &rbar[rs]);
        }
      }
      x = xn;
      rs = rs + 1 == S::ring ? 0 : rs + 1;
    }
    // node N: state, and the input of the copied last policy (K, dbias of node N-1, still in their ring slot) re-evaluated at x_N
    // (TimeTriggeredRollout.cpp:98-102); with nominal trajectories the copied policy acts on the absolute state: u = bias + alpha dbias + K x
    const double xabs = x + xnk;
    sm[S::sX + i] = xabs;
    __syncwarp();
    {
      const double* pol = sm + ((N - 1) % S::ring) * S::set + S::dyn;
      const int l = i < NU ? i : NU - 1;
      double u = a.alpha * pol[S::odb + l] + (NOM ? __ldcg(solp + (size_t)N * S::orec + S::obias + l) : 0.0);
#pragma unroll
      for (int j = 0; j < NX; ++j) u = fma(pol[l + NU * j], sm[S::sX + j], u);
      if (valid && i < NU) __stcg(uo + (size_t)N * NU + i, u);
    }
    if (valid) __stcg(xo + (size_t)N * NX + i, xabs);
    xfinite = xfinite && finite_bits(x);
    if (__ballot_sync(kFull, !xfinite) & gmask) bits |= O2C_STATUS_NONFINITE;
    if (valid && i == 0) a.status[prob] = bits;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// SLQ: the continuous-time Riccati flow map under fixed-step RK4, same row-per-lane mapping (nc = 0, LINE_SEARCH, reduced form,
// DIAGONAL_SHIFT). Per node k the projection with Hm = R (SLQ.cpp:183-208): R = L L', Pu = L^-T, B~ = B Pu, P~ = Pu'P, r~ = Pu'r;
// the flow map (ContinuousTimeRiccatiEquations.cpp:170-292) on the data lerped between the two nodes of the interval:
//   G~m = P~ + B~'S, G~v = r~ + B~'Sv,  dS/dz = Q + eps I + S A + (S A)' - G~m'G~m,  dSv/dz = q + S Hv + A'Sv - G~m'G~v,
//   ds/dz = c + Hv.Sv - 1/2 G~v.G~v;  the RK4 step schedule (boost::odeint integrate_times semantics) is precomputed on the host
//   (SlqStep). Lane i integrates row i of S (and Sv_i; s redundantly); A, B~, S A, G~m' are broadcast from shared memory.
//   Controller at an observed node (SLQ.cpp:127-169): K(:,i) = -L^-T (P~(:,i) + B~'S(:,i)), dbias = -L^-T (r~ + B~'Sv).
// Stage records arrive by TMA into a ring of three slots (nodes k, k+1 in use, node k-1 in flight).
// ---------------------------------------------------------------------------------------------------------------------
template <int NX, int NU>
struct SlqShape {
  using R = Shape<NX, NU, 0>;
  static constexpr int P = 32 / NX;
  static constexpr int rec = R::rec;
  static constexpr int sRec = 0, sAl = 3 * rec, sBl = sAl + cpad2(NX * NX),
                       sMs = sBl + cpad2(NX * NU), sGs = sMs + cpad2(NX * NX), sSv = sGs + cpad2(NX * NU), sHv = sSv + cpad2(NX),
                       sGv = sHv + cpad2(NX), slot = sGv + cpad2(NU);
  static constexpr int warp_doubles = P * slot + 2;
};

struct SlqArgs {
  const double* lq;
  const double* term;
  const double* x_nom;  // [batch][N+1][nx] or nullptr
  const double* u_nom;
  double* sol;
  int* status;
  const SlqStep* steps;
  const double* jump;  // [batch][jump_capacity] jump records (o2c_common.cuh: jump_rec) or nullptr
  int jump_capacity;
  int nsteps, N, begin, count;
  double eps;
};

template <int NX, int NU, bool EV>
__global__ void __launch_bounds__(64) slq_rpl_kernel(const SlqArgs a) {
  using S = SlqShape<NX, NU>;
  using R = typename S::R;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)warp * S::warp_doubles;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(wbase + S::P * S::slot);
  const int graw = lane / NX;
  const bool in_group = graw < S::P;
  const int gi = in_group ? graw : S::P - 1;
  const int i = in_group ? lane - graw * NX : NX - 1;
  const int gbase = gi * NX;
  double* sm = wbase + gi * S::slot;
  const int N = a.N, nodes = N + 1;
  const uint32_t recBytes = S::rec * sizeof(double);

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
  const int warps_total = gridDim.x * (blockDim.x >> 5);

  for (int base = (blockIdx.x * (blockDim.x >> 5) + warp) * S::P; base < a.count; base += warps_total * S::P) {
    const int nprob = (a.count - base) < S::P ? (a.count - base) : S::P;
    const bool valid = in_group && gi < nprob;
    const int prob = a.begin + base + (gi < nprob ? gi : nprob - 1);
    const double* term = a.term + (size_t)prob * R::trec;
    double* solp = a.sol + (size_t)prob * (N + 1) * R::orec;

    auto issue_node = [&](int k) {  // records of node k of every carried problem -> ring slot k % 3
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, recBytes * nprob);
        for (int g = 0; g < nprob; ++g)
          tma_load(wbase + g * S::slot + S::sRec + (k % 3) * S::rec, a.lq + ((size_t)(a.begin + base + g) * nodes + k) * S::rec, recBytes, bar);
      }
    };

    // state of the integration: row i of S, Sv_i, s
    double Srow[NX], Svi, sval;
#pragma unroll
    for (int j = 0; j < NX; ++j) Srow[j] = term[R::oQf + ((i <= j) ? i + NX * j : j + NX * i)];  // convert2Vector keeps the upper triangle
    Svi = term[R::oqf + i];
    sval = term[R::ocf];
    bool pd = true;

    // node-resident lane-private data: set 0 = node i0 (left end of the interval), set 1 = node i0 + 1
    double qr0[NX], qr1[NX], pt0[NU], pt1[NU], rt0[NU], rt1[NU], hv0 = 0.0, hv1 = 0.0, q0 = 0.0, q1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) qr0[j] = qr1[j] = 0.0;
#pragma unroll
    for (int l = 0; l < NU; ++l) pt0[l] = pt1[l] = rt0[l] = rt1[l] = 0.0;

    auto write_value = [&](int k) {
      if (valid) {
        double* out = solp + (size_t)k * R::orec;
#pragma unroll
        for (int j = 0; j < NX; ++j) out[R::oSm + i + NX * j] = Srow[j];
        out[R::oSv + i] = Svi;
        if (i == 0) out[R::os] = sval;
      }
    };
    // projection of node k (record in its ring slot) into set 0; B~ and L replace B and R inside the record (both are dead once
    // projected), which keeps the warp at 27 KB of shared memory: 8 one-warp CTAs per SM
    auto project_node = [&](int k) {
      double* rec = sm + S::sRec + (k % 3) * S::rec;
      double Lr[NU][NU];
#pragma unroll
      for (int l = 0; l < NU; ++l)
#pragma unroll
        for (int l2 = 0; l2 <= l; ++l2) Lr[l][l2] = rec[R::oR + l + NU * l2];
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const double d = Lr[j][j];
        pd = pd && (__double2hiint(d) > 0);
        const double rs = rsqrt_pivot(d);
        Lr[j][j] = rs;
#pragma unroll
        for (int l = j + 1; l < NU; ++l) Lr[l][j] *= rs;
#pragma unroll
        for (int l = j + 1; l < NU; ++l)
#pragma unroll
          for (int l2 = j + 1; l2 <= l; ++l2) Lr[l][l2] = fma(-Lr[l][j], Lr[l2][j], Lr[l][l2]);
      }
      double* Lm = rec + R::oR;
      double* Bt = rec + R::oB;
      __syncwarp();  // every lane has read R
      if (in_group && i == 0) {
#pragma unroll
        for (int l = 0; l < NU; ++l)
#pragma unroll
          for (int l2 = 0; l2 <= l; ++l2) Lm[l + NU * l2] = Lr[l][l2];
      }
      double bt[NU];
#pragma unroll
      for (int l = 0; l < NU; ++l) {  // B~_i = B_i L^-T, P~(:,i) = L^-1 P(:,i), r~ = L^-1 r: forward substitutions
        double vb = rec[R::oB + i + NX * l], vp = rec[R::oP + l + NU * i], vr = rec[R::or_ + l];
#pragma unroll
        for (int l2 = 0; l2 < l; ++l2) {
          vb = fma(-Lr[l][l2], bt[l2], vb);
          vp = fma(-Lr[l][l2], pt0[l2], vp);
          vr = fma(-Lr[l][l2], rt0[l2], vr);
        }
        bt[l] = vb * Lr[l][l];
        pt0[l] = vp * Lr[l][l];
        rt0[l] = vr * Lr[l][l];
        if (in_group) Bt[i + NX * l] = bt[l];  // row i in place (the left-over lanes mirror a row and must not race with its owner)
      }
#pragma unroll
      for (int j = 0; j < NX; ++j) qr0[j] = rec[R::oQ + ((i <= j) ? i + NX * j : j + NX * i)];
      hv0 = rec[R::oHv + i];
      q0 = rec[R::oq + i];
      c0 = rec[R::oc];
      __syncwarp();
    };
    // SLQ::calculateControllerWorker at node k (whose projection is set 0) from the current state
    auto controller = [&](int k) {
      const double* Lm = sm + S::sRec + (k % 3) * S::rec + R::oR;
      const double* Bt = sm + S::sRec + (k % 3) * S::rec + R::oB;
      sm[S::sSv + i] = Svi;
      __syncwarp();
      double kt[NU], db[NU], g[NU], gv[NU];
#pragma unroll
      for (int l = 0; l < NU; ++l) {
        g[l] = dot_col<NX>(Srow, Bt + NX * l, pt0[l]);
        double v = rt0[l];
        if (NX % 2 == 0) {
#pragma unroll
          for (int kk = 0; kk < NX; kk += 2) {
            const double2 bb = ld2(Bt + NX * l + kk), ss = ld2(sm + S::sSv + kk);
            v = fma(bb.x, ss.x, v);
            v = fma(bb.y, ss.y, v);
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < NX; ++kk) v = fma(Bt[NX * l + kk], sm[S::sSv + kk], v);
        }
        gv[l] = v;
      }
#pragma unroll
      for (int l = NU - 1; l >= 0; --l) {  // back substitution with L': x = L^-T (-g)
        double v = -g[l], vv = -gv[l];
#pragma unroll
        for (int l2 = l + 1; l2 < NU; ++l2) {
          v = fma(-Lm[l2 + NU * l], kt[l2], v);
          vv = fma(-Lm[l2 + NU * l], db[l2], vv);
        }
        kt[l] = v * Lm[l + NU * l];
        db[l] = vv * Lm[l + NU * l];
      }
      if (valid) {
        double* out = solp + (size_t)k * R::orec;
#pragma unroll
        for (int l = 0; l < NU; ++l) out[R::oK + l + NU * i] = kt[l];
        if (i < NU) {
          double dbi = db[0];
#pragma unroll
          for (int l = 1; l < NU; ++l) dbi = (i == l) ? db[l] : dbi;
          out[R::odb + i] = dbi;
          if (!a.x_nom) out[R::obias + i] = 0.0;
        }
      }
      if (a.x_nom) {  // bias = u_nom - K x_nom (GaussNewtonDDP.cpp:604-606); sAl is free between flow-map evaluations
        const double xni = __ldg(a.x_nom + ((size_t)prob * (N + 1) + k) * NX + i);
#pragma unroll
        for (int l = 0; l < NU; ++l) sm[S::sAl + l + NU * i] = kt[l] * xni;
        __syncwarp();
        const int l = i < NU ? i : NU - 1;
        double kx = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) kx += sm[S::sAl + l + NU * j];
        if (valid && i < NU) solp[(size_t)k * R::orec + R::obias + i] = __ldg(a.u_nom + ((size_t)prob * (N + 1) + k) * NU + i) - kx;
      }
      __syncwarp();
    };

    write_value(N);
    issue_node(N);
    mbar_wait(bar, parity);
    parity ^= 1u;
    project_node(N);
    controller(N);  // overwritten below by the copy of node N-1 (GaussNewtonDDP.cpp:609-618)
    if (N >= 1) issue_node(N - 1);
    int loaded_lo = N;

#pragma unroll 1
    for (int sidx = 0; sidx < a.nsteps; ++sidx) {
      const SlqStep sp = a.steps[sidx];
      const int i0 = sp.interval;
      if (i0 < loaded_lo) {  // the interval moved one node down: set 1 <- set 0, set 0 <- node i0
#pragma unroll
        for (int j = 0; j < NX; ++j) qr1[j] = qr0[j];
#pragma unroll
        for (int l = 0; l < NU; ++l) {
          pt1[l] = pt0[l];
          rt1[l] = rt0[l];
        }
        hv1 = hv0;
        q1 = q0;
        c1 = c0;
        mbar_wait(bar, parity);
        parity ^= 1u;
        project_node(i0);
        if (i0 >= 1) issue_node(i0 - 1);  // its ring slot held node i0 + 2, which is dead
        loaded_lo = i0;
      }
      const double* A0 = sm + S::sRec + (i0 % 3) * S::rec + R::oA;
      const double* A1 = sm + S::sRec + ((i0 + 1) % 3) * S::rec + R::oA;
      const double* Bt0 = sm + S::sRec + (i0 % 3) * S::rec + R::oB;
      const double* Bt1 = sm + S::sRec + ((i0 + 1) % 3) * S::rec + R::oB;
      const double h = sp.h;
      if (EV && sp.jump > 0) {  // (template variant: the event-free kernel is at its register limit)
        // node i0 is a pre-event node: instead of integrating, the value function crosses the event through
        // ContinuousTimeRiccatiEquations::computeJumpMap = riccatiTransversalityConditions on the event's jump model data
        // (SLQ.cpp:286-296, RiccatiTransversalityConditions.h:40-56): S- = Q_e + (S A_e)' A_e, Sv- = q_e + A_e'(Sv + S Hv_e),
        // s- = s + c_e + Hv_e.(Sv + S Hv_e / 2). Row i of S A_e per lane, columns through shared memory.
        const double* jr = a.jump + ((size_t)prob * a.jump_capacity + (sp.jump - 1)) * jump_rec(NX);
        const double* Ae = jr;
        const double* Hve = jr + jump_oHv(NX);
        double shv = 0.0;
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) shv = fma(Srow[kk], __ldg(Hve + kk), shv);
        sm[S::sSv + i] = Svi + shv;                               // w = Sv + S Hv_e
        sm[S::sHv + i] = __ldg(Hve + i) * (Svi + 0.5 * shv);      // this row's share of Hv_e.(Sv + S Hv_e / 2)
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          double v = 0.0;
#pragma unroll
          for (int kk = 0; kk < NX; ++kk) v = fma(Srow[kk], __ldg(Ae + kk + NX * j), v);
          sm[S::sMs + i + NX * j] = v;                            // (S A_e)(i, j)
        }
        __syncwarp();
        double mcol[NX];
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) mcol[kk] = sm[S::sMs + kk + NX * i];  // column i of S A_e
        double svn = __ldg(jr + jump_oq(NX) + i), ssum = 0.0;
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) {
          svn = fma(__ldg(Ae + kk + NX * i), sm[S::sSv + kk], svn);
          ssum += sm[S::sHv + kk];
        }
        __syncwarp();  // sMs is read; it now takes the upper triangle of S- so that both triangles carry the same numbers
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          if (j >= i) {  // convert2Vector keeps the upper triangle (row <= column)
            double v = 0.0;
#pragma unroll
            for (int kk = 0; kk < NX; ++kk) v = fma(mcol[kk], __ldg(Ae + kk + NX * j), v);
            sm[S::sMs + i + NX * j] = __ldg(jr + jump_oQ(NX) + i + NX * j) + v;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NX; ++j) Srow[j] = (j >= i) ? sm[S::sMs + i + NX * j] : sm[S::sMs + j + NX * i];
        Svi = svn;
        sval = sval + __ldg(jr + jump_oc(NX)) + ssum;
        __syncwarp();
        if (sp.observe_node >= 0) {
          write_value(sp.observe_node);
          controller(sp.observe_node);
        }
        continue;
      }
      // classic RK4 (boost::odeint runge_kutta4): k_s = f(y + c_s h k_{s-1}), y += h (k1 + 2 k2 + 2 k3 + k4) / 6
      double ys[NX], ysv = Svi, acc[NX], accv = Svi, accs = sval;
#pragma unroll
      for (int j = 0; j < NX; ++j) ys[j] = acc[j] = Srow[j];
      double prev_al = -1.0;  // lerp weights lie in [0, 1]
#pragma unroll 1
      for (int stg = 0; stg < 4; ++stg) {
        const double al = sp.alpha[stg], be = 1.0 - al;
        // lerped data: column i of A, row i of B~ -> shared (broadcast operands); lane-private pieces stay in registers. Stages 2 and 3 of
        // RK4 share their time, hence their data: the shared operands are still in place (warp-uniform test on the schedule)
        if (al != prev_al) {
          if (NX % 2 == 0) {
#pragma unroll
            for (int kk = 0; kk < NX; kk += 2) {
              const double2 v0 = ld2(A0 + kk + NX * i), v1 = ld2(A1 + kk + NX * i);
              *reinterpret_cast<double2*>(sm + S::sAl + kk + NX * i) = make_double2(fma(al, v0.x, be * v1.x), fma(al, v0.y, be * v1.y));
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < NX; ++kk) sm[S::sAl + kk + NX * i] = fma(al, A0[kk + NX * i], be * A1[kk + NX * i]);
          }
#pragma unroll
          for (int l = 0; l < NU; ++l) sm[S::sBl + i + NX * l] = fma(al, Bt0[i + NX * l], be * Bt1[i + NX * l]);
          sm[S::sHv + i] = fma(al, hv0, be * hv1);
          prev_al = al;
        }
        double ptl[NU];
#pragma unroll
        for (int l = 0; l < NU; ++l) ptl[l] = fma(al, pt0[l], be * pt1[l]);
        sm[S::sSv + i] = ysv;
        __syncwarp();
        // M_i = S_i A, G'_i = P~(:,i)' + S_i B~, G~v on lanes < nu
        double Mrow[NX], gt[NU];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          Mrow[j] = dot_col<NX>(ys, sm + S::sAl + NX * j, 0.0);
          sm[S::sMs + i + NX * j] = Mrow[j];
        }
#pragma unroll
        for (int l = 0; l < NU; ++l) {
          gt[l] = dot_col<NX>(ys, sm + S::sBl + NX * l, ptl[l]);
          sm[S::sGs + i + NX * l] = gt[l];
        }
        {
          const int l = i < NU ? i : NU - 1;
          double rl = fma(al, rt0[0], be * rt1[0]);
#pragma unroll
          for (int l2 = 1; l2 < NU; ++l2) rl = (l == l2) ? fma(al, rt0[l2], be * rt1[l2]) : rl;
          double v = rl;
#pragma unroll
          for (int kk = 0; kk < NX; ++kk) v = fma(sm[S::sBl + kk + NX * l], sm[S::sSv + kk], v);
          if (i < NU) sm[S::sGv + i] = v;
        }
        __syncwarp();
        // derivatives of this lane's row
        double gv[NU];
#pragma unroll
        for (int l = 0; l < NU; ++l) gv[l] = sm[S::sGv + l];
        double kS[NX], kv, ks;
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          double v = fma(al, qr0[j], be * qr1[j]) + ((j == i) ? a.eps : 0.0) + Mrow[j] + sm[S::sMs + j + NX * i];
#pragma unroll
          for (int l = 0; l < NU; ++l) v = fma(-gt[l], sm[S::sGs + j + NX * l], v);
          kS[j] = v;
        }
        {
          double v = fma(al, q0, be * q1);
          double hs = 0.0;
#pragma unroll
          for (int kk = 0; kk < NX; ++kk) {
            const double hvk = sm[S::sHv + kk], svk = sm[S::sSv + kk];
            v = fma(ys[kk], hvk, v);                       // S_i Hv
            v = fma(sm[S::sAl + kk + NX * i], svk, v);     // A(:,i)' Sv
            hs = fma(hvk, svk, hs);                        // Hv . Sv
          }
          double gg = 0.0;
#pragma unroll
          for (int l = 0; l < NU; ++l) {
            v = fma(-gt[l], gv[l], v);
            gg = fma(gv[l], gv[l], gg);
          }
          kv = v;
          ks = fma(al, c0, be * c1) + hs - 0.5 * gg;
        }
        __syncwarp();  // the shared operands of this stage are dead
        // accumulate and form the next stage's argument
        const double bw = h * ((stg == 0 || stg == 3) ? (1.0 / 6.0) : (1.0 / 3.0));
        const double cw = h * ((stg == 2) ? 1.0 : 0.5);
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          acc[j] = fma(bw, kS[j], acc[j]);
          ys[j] = fma(cw, kS[j], Srow[j]);
        }
        accv = fma(bw, kv, accv);
        accs = fma(bw, ks, accs);
        ysv = fma(cw, kv, Svi);
      }
#pragma unroll
      for (int j = 0; j < NX; ++j) Srow[j] = acc[j];
      Svi = accv;
      sval = accs;
      if (sp.observe_node >= 0) {
        write_value(sp.observe_node);
        controller(sp.observe_node);
      }
    }
    __syncwarp();
    if (valid && N >= 1) {  // node N of the controller := node N-1 (GaussNewtonDDP.cpp:609-618)
      const double* src = solp + (size_t)(N - 1) * R::orec;
      double* dst = solp + (size_t)N * R::orec;
      for (int e = i; e < R::oSm; e += NX) dst[e] = __ldcg(src + e);
    }
    bool finite = finite_bits(Svi) && finite_bits(sval);
#pragma unroll
    for (int j = 0; j < NX; ++j) finite = finite && finite_bits(Srow[j]);
    const unsigned gmask = (NX == 32) ? kFull : (((1u << NX) - 1u) << gbase);
    const unsigned bad_pd = __ballot_sync(kFull, !pd), bad_fin = __ballot_sync(kFull, !finite);
    const int bits = ((bad_pd & gmask) ? O2C_STATUS_CHOL_NOT_PD : 0) | ((bad_fin & gmask) ? O2C_STATUS_NONFINITE : 0);
    if (valid && i == 0) a.status[prob] = bits;
    __syncwarp();
  }
}

template <int NX, int NU, bool EV>
cudaError_t launch_slq(const SlqArgs& a, cudaStream_t stream) {
  using S = SlqShape<NX, NU>;
  constexpr int wpb = 1;
  const size_t smem = (size_t)S::warp_doubles * wpb * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(slq_rpl_kernel<NX, NU, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int num_sms = device_sm_count();
  if (num_sms <= 0) return cudaErrorInvalidDevice;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, slq_rpl_kernel<NX, NU, EV>, wpb * 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int per_cta = wpb * S::P;
  const int needed = (a.count + per_cta - 1) / per_cta;
  const int cap = num_sms * ctas_per_sm;
  slq_rpl_kernel<NX, NU, EV><<<needed < cap ? needed : cap, wpb * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Continuous rollout of the LQ model under the LinearController (SLQ): xdot = A(t) x + B(t) u(t,x) + Hv(t),
// u = lerp(bias + alpha dbias)(t) + lerp(K)(t) x, classic RK4 on the host-precomputed step schedule (RolloutStep: the timeSegment
// index / weight of every stage and observation time; TimeTriggeredRollout.cpp:46-115, LinearController.cpp:79-87,
// DDP_HelperFunctions.cpp:296-304). Lane i owns x_i; x and u are broadcast through shared memory. The {A|B|Hv} part of the stage
// records and the {K|dbias|bias} part of the solution records stream forwards through a three-slot TMA ring (the two nodes of the
// current time segment in use, the next one in flight); the lerps are applied on the fly to this lane's row.
// ---------------------------------------------------------------------------------------------------------------------
template <int NX, int NU>
struct RoShape {
  using R = Shape<NX, NU, 0>;
  static constexpr int P = 32 / NX;
  static constexpr int dyn = R::oq;      // { A | B | Hv } doubles of a stage record
  static constexpr int pol = R::oSm;     // { K | dbias | bias } doubles of a solution record
  static constexpr int sDyn = 0, sPol = 3 * dyn, sX = sPol + 3 * pol, sU = sX + cpad2(NX), sDx = sU + cpad2(NU), slot = sDx + cpad2(NX);
  static constexpr int warp_doubles = P * slot + 2;
};

struct RoArgs {
  const double* lq;
  const double* sol;
  const double* x0;
  const double* x_nom;  // [batch][N+1][nx] or nullptr: xdot = A (x - x_nom(t)) + B (u - u_nom(t)) + Hv with lerped nominal trajectories
  const double* u_nom;
  double* xs;
  double* us;
  int* status;
  const RolloutStep* steps;
  const double* alphas;
  const double* jump;  // [batch][jump_capacity] jump records or nullptr
  int jump_capacity;
  int nsteps, first_idx, out_nodes, N, batch, begin, count;
  double first_alpha;
};

template <int NX, int NU>
__global__ void __launch_bounds__(128) rollout_cont_rpl_kernel(const RoArgs a) {
  using S = RoShape<NX, NU>;
  using R = typename S::R;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)warp * S::warp_doubles;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(wbase + S::P * S::slot);
  const int graw = lane / NX;
  const bool in_group = graw < S::P;
  const int gi = in_group ? graw : S::P - 1;
  const int i = in_group ? lane - graw * NX : NX - 1;
  const int gbase = gi * NX;
  double* sm = wbase + gi * S::slot;
  const int N = a.N, nodes = N + 1;
  const uint32_t dynBytes = S::dyn * sizeof(double), polBytes = S::pol * sizeof(double);
  const int ia = blockIdx.y;
  const double alpha = a.alphas[ia];

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
  const int warps_total = gridDim.x * (blockDim.x >> 5);

  for (int base = (blockIdx.x * (blockDim.x >> 5) + warp) * S::P; base < a.count; base += warps_total * S::P) {
    const int nprob = (a.count - base) < S::P ? (a.count - base) : S::P;
    const bool valid = in_group && gi < nprob;
    const int prob = a.begin + base + (gi < nprob ? gi : nprob - 1);
    double* xo = a.xs + ((size_t)ia * a.batch + prob) * (size_t)a.out_nodes * NX;
    double* uo = a.us + ((size_t)ia * a.batch + prob) * (size_t)a.out_nodes * NU;

    auto issue_node = [&](int k) {  // {A|B|Hv} and {K|dbias|bias} of node k of every carried problem -> ring slot k % 3
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, (dynBytes + polBytes) * nprob);
        for (int g = 0; g < nprob; ++g) {
          const size_t pb = (size_t)(a.begin + base + g);
          tma_load(wbase + g * S::slot + S::sDyn + (k % 3) * S::dyn, a.lq + (pb * nodes + k) * R::rec, dynBytes, bar);
          tma_load(wbase + g * S::slot + S::sPol + (k % 3) * S::pol, a.sol + (pb * nodes + k) * R::orec, polBytes, bar);
        }
      }
    };
    int loaded_hi = -1, issued_hi = -1;
    // make nodes <= q resident; called with q = idx + 1 of the time segment about to be used (idx is non-decreasing along the
    // schedule, so when node q has landed node q - 2 is dead and its slot takes the prefetch of node q + 1)
    auto ensure = [&](int q) {
      q = q < N ? q : N;
      while (loaded_hi < q) {
        if (issued_hi == loaded_hi) {
          issue_node(loaded_hi + 1);
          issued_hi = loaded_hi + 1;
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        loaded_hi += 1;
        if (loaded_hi >= q && loaded_hi + 1 <= N) {
          __syncwarp();
          issue_node(loaded_hi + 1);
          issued_hi = loaded_hi + 1;
        }
      }
    };
    // The rows this lane works with — row i of A, B, Hv and (lanes < nu) row l of K, bias, dbias — of the two nodes of the current time
    // segment live in REGISTERS: the segment index never decreases along the schedule, so each node's rows cross the shared-memory pipe
    // once (when the segment advances, lo <- hi and hi <- node idx + 1) instead of once per RK4 stage; the lerp runs on registers.
    const int lrow = i < NU ? i : NU - 1;
    double alo[NX], ahi[NX], blo[NU], bhi[NU], klo[NX], khi[NX];
    double hvlo = 0.0, hvhi = 0.0, fflo = 0.0, ffhi = 0.0;  // ff = bias + alpha dbias
    int cur = -2;                                            // node held in the lo registers
    auto load_rows = [&](int q, double (&ar)[NX], double (&br)[NU], double (&kr)[NX], double& hv, double& ff) {
      const double* d = sm + S::sDyn + (q % 3) * S::dyn;
      const double* p = sm + S::sPol + (q % 3) * S::pol;
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        ar[j] = d[R::oA + i + NX * j];
        kr[j] = p[R::oK + lrow + NU * j];
      }
#pragma unroll
      for (int l = 0; l < NU; ++l) br[l] = d[R::oB + i + NX * l];
      hv = d[R::oHv + i];
      ff = p[R::obias + lrow] + alpha * p[R::odb + lrow];
    };
    auto seek = [&](int idx) {  // warp-uniform: idx comes from the schedule shared by the batch
      if (idx == cur) return;
      ensure(idx + 1);
      const int hi = idx + 1 < N ? idx + 1 : N;
      if (idx == cur + 1) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          alo[j] = ahi[j];
          klo[j] = khi[j];
        }
#pragma unroll
        for (int l = 0; l < NU; ++l) blo[l] = bhi[l];
        hvlo = hvhi;
        fflo = ffhi;
      } else {
        load_rows(idx, alo, blo, klo, hvlo, fflo);
      }
      load_rows(hi, ahi, bhi, khi, hvhi, ffhi);
      cur = idx;
    };
    // u(t, x) on lanes < nu from the x in shared memory; result broadcast through shared memory
    auto policy = [&](int idx, double w0) {
      seek(idx);
      const double w1 = 1.0 - w0;
      double u = w0 * fflo + w1 * ffhi;
#pragma unroll
      for (int j = 0; j < NX; ++j) u = fma(fma(w0, klo[j], w1 * khi[j]), sm[S::sX + j], u);
      if (i < NU) sm[S::sU + i] = u;
      __syncwarp();
    };
    // dx_i/dt at (idx, w0) for the x in shared memory
    auto flow = [&](int idx, double w0) -> double {
      policy(idx, w0);
      const double w1 = 1.0 - w0;
      const double* xv = sm + S::sX;
      if (a.x_nom) {  // deviations from the lerped nominal trajectories (TimeTriggeredRollout on the LQ model of an SLQ iteration)
        const double* xn = a.x_nom + ((size_t)prob * (N + 1) + idx) * NX;
        const double* un = a.u_nom + ((size_t)prob * (N + 1) + idx) * NU;
        sm[S::sDx + i] = sm[S::sX + i] - fma(w0, __ldg(xn + i), w1 * __ldg(xn + NX + i));
        __syncwarp();
        if (i < NU) sm[S::sU + i] -= fma(w0, __ldg(un + i), w1 * __ldg(un + NU + i));
        __syncwarp();
        xv = sm + S::sDx;
      }
      double acc = fma(w0, hvlo, w1 * hvhi);
#pragma unroll
      for (int j = 0; j < NX; ++j) acc = fma(fma(w0, alo[j], w1 * ahi[j]), xv[j], acc);
#pragma unroll
      for (int l = 0; l < NU; ++l) acc = fma(fma(w0, blo[l], w1 * bhi[l]), sm[S::sU + l], acc);
      __syncwarp();  // x, u in shared memory are dead
      return acc;
    };

    double x = a.x0[(size_t)prob * NX + i];
    bool finite = true;
    auto observe = [&](int o, int idx, double w0) {
      sm[S::sX + i] = x;
      __syncwarp();
      policy(idx, w0);
      if (valid) {
        __stcg(xo + (size_t)o * NX + i, x);
        if (i < NU) __stcg(uo + (size_t)o * NU + i, sm[S::sU + i]);
      }
      finite = finite && finite_bits(x);
      __syncwarp();
    };
    observe(0, a.first_idx, a.first_alpha);
#pragma unroll 1
    for (int sidx = 0; sidx < a.nsteps; ++sidx) {
      const RolloutStep sp = a.steps[sidx];
      const double h = sp.h;
      if (sp.jump > 0) {  // an event (TimeTriggeredRollout.cpp:104-108): x+ = x_nom(post) + A_e (x - x_nom(pre)) + Hv_e
        const double* jr = a.jump + ((size_t)prob * a.jump_capacity + (sp.jump - 1)) * jump_rec(NX);
        sm[S::sX + i] = x - (a.x_nom ? __ldg(a.x_nom + ((size_t)prob * (N + 1) + sp.pre_node) * NX + i) : 0.0);
        __syncwarp();
        double xn = __ldg(jr + jump_oHv(NX) + i) + (a.x_nom ? __ldg(a.x_nom + ((size_t)prob * (N + 1) + sp.pre_node + 1) * NX + i) : 0.0);
#pragma unroll
        for (int kk = 0; kk < NX; ++kk) xn = fma(__ldg(jr + i + NX * kk), sm[S::sX + kk], xn);
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
        sm[S::sX + i] = xs;
        __syncwarp();
        const double kx = flow((stg == 0 ? sp.idx[0] : (stg == 1 ? sp.idx[1] : (stg == 2 ? sp.idx[2] : sp.idx[3]))), (stg == 0 ? sp.alpha[0] : (stg == 1 ? sp.alpha[1] : (stg == 2 ? sp.alpha[2] : sp.alpha[3]))));  // (constant indices: the step stays in registers)
        acc = fma(h * ((stg == 0 || stg == 3) ? (1.0 / 6.0) : (1.0 / 3.0)), kx, acc);
        xs = fma(h * ((stg == 2) ? 1.0 : 0.5), kx, x);
      }
      x = acc;
      observe(sidx + 1, sp.obs_idx, sp.obs_alpha);
    }
    // drain a prefetch that is still in flight before the ring is reused by the next round
    if (issued_hi > loaded_hi) {
      mbar_wait(bar, parity);
      parity ^= 1u;
    }
    const unsigned gmask = (NX == 32) ? kFull : (((1u << NX) - 1u) << gbase);
    if ((__ballot_sync(kFull, !finite) & gmask) && valid && i == 0) atomicOr(a.status + prob, O2C_STATUS_NONFINITE);
    __syncwarp();
  }
}

template <int NX, int NU>
cudaError_t launch_ro(const RoArgs& a, int n_alpha, cudaStream_t stream) {
  using S = RoShape<NX, NU>;
  constexpr int wpb = 4;
  const size_t smem = (size_t)S::warp_doubles * wpb * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(rollout_cont_rpl_kernel<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int num_sms = device_sm_count();
  if (num_sms <= 0) return cudaErrorInvalidDevice;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, rollout_cont_rpl_kernel<NX, NU>, wpb * 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int per_cta = wpb * S::P;
  const int needed = (a.count + per_cta - 1) / per_cta;
  const int cap = num_sms * ctas_per_sm;
  dim3 grid(needed < cap ? needed : cap, n_alpha);
  rollout_cont_rpl_kernel<NX, NU><<<grid, wpb * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int NX, int NU, int NC>
bool layout_matches(const Layout& L) {
  using S = Shape<NX, NU, NC>;
  return L.n == NX && L.m == NU && L.ncmax == NC && (NC == 0 || (L.oC == S::oC && L.oD == S::oD && L.oe == S::oe)) && L.rec == S::rec && L.oA == S::oA && L.oB == S::oB && L.oHv == S::oHv && L.oq == S::oq &&
         L.or_ == S::or_ && L.oc == S::oc && L.oQ == S::oQ && L.oP == S::oP && L.oR == S::oR && L.orec == S::orec && L.oK == S::oK &&
         L.odb == S::odb && L.obias == S::obias && L.oSm == S::oSm && L.oSv == S::oSv && L.os == S::os && L.trec == S::trec &&
         L.oQf == S::oQf && L.oqf == S::oqf && L.ocf == S::ocf;
}

template <int NX, int NU, int NC, bool NOM, bool EV>
cudaError_t launch(const Args& a, cudaStream_t stream) {
  using S = Shape<NX, NU, NC>;
  constexpr int wpb = (NX == 10) ? 1 : 2;
  const size_t smem = (size_t)S::warp_doubles * wpb * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(ilqr_rpl_kernel<NX, NU, NC, NOM, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int num_sms = device_sm_count();
  if (num_sms <= 0) return cudaErrorInvalidDevice;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, ilqr_rpl_kernel<NX, NU, NC, NOM, EV>, wpb * 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int per_cta = wpb * S::P;
  const int needed = (a.count + per_cta - 1) / per_cta;
  const int cap = num_sms * ctas_per_sm;  // persistent warps: one resident wave, static stride over the problem index
  ilqr_rpl_kernel<NX, NU, NC, NOM, EV><<<needed < cap ? needed : cap, wpb * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

bool settings_match(const SolverSettings& st, const DeviceBuffers& buf, const Layout& L) {
  // reduced and full Riccati form alike: under LINE_SEARCH they are the same map (RiccatiTest.cpp:87-105 holds them equal to 1e-9)
  return st.algorithm == O2C_ALG_ILQR && st.strategy == O2C_STRATEGY_LINE_SEARCH && st.hc == O2C_HC_DIAGONAL_SHIFT &&
         (buf.x_nom == nullptr) == (buf.u_nom == nullptr) && L.N >= 1;
}

}  // namespace

bool rpl_slq_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  if (!(st.algorithm == O2C_ALG_SLQ && st.strategy == O2C_STRATEGY_LINE_SEARCH && st.hc == O2C_HC_DIAGONAL_SHIFT &&
        (buf.x_nom == nullptr) == (buf.u_nom == nullptr) && L.N >= 1 && L.nodes == L.N + 1))
    return false;
  if (L.ncmax != 0) return false;
  return layout_matches<12, 4, 0>(L) || layout_matches<10, 3, 0>(L) || layout_matches<4, 1, 0>(L);
}

cudaError_t launch_slq_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin,
                           int count, cudaStream_t stream) {
  if (!rpl_slq_supported(L, st, buf)) return cudaErrorNotSupported;
  SlqArgs a{};
  a.lq = buf.lq;
  a.term = buf.term;
  a.x_nom = buf.x_nom;
  a.u_nom = buf.u_nom;
  a.sol = buf.sol;
  a.status = buf.status;
  a.steps = steps;
  a.jump = buf.jump;
  a.jump_capacity = buf.jump_capacity;
  a.nsteps = nsteps;
  a.N = L.N;
  a.begin = begin;
  a.count = count;
  a.eps = st.eps;
  const bool ev = buf.event != nullptr;
  if (layout_matches<12, 4, 0>(L)) return ev ? launch_slq<12, 4, true>(a, stream) : launch_slq<12, 4, false>(a, stream);
  if (layout_matches<10, 3, 0>(L)) return ev ? launch_slq<10, 3, true>(a, stream) : launch_slq<10, 3, false>(a, stream);
  return ev ? launch_slq<4, 1, true>(a, stream) : launch_slq<4, 1, false>(a, stream);
}

bool rpl_rollout_cont_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  return st.algorithm == O2C_ALG_SLQ && (buf.x_nom == nullptr) == (buf.u_nom == nullptr) && L.N >= 1 &&
         L.nodes == L.N + 1 && L.ncmax == 0 && (layout_matches<12, 4, 0>(L) || layout_matches<10, 3, 0>(L) || layout_matches<4, 1, 0>(L));
}

cudaError_t launch_rollout_cont_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const RolloutStep* steps, int nsteps,
                                    int first_idx, double first_alpha, int out_nodes, const double* alphas_dev, int n_alpha, int batch,
                                    int begin, int count, cudaStream_t stream) {
  if (!rpl_rollout_cont_supported(L, st, buf)) return cudaErrorNotSupported;
  RoArgs a{};
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
  if (layout_matches<12, 4, 0>(L)) return launch_ro<12, 4>(a, n_alpha, stream);
  if (layout_matches<10, 3, 0>(L)) return launch_ro<10, 3>(a, n_alpha, stream);
  return launch_ro<4, 1>(a, n_alpha, stream);
}

bool rpl_ilqr_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  if (!settings_match(st, buf, L)) return false;
  return layout_matches<10, 3, 0>(L) || layout_matches<4, 1, 0>(L) || layout_matches<9, 9, 3>(L) || layout_matches<12, 4, 0>(L);
}

cudaError_t launch_ilqr_rpl(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, bool with_rollout, double alpha, int begin,
                            int count, cudaStream_t stream, int* launches) {
  if (!rpl_ilqr_supported(L, st, buf)) return cudaErrorNotSupported;
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
  a.N = L.N;
  a.begin = begin;
  a.count = count;
  a.with_rollout = with_rollout ? 1 : 0;
  a.eps = st.eps;
  a.alpha = alpha;
  if (launches) *launches = 1;
  const bool nom = buf.x_nom != nullptr, ev = buf.event != nullptr;
#define O2C_RPL_DISPATCH(NX, NU, NC)                                                                                  \
  return ev ? (nom ? launch<NX, NU, NC, true, true>(a, stream) : launch<NX, NU, NC, false, true>(a, stream))         \
            : (nom ? launch<NX, NU, NC, true, false>(a, stream) : launch<NX, NU, NC, false, false>(a, stream))
  if (layout_matches<10, 3, 0>(L)) O2C_RPL_DISPATCH(10, 3, 0);
  if (layout_matches<9, 9, 3>(L)) O2C_RPL_DISPATCH(9, 9, 3);
  if (layout_matches<12, 4, 0>(L)) O2C_RPL_DISPATCH(12, 4, 0);
  O2C_RPL_DISPATCH(4, 1, 0);
#undef O2C_RPL_DISPATCH
}

}  // namespace o2c
