// riccati_dmma.cu — shape-specialised ILQR sweep + fused LQ rollout for nx = nu = 24 (the legged-robot shape), FP64, sm_100a.
//
// One CTA (3 GEMM warps + 1 "vector" warp) owns one problem's time-sequential sweep:
//   * the per-node operand block {A | B | Hv | q | r | c} (9.8 KB) is staged into shared memory by one TMA bulk copy per node
//     (cp.async.bulk + mbarrier), double buffered, a full stage ahead of its use; the cost Hessians Q, P, R are only ever
//     accumulator initial values, so they are read straight from L2 into DMMA accumulator fragments (the producer
//     L2-prefetches them together with the bulk copy);
//   * every 24x24x24 contraction is written in the "TN" form D = X'Y (both operands column-major, the contracted index
//     contiguous), which makes every operand fragment one conflict-free 16-byte LDS and every accumulator fragment one
//     conflict-free 16-byte STS with the dense ld = 24 the TMA copy produces; the products run on the FP64 tensor pipe
//     (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no FP64 kind);
//   * the vector warp does the Cholesky of Hm = R + B'SB, the triangular inverse, all matrix-vector terms and the TMA issue.
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
// Rollout (fused, vector warp): du_k = K_k dx_k + alpha dbias_k ; dx_{k+1} = A_k dx_k + B_k du_k + Hv_k
//                                            DDP_HelperFunctions.cpp:125-138, 296-304; LinearController.cpp:79-87
#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr int kN = 24;            // nx == nu
constexpr int kMat = kN * kN;     // 576
constexpr int kThreads = 128;
constexpr int kGemmWarps = 3;
constexpr int kOperand = 2 * kMat + 3 * kN + 2;  // {A|B|Hv|q|r|c,pad} doubles staged by TMA (= Layout::oQ for n = m = 24)
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy, L2 prefetch, DMMA
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe. Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4], d = D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}
// one 8-deep k block of the TN product D(tile i,j) += X(kblock, i)' Y(kblock, j): the 16-byte fragment holds rows {2c, 2c+1} of the
// k block, so the two DMMAs contract k = {0,2,4,6} and {1,3,5,7} (the same permutation on both operands).
__device__ __forceinline__ void dmma2(double2& d, const double2& y, const double2& x) {
  dmma(d, y.x, x.x);
  dmma(d, y.y, x.y);
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, const double2& v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stg2(double* p, const double2& v) { __stcg(reinterpret_cast<double2*>(p), v); }
__device__ __forceinline__ double2 add2(const double2& a, const double2& b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 sub2(const double2& a, const double2& b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ bool finite2(const double2& v) { return isfinite(v.x) && isfinite(v.y); }

// tile offset inside a column-major 24x24 matrix: rows 8*ib.., columns 8*jb..
__device__ __forceinline__ constexpr int tile(int ib, int jb) { return 8 * ib + 8 * kN * jb; }

// the six lower tiles (ib >= jb) of a symmetric 24x24 result, two per GEMM warp
template <int W>
struct LowerTiles;
template <>
struct LowerTiles<0> {
  static constexpr int i0 = 0, j0 = 0, i1 = 1, j1 = 0;
};
template <>
struct LowerTiles<1> {
  static constexpr int i0 = 2, j0 = 0, i1 = 1, j1 = 1;
};
template <>
struct LowerTiles<2> {
  static constexpr int i0 = 2, j0 = 1, i1 = 2, j1 = 2;
};

struct __align__(16) Smem {
  double in[2][kOperand];  // TMA destination: {A | B | Hv | q | r | c}
  double S[kMat];          // value function of node k+1 (symmetric, both triangles)
  double SA[kMat];         // S A;   later Li  = L^-1
  double SB[kMat];         // S B;   later LiT = L^-T
  double H[kMat];          // Hm (lower tiles) -> L (strictly lower) with 1/L_jj on the diagonal
  double G[kMat];          // G = P + B'SA -> Y = L^-1 G
  double Sv[kN], w[kN], Gv[kN], tv[kN], Yv[kN], xb[kN], ub[kN];
  unsigned long long full[4];
  int flags;
};

struct Args {
  const double* lq;
  const double* term;
  const double* x0;
  double* sol;
  double* xs;
  double* us;
  int* status;
  int rec, orec, N;
  int oQ, oP, oR, oHv, oq, or_, oc;
  int oK, odb, obias, oSm, oSv, os;
  int oQf, oqf, ocf, trec;
  int begin, batch, with_rollout;
  double eps, alpha;
};

// ---------------------------------------------------------------------------------------------------------------------
// GEMM-warp phases (w = warp index 0..2, lo = lane offset 2*(lane%4) + 24*(lane/4) of the fragment inside a tile)
// ---------------------------------------------------------------------------------------------------------------------
// [SA | SB](:, block w) = S' [A | B](:, block w)
__device__ __forceinline__ void phase1_gemm(const Smem& sm, const double* A, const double* B, double* SA, double* SB, int w, int lo) {
  double2 aA[3] = {zero2(), zero2(), zero2()}, aB[3] = {zero2(), zero2(), zero2()};
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) {
    const double2 ya = ld2(A + lo + 8 * kb + 8 * kN * w), yb = ld2(B + lo + 8 * kb + 8 * kN * w);
#pragma unroll
    for (int ib = 0; ib < 3; ++ib) {
      const double2 x = ld2(sm.S + lo + tile(kb, ib));
      dmma2(aA[ib], ya, x);
      dmma2(aB[ib], yb, x);
    }
  }
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) {
    st2(SA + lo + 8 * ib + 8 * kN * w, aA[ib]);
    st2(SB + lo + 8 * ib + 8 * kN * w, aB[ib]);
  }
}

// G(:, block W) = P + B'(SA)(:, block W);  two lower tiles of Hm = R + B'(SB)
template <int W>
__device__ __forceinline__ void phase2_gemm(Smem& sm, const double* B, const double* Pg, const double* Rg, int lo) {
  using T = LowerTiles<W>;
  double2 pG[3], pH[2];
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) pG[ib] = ldg2(Pg + lo + tile(ib, W));
  pH[0] = ldg2(Rg + lo + tile(T::i0, T::j0));
  pH[1] = ldg2(Rg + lo + tile(T::i1, T::j1));
  double2 aG[3] = {zero2(), zero2(), zero2()}, aH[2] = {zero2(), zero2()};
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) {
    double2 x[3];
#pragma unroll
    for (int ib = 0; ib < 3; ++ib) x[ib] = ld2(B + lo + tile(kb, ib));
    const double2 ysa = ld2(sm.SA + lo + tile(kb, W));
    const double2 y0 = ld2(sm.SB + lo + tile(kb, T::j0));
    const double2 y1 = (T::j1 == T::j0) ? y0 : ld2(sm.SB + lo + tile(kb, T::j1));
#pragma unroll
    for (int ib = 0; ib < 3; ++ib) dmma2(aG[ib], ysa, x[ib]);
    dmma2(aH[0], y0, x[T::i0]);
    dmma2(aH[1], y1, x[T::i1]);
  }
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) st2(sm.G + lo + tile(ib, W), add2(aG[ib], pG[ib]));
  st2(sm.H + lo + tile(T::i0, T::j0), add2(aH[0], pH[0]));
  st2(sm.H + lo + tile(T::i1, T::j1), add2(aH[1], pH[1]));
}

// two lower tiles of T = Q + eps I + A'(SA), kept in registers until phase 5
template <int W>
__device__ __forceinline__ void phase3_gemm(const Smem& sm, const double* A, const double* Qg, double eps, int lo, int r, int c, double2 (&aT)[2]) {
  using T = LowerTiles<W>;
  const double2 q0 = ldg2(Qg + lo + tile(T::i0, T::j0)), q1 = ldg2(Qg + lo + tile(T::i1, T::j1));
  aT[0] = zero2();
  aT[1] = zero2();
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) {
    const double2 x0 = ld2(A + lo + tile(kb, T::i0));
    const double2 x1 = (T::i1 == T::i0) ? x0 : ld2(A + lo + tile(kb, T::i1));
    const double2 y0 = ld2(sm.SA + lo + tile(kb, T::j0));
    const double2 y1 = (T::j1 == T::j0) ? y0 : ld2(sm.SA + lo + tile(kb, T::j1));
    dmma2(aT[0], y0, x0);
    dmma2(aT[1], y1, x1);
  }
  aT[0] = add2(aT[0], q0);
  aT[1] = add2(aT[1], q1);
  const double ex = (2 * c == r) ? eps : 0.0, ey = (2 * c + 1 == r) ? eps : 0.0;
  if (T::i0 == T::j0) {
    aT[0].x += ex;
    aT[0].y += ey;
  }
  if (T::i1 == T::j1) {
    aT[1].x += ex;
    aT[1].y += ey;
  }
}

// Y(:, block W) = Li G(:, block W) in place (X = LiT, upper block triangular), and two lower tiles of Li = (LiT)' via DMMA with identity
template <int W>
__device__ __forceinline__ void phase4_gemm(Smem& sm, int lo, int r, int c) {
  using T = LowerTiles<W>;
  double2 g[3];
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) g[kb] = ld2(sm.G + lo + tile(kb, W));
  double2 aY[3] = {zero2(), zero2(), zero2()};
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) {
#pragma unroll
    for (int kb = 0; kb <= ib; ++kb) dmma2(aY[ib], g[kb], ld2(sm.SB + lo + tile(kb, ib)));
  }
  // Li(tile i,j) = sum_k LiT[k][i] I[k][j]: only the k block j contributes
  const double2 id = make_double2((2 * c == r) ? 1.0 : 0.0, (2 * c + 1 == r) ? 1.0 : 0.0);
  double2 t0 = zero2(), t1 = zero2();
  dmma2(t0, id, ld2(sm.SB + lo + tile(T::j0, T::i0)));
  dmma2(t1, id, ld2(sm.SB + lo + tile(T::j1, T::i1)));
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) st2(sm.G + lo + tile(ib, W), aY[ib]);
  st2(sm.SA + lo + tile(T::i0, T::j0), t0);
  st2(sm.SA + lo + tile(T::i1, T::j1), t1);
}

// S = T - Y'Y (two lower tiles, mirrored into both triangles; shared + global) and K(:, block W) = -Li' Y(:, block W) (global)
template <int W>
__device__ __forceinline__ bool phase5_gemm(Smem& sm, const double2 (&aT)[2], double* Smg, double* Kg, double* Kg2, int lo, int r, int c) {
  using T = LowerTiles<W>;
  double2 aS[2] = {zero2(), zero2()}, aK[3] = {zero2(), zero2(), zero2()};
  double2 yk[3];
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) {
    yk[kb] = ld2(sm.G + lo + tile(kb, W));
    const double2 x0 = ld2(sm.G + lo + tile(kb, T::i0));
    const double2 x1 = (T::i1 == T::i0) ? x0 : ld2(sm.G + lo + tile(kb, T::i1));
    const double2 y0 = ld2(sm.G + lo + tile(kb, T::j0));
    const double2 y1 = (T::j1 == T::j0) ? y0 : ld2(sm.G + lo + tile(kb, T::j1));
    dmma2(aS[0], y0, x0);
    dmma2(aS[1], y1, x1);
  }
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) {
#pragma unroll
    for (int kb = ib; kb < 3; ++kb) dmma2(aK[ib], yk[kb], ld2(sm.SA + lo + tile(kb, ib)));
  }
  const double2 s0 = sub2(aT[0], aS[0]), s1 = sub2(aT[1], aS[1]);
  st2(sm.S + lo + tile(T::i0, T::j0), s0);
  st2(sm.S + lo + tile(T::i1, T::j1), s1);
  stg2(Smg + lo + tile(T::i0, T::j0), s0);
  stg2(Smg + lo + tile(T::i1, T::j1), s1);
  const int mo = r + kN * 2 * c;  // transposed element (j, i) of the fragment's first value; the second is one column further
  if (T::i0 != T::j0) {
    double* p = sm.S + mo + tile(T::j0, T::i0);
    p[0] = s0.x;
    p[kN] = s0.y;
    double* gq = Smg + mo + tile(T::j0, T::i0);
    __stcg(gq, s0.x);
    __stcg(gq + kN, s0.y);
  }
  if (T::i1 != T::j1) {
    double* p = sm.S + mo + tile(T::j1, T::i1);
    p[0] = s1.x;
    p[kN] = s1.y;
    double* gq = Smg + mo + tile(T::j1, T::i1);
    __stcg(gq, s1.x);
    __stcg(gq + kN, s1.y);
  }
#pragma unroll
  for (int ib = 0; ib < 3; ++ib) {
    const double2 kv = make_double2(-aK[ib].x, -aK[ib].y);
    stg2(Kg + lo + tile(ib, W), kv);
    if (Kg2) stg2(Kg2 + lo + tile(ib, W), kv);
  }
  return finite2(s0) && finite2(s1);
}

// ---------------------------------------------------------------------------------------------------------------------
// vector-warp pieces
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  return v;
}
__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// z[i] = sum_k M[i + 24 k] v[k] for lane i < 24 (M column-major in shared memory, v in shared memory, broadcast reads)
__device__ __forceinline__ double matvec_rows(const double* M, const double* v, int li) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < kN; k += 4) {
    const double2 v01 = ld2(v + k), v23 = ld2(v + k + 2);
    a0 = fma(M[li + kN * k], v01.x, a0);
    a1 = fma(M[li + kN * (k + 1)], v01.y, a1);
    a2 = fma(M[li + kN * (k + 2)], v23.x, a2);
    a3 = fma(M[li + kN * (k + 3)], v23.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// z[j] = sum_k M[k + 24 j] v[k] (transposed product) with the fragment access pattern: every lane of quad r ends up with z[8 jb + r]
__device__ __forceinline__ void matvec_cols(const double* M, const double* v, int lo, int c, double (&z)[3]) {
  double2 vf[3];
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) vf[kb] = ld2(v + 8 * kb + 2 * c);
#pragma unroll
  for (int jb = 0; jb < 3; ++jb) {
    double p = 0.0;
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) {
      const double2 mv = ld2(M + lo + tile(kb, jb));
      p = fma(mv.x, vf[kb].x, p);
      p = fma(mv.y, vf[kb].y, p);
    }
    z[jb] = quad_sum(p);
  }
}
__device__ __forceinline__ double pick3(const double (&z)[3], int c) { return c == 0 ? z[0] : (c == 1 ? z[1] : z[2]); }

// Cholesky Hm = L L' (rows in registers, one row per lane, scaled column broadcast through shared memory), then rows of L^-T by the
// same right-looking recurrence applied to the identity; lane 24 carries Gv through it and ends with Yv = L^-1 Gv.
__device__ __forceinline__ bool cholesky_inverse(Smem& sm, int lane) {
  const int li = lane < kN ? lane : kN - 1;
  double* Hs = sm.H;
  double h[kN];
#pragma unroll
  for (int k = 0; k < kN; ++k) h[k] = Hs[li + kN * k];
  bool pd = true;
#pragma unroll
  for (int j = 0; j < kN; ++j) {
    const double d = __shfl_sync(kFull, h[j], j);
    const bool ok = d > 0.0;
    pd = pd && ok;
    const double rs = ok ? rsqrt(d) : __longlong_as_double(0x7ff8000000000000LL);  // NaNs propagate like the reference's LLT
    const double l = h[j] * rs;
    if (lane < kN) Hs[lane + kN * j] = (lane == j) ? rs : l;
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < kN; ++k) h[k] = fma(-l, Hs[k + kN * j], h[k]);
  }
  // rows of L^-T: g <- e_lane' L^-T (lanes < 24); lane 24: g <- Gv' L^-T = Yv'
#pragma unroll
  for (int k = 0; k < kN; ++k) h[k] = (lane == kN) ? sm.Gv[k] : ((k == lane) ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < kN; ++j) {
    const double a = h[j] * Hs[j + kN * j];
    h[j] = a;
#pragma unroll
    for (int k = j + 1; k < kN; ++k) h[k] = fma(-a, Hs[k + kN * j], h[k]);
  }
  if (lane < kN) {
#pragma unroll
    for (int k = 0; k < kN; ++k) sm.SB[lane + kN * k] = h[k];  // LiT, column-major: exact zeros below the diagonal
  } else if (lane == kN) {
#pragma unroll
    for (int k = 0; k < kN; ++k) sm.Yv[k] = h[k];
  }
  return pd;
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 4) ilqr_dmma_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  const int lo = 2 * c + kN * r;
  const int li = lane < kN ? lane : kN - 1;
  const int prob = a.begin + blockIdx.x;
  const int N = a.N;
  const double* lqp = a.lq + (size_t)prob * N * a.rec;
  const double* term = a.term + (size_t)prob * a.trec;
  double* solp = a.sol + (size_t)prob * (N + 1) * a.orec;
  const uint32_t opBytes = kOperand * sizeof(double);
  const uint32_t hessBytes = (uint32_t)(a.rec - a.oQ) * sizeof(double);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&sm.full[i], 1);
    sm.flags = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // producer prologue: nodes N-1 and N-2
  if (warp == 3 && lane == 0) {
    for (int k = N - 1; k >= 0 && k >= N - 2; --k) {
      mbar_expect_tx(&sm.full[k & 1], opBytes);
      tma_load(sm.in[k & 1], lqp + (size_t)k * a.rec, opBytes, &sm.full[k & 1]);
      l2_prefetch(lqp + (size_t)k * a.rec + a.oQ, hessBytes);
    }
  }
  // terminal condition: valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526)
  {
    double* outN = solp + (size_t)N * a.orec;
    for (int i = threadIdx.x; i < kMat; i += kThreads) {
      const double v = term[a.oQf + i];
      sm.S[i] = v;
      outN[a.oSm + i] = v;
    }
    if (threadIdx.x < kN) {
      const double v = term[a.oqf + threadIdx.x];
      sm.Sv[threadIdx.x] = v;
      outN[a.oSv + threadIdx.x] = v;
    }
    if (threadIdx.x == 0) outN[a.os] = term[a.ocf];
  }
  double sval = term[a.ocf];  // s of node k+1 (tracked by the vector warp)
  bool finite = true, pd = true;
  __syncthreads();

  double2 aT[2];
  uint32_t phbits = 0;  // bit q = parity of the next phase to wait for on full[q]
  for (int k = N - 1; k >= 0; --k) {
    const int b = k & 1;
    const double* in = sm.in[b];
    const double* A = in;
    const double* B = in + kMat;
    const double* Hv = in + 2 * kMat;
    const double* qv = Hv + kN;
    const double* rv = qv + kN;
    const double* rec = lqp + (size_t)k * a.rec;
    double* out = solp + (size_t)k * a.orec;
    double* out2 = (k == N - 1) ? solp + (size_t)N * a.orec : nullptr;  // node N := node N-1 (GaussNewtonDDP.cpp:609-618)
    mbar_wait(&sm.full[b], (phbits >> b) & 1u);
    phbits ^= 1u << b;
    double spart = 0.0;

    // ---- phase 1: SA, SB | w = Sv + S Hv ----
    if (warp < kGemmWarps) {
      phase1_gemm(sm, A, B, sm.SA, sm.SB, warp, lo);
    } else {
      const double shv = matvec_rows(sm.S, Hv, li);
      const double wv = sm.Sv[li] + shv;
      if (lane < kN) {
        sm.w[lane] = wv;
        spart = Hv[lane] * (wv - 0.5 * shv);
      }
    }
    __syncthreads();
    // ---- phase 2: G, Hm | Gv = r + B'w, tv = q + A'w ----
    if (warp == 0) {
      phase2_gemm<0>(sm, B, rec + a.oP, rec + a.oR, lo);
    } else if (warp == 1) {
      phase2_gemm<1>(sm, B, rec + a.oP, rec + a.oR, lo);
    } else if (warp == 2) {
      phase2_gemm<2>(sm, B, rec + a.oP, rec + a.oR, lo);
    } else {
      double zB[3], zA[3];
      matvec_cols(B, sm.w, lo, c, zB);
      matvec_cols(A, sm.w, lo, c, zA);
      if (c < 3) {
        sm.Gv[8 * c + r] = rv[8 * c + r] + pick3(zB, c);
        sm.tv[8 * c + r] = qv[8 * c + r] + pick3(zA, c);
      }
    }
    __syncthreads();
    // ---- phase 3: T = Q + eps I + A'(SA) | Cholesky, L^-T, Yv ----
    if (warp == 0) {
      phase3_gemm<0>(sm, A, rec + a.oQ, a.eps, lo, r, c, aT);
    } else if (warp == 1) {
      phase3_gemm<1>(sm, A, rec + a.oQ, a.eps, lo, r, c, aT);
    } else if (warp == 2) {
      phase3_gemm<2>(sm, A, rec + a.oQ, a.eps, lo, r, c, aT);
    } else {
      pd = cholesky_inverse(sm, lane) && pd;
    }
    __syncthreads();
    // ---- phase 4: Y = Li G, Li | dbias = -L^-T Yv, s ----
    if (warp == 0) {
      phase4_gemm<0>(sm, lo, r, c);
    } else if (warp == 1) {
      phase4_gemm<1>(sm, lo, r, c);
    } else if (warp == 2) {
      phase4_gemm<2>(sm, lo, r, c);
    } else {
      const double kv = -matvec_rows(sm.SB, sm.Yv, li);
      if (lane < kN) {
        __stcg(out + a.odb + lane, kv);
        __stcg(out + a.obias + lane, 0.0);
        if (out2) {
          __stcg(out2 + a.odb + lane, kv);
          __stcg(out2 + a.obias + lane, 0.0);
        }
        const double yv = sm.Yv[lane];
        spart = fma(-0.5 * yv, yv, spart);
      }
      sval = sval + in[2 * kMat + 3 * kN] + warp_sum_all(spart);
      if (lane == 0) __stcg(out + a.os, sval);
    }
    __syncthreads();
    // ---- phase 5: S = T - Y'Y, K = -Li'Y | Sv = tv - Y'Yv ----
    if (warp == 0) {
      finite = phase5_gemm<0>(sm, aT, out + a.oSm, out + a.oK, out2 ? out2 + a.oK : nullptr, lo, r, c) && finite;
    } else if (warp == 1) {
      finite = phase5_gemm<1>(sm, aT, out + a.oSm, out + a.oK, out2 ? out2 + a.oK : nullptr, lo, r, c) && finite;
    } else if (warp == 2) {
      finite = phase5_gemm<2>(sm, aT, out + a.oSm, out + a.oK, out2 ? out2 + a.oK : nullptr, lo, r, c) && finite;
    } else {
      double zY[3];
      matvec_cols(sm.G, sm.Yv, lo, c, zY);
      if (c < 3) {
        const double v = sm.tv[8 * c + r] - pick3(zY, c);
        sm.Sv[8 * c + r] = v;
        __stcg(out + a.oSv + 8 * c + r, v);
        finite = finite && isfinite(v);
      }
      finite = finite && isfinite(sval);
    }
    __syncthreads();
    // buffer b is free: stage node k-2 into it (one full stage ahead of its use)
    if (warp == 3 && lane == 0 && k >= 2) {
      mbar_expect_tx(&sm.full[b], opBytes);
      tma_load(sm.in[b], lqp + (size_t)(k - 2) * a.rec, opBytes, &sm.full[b]);
      l2_prefetch(lqp + (size_t)(k - 2) * a.rec + a.oQ, hessBytes);
    }
  }

  // ---- status ----
  {
    int bits = 0;
    if (!__all_sync(kFull, pd)) bits |= O2C_STATUS_CHOL_NOT_PD;
    if (!__all_sync(kFull, finite)) bits |= O2C_STATUS_NONFINITE;
    if (lane == 0 && bits) atomicOr(&sm.flags, bits);
  }
  __syncthreads();
  if (warp != 3) return;
  if (!a.with_rollout) {
    if (lane == 0) a.status[prob] = sm.flags;
    return;
  }

  // ---- fused forward rollout of the LQ model (vector warp). Operand ring of 4 nodes over the (now free) sweep buffers. ----
  double* ring[4] = {sm.in[0], sm.in[1], sm.S, sm.S + kOperand};  // S..SB span 3*576 >= 2*1226 doubles
  // barriers 0,1 continue with the phase parities tracked during the sweep (every issued copy has been waited on); 2,3 are fresh
  if (lane == 0) {
    fence_proxy_async();  // the ring reuses buffers that were written through the generic proxy
    for (int k = 0; k < 4 && k < N; ++k) {
      mbar_expect_tx(&sm.full[k], opBytes);
      tma_load(ring[k], lqp + (size_t)k * a.rec, opBytes, &sm.full[k]);
    }
    for (int k = 4; k < 12 && k < N; ++k) l2_prefetch(lqp + (size_t)k * a.rec, opBytes);
    for (int k = 0; k < 8 && k < N; ++k) l2_prefetch(solp + (size_t)k * a.orec + a.oK, (kMat + kN) * sizeof(double));
  }
  __syncwarp();
  double* xo = a.xs + (size_t)prob * (N + 1) * kN;
  double* uo = a.us + (size_t)prob * (N + 1) * kN;
  double x = a.x0[(size_t)prob * kN + li];
  if (lane < kN) sm.xb[lane] = x;
  __syncwarp();
  bool xfinite = true;
  for (int k = 0; k < N; ++k) {
    const double* Kg = solp + (size_t)k * a.orec + a.oK;
    // u = alpha dbias + K x
    double u0 = a.alpha * __ldcg(solp + (size_t)k * a.orec + a.odb + li), u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
    for (int j = 0; j < kN; j += 4) {
      const double2 x01 = ld2(sm.xb + j), x23 = ld2(sm.xb + j + 2);
      u0 = fma(__ldcg(Kg + li + kN * j), x01.x, u0);
      u1 = fma(__ldcg(Kg + li + kN * (j + 1)), x01.y, u1);
      u2 = fma(__ldcg(Kg + li + kN * (j + 2)), x23.x, u2);
      u3 = fma(__ldcg(Kg + li + kN * (j + 3)), x23.y, u3);
    }
    const double u = (u0 + u1) + (u2 + u3);
    if (lane < kN) {
      sm.ub[lane] = u;
      __stcg(xo + (size_t)k * kN + lane, x);
      __stcg(uo + (size_t)k * kN + lane, u);
    }
    xfinite = xfinite && isfinite(x);
    mbar_wait(&sm.full[k & 3], (phbits >> (k & 3)) & 1u);
    phbits ^= 1u << (k & 3);
    const double* in = ring[k & 3];
    __syncwarp();
    const double xn = in[2 * kMat + li] + matvec_rows(in, sm.xb, li) + matvec_rows(in + kMat, sm.ub, li);
    __syncwarp();
    x = xn;
    if (lane < kN) sm.xb[lane] = x;
    __syncwarp();
    if (lane == 0) {
      if (k + 4 < N) {
        mbar_expect_tx(&sm.full[k & 3], opBytes);
        tma_load(ring[k & 3], lqp + (size_t)(k + 4) * a.rec, opBytes, &sm.full[k & 3]);
      }
      if (k + 12 < N) l2_prefetch(lqp + (size_t)(k + 12) * a.rec, opBytes);
      if (k + 8 < N) l2_prefetch(solp + (size_t)(k + 8) * a.orec + a.oK, (kMat + kN) * sizeof(double));
    }
  }
  // node N: state, and the input of the copied last policy re-evaluated at x_N (TimeTriggeredRollout.cpp:98-102)
  {
    const double* Kg = solp + (size_t)N * a.orec + a.oK;
    double u0 = a.alpha * __ldcg(solp + (size_t)N * a.orec + a.odb + li), u1 = 0.0;
#pragma unroll
    for (int j = 0; j < kN; j += 2) {
      const double2 x01 = ld2(sm.xb + j);
      u0 = fma(__ldcg(Kg + li + kN * j), x01.x, u0);
      u1 = fma(__ldcg(Kg + li + kN * (j + 1)), x01.y, u1);
    }
    if (lane < kN) {
      __stcg(xo + (size_t)N * kN + lane, x);
      __stcg(uo + (size_t)N * kN + lane, u0 + u1);
    }
    xfinite = xfinite && isfinite(x);
  }
  const bool allfinite = __all_sync(kFull, xfinite);
  if (lane == 0) a.status[prob] = sm.flags | (allfinite ? 0 : O2C_STATUS_NONFINITE);
}

}  // namespace

bool fast_ilqr_supported(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf) {
  return L.n == kN && L.m == kN && L.ncmax == 0 && st.algorithm == O2C_ALG_ILQR && st.reduced && st.strategy == O2C_STRATEGY_LINE_SEARCH &&
         st.hc == O2C_HC_DIAGONAL_SHIFT && buf.x_nom == nullptr && buf.u_nom == nullptr && L.N >= 1 && L.oQ == kOperand;
}

cudaError_t launch_ilqr_fast(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, bool with_rollout, double alpha, int batch,
                             int begin, int count, cudaStream_t stream, int* launches) {
  if (!fast_ilqr_supported(L, st, buf)) return cudaErrorNotSupported;
  static bool configured = false;
  const size_t smem = sizeof(Smem);
  cudaError_t e = cudaFuncSetAttribute(ilqr_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (!configured) {
    cudaFuncSetAttribute(ilqr_dmma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  Args a{};
  a.lq = buf.lq;
  a.term = buf.term;
  a.x0 = buf.x0;
  a.sol = buf.sol;
  a.xs = buf.xs;
  a.us = buf.us;
  a.status = buf.status;
  a.rec = L.rec;
  a.orec = L.orec;
  a.N = L.N;
  a.oQ = L.oQ;
  a.oP = L.oP;
  a.oR = L.oR;
  a.oHv = L.oHv;
  a.oq = L.oq;
  a.or_ = L.or_;
  a.oc = L.oc;
  a.oK = L.oK;
  a.odb = L.odb;
  a.obias = L.obias;
  a.oSm = L.oSm;
  a.oSv = L.oSv;
  a.os = L.os;
  a.oQf = L.oQf;
  a.oqf = L.oqf;
  a.ocf = L.ocf;
  a.trec = L.trec;
  a.begin = begin;
  a.batch = batch;
  a.with_rollout = with_rollout ? 1 : 0;
  a.eps = st.eps;
  a.alpha = alpha;
  ilqr_dmma_kernel<<<count, kThreads, smem, stream>>>(a);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace o2c
