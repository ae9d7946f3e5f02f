// wpp_tiles.cuh — building blocks of the warp-per-problem DMMA kernels for 24 x 24 blocks (the legged-robot shape): FP64 tensor-pipe
// fragments, the conflict-free scratch layout, and the blocked Cholesky / triangular inverse. Shared by the ILQR sweep (riccati_wpp.cu)
// and the SLQ kernels (slq_wpp.cu). Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once

#include "o2c_common.cuh"

namespace o2c {
namespace {

constexpr int kN = 24;            // nx == nu
constexpr int kMat = kN * kN;     // 576
constexpr int kLd = 26;           // leading dimension of the scratch matrix (conflict-free 8-byte transposed access)
// Residency: one CTA per SM whose warp count is chosen per launch (1..12: 168 registers per thread and 15.9 KB of shared memory per warp
// allow three warps on each of the four schedulers; a fourth would need <= 128 registers: 4 x 32 x 144 > the 16 K registers of a scheduler).
constexpr int kMaxWarps = 12;
constexpr int kOperand = 2 * kMat + 3 * kN + 2 + 6;  // {A|B|Hv|q|r|c,pad} doubles staged by TMA (= Layout::oQ for n = m = 24, 128-byte aligned)
constexpr unsigned kFull = 0xffffffffu;
// record offsets of make_layout(24, 24, 0, N, ILQR) as compile-time constants (address arithmetic folds into the instructions'
// immediate fields); wpp_ilqr_supported() checks that the handle's layout is exactly this one
constexpr int kRec = kOperand + 3 * kMat;  // 2960: { A | B | Hv | q | r | c,pad | Q | P | R }
constexpr int kOQ = kOperand, kOP = kOperand + kMat, kOR = kOperand + 2 * kMat;
constexpr int kOK = 0, kOdb = kMat, kObias = kMat + kN, kOSm = kMat + 2 * kN, kOSv = 2 * kMat + 2 * kN, kOs = 2 * kMat + 3 * kN;
constexpr int kORec = 2 * kMat + 3 * kN + 2 + 6;  // 1232: { K | dbias | bias | Sm | Sv | s,pad }

// ---------------------------------------------------------------------------------------------------------------------
// DMMA and fragment helpers (the mbarrier / TMA / prefetch wrappers live in o2c_common.cuh)
// ---------------------------------------------------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe. Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4], d = D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}
// acc(Z tile) += X(kblock, i)' Y(kblock, j) for one 8-deep k block; x, y are "op" fragments (rows {2c, 2c+1} of the k block, column r),
// so the two DMMAs contract k = {0,2,4,6} and {1,3,5,7} (the same permutation on both operands).
__device__ __forceinline__ void dmma2(double2& d, const double2& x, const double2& y) {
  dmma(d, x.x, y.x);
  dmma(d, x.y, y.y);
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, const double2& v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stg2(double* p, const double2& v) { __stcg(reinterpret_cast<double2*>(p), v); }
__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }
// sign flip on the integer pipe (keeps the FP64 pipe for the contractions)
__device__ __forceinline__ double neg(double v) { return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v)); }
__device__ __forceinline__ double2 neg2(const double2& v) { return make_double2(neg(v.x), neg(v.y)); }
// exponent field all ones <=> Inf or NaN, tested on the integer pipe
__device__ __forceinline__ bool finite_bits(double v) { return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000; }
__device__ __forceinline__ bool finite2(const double2& v) { return finite_bits(v.x) && finite_bits(v.y); }

// 1/sqrt(d) for a normal positive d: MUFU.RSQ64H seed (>= 20 good bits) and one third-order step, relative error below 1e-17
// (the library rsqrt spends ~16 instructions on denormal / special-case handling the pivot test has already excluded)
__device__ __forceinline__ double rsqrt_pivot(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double ye = y * e;
  return fma(ye, p, y);
}

// tile offsets: rows 8*rb.., columns 8*cb.. of a column-major matrix with leading dimension 24 (staged / global) or 26 (scratch)
__device__ __forceinline__ constexpr int t24(int rb, int cb) { return 8 * rb + 8 * kN * cb; }
__device__ __forceinline__ constexpr int t26(int rb, int cb) { return 8 * rb + 8 * kLd * cb; }
// index of lower tile (ib >= jb) in a packed array of six
__device__ __forceinline__ constexpr int lt(int ib, int jb) { return ib * (ib + 1) / 2 + jb; }

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
__device__ __forceinline__ double pick3(const double (&z)[3], int c) { return c == 0 ? z[0] : (c == 1 ? z[1] : z[2]); }

// "transposed" fragment of the scratch: lane (r,c) gets M[8rb + r][8cb + 2c .. 2c+1] (two conflict-free 8-byte loads)
__device__ __forceinline__ double2 tfrag(const double* W, int rb, int cb, int r, int c) {
  const double* p = W + (8 * rb + r) + kLd * (8 * cb + 2 * c);
  return make_double2(p[0], p[kLd]);
}
__device__ __forceinline__ void tput(double* W, int rb, int cb, int r, int c, const double2& v) {
  double* p = W + (8 * rb + r) + kLd * (8 * cb + 2 * c);
  p[0] = v.x;
  p[kLd] = v.y;
}

// z[8 jb + r] = sum_k M[k + LD j] v[k] (transposed product) with the operand-fragment access pattern; every lane of quad r gets z[jb].
// UPPER: M is block upper triangular (only tiles kb <= jb are read).
template <int LD, bool UPPER, int NB = 3>
__device__ __forceinline__ void matvec_cols(const double* M, const double* v, int r, int c, double (&z)[NB]) {
  double2 vf[NB];
#pragma unroll
  for (int kb = 0; kb < NB; ++kb) vf[kb] = ld2(v + 8 * kb + 2 * c);
#pragma unroll
  for (int jb = 0; jb < NB; ++jb) {
    double p = 0.0;
#pragma unroll
    for (int kb = 0; kb < (UPPER ? jb + 1 : NB); ++kb) {
      const double2 mv = ld2(M + 2 * c + LD * r + 8 * kb + 8 * LD * jb);
      p = fma(mv.x, vf[kb].x, p);
      p = fma(mv.y, vf[kb].y, p);
    }
    z[jb] = quad_sum(p);
  }
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

// ---------------------------------------------------------------------------------------------------------------------
// Blocked factorisation of Hm (24x24 in 8x8 blocks). In: the six lower tiles of Hm as accumulator fragments. Out: the scratch
// holds L^-T column-major in its block upper triangle (exact zeros below the diagonal inside the diagonal tiles; the strictly
// lower tiles hold left-over panels of L and are never read again).
//   * each block column is factorised one row per lane (rows 8b..23: the unblocked right-looking recurrence restricted to the
//     8 columns of the block, which also solves the panel below the diagonal block), 8 pivots per block;
//   * the trailing tiles are updated on the tensor pipe (Hm_ij -= L_ib L_jb');
//   * the three diagonal blocks are inverted together, one row of L_bb^-T per lane (8 steps), and the off-diagonal blocks of
//     L^-1 follow from three small DMMA chains:  Li10 = -Li11 (L10 Li00),  Li21 = -Li22 (L21 Li11),
//     Li20 = -Li22 (L20 Li00 + L21 Li10).
// ---------------------------------------------------------------------------------------------------------------------
// NB: the matrix is 8 NB x 8 NB (NB = 3: Hm; NB = 1, 2: the Gram matrix M = Z'Z of up to 16 equality constraints, which lives in the
// leading rows / columns of its own scratch). CLAMP: pivots are floored instead of being tested for positivity — the reference clamps
// the diagonal of the R factor of its constraint QR, |Rc_jj| = L_M,jj, to 1e-9 (LinearAlgebra.cpp:38-47), i.e. the pivot L_M,jj^2 to
// 1e-18. The Gram form resolves a pivot only down to rounding of its diagonal entry (a dependent row leaves ~1e-16 M_jj instead of
// 0), so the floor is max(1e-18, 1e-13 M_jj); the return value then says "no pivot was clamped" (full row rank to working accuracy).
template <int NB, bool CLAMP>
__device__ __forceinline__ bool factor_blocks(double* W, int lane, int r, int c) {
  constexpr int kRows = 8 * NB;
  const int li = lane < kRows ? lane : kRows - 1;
  bool pd = true;
  const double dorig = CLAMP ? W[li + kLd * li] : 0.0;  // the lane's own diagonal entry before the elimination
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    double* col = W + kLd * 8 * b;  // column 8b of the scratch
    const bool owner = lane >= 8 * b && lane < kRows;
    double g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = col[li + kLd * k];
    // pivots in pairs (j, j+1): the second pivot of a pair is d1 = c - b^2/d0 with [d0 b; b c] the leading 2x2 block, so
    // 1/sqrt(d1) = rsqrt(d0 c - b^2) sqrt(d0) does not wait for rsqrt(d0): two independent MUFU chains per round trip through
    // shared memory instead of one (same conditioning: both forms subtract b^2 (/d0) from (d0) c).
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      double d0 = __shfl_sync(kFull, g[j], 8 * b + j);
      const double bq = __shfl_sync(kFull, g[j], 8 * b + j + 1);
      const double cq = __shfl_sync(kFull, g[j + 1], 8 * b + j + 1);
      double floor1 = 0.0;
      if (CLAMP) {
        const double floor0 = fmax(1e-18, 1e-13 * __shfl_sync(kFull, dorig, 8 * b + j));
        floor1 = fmax(1e-18, 1e-13 * __shfl_sync(kFull, dorig, 8 * b + j + 1));
        if (!(d0 >= floor0)) {
          pd = false;
          d0 = floor0;
        }
      }
      double det = fma(d0, cq, -bq * bq);
      if (CLAMP) {
        if (!(det >= floor1 * d0)) {  // second pivot det / d0 below the floor
          pd = false;
          det = floor1 * d0;
        }
      } else {
        pd = pd && (__double2hiint(d0) > 0) && (__double2hiint(det) > 0);  // integer pipe, off the dependency chain; a non-positive
      }                                                                    // pivot makes rs NaN / Inf, which then propagates like the
      const double rs0 = rsqrt_pivot(d0);                                  // NaNs of the reference's LLT
      const double rdet = rsqrt_pivot(det);
      const double rs1 = rdet * (d0 * rs0);
      const double l0 = g[j] * rs0;
      const double lb = bq * rs0;  // L[j+1][j]
      const double l1 = fma(-l0, lb, g[j + 1]) * rs1;
      if (owner) {
        col[lane + kLd * j] = (lane == 8 * b + j) ? rs0 : l0;  // 1/L_jj on the diagonal
        col[lane + kLd * (j + 1)] = (lane == 8 * b + j + 1) ? rs1 : l1;
      }
      __syncwarp();
#pragma unroll
      for (int k = j + 2; k < 8; ++k) {
        g[k] = fma(-l0, col[8 * b + k + kLd * j], g[k]);
        g[k] = fma(-l1, col[8 * b + k + kLd * (j + 1)], g[k]);
      }
    }
    // trailing tiles (in the scratch): Hm_ij -= L_ib L_jb'
    if (NB >= 2 && b == 0) {
      const double2 f1 = tfrag(W, 1, 0, r, c);
      const double2 n1 = neg2(f1);
      double2 h11 = tfrag(W, 1, 1, r, c);
      dmma2(h11, n1, f1);
      tput(W, 1, 1, r, c, h11);
      if (NB == 3) {
        const double2 f2 = tfrag(W, 2, 0, r, c);
        const double2 n2 = neg2(f2);
        double2 h21 = tfrag(W, 2, 1, r, c), h22 = tfrag(W, 2, 2, r, c);
        dmma2(h21, n2, f1);
        dmma2(h22, n2, f2);
        tput(W, 2, 1, r, c, h21);
        tput(W, 2, 2, r, c, h22);
      }
    } else if (NB == 3 && b == 1) {
      const double2 f2 = tfrag(W, 2, 1, r, c);
      double2 h22 = tfrag(W, 2, 2, r, c);
      dmma2(h22, neg2(f2), f2);
      tput(W, 2, 2, r, c, h22);
    }
    __syncwarp();
  }
  // diagonal blocks: lane (8 bi + ii) computes row ii of L_bb^-T by the recurrence applied to e_ii
  {
    const int bi = li >> 3, ii = li & 7;
    double* dg = W + 8 * bi + kLd * 8 * bi;
    double g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = (k == ii) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double a = g[j] * dg[j + kLd * j];
      g[j] = a;
#pragma unroll
      for (int k = j + 1; k < 8; ++k) g[k] = fma(-a, dg[k + kLd * j], g[k]);
    }
    __syncwarp();
    if (lane < kRows) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dg[ii + kLd * k] = g[k];
    }
    __syncwarp();
  }
  // off-diagonal blocks of L^-1 (stored transposed into the upper tiles)
  if (NB >= 2) {
    const double2 li00 = tfrag(W, 0, 0, r, c);                        // op fragment of Li00
    const double2 l10 = tfrag(W, 1, 0, r, c);                         // op fragment of L10'
    const double2 lt11 = ld2(W + 2 * c + kLd * r + t26(1, 1));        // op fragment of Li11'
    double2 x10 = zero2();
    dmma2(x10, li00, l10);  // (L10 Li00)' as accumulator = op fragment of L10 Li00
    double2 i10 = zero2();
    dmma2(i10, lt11, x10);
    i10 = neg2(i10);
    if (NB == 3) {
      const double2 li11 = tfrag(W, 1, 1, r, c);
      const double2 l20 = tfrag(W, 2, 0, r, c), l21 = tfrag(W, 2, 1, r, c);  // op fragments of L20', L21'
      const double2 lt22 = ld2(W + 2 * c + kLd * r + t26(2, 2));             // op fragment of Li22'
      double2 x21 = zero2(), x20 = zero2();
      dmma2(x21, li11, l21);
      dmma2(x20, li00, l20);
      double2 i21 = zero2();
      dmma2(i21, lt22, x21);
      i21 = neg2(i21);
      st2(W + 2 * c + kLd * r + t26(0, 1), i10);  // acc(Li10) = op fragment of (L^-T) tile (0,1)
      st2(W + 2 * c + kLd * r + t26(1, 2), i21);
      __syncwarp();
      dmma2(x20, tfrag(W, 0, 1, r, c), l21);  // + (L21 Li10)': op fragment of Li10 = transposed read of tile (0,1)
      double2 i20 = zero2();
      dmma2(i20, lt22, x20);
      st2(W + 2 * c + kLd * r + t26(0, 2), neg2(i20));
    } else {
      st2(W + 2 * c + kLd * r + t26(0, 1), i10);
    }
    __syncwarp();
  }
  return pd;
}
// Hm (24 x 24). `li` is kept in the signature for the callers' sake (lane clamped to 23).
__device__ __forceinline__ bool factor_hm(double* W, int lane, int li, int r, int c) {
  (void)li;
  return factor_blocks<3, false>(W, lane, r, c);
}

}  // namespace
}  // namespace o2c
