// riccati_generic.cu — shape-generic backward-pass kernels (one warp owns one problem's time-sequential sweep).
//
// These kernels cover EVERY configuration of the ABI (any nx/nu, state-input equality constraints, LINE_SEARCH or
// LEVENBERG_MARQUARDT modification, reduced or full Riccati form, DIAGONAL_SHIFT or GERSHGORIN correction, ILQR and SLQ-RK4).
// They follow the reference's projected formulation stage by stage:
//   Hm = R + B'SB (+ mu B'B)                      ILQR::computeHamiltonianHessian            ocs2_ddp/src/ILQR.cpp:217-222
//   Ui = U^-1, Hm = U'U                           LinearAlgebra::computeInverseMatrixUUT     ocs2_core/src/misc/LinearAlgebra.cpp:119-124
//   Ddagger, Pu by Householder QR                 LinearAlgebra::computeConstraintProjection LinearAlgebra.cpp:129-155
//   projected LQ                                  projectLQ / changeOfInputVariables         DDP_HelperFunctions.cpp:143-201, ChangeOfInputVariables.cpp:34-108
//   dQ, dGm, dGv                                  computeRiccatiModification                 LineSearchStrategy.cpp:294-312, LevenbergMarquardtStrategy.cpp:230-240
//   one-step map / flow map                       DiscreteTimeRiccatiEquations.cpp:65-154, ContinuousTimeRiccatiEquations.cpp:170-292
//   K, bias, dbias                                ILQR/SLQ::calculateControllerWorker         ILQR.cpp:162-181, SLQ.cpp:127-169
// The shape-specialised warp-per-problem DMMA/TMA kernel for the unconstrained LINE_SEARCH/reduced/DIAGONAL_SHIFT ILQR sweep of the
// legged shape lives in riccati_wpp.cu; o2c_backward / o2c_solve dispatch to it when the configuration matches.
#include <cfloat>

#include "o2c_common.cuh"

namespace o2c {
namespace {

struct Dims {
  int n, m, ncmax;
};

// per-warp workspace: pointers into the warp's shared-memory slice
struct Work {
  double* rec;  // stage record, transformed in place (A -> A~, Q -> Q~, Hv -> Hv~, q -> q~)
  double *A, *B, *Q, *P, *R, *Hv, *q, *r, *C, *D, *e;
  double *Hm, *Ui, *Qm, *Mx, *Rci, *Ddag, *PuBuf, *Px, *T;
  double *Bt, *Pt, *Rt, *SA, *SB, *Gm, *Km, *dGm, *dQ, *Kout, *Ev, *ew;
  double *u0, *tv, *rt, *Gv, *Lv, *dGv, *w, *SHv, *kout, *bout, *tau, *dqd, *HmLv, *dinv;
  double ct;        // projected c
  const double* Pu; // m x p (aliases Ui when nc == 0)
  const double* Tm; // P + R Px (aliases P when nc == 0)
  const double* tvp;
  int p;
};

struct WorkSizes {
  int total;  // doubles per warp
};

__host__ __device__ inline int work_doubles(const Layout& L, bool full, bool lm, bool gersh) {
  const int n = L.n, m = L.m, nc = L.ncmax;
  const int mx = n > m ? n : m;
  int t = L.rec;
  t += 2 * m * m;  // Hm, Ui
  if (nc > 0) t += m * m /*Qm*/ + m * nc /*Mx*/ + nc * nc /*Rci*/ + m * nc /*Ddag*/ + m * m /*Pu*/ + 2 * m * n /*Px,T*/;
  t += n * m + m * n;             // Bt, Pt
  if (full) t += m * m;           // Rt
  t += n * n + mx * m + 2 * m * n;  // SA, SB, Gm, Km
  if (lm) t += m * n;             // dGm
  if (gersh) t += 2 * n * n + n;  // dQ, and the eigenvector / eigenvalue scratch of EIGENVALUE_MODIFICATION
  t += m * n;                     // Kout
  t += 10 * m + 3 * n + nc;       // u0 tv rt Gv Lv dGv kout bout HmLv dinv | w SHv dqd | tau
  return (t + 1) & ~1;
}

__device__ __forceinline__ void carve(Work& W, double* base, const Layout& L, bool full, bool lm, bool gersh) {
  const int n = L.n, m = L.m, nc = L.ncmax;
  const int mx = n > m ? n : m;
  double* p = base;
  auto take = [&](int cnt) {
    double* r = p;
    p += cnt;
    return r;
  };
  W.rec = take(L.rec);
  W.A = W.rec + L.oA;
  W.B = W.rec + L.oB;
  W.Q = W.rec + L.oQ;
  W.P = W.rec + L.oP;
  W.R = W.rec + L.oR;
  W.Hv = W.rec + L.oHv;
  W.q = W.rec + L.oq;
  W.r = W.rec + L.or_;
  W.C = W.rec + L.oC;
  W.D = W.rec + L.oD;
  W.e = W.rec + L.oe;
  W.Hm = take(m * m);
  W.Ui = take(m * m);
  if (nc > 0) {
    W.Qm = take(m * m);
    W.Mx = take(m * nc);
    W.Rci = take(nc * nc);
    W.Ddag = take(m * nc);
    W.PuBuf = take(m * m);
    W.Px = take(m * n);
    W.T = take(m * n);
  } else {
    W.Qm = W.Mx = W.Rci = W.Ddag = W.PuBuf = W.Px = W.T = nullptr;
  }
  W.Bt = take(n * m);
  W.Pt = take(m * n);
  W.Rt = full ? take(m * m) : nullptr;
  W.SA = take(n * n);
  W.SB = take(mx * m);
  W.Gm = take(m * n);
  W.Km = take(m * n);
  W.dGm = lm ? take(m * n) : nullptr;
  W.dQ = gersh ? take(n * n) : nullptr;
  W.Ev = gersh ? take(n * n) : nullptr;
  W.ew = gersh ? take(n) : nullptr;
  W.Kout = take(m * n);
  W.u0 = take(m);
  W.tv = take(m);
  W.rt = take(m);
  W.Gv = take(m);
  W.Lv = take(m);
  W.dGv = take(m);
  W.kout = take(m);
  W.bout = take(m);
  W.HmLv = take(m);
  W.w = take(n);
  W.SHv = take(n);
  W.dqd = take(n);
  W.tau = take(nc);
  W.dinv = take(m);
}

// in-place Cholesky (lower) of the m x m matrix H (ld m). Returns false if a pivot is not positive (NaNs propagate like the reference).
__device__ __forceinline__ bool warp_cholesky(int m, double* H) {
  const int lane = lane_id();
  bool ok = true;
  if ((m & 7) == 0 && m >= 16) {
    // blocked right-looking variant for tile-aligned sizes: eight columns are factorised with their panel by rank-1 updates restricted to
    // the block, the trailing matrix takes ONE rank-8 update through the warp GEMM (FP64 tensor pipe). Same pivots, same NaN propagation;
    // the strictly upper triangle of the trailing blocks is scratch (nothing reads it: U = L' is taken from the lower triangle).
    for (int b0 = 0; b0 < m; b0 += 8) {
      const int b1 = b0 + 8;
      for (int j = b0; j < b1; ++j) {
        double d = H[j + j * m];
        if (!(d > 0.0)) {
          ok = false;
          d = __longlong_as_double(0x7ff8000000000000LL);
        }
        const double rs = rsqrt(d);
        const double ljj = d * rs;
        __syncwarp();
        for (int i = j + 1 + lane; i < m; i += 32) H[i + j * m] *= rs;
        if (lane == 0) H[j + j * m] = ljj;
        __syncwarp();
        for (int k = j + 1 + (lane >> 3); k < b1; k += 4) {
          const double lkj = H[k + j * m];
          for (int i = k + (lane & 7); i < m; i += 8) H[i + k * m] = fma(-H[i + j * m], lkj, H[i + k * m]);
        }
        __syncwarp();
      }
      if (b1 < m) wgemm<false, true>(m - b1, m - b1, 8, -1.0, H + b1 + b0 * m, m, H + b1 + b0 * m, m, 1.0, H + b1 + b1 * m, m);
    }
    return ok;
  }
  for (int j = 0; j < m; ++j) {
    double d = H[j + j * m];
    if (!(d > 0.0)) {
      ok = false;
      d = __longlong_as_double(0x7ff8000000000000LL);
    }
    const double rs = rsqrt(d);  // one reciprocal square root per pivot instead of a square root and a division
    const double ljj = d * rs;
    __syncwarp();
    for (int i = j + 1 + lane; i < m; i += 32) H[i + j * m] *= rs;
    if (lane == 0) H[j + j * m] = ljj;
    __syncwarp();
    // trailing update of the lower triangle: four columns at a time, eight row lanes per column (no index divisions)
    for (int k = j + 1 + (lane >> 3); k < m; k += 4) {
      const double lkj = H[k + j * m];
      for (int i = k + (lane & 7); i < m; i += 8) H[i + k * m] = fma(-H[i + j * m], lkj, H[i + k * m]);
    }
    __syncwarp();
  }
  return ok;
}

// X = U^-1 for an upper-triangular U given by U(i,k) = Ub[i*si + k*sk]; X is dim x dim, ld dim, strictly lower part zeroed.
__device__ __forceinline__ void warp_upper_inverse(int dim, const double* Ub, int si, int sk, double* X, double* dinv) {
  for (int i = lane_id(); i < dim; i += 32) dinv[i] = 1.0 / Ub[i * si + i * sk];  // the divisions, once and in parallel
  __syncwarp();
  if (sk == 1 && (dim & 7) == 0 && dim >= 16 && dim <= 32) {
    // blocked variant for tile-aligned sizes (U = L' of a Cholesky factor stored lower, ld si): the 8x8 diagonal blocks are inverted side by
    // side by back substitution (one column per lane: a chain of at most 28 fmas instead of dim^2 / 2), the off-diagonal blocks follow from
    // X_ab = -X_aa (sum_{a < c <= b} U_ac X_cb) through the warp GEMM; the (zero) lower block (b, a) serves as the scratch of the sum
    const int lane = lane_id(), nb = dim >> 3;
    if (lane < dim) {
      const int j = lane, b0 = j & ~7;
      for (int i = j; i >= b0; --i) {
        double v = (i == j) ? 1.0 : 0.0;
        for (int k = i + 1; k <= j; ++k) v = fma(-Ub[i * si + k], X[k + j * dim], v);
        X[i + j * dim] = v * dinv[i];
      }
      for (int i = j + 1; i < b0 + 8; ++i) X[i + j * dim] = 0.0;  // the diagonal blocks enter the products below as full 8x8 tiles
    }
    __syncwarp();
    for (int b = 1; b < nb; ++b)
      for (int a = b - 1; a >= 0; --a) {
        double* S = X + 8 * b + 8 * a * dim;
        for (int c = a + 1; c <= b; ++c)
          wgemm<true, false>(8, 8, 8, 1.0, Ub + 8 * c + 8 * a * si, si, X + 8 * c + 8 * b * dim, dim, c == a + 1 ? 0.0 : 1.0, S, dim);
        wgemm<false, false>(8, 8, 8, -1.0, X + 8 * a + 8 * a * dim, dim, S, dim, 0.0, X + 8 * a + 8 * b * dim, dim);
      }
    for (int idx = lane; idx < dim * dim; idx += 32)
      if ((idx % dim) > (idx / dim)) X[idx] = 0.0;
    __syncwarp();
    return;
  }
  for (int j = lane_id(); j < dim; j += 32) {
    for (int i = j; i >= 0; --i) {
      double v = (i == j) ? 1.0 : 0.0;
      for (int k = i + 1; k <= j; ++k) v = fma(-Ub[i * si + k * sk], X[k + j * dim], v);
      X[i + j * dim] = v * dinv[i];
    }
    for (int i = j + 1; i < dim; ++i) X[i + j * dim] = 0.0;
  }
  __syncwarp();
}

// Householder QR of Mx (rows x cols, ld rows) restating Eigen::HouseholderQR; Q (rows x rows) formed explicitly.
__device__ __forceinline__ void warp_householder_qr(int rows, int cols, double* Mx, double* tau, double* Qm) {
  const int lane = lane_id();
  for (int j = 0; j < cols; ++j) {
    double part = 0.0;
    for (int i = j + 1 + lane; i < rows; i += 32) part = fma(Mx[i + j * rows], Mx[i + j * rows], part);
    const double tailSq = warp_sum(part);
    const double c0 = Mx[j + j * rows];
    double beta, tj;
    __syncwarp();
    if (tailSq <= DBL_MIN) {
      tj = 0.0;
      beta = c0;
      for (int i = j + 1 + lane; i < rows; i += 32) Mx[i + j * rows] = 0.0;
    } else {
      beta = sqrt(fma(c0, c0, tailSq));
      if (c0 >= 0.0) beta = -beta;
      const double denom = c0 - beta;
      for (int i = j + 1 + lane; i < rows; i += 32) Mx[i + j * rows] /= denom;
      tj = (beta - c0) / beta;
    }
    if (lane == 0) {
      Mx[j + j * rows] = beta;
      tau[j] = tj;
    }
    __syncwarp();
    for (int c = j + 1 + lane; c < cols; c += 32) {
      double wv = Mx[j + c * rows];
      for (int i = j + 1; i < rows; ++i) wv = fma(Mx[i + j * rows], Mx[i + c * rows], wv);
      wv *= tj;
      Mx[j + c * rows] -= wv;
      for (int i = j + 1; i < rows; ++i) Mx[i + c * rows] = fma(-Mx[i + j * rows], wv, Mx[i + c * rows]);
    }
    __syncwarp();
  }
  for (int idx = lane; idx < rows * rows; idx += 32) Qm[idx] = (idx % rows == idx / rows) ? 1.0 : 0.0;
  __syncwarp();
  for (int j = cols - 1; j >= 0; --j) {
    const double tj = tau[j];
    for (int c = lane; c < rows; c += 32) {
      double wv = Qm[j + c * rows];
      for (int i = j + 1; i < rows; ++i) wv = fma(Mx[i + j * rows], Qm[i + c * rows], wv);
      wv *= tj;
      Qm[j + c * rows] -= wv;
      for (int i = j + 1; i < rows; ++i) Qm[i + c * rows] = fma(-Mx[i + j * rows], wv, Qm[i + c * rows]);
    }
    __syncwarp();
  }
}

// LinearAlgebra::makePsdEigenvalue (ocs2_core/src/misc/LinearAlgebra.cpp:52-72) on the symmetric matrix whose lower triangle is in M:
// R = V max(lambda, eps) V' if an eigenvalue is below eps, sym(M) otherwise. A, V: n x n scratch, w: n scratch; the result lands in A.
// The common case (spectrum above eps) is detected by a Cholesky of M - eps I, the rare one runs cyclic Jacobi rotations
// (Eigen's tridiagonal QR is not restated: V max(lambda, eps) V' does not depend on the algorithm beyond rounding).
__device__ __forceinline__ void warp_make_psd_eigenvalue(int n, const double* M, double eps, double* A, double* V, double* w) {
  const int lane = lane_id();
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx % n, j = idx / n;
    const double v = (i >= j) ? M[i + j * n] : M[j + i * n];
    A[idx] = v;
    V[idx] = v - ((i == j) ? eps : 0.0);
  }
  __syncwarp();
  if (warp_cholesky(n, V)) {  // every eigenvalue is above eps: only the symmetrisation remains
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx % n, j = idx / n;
      A[idx] = 0.5 * (M[i + j * n] + M[j + i * n]);
    }
    __syncwarp();
    return;
  }
  for (int idx = lane; idx < n * n; idx += 32) V[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int idx = lane; idx < n * n; idx += 32) {
      const double v = A[idx] * A[idx];
      if (idx % n == idx / n)
        dg += v;
      else
        off += v;
    }
    off = warp_sum(off);
    dg = warp_sum(dg);
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + q * n];
        if (apq == 0.0) continue;  // uniform over the warp: every lane reads the same element
        const double theta = (A[q + q * n] - A[p + p * n]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        __syncwarp();
        for (int k = lane; k < n; k += 32) {  // A <- A J, V <- V J (columns p, q)
          const double akp = A[k + p * n], akq = A[k + q * n];
          A[k + p * n] = c * akp - sn * akq;
          A[k + q * n] = sn * akp + c * akq;
          const double vkp = V[k + p * n], vkq = V[k + q * n];
          V[k + p * n] = c * vkp - sn * vkq;
          V[k + q * n] = sn * vkp + c * vkq;
        }
        __syncwarp();
        for (int k = lane; k < n; k += 32) {  // A <- J' A (rows p, q)
          const double apk = A[p + k * n], aqk = A[q + k * n];
          A[p + k * n] = c * apk - sn * aqk;
          A[q + k * n] = sn * apk + c * aqk;
        }
        __syncwarp();
      }
  }
  for (int i = lane; i < n; i += 32) w[i] = fmax(A[i + i * n], eps);
  __syncwarp();
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx % n, j = idx / n;
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += V[i + k * n] * w[k] * V[j + k * n];
    A[idx] = acc;
  }
  __syncwarp();
}

// GaussNewtonDDP::computeProjectionAndRiccatiModification on the record loaded in W.rec. Snext == nullptr: Hm = R (SLQ).
// On return: A~ = W.A, Hv~ = W.Hv, Q~ = W.Q, q~ = W.q, c~ = W.ct, B~ = W.Bt, P~ = W.Pt (ld p), R~ = W.Rt (ld p, full form only),
// r~ = W.rt, Pu = W.Pu, Px = W.Px / u0 = W.u0 (nc > 0), dQ diag in W.dqd (DIAGONAL_SHIFT) or full in W.dQ (GERSHGORIN),
// dGm/dGv (LM).
__device__ __forceinline__ int project_stage(Work& W, const Layout& L, const SolverSettings& st, int nc, const double* Snext) {
  const int n = L.n, m = L.m, ldc = L.ncmax > 0 ? L.ncmax : 1;
  const int lane = lane_id();
  const int p = m - nc;
  int status = 0;
  W.p = p;
  const bool lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool full = !st.reduced;

  // ---- Hm ----
  wcopy(m * m, W.R, W.Hm);
  if (Snext != nullptr) {
    wgemm<false, false>(n, m, n, 1.0, Snext, n, W.B, n, 0.0, W.SB, n);   // S B
    wgemm<true, false>(m, m, n, 1.0, W.B, n, W.SB, n, 1.0, W.Hm, m);     // Hm += B' (S B)
  }
  if (lm) wgemm<true, false>(m, m, n, st.mu, W.B, n, W.B, n, 1.0, W.Hm, m);
  // ---- Ui ----
  if (!warp_cholesky(m, W.Hm)) status |= O2C_STATUS_CHOL_NOT_PD;
  warp_upper_inverse(m, W.Hm, m, 1, W.Ui, W.dinv);  // U(i,k) = L(k,i) = Hm[k + i*m]
  // ---- projectors ----
  if (nc == 0) {
    W.Pu = W.Ui;
    W.Tm = W.P;
    W.tvp = W.r;
  } else {
    wgemm<true, true>(m, nc, m, 1.0, W.Ui, m, W.D, ldc, 0.0, W.Mx, m);  // Ui' D'
    warp_householder_qr(m, nc, W.Mx, W.tau, W.Qm);
    // setTriangularMinimumEigenvalues (weakEpsilon = 1e-9)
    bool clamped = false;
    if (lane < nc) {
      double ev = W.Mx[lane + lane * m];
      const double cl = (ev < 0.0) ? fmin(-1e-9, ev) : fmax(1e-9, ev);
      clamped = (cl != ev);
      W.Mx[lane + lane * m] = cl;
    }
    if (__any_sync(0xffffffffu, clamped)) status |= O2C_STATUS_CONSTRAINT_RANK;
    __syncwarp();
    warp_upper_inverse(nc, W.Mx, 1, m, W.Rci, W.dinv);  // Rc(i,k) = Mx[i + k*m]
    // tmp (m x nc) = Qc * RcInv'  -> reuse SB as scratch (max(n,m)*m >= m*nc)
    wgemm<false, true>(m, nc, nc, 1.0, W.Qm, m, W.Rci, nc, 0.0, W.SB, m);
    wgemm<false, false>(m, nc, m, 1.0, W.Ui, m, W.SB, m, 0.0, W.Ddag, m);
    wgemm<false, false>(m, p, m, 1.0, W.Ui, m, W.Qm + nc * m, m, 0.0, W.PuBuf, m);
    W.Pu = W.PuBuf;
    // Px = -Ddag C, u0 = -Ddag e
    wgemm<false, false>(m, n, nc, -1.0, W.Ddag, m, W.C, ldc, 0.0, W.Px, m);
    wgemm<false, false>(m, 1, nc, -1.0, W.Ddag, m, W.e, ldc, 0.0, W.u0, m);
    // T = P + R Px ; tv = r + R u0
    wcopy(m * n, W.P, W.T);
    wgemm<false, false>(m, n, m, 1.0, W.R, m, W.Px, m, 1.0, W.T, m);
    wcopy(m, W.r, W.tv);
    wgemm<false, false>(m, 1, m, 1.0, W.R, m, W.u0, m, 1.0, W.tv, m);
    W.Tm = W.T;
    W.tvp = W.tv;
    // Q~ = Q + P'Px + Px'T ; q~ = q + P'u0 + Px'tv ; c~ = c + 1/2 u0.(tv + r)
    wgemm<true, false>(n, n, m, 1.0, W.P, m, W.Px, m, 1.0, W.Q, n);
    wgemm<true, false>(n, n, m, 1.0, W.Px, m, W.T, m, 1.0, W.Q, n);
    wgemm<true, false>(n, 1, m, 1.0, W.P, m, W.u0, m, 1.0, W.q, n);
    wgemm<true, false>(n, 1, m, 1.0, W.Px, m, W.tv, m, 1.0, W.q, n);
    double acc = 0.0;
    for (int i = lane; i < m; i += 32) acc = fma(W.u0[i], W.tv[i] + W.r[i], acc);
    W.ct += 0.5 * warp_sum(acc);
    // A~ = A + B Px ; Hv~ = Hv + B u0
    wgemm<false, false>(n, n, m, 1.0, W.B, n, W.Px, m, 1.0, W.A, n);
    wgemm<false, false>(n, 1, m, 1.0, W.B, n, W.u0, m, 1.0, W.Hv, n);
  }
  // B~ = B Pu ; P~ = Pu' T ; r~ = Pu' tv ; R~ = Pu' R Pu (full form only)
  wgemm<false, false>(n, p, m, 1.0, W.B, n, W.Pu, m, 0.0, W.Bt, n);
  wgemm<true, false>(p, n, m, 1.0, W.Pu, m, W.Tm, m, 0.0, W.Pt, p);
  wgemm<true, false>(p, 1, m, 1.0, W.Pu, m, W.tvp, m, 0.0, W.rt, p);
  if (full) {
    wgemm<false, false>(m, p, m, 1.0, W.R, m, W.Pu, m, 0.0, W.SB, m);
    wgemm<true, false>(p, p, m, 1.0, W.Pu, m, W.SB, m, 0.0, W.Rt, p);
  }
  // ---- Riccati modification ----
  if (!lm) {
    if (st.hc == O2C_HC_DIAGONAL_SHIFT) {
      // dQ = (M + eps I) - M with M = Q~ - P~'P~ : only the diagonal can be non-zero
      for (int i = lane; i < n; i += 32) {
        double mii = W.Q[i + i * n];
        double pp = 0.0;
        for (int k = 0; k < p; ++k) pp = fma(W.Pt[k + i * p], W.Pt[k + i * p], pp);
        mii -= pp;
        W.dqd[i] = (mii + st.eps) - mii;
      }
      __syncwarp();
    } else if (st.hc == O2C_HC_EIGENVALUE_MODIFICATION) {
      // EIGENVALUE: dQ = makePsdEigenvalue(M) - M with M = Q~ - P~'P~ (LineSearchStrategy.cpp:294-312). SA is free here: holds M.
      wcopy(n * n, W.Q, W.SA);
      wgemm<true, false>(n, n, p, -1.0, W.Pt, p, W.Pt, p, 1.0, W.SA, n);
      warp_make_psd_eigenvalue(n, W.SA, st.eps, W.dQ, W.Ev, W.ew);
      for (int idx = lane; idx < n * n; idx += 32) W.dQ[idx] -= W.SA[idx];
      __syncwarp();
    } else {
      // GERSHGORIN: M -> sym(M), M_ii = max(M_ii, R_i + eps); dQ = that - M (LinearAlgebra.cpp:77-85). SA is free here: holds M.
      wcopy(n * n, W.Q, W.SA);
      wgemm<true, false>(n, n, p, -1.0, W.Pt, p, W.Pt, p, 1.0, W.SA, n);
      for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx % n, j = idx / n;
        W.dQ[idx] = 0.5 * (W.SA[i + j * n] + W.SA[j + i * n]);
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        double colAbs = 0.0;
        for (int k = 0; k < n; ++k) colAbs += fabs(W.dQ[k + i * n]);
        const double dii = W.dQ[i + i * n];
        W.dqd[i] = fmax(dii, (colAbs - fabs(dii)) + st.eps);
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) W.dQ[i + i * n] = W.dqd[i];
      __syncwarp();
      for (int idx = lane; idx < n * n; idx += 32) W.dQ[idx] -= W.SA[idx];
      __syncwarp();
    }
  } else {
    wgemm<true, false>(p, n, n, st.mu, W.Bt, n, W.A, n, 0.0, W.dGm, p);
    wgemm<true, false>(p, 1, n, st.mu, W.Bt, n, W.Hv, n, 0.0, W.dGv, p);
  }
  return status;
}

__device__ __forceinline__ double dq_at(const Work& W, const SolverSettings& st, int n, int i, int j) {
  if (st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT) return 0.0;
  if (st.hc == O2C_HC_DIAGONAL_SHIFT) return (i == j) ? W.dqd[i] : 0.0;
  return W.dQ[i + j * n];
}

__device__ __forceinline__ void load_record(const Layout& L, const double* __restrict__ g, double* s) {
  const double2* g2 = reinterpret_cast<const double2*>(g);
  double2* s2 = reinterpret_cast<double2*>(s);
  for (int i = lane_id(); i < L.rec / 2; i += 32) s2[i] = __ldg(g2 + i);
  __syncwarp();
}

// un-projection of the controller and write of one output node: K = Px + Pu K~ ; dbias = u0 + Pu L~ ; bias = u_nom - K x_nom
__device__ __forceinline__ bool emit_controller(const Work& W, const Layout& L, int nc, const double* Km, const double* Lv, const double* xnom,
                                const double* unom, double* out) {
  const int n = L.n, m = L.m, p = W.p, lane = lane_id();
  bool finite = true;
  for (int idx = lane; idx < m * n; idx += 32) W.Kout[idx] = (nc > 0) ? W.Px[idx] : 0.0;
  __syncwarp();
  wgemm<false, false>(m, n, p, 1.0, W.Pu, m, Km, p, 1.0, W.Kout, m);
  for (int idx = lane; idx < m * n; idx += 32) {
    const double acc = W.Kout[idx];
    out[L.oK + idx] = acc;
    finite = finite && isfinite(acc);
  }
  __syncwarp();
  for (int i = lane; i < m; i += 32) {
    double acc = (nc > 0) ? W.u0[i] : 0.0;
    for (int k = 0; k < p; ++k) acc = fma(W.Pu[i + k * m], Lv[k], acc);
    out[L.odb + i] = acc;
    finite = finite && isfinite(acc);
    double b = unom ? unom[i] : 0.0;
    if (xnom)
      for (int j = 0; j < n; ++j) b = fma(-W.Kout[i + j * m], xnom[j], b);
    out[L.obias + i] = b;
  }
  __syncwarp();
  return __all_sync(0xffffffffu, finite);
}

__device__ __forceinline__ void copy_last_controller(const Layout& L, double* solp) {
  // GaussNewtonDDP::calculateController, GaussNewtonDDP.cpp:609-618
  const double* src = solp + (size_t)(L.N - 1) * L.orec;
  double* dst = solp + (size_t)L.N * L.orec;
  for (int i = lane_id(); i < L.oSm; i += 32) dst[i] = src[i];  // K, dbias, bias precede Sm in the record
}

// Shape specialisation: CN > 0 fixes (n, m, ncmax) at compile time (loops unroll, index arithmetic folds); STD fixes the settings to
// the benchmark configuration (reduced form, LINE_SEARCH, DIAGONAL_SHIFT) so the other variants' branches disappear. <0,0,0,false>
// is the fully run-time version that serves every other shape / setting.
template <int CN, int CM, int CNC, bool STD>
__device__ __forceinline__ void specialise(Layout& L, SolverSettings& st) {
  if (CN > 0) {
    L.n = CN;
    L.m = CM;
    L.ncmax = CNC;
  }
  if (STD) {
    st.reduced = 1;
    st.strategy = O2C_STRATEGY_LINE_SEARCH;
    st.hc = O2C_HC_DIAGONAL_SHIFT;
  }
}

template <int CN, int CM, int CNC, bool STD>
__global__ void __launch_bounds__(128) ilqr_generic_kernel(Layout L, SolverSettings st, DeviceBuffers buf, int begin, int count,
                                                           int warp_doubles) {
  specialise<CN, CM, CNC, STD>(L, st);
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int local = blockIdx.x * wpb + warp;
  if (local >= count) return;
  const int prob = begin + local;
  const int n = L.n, m = L.m;
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  double* base = smem + (size_t)warp * warp_doubles;
  Work W;
  carve(W, base + 2 * n * n + 2 * n, L, full, lm, gersh);
  double* Sa = base;              // S of node k+1
  double* Sb = base + n * n;      // S of node k
  double* Sva = base + 2 * n * n;
  double* Svb = Sva + n;
  double snext;

  const double* term = buf.term + (size_t)prob * L.trec;
  double* solp = buf.sol + (size_t)prob * (L.N + 1) * L.orec;
  int status = 0;
  // valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526)
  for (int i = lane; i < n * n; i += 32) {
    Sa[i] = term[L.oQf + i];
    solp[(size_t)L.N * L.orec + L.oSm + i] = Sa[i];
  }
  for (int i = lane; i < n; i += 32) {
    Sva[i] = term[L.oqf + i];
    solp[(size_t)L.N * L.orec + L.oSv + i] = Sva[i];
  }
  snext = term[L.ocf];
  if (lane == 0) solp[(size_t)L.N * L.orec + L.os] = snext;
  __syncwarp();

  for (int k = L.N - 1; k >= 0; --k) {
    load_record(L, buf.lq + ((size_t)prob * L.nodes + k) * L.rec, W.rec);
    W.ct = W.rec[L.oc];
    const int nc = (L.ncmax > 0) ? (buf.nc ? buf.nc[(size_t)prob * L.nodes + k] : L.ncmax) : 0;
    const bool is_event = buf.event != nullptr && buf.event[(size_t)prob * L.nodes + k] != 0;
    double sv;
    if (is_event) {
      // ---- pre-event node (ILQR.cpp:263-295): riccatiTransversalityConditions (RiccatiTransversalityConditions.h:40-56) on the jump
      // model data held in A, Hv, Q, q, c; then the controller entry from the regular data projected with Sm = 0 (Hm = R) ----
      wgemm<false, false>(n, 1, n, 1.0, Sa, n, W.Hv, n, 0.0, W.SHv, n);
      wgemm<true, false>(n, n, n, 1.0, Sa, n, W.A, n, 0.0, W.SA, n);  // Sm' A_e
      for (int i = lane; i < n; i += 32) W.w[i] = Sva[i] + W.SHv[i];
      __syncwarp();
      wcopy(n * n, W.Q, Sb);
      wgemm<true, false>(n, n, n, 1.0, W.SA, n, W.A, n, 1.0, Sb, n);  // Q_e + (Sm' A_e)' A_e
      wcopy(n, W.q, Svb);
      wgemm<true, false>(n, 1, n, 1.0, W.A, n, W.w, n, 1.0, Svb, n);  // q_e + A_e' (Sv + Sm Hv_e)
      sv = snext + W.ct + wdot(n, W.Hv, W.w) - 0.5 * wdot(n, W.Hv, W.SHv);
      __syncwarp();
      status |= project_stage(W, L, st, nc, nullptr);
      const int pe = W.p;
      for (int i = lane; i < pe * n; i += 32) W.Km[i] = -W.Pt[i] - (lm ? W.dGm[i] : 0.0);
      for (int i = lane; i < pe; i += 32) W.Lv[i] = -W.rt[i] - (lm ? W.dGv[i] : 0.0);
      __syncwarp();
      wgemm<true, false>(pe, n, n, -1.0, W.Bt, n, Sb, n, 1.0, W.Km, pe);
      wgemm<true, false>(pe, 1, n, -1.0, W.Bt, n, Svb, n, 1.0, W.Lv, pe);
    } else {
    status |= project_stage(W, L, st, nc, Sa);
    const int p = W.p;
    // ---- DiscreteTimeRiccatiEquations::computeMapILQR ----
    wgemm<false, false>(n, 1, n, 1.0, Sa, n, W.Hv, n, 0.0, W.SHv, n);
    wgemm<false, false>(n, n, n, 1.0, Sa, n, W.A, n, 0.0, W.SA, n);
    for (int i = lane; i < n; i += 32) W.w[i] = Sva[i] + W.SHv[i];
    __syncwarp();
    wcopy(p * n, W.Pt, W.Gm);
    wgemm<true, false>(p, n, n, 1.0, W.Bt, n, W.SA, n, 1.0, W.Gm, p);
    wcopy(p, W.rt, W.Gv);
    wgemm<true, false>(p, 1, n, 1.0, W.Bt, n, W.w, n, 1.0, W.Gv, p);
    for (int i = lane; i < p * n; i += 32) W.Km[i] = -W.Gm[i] - (lm ? W.dGm[i] : 0.0);
    for (int i = lane; i < p; i += 32) W.Lv[i] = -W.Gv[i] - (lm ? W.dGv[i] : 0.0);
    __syncwarp();
    double* HmKm = W.Kout;  // scratch until emit_controller
    if (full) {
      // projectedHm = R~ + (S B~)' B~ (overwrites Rt); HmKm; HmLv
      wgemm<false, false>(n, p, n, 1.0, Sa, n, W.Bt, n, 0.0, W.SB, n);
      wgemm<true, false>(p, p, n, 1.0, W.SB, n, W.Bt, n, 1.0, W.Rt, p);
      wgemm<false, false>(p, n, p, 1.0, W.Rt, p, W.Km, p, 0.0, HmKm, p);
      wgemm<false, false>(p, 1, p, 1.0, W.Rt, p, W.Lv, p, 0.0, W.HmLv, p);
    }
    // Sm = Q + dQ + (S A)'A + K~'G~ [+ (K~'G~)' + K~'(H~ K~)] through the warp GEMM (tensor pipe when the shape is made of whole tiles)
    for (int idx = lane; idx < n * n; idx += 32) Sb[idx] = W.Q[idx] + dq_at(W, st, n, idx % n, idx / n);
    __syncwarp();
    wgemm<true, false>(n, n, n, 1.0, W.SA, n, W.A, n, 1.0, Sb, n);
    if (!full) {
      wgemm<true, false>(n, n, p, 1.0, W.Km, p, W.Gm, p, 1.0, Sb, n);
    } else {
      wgemm<true, false>(n, n, p, 1.0, W.Km, p, W.Gm, p, 0.0, W.SA, n);  // K~'G~ (S A is dead from here on)
      for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx % n, j = idx / n;
        Sb[idx] += W.SA[idx] + W.SA[j + i * n];
      }
      __syncwarp();
      wgemm<true, false>(n, n, p, 1.0, W.Km, p, HmKm, p, 1.0, Sb, n);
    }
    // Sv
    for (int i = lane; i < n; i += 32) {
      double acc = W.q[i];
      for (int kk = 0; kk < n; ++kk) acc = fma(W.A[kk + i * n], W.w[kk], acc);
      for (int l = 0; l < p; ++l) acc = fma(W.Gm[l + i * p], W.Lv[l], acc);
      if (full)
        for (int l = 0; l < p; ++l) acc += W.Km[l + i * p] * W.Gv[l] + HmKm[l + i * p] * W.Lv[l];
      Svb[i] = acc;
    }
    // s
    sv = snext + W.ct + wdot(n, W.Hv, W.w) - 0.5 * wdot(n, W.Hv, W.SHv);
    if (st.reduced)
      sv += 0.5 * wdot(p, W.Lv, W.Gv);
    else
      sv += wdot(p, W.Lv, W.Gv) + 0.5 * wdot(p, W.Lv, W.HmLv);
    }
    __syncwarp();
    // ---- outputs of node k ----
    double* out = solp + (size_t)k * L.orec;
    bool finite = true;
    for (int i = lane; i < n * n; i += 32) {
      out[L.oSm + i] = Sb[i];
      finite = finite && isfinite(Sb[i]);
    }
    for (int i = lane; i < n; i += 32) out[L.oSv + i] = Svb[i];
    if (lane == 0) out[L.os] = sv;
    if (!__all_sync(0xffffffffu, finite)) status |= O2C_STATUS_NONFINITE;
    const double* xn = buf.x_nom ? buf.x_nom + ((size_t)prob * (L.N + 1) + k) * n : nullptr;
    const double* un = buf.u_nom ? buf.u_nom + ((size_t)prob * (L.N + 1) + k) * m : nullptr;
    if (!emit_controller(W, L, nc, W.Km, W.Lv, xn, un, out)) status |= O2C_STATUS_NONFINITE;
    // swap
    double* t = Sa;
    Sa = Sb;
    Sb = t;
    t = Sva;
    Sva = Svb;
    Svb = t;
    snext = sv;
    __syncwarp();
  }
  __syncwarp();
  if (L.N >= 1) copy_last_controller(L, solp);
  if (lane == 0) buf.status[prob] = status;
}

// ---------------------------------------------------------------------------------------------------------------------
// SLQ: per-node projection with Hm = R, Riccati flow map, RK4 with boost::odeint integrate_times semantics (step schedule
// precomputed on the host), controller. One warp per problem; two projected nodes are kept resident and lerped on the fly.
// ---------------------------------------------------------------------------------------------------------------------
struct ProjSet {
  double *At, *Bt, *Hvt, *Qt, *Pt, *Rt, *qt, *rt, *dQ, *dGm, *dGv, *Px, *u0, *Pu;
  double ct;
  int p, nc;
};

__host__ __device__ inline int projset_doubles(const Layout& L, bool full, bool lm, bool gersh, bool with_unproject) {
  const int n = L.n, m = L.m;
  int t = n * n + n * m + n + n * n + m * n + n + m;  // At Bt Hvt Qt Pt qt rt
  if (full) t += m * m;
  t += gersh ? n * n : n;  // dQ (full or diagonal)
  if (lm) t += m * n + m;
  if (with_unproject) t += m * n + m + m * m;  // Px u0 Pu
  return (t + 1) & ~1;
}

__device__ __forceinline__ void carve_proj(ProjSet& S, double* base, const Layout& L, bool full, bool lm, bool gersh, bool with_unproject) {
  const int n = L.n, m = L.m;
  double* p = base;
  auto take = [&](int cnt) {
    double* r = p;
    p += cnt;
    return r;
  };
  S.At = take(n * n);
  S.Bt = take(n * m);
  S.Hvt = take(n);
  S.Qt = take(n * n);
  S.Pt = take(m * n);
  S.qt = take(n);
  S.rt = take(m);
  S.Rt = full ? take(m * m) : nullptr;
  S.dQ = take(gersh ? n * n : n);
  S.dGm = lm ? take(m * n) : nullptr;
  S.dGv = lm ? take(m) : nullptr;
  if (with_unproject) {
    S.Px = take(m * n);
    S.u0 = take(m);
    S.Pu = take(m * m);
  } else {
    S.Px = S.u0 = S.Pu = nullptr;
  }
}

__device__ __forceinline__ void store_proj(const Work& W, const Layout& L, const SolverSettings& st, int nc, ProjSet& S) {
  const int n = L.n, m = L.m, p = W.p;
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  S.p = p;
  S.nc = nc;
  S.ct = W.ct;
  wcopy(n * n, W.A, S.At);
  wcopy(n * p, W.Bt, S.Bt);
  wcopy(n, W.Hv, S.Hvt);
  wcopy(n * n, W.Q, S.Qt);
  wcopy(p * n, W.Pt, S.Pt);
  wcopy(n, W.q, S.qt);
  wcopy(p, W.rt, S.rt);
  if (full) wcopy(p * p, W.Rt, S.Rt);
  if (gersh)
    wcopy(n * n, W.dQ, S.dQ);
  else if (!lm)
    wcopy(n, W.dqd, S.dQ);
  else
    wfill(n, 0.0, S.dQ);
  if (lm) {
    wcopy(p * n, W.dGm, S.dGm);
    wcopy(p, W.dGv, S.dGv);
  }
  if (nc > 0) {
    wcopy(m * n, W.Px, S.Px);
    wcopy(m, W.u0, S.u0);
  }
  wcopy(m * p, W.Pu, S.Pu);
}

// alpha * lhs + (1 - alpha) * rhs on every projected field (LinearInterpolation::interpolate). If the projected input dimension
// differs between the two nodes the reference takes the nearer node for input-sized fields.
__device__ __forceinline__ void lerp_proj(const Layout& L, bool full, bool lm, bool gersh, double a, const ProjSet& Lh, const ProjSet& Rh, ProjSet& O) {
  const int n = L.n, lane = lane_id();
  const double b = 1.0 - a;
  const bool same = (Lh.p == Rh.p);
  const ProjSet& pick = (a > 0.5) ? Lh : Rh;
  const int p = same ? Lh.p : pick.p;
  O.p = p;
  O.ct = a * Lh.ct + b * Rh.ct;
  for (int i = lane; i < n * n; i += 32) {
    O.At[i] = a * Lh.At[i] + b * Rh.At[i];
    O.Qt[i] = a * Lh.Qt[i] + b * Rh.Qt[i];
  }
  for (int i = lane; i < (gersh ? n * n : n); i += 32) O.dQ[i] = a * Lh.dQ[i] + b * Rh.dQ[i];
  for (int i = lane; i < n; i += 32) {
    O.Hvt[i] = a * Lh.Hvt[i] + b * Rh.Hvt[i];
    O.qt[i] = a * Lh.qt[i] + b * Rh.qt[i];
  }
  if (same) {
    for (int i = lane; i < n * p; i += 32) {
      O.Bt[i] = a * Lh.Bt[i] + b * Rh.Bt[i];
      O.Pt[i] = a * Lh.Pt[i] + b * Rh.Pt[i];
      if (lm) O.dGm[i] = a * Lh.dGm[i] + b * Rh.dGm[i];
    }
    for (int i = lane; i < p; i += 32) {
      O.rt[i] = a * Lh.rt[i] + b * Rh.rt[i];
      if (lm) O.dGv[i] = a * Lh.dGv[i] + b * Rh.dGv[i];
    }
    if (full)
      for (int i = lane; i < p * p; i += 32) O.Rt[i] = a * Lh.Rt[i] + b * Rh.Rt[i];
  } else {
    for (int i = lane; i < n * p; i += 32) {
      O.Bt[i] = pick.Bt[i];
      O.Pt[i] = pick.Pt[i];
      if (lm) O.dGm[i] = pick.dGm[i];
    }
    for (int i = lane; i < p; i += 32) {
      O.rt[i] = pick.rt[i];
      if (lm) O.dGv[i] = pick.dGv[i];
    }
    if (full)
      for (int i = lane; i < p * p; i += 32) O.Rt[i] = pick.Rt[i];
  }
  __syncwarp();
}

// y layout: upper triangle of Sm column-wise | Sv | s  (ContinuousTimeRiccatiEquations::convert2Vector, :55-81)
__device__ __forceinline__ int tri_index(int row, int col) { return col * (col + 1) / 2 + row; }  // row <= col

struct FlowWork {
  double *Sm, *Gm, *Km, *StA, *RmKm, *Gv, *Lv, *RmLv;
};

// ContinuousTimeRiccatiEquations::computeFlowMapSLQ
__device__ __forceinline__ void flow_map(const Layout& L, const SolverSettings& st, const ProjSet& P, const FlowWork& F, const double* y, double* dy) {
  const int n = L.n, p = P.p, lane = lane_id();
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  const int ntri = n * (n + 1) / 2;
  const double* Sv = y + ntri;
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx % n, j = idx / n;
    F.Sm[idx] = (i <= j) ? y[tri_index(i, j)] : y[tri_index(j, i)];
  }
  __syncwarp();
  wcopy(p * n, P.Pt, F.Gm);
  wgemm<true, false>(p, n, n, 1.0, P.Bt, n, F.Sm, n, 1.0, F.Gm, p);
  wcopy(p, P.rt, F.Gv);
  wgemm<true, false>(p, 1, n, 1.0, P.Bt, n, Sv, n, 1.0, F.Gv, p);
  for (int i = lane; i < p * n; i += 32) F.Km[i] = -(F.Gm[i] + (lm ? P.dGm[i] : 0.0));
  for (int i = lane; i < p; i += 32) F.Lv[i] = -(F.Gv[i] + (lm ? P.dGv[i] : 0.0));
  __syncwarp();
  wgemm<true, false>(n, n, n, 1.0, F.Sm, n, P.At, n, 0.0, F.StA, n);
  if (full) {
    wgemm<false, false>(p, n, p, 1.0, P.Rt, p, F.Km, p, 0.0, F.RmKm, p);
    wgemm<false, false>(p, 1, p, 1.0, P.Rt, p, F.Lv, p, 0.0, F.RmLv, p);
  }
  // dSv first: it is the last reader of S, whose storage then takes K~'G~ (+ K~'(R~ K~) / 2 in the full form)
  for (int i = lane; i < n; i += 32) {
    double acc = P.qt[i];
    for (int k = 0; k < n; ++k) acc = fma(F.Sm[k + i * n], P.Hvt[k], acc);
    for (int k = 0; k < n; ++k) acc = fma(P.At[k + i * n], Sv[k], acc);
    for (int l = 0; l < p; ++l) acc = fma(F.Gm[l + i * p], F.Lv[l], acc);
    if (full)
      for (int l = 0; l < p; ++l) acc += F.Km[l + i * p] * F.Gv[l] + F.RmKm[l + i * p] * F.Lv[l];
    dy[ntri + i] = acc;
  }
  __syncwarp();
  wgemm<true, false>(n, n, p, 1.0, F.Km, p, F.Gm, p, 0.0, F.Sm, n);
  if (full) wgemm<true, false>(n, n, p, 0.5, F.Km, p, F.RmKm, p, 1.0, F.Sm, n);  // symmetric: its two halves re-join in T + T' below
  // dSm (upper triangle only is packed)
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx % n, j = idx / n;
    if (i > j) continue;
    const double dq = gersh ? P.dQ[idx] : ((i == j) ? P.dQ[i] : 0.0);
    double acc = P.Qt[idx] + (dq + F.StA[i + j * n] + F.StA[j + i * n]);
    acc += full ? (F.Sm[i + j * n] + F.Sm[j + i * n]) : F.Sm[i + j * n];
    dy[tri_index(i, j)] = acc;
  }
  double ds = P.ct + wdot(n, P.Hvt, Sv);
  if (st.reduced)
    ds += 0.5 * wdot(p, F.Lv, F.Gv);
  else
    ds += wdot(p, F.Lv, F.Gv) + 0.5 * wdot(p, F.Lv, F.RmLv);
  if (lane == 0) dy[ntri + n] = ds;
  __syncwarp();
}

template <int CN, int CM, int CNC, bool STD>
__global__ void __launch_bounds__(128) slq_generic_kernel(Layout L, SolverSettings st, DeviceBuffers buf, const SlqStep* __restrict__ steps,
                                                          int nsteps, int begin, int count, int warp_doubles) {
  specialise<CN, CM, CNC, STD>(L, st);
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int local = blockIdx.x * wpb + warp;
  if (local >= count) return;
  const int prob = begin + local;
  const int n = L.n, m = L.m, N = L.N;
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  const int ntri = n * (n + 1) / 2, dim = ntri + n + 1;
  const int dimp = (dim + 1) & ~1;

  double* base = smem + (size_t)warp * warp_doubles;
  Work W;
  carve(W, base, L, full, lm, gersh);
  double* p = base + work_doubles(L, full, lm, gersh);
  ProjSet node[2], cur;
  const int psz = projset_doubles(L, full, lm, gersh, true);
  carve_proj(node[0], p, L, full, lm, gersh, true);
  p += psz;
  carve_proj(node[1], p, L, full, lm, gersh, true);
  p += psz;
  carve_proj(cur, p, L, full, lm, gersh, false);
  p += projset_doubles(L, full, lm, gersh, false);
  double* y = p;
  double* k1 = y + dimp;
  double* k2 = k1 + dimp;
  double* k3 = k2 + dimp;
  double* k4 = k3 + dimp;
  double* yt = k4 + dimp;
  p = yt + dimp;
  FlowWork F;
  F.Sm = p;
  p += n * n;
  F.StA = p;
  p += n * n;
  F.Gm = p;
  p += m * n;
  F.Km = p;
  p += m * n;
  F.RmKm = p;
  p += m * n;
  F.Gv = p;
  p += m;
  F.Lv = p;
  p += m;
  F.RmLv = p;
  p += m;

  const double* term = buf.term + (size_t)prob * L.trec;
  double* solp = buf.sol + (size_t)prob * (N + 1) * L.orec;
  int status = 0;

  auto project_node = [&](int k, ProjSet& dst) {
    load_record(L, buf.lq + ((size_t)prob * L.nodes + k) * L.rec, W.rec);
    W.ct = W.rec[L.oc];
    const int nc = (L.ncmax > 0) ? (buf.nc ? buf.nc[(size_t)prob * L.nodes + k] : L.ncmax) : 0;
    status |= project_stage(W, L, st, nc, nullptr);
    store_proj(W, L, st, nc, dst);
  };
  // SLQ::calculateControllerWorker at node k from the resident projected node and y = value function of node k
  auto controller = [&](int k, const ProjSet& S) {
    const int pk = S.p;
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx % n, j = idx / n;
      F.Sm[idx] = (i <= j) ? y[tri_index(i, j)] : y[tri_index(j, i)];
    }
    __syncwarp();
    for (int i = lane; i < pk * n; i += 32) F.Km[i] = -((lm ? S.dGm[i] : 0.0) + S.Pt[i]);
    __syncwarp();
    wgemm<true, false>(pk, n, n, -1.0, S.Bt, n, F.Sm, n, 1.0, F.Km, pk);
    for (int i = lane; i < pk; i += 32) F.Lv[i] = -((lm ? S.dGv[i] : 0.0) + S.rt[i]);
    __syncwarp();
    wgemm<true, false>(pk, 1, n, -1.0, S.Bt, n, y + ntri, n, 1.0, F.Lv, pk);
    // reuse emit_controller through a Work view of the stored projectors
    Work V = W;
    V.p = pk;
    V.Pu = S.Pu;
    V.Px = S.Px;
    V.u0 = S.u0;
    const double* xn = buf.x_nom ? buf.x_nom + ((size_t)prob * (N + 1) + k) * n : nullptr;
    const double* un = buf.u_nom ? buf.u_nom + ((size_t)prob * (N + 1) + k) * m : nullptr;
    if (!emit_controller(V, L, S.nc, F.Km, F.Lv, xn, un, solp + (size_t)k * L.orec)) status |= O2C_STATUS_NONFINITE;
  };
  auto write_value = [&](int k) {
    double* out = solp + (size_t)k * L.orec;
    bool finite = true;
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx % n, j = idx / n;
      const double v = (i <= j) ? y[tri_index(i, j)] : y[tri_index(j, i)];
      out[L.oSm + idx] = v;
      finite = finite && isfinite(v);
    }
    for (int i = lane; i < n; i += 32) out[L.oSv + i] = y[ntri + i];
    if (lane == 0) out[L.os] = y[ntri + n];
    if (!__all_sync(0xffffffffu, finite)) status |= O2C_STATUS_NONFINITE;
  };

  // terminal condition: allSsFinal = convert2Vector(finalValueFunction) (SLQ.cpp:228)
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx % n, j = idx / n;
    if (i <= j) y[tri_index(i, j)] = term[L.oQf + idx];
  }
  for (int i = lane; i < n; i += 32) y[ntri + i] = term[L.oqf + i];
  if (lane == 0) y[ntri + n] = term[L.ocf];
  __syncwarp();
  write_value(N);
  // node N is resident in slot N & 1
  project_node(N, node[N & 1]);
  controller(N, node[N & 1]);  // overwritten below by the copy of node N-1 (GaussNewtonDDP.cpp:609-618)

  int loaded_lo = N;  // lowest node index whose projection is resident
  for (int sidx = 0; sidx < nsteps; ++sidx) {
    const SlqStep sp = steps[sidx];
    const int i0 = sp.interval;
    if (i0 < loaded_lo) {
      project_node(i0, node[i0 & 1]);
      loaded_lo = i0;
    }
    const ProjSet& Lh = node[i0 & 1];
    const ProjSet& Rh = node[(i0 + 1) & 1];
    const double h = sp.h;
    if (sp.jump > 0) {
      // node i0 is a pre-event node: the segments on either side are integrated separately and joined by
      // ContinuousTimeRiccatiEquations::computeJumpMap = riccatiTransversalityConditions on the event's jump model data
      // (SLQ.cpp:286-296, ContinuousTimeRiccatiEquations.cpp:135-147, RiccatiTransversalityConditions.h:40-56)
      const double* jr = buf.jump + ((size_t)prob * buf.jump_capacity + (sp.jump - 1)) * jump_rec(n);
      const double* Ae = jr;
      const double* Hve = jr + jump_oHv(n);
      for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx % n, j = idx / n;
        F.Sm[idx] = (i <= j) ? y[tri_index(i, j)] : y[tri_index(j, i)];
      }
      __syncwarp();
      wgemm<true, false>(n, n, n, 1.0, F.Sm, n, Ae, n, 0.0, F.StA, n);   // SmTransAm = Sm' A_e
      wgemm<false, false>(n, 1, n, 1.0, F.Sm, n, Hve, n, 0.0, k1, n);    // SmHv
      double part = 0.0;
      for (int i = lane; i < n; i += 32) {
        part += Hve[i] * (y[ntri + i] + 0.5 * k1[i]);
        k1[n + i] = y[ntri + i] + k1[i];                                  // Sv + SmHv
      }
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      __syncwarp();
      wcopy(n * n, jr + jump_oQ(n), F.Sm);
      wgemm<true, false>(n, n, n, 1.0, F.StA, n, Ae, n, 1.0, F.Sm, n);   // Sm- = Q_e + SmTransAm' A_e
      wcopy(n, jr + jump_oq(n), k2);
      wgemm<true, false>(n, 1, n, 1.0, Ae, n, k1 + n, n, 1.0, k2, n);    // Sv- = q_e + A_e' (Sv + SmHv)
      for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx % n, j = idx / n;
        if (i <= j) y[tri_index(i, j)] = F.Sm[idx];                       // convert2Vector keeps the upper triangle
      }
      for (int i = lane; i < n; i += 32) y[ntri + i] = k2[i];
      if (lane == 0) y[ntri + n] = y[ntri + n] + jr[jump_oc(n)] + part;
      __syncwarp();
      if (sp.observe_node >= 0) {
        write_value(sp.observe_node);
        controller(sp.observe_node, node[sp.observe_node & 1]);
      }
      continue;
    }
    // classic RK4 (boost::odeint runge_kutta4)
    lerp_proj(L, full, lm, gersh, sp.alpha[0], Lh, Rh, cur);
    flow_map(L, st, cur, F, y, k1);
    for (int i = lane; i < dim; i += 32) yt[i] = y[i] + (h * 0.5) * k1[i];
    __syncwarp();
    lerp_proj(L, full, lm, gersh, sp.alpha[1], Lh, Rh, cur);
    flow_map(L, st, cur, F, yt, k2);
    for (int i = lane; i < dim; i += 32) yt[i] = y[i] + (h * 0.5) * k2[i];
    __syncwarp();
    lerp_proj(L, full, lm, gersh, sp.alpha[2], Lh, Rh, cur);
    flow_map(L, st, cur, F, yt, k3);
    for (int i = lane; i < dim; i += 32) yt[i] = y[i] + h * k3[i];
    __syncwarp();
    lerp_proj(L, full, lm, gersh, sp.alpha[3], Lh, Rh, cur);
    flow_map(L, st, cur, F, yt, k4);
    const double b1 = h * (1.0 / 6.0), b2 = h * (1.0 / 3.0);
    for (int i = lane; i < dim; i += 32) y[i] = y[i] + b1 * k1[i] + b2 * k2[i] + b2 * k3[i] + b1 * k4[i];
    __syncwarp();
    if (sp.observe_node >= 0) {
      write_value(sp.observe_node);
      controller(sp.observe_node, node[sp.observe_node & 1]);
    }
  }
  __syncwarp();
  if (N >= 1) copy_last_controller(L, solp);
  if (lane == 0) buf.status[prob] = status;
}

int slq_warp_doubles(const Layout& L, bool full, bool lm, bool gersh) {
  const int n = L.n, m = L.m;
  const int dim = n * (n + 1) / 2 + n + 1, dimp = (dim + 1) & ~1;
  int t = work_doubles(L, full, lm, gersh) + 2 * projset_doubles(L, full, lm, gersh, true) + projset_doubles(L, full, lm, gersh, false);
  t += 6 * dimp + 2 * n * n + 3 * m * n + 3 * m;
  return (t + 1) & ~1;
}

template <class Kernel>
cudaError_t configure(Kernel kernel, int warp_doubles, int& wpb, size_t& smem) {
  const size_t per_warp = (size_t)warp_doubles * sizeof(double);
  const size_t cap = 227 * 1024;
  if (per_warp > cap) return cudaErrorInvalidConfiguration;
  wpb = (int)(cap / per_warp);
  if (wpb > 4) wpb = 4;
  smem = per_warp * wpb;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

namespace {
bool std_settings(const SolverSettings& st) {
  return st.reduced && st.strategy == O2C_STRATEGY_LINE_SEARCH && st.hc == O2C_HC_DIAGONAL_SHIFT;
}
template <int CN, int CM, int CNC, bool STD>
cudaError_t run_ilqr(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles,
                     cudaStream_t stream) {
  int wpb;
  size_t smem;
  cudaError_t e = configure(ilqr_generic_kernel<CN, CM, CNC, STD>, warp_doubles, wpb, smem);
  if (e != cudaSuccess) return e;
  const int grid = (count + wpb - 1) / wpb;
  ilqr_generic_kernel<CN, CM, CNC, STD><<<grid, wpb * 32, smem, stream>>>(L, st, buf, begin, count, warp_doubles);
  return cudaGetLastError();
}
template <int CN, int CM, int CNC, bool STD>
cudaError_t run_slq(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin,
                    int count, int warp_doubles, cudaStream_t stream) {
  int wpb;
  size_t smem;
  cudaError_t e = configure(slq_generic_kernel<CN, CM, CNC, STD>, warp_doubles, wpb, smem);
  if (e != cudaSuccess) return e;
  const int grid = (count + wpb - 1) / wpb;
  slq_generic_kernel<CN, CM, CNC, STD><<<grid, wpb * 32, smem, stream>>>(L, st, buf, steps, nsteps, begin, count, warp_doubles);
  return cudaGetLastError();
}
}  // namespace

// The kernel instantiations are spread over several translation units that nvcc compiles in parallel (one instantiation with
// compile-time dimensions takes about a minute): riccati_generic_p1.cu ... p3.cu include this file with O2C_GENERIC_PART set and
// define the named-shape runners declared here; this file (part 0) keeps the run-time-dimension kernels and the public launchers.
#ifndef O2C_GENERIC_PART
#define O2C_GENERIC_PART 0
#endif
cudaError_t run_ilqr_generic_10_3_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream);
cudaError_t run_ilqr_generic_4_1_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream);
cudaError_t run_ilqr_generic_9_9_3(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream);
cudaError_t run_slq_generic_12_4_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin, int count,
                                   int warp_doubles, cudaStream_t stream);
cudaError_t run_slq_generic_any(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin, int count,
                                int warp_doubles, cudaStream_t stream);

#if O2C_GENERIC_PART == 1
cudaError_t run_ilqr_generic_10_3_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream) {
  return run_ilqr<10, 3, 0, true>(L, st, buf, begin, count, warp_doubles, stream);
}
cudaError_t run_ilqr_generic_4_1_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream) {
  return run_ilqr<4, 1, 0, true>(L, st, buf, begin, count, warp_doubles, stream);
}
#elif O2C_GENERIC_PART == 2
cudaError_t run_ilqr_generic_9_9_3(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count, int warp_doubles, cudaStream_t stream) {
  return run_ilqr<9, 9, 3, true>(L, st, buf, begin, count, warp_doubles, stream);
}
#elif O2C_GENERIC_PART == 4
cudaError_t run_slq_generic_any(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin, int count,
                                int warp_doubles, cudaStream_t stream) {
  return run_slq<0, 0, 0, false>(L, st, buf, steps, nsteps, begin, count, warp_doubles, stream);
}
#elif O2C_GENERIC_PART == 3
cudaError_t run_slq_generic_12_4_0(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps, int begin, int count,
                                   int warp_doubles, cudaStream_t stream) {
  return run_slq<12, 4, 0, true>(L, st, buf, steps, nsteps, begin, count, warp_doubles, stream);
}
#else
// which compiled variant serves this configuration (diagnostics: o2c_kernel_variant)
const char* generic_variant_name(const Layout& L, const SolverSettings& st) {
  const bool s = std_settings(st);
  if (st.algorithm == O2C_ALG_SLQ) return (s && L.n == 12 && L.m == 4 && L.ncmax == 0) ? "slq_generic_kernel<12,4,0>" : "slq_generic_kernel";
  if (s && L.n == 10 && L.m == 3 && L.ncmax == 0) return "ilqr_generic_kernel<10,3,0>";
  if (s && L.n == 9 && L.m == 9 && L.ncmax == 3) return "ilqr_generic_kernel<9,9,3>";
  if (s && L.n == 4 && L.m == 1 && L.ncmax == 0) return "ilqr_generic_kernel<4,1,0>";
  return "ilqr_generic_kernel";
}

cudaError_t launch_ilqr_generic(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, int begin, int count,
                                cudaStream_t stream) {
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  const int wd = work_doubles(L, full, lm, gersh) + 2 * L.n * L.n + 2 * L.n;
  const int warp_doubles = (wd + 1) & ~1;
  if (std_settings(st)) {  // the named BASELINE shapes get compile-time dimensions
    if (L.n == 10 && L.m == 3 && L.ncmax == 0) return run_ilqr_generic_10_3_0(L, st, buf, begin, count, warp_doubles, stream);
    if (L.n == 9 && L.m == 9 && L.ncmax == 3) return run_ilqr_generic_9_9_3(L, st, buf, begin, count, warp_doubles, stream);
    if (L.n == 4 && L.m == 1 && L.ncmax == 0) return run_ilqr_generic_4_1_0(L, st, buf, begin, count, warp_doubles, stream);
  }
  return run_ilqr<0, 0, 0, false>(L, st, buf, begin, count, warp_doubles, stream);
}

cudaError_t launch_slq_generic(const Layout& L, const SolverSettings& st, const DeviceBuffers& buf, const SlqStep* steps, int nsteps,
                               int begin, int count, cudaStream_t stream) {
  const bool full = !st.reduced, lm = st.strategy == O2C_STRATEGY_LEVENBERG_MARQUARDT;
  const bool gersh = !lm && (st.hc == O2C_HC_GERSHGORIN_MODIFICATION || st.hc == O2C_HC_EIGENVALUE_MODIFICATION);  // full dQ matrix
  const int warp_doubles = slq_warp_doubles(L, full, lm, gersh);
  if (std_settings(st) && L.n == 12 && L.m == 4 && L.ncmax == 0)
    return run_slq_generic_12_4_0(L, st, buf, steps, nsteps, begin, count, warp_doubles, stream);
  return run_slq_generic_any(L, st, buf, steps, nsteps, begin, count, warp_doubles, stream);
}

#endif  // O2C_GENERIC_PART

}  // namespace o2c
