/*
 * lq_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE). See lq_oracle.h for scope and parity pinning.
 *
 * Every function restates one reference function operation-for-operation (same intermediate quantities, same order
 * of the matrix expressions) using plain loops instead of Eigen. Citations are relative to /root/reference.
 * Build: g++ -O2 -std=c++17 -fPIC -shared -pthread -o liblq_oracle.so lq_oracle.cpp   (see oracle/Makefile)
 */
#include "lq_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

using vec = std::vector<double>;

// ---------------------------------------------------------------------------------------------------------------------
// tiny column-major dense kernels.  op(A) is M x K, op(B) is K x N, C is M x N.  beta in {0,1}.
// ---------------------------------------------------------------------------------------------------------------------
inline void gemm_nn(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
                    int ldc) {
  for (int j = 0; j < N; ++j) {
    double* c = C + (size_t)j * ldc;
    if (beta == 0.0) {
      for (int i = 0; i < M; ++i) c[i] = 0.0;
    }
    for (int k = 0; k < K; ++k) {
      const double b = alpha * B[k + (size_t)j * ldb];
      const double* a = A + (size_t)k * lda;
      for (int i = 0; i < M; ++i) c[i] += a[i] * b;
    }
  }
}
// C = beta*C + alpha * A^T * B, A is K x M
inline void gemm_tn(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
                    int ldc) {
  for (int j = 0; j < N; ++j) {
    const double* b = B + (size_t)j * ldb;
    for (int i = 0; i < M; ++i) {
      const double* a = A + (size_t)i * lda;
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += a[k] * b[k];
      double& c = C[i + (size_t)j * ldc];
      c = (beta == 0.0 ? 0.0 : c) + alpha * acc;
    }
  }
}
// C = beta*C + alpha * A * B^T, B is N x K
inline void gemm_nt(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
                    int ldc) {
  for (int j = 0; j < N; ++j) {
    double* c = C + (size_t)j * ldc;
    if (beta == 0.0) {
      for (int i = 0; i < M; ++i) c[i] = 0.0;
    }
    for (int k = 0; k < K; ++k) {
      const double b = alpha * B[j + (size_t)k * ldb];
      const double* a = A + (size_t)k * lda;
      for (int i = 0; i < M; ++i) c[i] += a[i] * b;
    }
  }
}
inline double dot(int n, const double* a, const double* b) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc += a[i] * b[i];
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Eigen::LLT (lower) restated: unblocked right-looking Cholesky. Returns false when a pivot is not positive
// (Eigen reports NumericalIssue; the reference ignores it and carries on with what was factored).
// ---------------------------------------------------------------------------------------------------------------------
bool cholesky_lower(int m, double* L, int ld) {
  bool ok = true;
  for (int j = 0; j < m; ++j) {
    double d = L[j + (size_t)j * ld];
    for (int k = 0; k < j; ++k) d -= L[j + (size_t)k * ld] * L[j + (size_t)k * ld];
    if (!(d > 0.0)) {
      ok = false;
      d = std::numeric_limits<double>::quiet_NaN();
    }
    const double ljj = std::sqrt(d);
    L[j + (size_t)j * ld] = ljj;
    for (int i = j + 1; i < m; ++i) {
      double v = L[i + (size_t)j * ld];
      for (int k = 0; k < j; ++k) v -= L[i + (size_t)k * ld] * L[j + (size_t)k * ld];
      L[i + (size_t)j * ld] = v / ljj;
    }
  }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < j; ++i) L[i + (size_t)j * ld] = 0.0;
  return ok;
}

// solve U X = I in place for upper-triangular U (m x m, ld) -> X = U^-1 (upper triangular)
void upper_inverse(int m, const double* U, int ldu, double* X, int ldx) {
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < m; ++i) X[i + (size_t)j * ldx] = (i == j) ? 1.0 : 0.0;
    for (int i = j; i >= 0; --i) {
      double v = X[i + (size_t)j * ldx];
      for (int k = i + 1; k <= j; ++k) v -= U[i + (size_t)k * ldu] * X[k + (size_t)j * ldx];
      X[i + (size_t)j * ldx] = v / U[i + (size_t)i * ldu];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Eigen::HouseholderQR restated (unblocked): M (rows x cols, rows >= cols) -> R in the upper triangle, Q formed
// explicitly (rows x rows). The sign convention (beta = -sign(x0)*|x|) follows Eigen's makeHouseholder.
// ---------------------------------------------------------------------------------------------------------------------
void householder_qr(int rows, int cols, double* M, int ld, double* Q /* rows x rows, ld rows */) {
  vec tau(cols, 0.0);
  vec v(rows);
  for (int j = 0; j < cols; ++j) {
    // makeHouseholder on M(j:rows, j)
    double tailSq = 0.0;
    for (int i = j + 1; i < rows; ++i) tailSq += M[i + (size_t)j * ld] * M[i + (size_t)j * ld];
    const double c0 = M[j + (size_t)j * ld];
    double beta;
    if (tailSq <= std::numeric_limits<double>::min()) {
      tau[j] = 0.0;
      beta = c0;
      for (int i = j + 1; i < rows; ++i) M[i + (size_t)j * ld] = 0.0;
    } else {
      beta = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0.0) beta = -beta;
      for (int i = j + 1; i < rows; ++i) M[i + (size_t)j * ld] /= (c0 - beta);
      tau[j] = (beta - c0) / beta;
    }
    M[j + (size_t)j * ld] = beta;
    // apply H = I - tau [1;v][1;v]^T to the trailing columns
    for (int cidx = j + 1; cidx < cols; ++cidx) {
      double w = M[j + (size_t)cidx * ld];
      for (int i = j + 1; i < rows; ++i) w += M[i + (size_t)j * ld] * M[i + (size_t)cidx * ld];
      w *= tau[j];
      M[j + (size_t)cidx * ld] -= w;
      for (int i = j + 1; i < rows; ++i) M[i + (size_t)cidx * ld] -= M[i + (size_t)j * ld] * w;
    }
  }
  // Q = H_0 H_1 ... H_{cols-1}: apply reflectors to the identity from the last to the first
  for (int j = 0; j < rows; ++j)
    for (int i = 0; i < rows; ++i) Q[i + (size_t)j * rows] = (i == j) ? 1.0 : 0.0;
  for (int j = cols - 1; j >= 0; --j) {
    for (int cidx = 0; cidx < rows; ++cidx) {
      double w = Q[j + (size_t)cidx * rows];
      for (int i = j + 1; i < rows; ++i) w += M[i + (size_t)j * ld] * Q[i + (size_t)cidx * rows];
      w *= tau[j];
      Q[j + (size_t)cidx * rows] -= w;
      for (int i = j + 1; i < rows; ++i) Q[i + (size_t)cidx * rows] -= M[i + (size_t)j * ld] * w;
    }
  }
}

// LinearAlgebra::makePsdGershgorin, ocs2_core/src/misc/LinearAlgebra.cpp:77-85
void make_psd_gershgorin(int n, double* M, double minEigenvalue) {
  vec T((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) T[i + (size_t)j * n] = 0.5 * (M[i + (size_t)j * n] + M[j + (size_t)i * n]);
  std::copy(T.begin(), T.end(), M);
  for (int i = 0; i < n; ++i) {
    double colAbs = 0.0;
    for (int k = 0; k < n; ++k) colAbs += std::fabs(M[k + (size_t)i * n]);
    const double Ri = colAbs - std::fabs(M[i + (size_t)i * n]);
    M[i + (size_t)i * n] = std::max(M[i + (size_t)i * n], Ri + minEigenvalue);
  }
}

// Symmetric eigen-decomposition by cyclic Jacobi rotations: A (lower triangle read, like Eigen::SelfAdjointEigenSolver) = V diag(w) V'.
// Eigen's tridiagonal-QR algorithm is not restated; V max(w, eps) V' does not depend on the algorithm beyond rounding.
void jacobi_eigh(int n, const double* Ain, double* w, double* V) {
  vec A((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A[i + (size_t)j * n] = (i >= j) ? Ain[i + (size_t)j * n] : Ain[j + (size_t)i * n];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) V[i + (size_t)j * n] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += A[i + (size_t)j * n] * A[i + (size_t)j * n];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + (size_t)q * n];
        if (apq == 0.0) continue;
        const double theta = (A[q + (size_t)q * n] - A[p + (size_t)p * n]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {  // A <- A J (columns p, q)
          const double akp = A[k + (size_t)p * n], akq = A[k + (size_t)q * n];
          A[k + (size_t)p * n] = c * akp - sn * akq;
          A[k + (size_t)q * n] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // A <- J' A (rows p, q)
          const double apk = A[p + (size_t)k * n], aqk = A[q + (size_t)k * n];
          A[p + (size_t)k * n] = c * apk - sn * aqk;
          A[q + (size_t)k * n] = sn * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {  // V <- V J
          const double vkp = V[k + (size_t)p * n], vkq = V[k + (size_t)q * n];
          V[k + (size_t)p * n] = c * vkp - sn * vkq;
          V[k + (size_t)q * n] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i + (size_t)i * n];
}

// LinearAlgebra::makePsdEigenvalue, ocs2_core/src/misc/LinearAlgebra.cpp:52-72
void make_psd_eigenvalue(int n, double* M, double minEigenvalue) {
  vec w(n), V((size_t)n * n);
  jacobi_eigh(n, M, w.data(), V.data());
  bool hasNegativeEigenValue = false;
  for (int j = 0; j < n; ++j)
    if (w[j] < minEigenvalue) {
      hasNegativeEigenValue = true;
      w[j] = minEigenvalue;
    }
  if (hasNegativeEigenValue) {  // V diag(lambda) V^-1 with V orthogonal
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += V[i + (size_t)k * n] * w[k] * V[j + (size_t)k * n];
        M[i + (size_t)j * n] = acc;
      }
  } else {
    vec T((size_t)n * n);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) T[i + (size_t)j * n] = 0.5 * (M[i + (size_t)j * n] + M[j + (size_t)i * n]);
    std::copy(T.begin(), T.end(), M);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// projected stage storage
// ---------------------------------------------------------------------------------------------------------------------
struct Projected {
  int n = 0, m = 0, p = 0, nc = 0;
  vec At, Bt, Hvt, Qt, Pt, Rt, qt, rt, Cmt, Evt, Pu, Ddag, Hm, dQ, dGm, dGv;
  double ct = 0.0;
  void resize(int n_, int m_) {
    n = n_;
    m = m_;
    At.resize((size_t)n * n);
    Bt.resize((size_t)n * m);
    Hvt.resize(n);
    Qt.resize((size_t)n * n);
    Pt.resize((size_t)m * n);
    Rt.resize((size_t)m * m);
    qt.resize(n);
    rt.resize(m);
    Cmt.resize((size_t)m * n);
    Evt.resize(m);
    Pu.resize((size_t)m * m);
    Ddag.resize((size_t)m * std::max(m, 1));
    Hm.resize((size_t)m * m);
    dQ.resize((size_t)n * n);
    dGm.resize((size_t)m * n);
    dGv.resize(m);
  }
};

struct StageIn {
  int n, m, nc, ldc;
  const double *A, *B, *Hv, *Q, *P, *R, *q, *r;
  double c;
  const double *C, *D, *e;
};

// GaussNewtonDDP::computeProjectionAndRiccatiModification, ocs2_ddp/src/GaussNewtonDDP.cpp:734-750
int project_stage(const orc_settings& st, const StageIn& in, const double* Sm, Projected& pr) {
  const int n = in.n, m = in.m, nc = in.nc, p = m - nc;
  int status = ORC_STATUS_OK;
  pr.p = p;
  pr.nc = nc;

  // ---- Hamiltonian Hessian: ILQR::computeHamiltonianHessian (ILQR.cpp:217-222) / SLQ (SLQ.cpp:206-208) ----
  vec& Hm = pr.Hm;
  std::copy(in.R, in.R + (size_t)m * m, Hm.begin());
  if (Sm != nullptr) {
    vec BtS((size_t)m * n);
    gemm_tn(m, n, n, 1.0, in.B, n, Sm, n, 0.0, BtS.data(), m);   // BmTransSm = B^T Sm
    gemm_nn(m, m, n, 1.0, BtS.data(), m, in.B, n, 1.0, Hm.data(), m);  // Hm += BmTransSm * B
  }
  if (st.strategy == ORC_STRATEGY_LM) {
    // LevenbergMarquardtStrategy::augmentHamiltonianHessian (LevenbergMarquardtStrategy.cpp:245-249)
    gemm_tn(m, m, n, st.lm_riccati_multiple, in.B, n, in.B, n, 1.0, Hm.data(), m);
  }

  // ---- projectors: GaussNewtonDDP::computeProjections (GaussNewtonDDP.cpp:755-782) ----
  vec Ui((size_t)m * m);
  if (orc_inverse_uut(m, Hm.data(), Ui.data()) != 0) status |= ORC_STATUS_CHOL_NOT_PD;
  if (nc == 0) {
    std::copy(Ui.begin(), Ui.end(), pr.Pu.begin());  // constraintNullProjector = HmInvUmUmT
  } else {
    vec RcInv((size_t)nc * nc);
    orc_constraint_projection(m, nc, in.D, in.ldc, Ui.data(), pr.Ddag.data(), RcInv.data(), pr.Pu.data());
  }
  const double* Pu = pr.Pu.data();  // m x p

  // ---- projectLQ (DDP_HelperFunctions.cpp:143-201) + changeOfInputVariables (ChangeOfInputVariables.cpp:34-108) ----
  vec Px, u0;  // Px = -Ddag*C (m x n), u0 = -Ddag*e (m)
  const bool hasPx = nc > 0;
  if (hasPx) {
    gemm_nn(m, 1, nc, 1.0, pr.Ddag.data(), m, in.e, nc, 0.0, pr.Evt.data(), m);        // EvProjected = Ddag * e
    gemm_nn(m, n, nc, 1.0, pr.Ddag.data(), m, in.C, in.ldc, 0.0, pr.Cmt.data(), m);    // CmProjected = Ddag * C
    Px.resize((size_t)m * n);
    u0.resize(m);
    for (size_t i = 0; i < Px.size(); ++i) Px[i] = -pr.Cmt[i];
    for (int i = 0; i < m; ++i) u0[i] = -pr.Evt[i];
  } else {
    std::fill(pr.Evt.begin(), pr.Evt.end(), 0.0);
    std::fill(pr.Cmt.begin(), pr.Cmt.end(), 0.0);
  }

  // dynamics: A~ = A + B Px ; B~ = B Pu ; Hv~ = Hv + B u0   (ChangeOfInputVariables.cpp:91-108, DDP_HelperFunctions.cpp:194-195)
  std::copy(in.A, in.A + (size_t)n * n, pr.At.begin());
  if (hasPx) gemm_nn(n, n, m, 1.0, in.B, n, Px.data(), m, 1.0, pr.At.data(), n);
  gemm_nn(n, p, m, 1.0, in.B, n, Pu, m, 0.0, pr.Bt.data(), n);
  std::copy(in.Hv, in.Hv + n, pr.Hvt.begin());
  if (hasPx) gemm_nn(n, 1, m, 1.0, in.B, n, u0.data(), m, 1.0, pr.Hvt.data(), n);

  // cost
  vec P_plus_R_Px(in.P, in.P + (size_t)m * n);  // shared term 1
  if (hasPx) gemm_nn(m, n, m, 1.0, in.R, m, Px.data(), m, 1.0, P_plus_R_Px.data(), m);
  vec r_plus_R_u0(in.r, in.r + m);  // shared term 2
  if (hasPx) gemm_nn(m, 1, m, 1.0, in.R, m, u0.data(), m, 1.0, r_plus_R_u0.data(), m);
  std::copy(in.Q, in.Q + (size_t)n * n, pr.Qt.begin());
  std::copy(in.q, in.q + n, pr.qt.begin());
  pr.ct = in.c;
  if (hasPx) {
    gemm_tn(n, n, m, 1.0, in.P, m, Px.data(), m, 1.0, pr.Qt.data(), n);            // Q += P^T Px
    gemm_tn(n, n, m, 1.0, Px.data(), m, P_plus_R_Px.data(), m, 1.0, pr.Qt.data(), n);  // Q += Px^T (P + R Px)
    gemm_tn(n, 1, m, 1.0, in.P, m, u0.data(), m, 1.0, pr.qt.data(), n);            // q += P^T u0
    gemm_tn(n, 1, m, 1.0, Px.data(), m, r_plus_R_u0.data(), m, 1.0, pr.qt.data(), n);  // q += Px^T (r + R u0)
    double acc = 0.0;
    for (int i = 0; i < m; ++i) acc += u0[i] * (r_plus_R_u0[i] + in.r[i]);
    pr.ct += 0.5 * acc;  // c += 1/2 u0^T ((R u0 + r) + r)
  }
  gemm_tn(p, n, m, 1.0, Pu, m, P_plus_R_Px.data(), m, 0.0, pr.Pt.data(), p);  // P~ = Pu^T (P + R Px)
  {
    vec R_Pu((size_t)m * p);
    gemm_nn(m, p, m, 1.0, in.R, m, Pu, m, 0.0, R_Pu.data(), m);
    gemm_tn(p, p, m, 1.0, Pu, m, R_Pu.data(), m, 0.0, pr.Rt.data(), p);  // R~ = Pu^T R Pu
  }
  gemm_tn(p, 1, m, 1.0, Pu, m, r_plus_R_u0.data(), m, 0.0, pr.rt.data(), p);  // r~ = Pu^T (r + R u0)

  // ---- Riccati modification ----
  if (st.strategy == ORC_STRATEGY_LINE_SEARCH) {
    // LineSearchStrategy::computeRiccatiModification (LineSearchStrategy.cpp:294-312)
    vec Mq(pr.Qt.begin(), pr.Qt.begin() + (size_t)n * n);  // Q_minus_PTRinvP
    gemm_tn(n, n, p, -1.0, pr.Pt.data(), p, pr.Pt.data(), p, 1.0, Mq.data(), n);
    std::copy(Mq.begin(), Mq.end(), pr.dQ.begin());
    if (orc_shift_hessian(st.hessian_correction, n, pr.dQ.data(), st.hessian_multiple) != 0) status |= ORC_STATUS_NONFINITE;
    for (size_t i = 0; i < (size_t)n * n; ++i) pr.dQ[i] -= Mq[i];
    std::fill(pr.dGv.begin(), pr.dGv.end(), 0.0);
    std::fill(pr.dGm.begin(), pr.dGm.end(), 0.0);
  } else {
    // LevenbergMarquardtStrategy::computeRiccatiModification (LevenbergMarquardtStrategy.cpp:230-240)
    std::fill(pr.dQ.begin(), pr.dQ.end(), 0.0);
    gemm_tn(p, 1, n, st.lm_riccati_multiple, pr.Bt.data(), n, pr.Hvt.data(), n, 0.0, pr.dGv.data(), p);
    gemm_tn(p, n, n, st.lm_riccati_multiple, pr.Bt.data(), n, pr.At.data(), n, 0.0, pr.dGm.data(), p);
  }
  return status;
}

struct ProjPtrs {
  const double *At, *Bt, *Hvt, *Qt, *Pt, *Rt, *qt, *rt, *dQ, *dGm, *dGv;
  double ct;
};
ProjPtrs ptrs_of(const Projected& pr) {
  return {pr.At.data(), pr.Bt.data(), pr.Hvt.data(), pr.Qt.data(), pr.Pt.data(), pr.Rt.data(),
          pr.qt.data(), pr.rt.data(), pr.dQ.data(),  pr.dGm.data(), pr.dGv.data(), pr.ct};
}

// DiscreteTimeRiccatiEquations::computeMapILQR, ocs2_ddp/src/riccati_equations/DiscreteTimeRiccatiEquations.cpp:65-154
void compute_map(bool reduced, int n, int p, const ProjPtrs& pr, const double* SmNext, const double* SvNext, double sNext, double* Km,
                 double* Lv, double* Sm, double* Sv, double* s) {
  vec SmHv(n), SmAm((size_t)n * n), SmBm((size_t)n * p), SvPlus(n), Gm((size_t)p * n), Gv(p), KtG((size_t)n * n);
  gemm_nn(n, 1, n, 1.0, SmNext, n, pr.Hvt, n, 0.0, SmHv.data(), n);   // Sm_projectedHv
  gemm_nn(n, n, n, 1.0, SmNext, n, pr.At, n, 0.0, SmAm.data(), n);    // Sm_projectedAm
  gemm_nn(n, p, n, 1.0, SmNext, n, pr.Bt, n, 0.0, SmBm.data(), n);    // Sm_projectedBm
  for (int i = 0; i < n; ++i) SvPlus[i] = SvNext[i] + SmHv[i];         // Sv_plus_Sm_projectedHv
  std::copy(pr.Pt, pr.Pt + (size_t)p * n, Gm.begin());
  gemm_tn(p, n, n, 1.0, pr.Bt, n, SmAm.data(), n, 1.0, Gm.data(), p);  // projectedGm = Pm + Bm^T Sm Am
  std::copy(pr.rt, pr.rt + p, Gv.begin());
  gemm_tn(p, 1, n, 1.0, pr.Bt, n, SvPlus.data(), n, 1.0, Gv.data(), p);  // projectedGv = Rv + Bm^T (Sv + Sm Hv)
  for (size_t i = 0; i < (size_t)p * n; ++i) Km[i] = -Gm[i] - pr.dGm[i];
  for (int i = 0; i < p; ++i) Lv[i] = -Gv[i] - pr.dGv[i];
  gemm_tn(n, n, p, 1.0, Km, p, Gm.data(), p, 0.0, KtG.data(), n);  // projectedKm_T_projectedGm
  vec Hm, HmKm, HmLv;
  if (!reduced) {
    Hm.assign(pr.Rt, pr.Rt + (size_t)p * p);
    gemm_tn(p, p, n, 1.0, SmBm.data(), n, pr.Bt, n, 1.0, Hm.data(), p);  // projectedHm = Rm + (Sm Bm)^T Bm
    HmKm.resize((size_t)p * n);
    HmLv.resize(p);
    gemm_nn(p, n, p, 1.0, Hm.data(), p, Km, p, 0.0, HmKm.data(), p);
    gemm_nn(p, 1, p, 1.0, Hm.data(), p, Lv, p, 0.0, HmLv.data(), p);
  }
  // Sm
  for (size_t i = 0; i < (size_t)n * n; ++i) Sm[i] = pr.Qt[i] + pr.dQ[i];
  gemm_tn(n, n, n, 1.0, SmAm.data(), n, pr.At, n, 1.0, Sm, n);  // += (Sm Am)^T Am
  if (reduced) {
    for (size_t i = 0; i < (size_t)n * n; ++i) Sm[i] += KtG[i];
  } else {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) Sm[i + (size_t)j * n] += KtG[i + (size_t)j * n] + KtG[j + (size_t)i * n];
    gemm_tn(n, n, p, 1.0, Km, p, HmKm.data(), p, 1.0, Sm, n);
  }
  // Sv
  std::copy(pr.qt, pr.qt + n, Sv);
  gemm_tn(n, 1, n, 1.0, pr.At, n, SvPlus.data(), n, 1.0, Sv, n);
  gemm_tn(n, 1, p, 1.0, Gm.data(), p, Lv, p, 1.0, Sv, n);
  if (!reduced) {
    gemm_tn(n, 1, p, 1.0, Km, p, Gv.data(), p, 1.0, Sv, n);
    gemm_tn(n, 1, p, 1.0, HmKm.data(), p, Lv, p, 1.0, Sv, n);
  }
  // s
  double sv = sNext + pr.ct;
  sv += dot(n, pr.Hvt, SvPlus.data());
  sv -= 0.5 * dot(n, pr.Hvt, SmHv.data());
  if (reduced) {
    sv += 0.5 * dot(p, Lv, Gv.data());
  } else {
    sv += dot(p, Lv, Gv.data());
    sv += 0.5 * dot(p, Lv, HmLv.data());
  }
  *s = sv;
}

// ContinuousTimeRiccatiEquations::computeFlowMapSLQ on interpolated data, ContinuousTimeRiccatiEquations.cpp:170-292
void flow_map_slq(bool reduced, int n, int p, const ProjPtrs& pr, const double* allSs, double* dallSs) {
  vec Sm((size_t)n * n), Sv(n), dSm((size_t)n * n), dSv(n);
  double s, ds;
  orc_unflatten(n, allSs, Sm.data(), Sv.data(), &s);
  ds = pr.ct;
  std::copy(pr.qt, pr.qt + n, dSv.begin());
  std::copy(pr.Qt, pr.Qt + (size_t)n * n, dSm.begin());
  vec Gv(pr.rt, pr.rt + p), Gm(pr.Pt, pr.Pt + (size_t)p * n), Km((size_t)p * n), Lv(p);
  gemm_tn(p, n, n, 1.0, pr.Bt, n, Sm.data(), n, 1.0, Gm.data(), p);  // Gm = Pm + Bm^T Sm
  gemm_tn(p, 1, n, 1.0, pr.Bt, n, Sv.data(), n, 1.0, Gv.data(), p);  // Gv = Rv + Bm^T Sv
  for (size_t i = 0; i < (size_t)p * n; ++i) Km[i] = -(Gm[i] + pr.dGm[i]);
  for (int i = 0; i < p; ++i) Lv[i] = -(Gv[i] + pr.dGv[i]);
  vec StA((size_t)n * n), KtG((size_t)n * n), RmKm, RmLv;
  gemm_tn(n, n, n, 1.0, Sm.data(), n, pr.At, n, 0.0, StA.data(), n);  // SmTrans_projectedAm
  gemm_tn(n, n, p, 1.0, Km.data(), p, Gm.data(), p, 0.0, KtG.data(), n);
  if (!reduced) {
    RmKm.resize((size_t)p * n);
    RmLv.resize(p);
    gemm_nn(p, n, p, 1.0, pr.Rt, p, Km.data(), p, 0.0, RmKm.data(), p);
    gemm_nn(p, 1, p, 1.0, pr.Rt, p, Lv.data(), p, 0.0, RmLv.data(), p);
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) dSm[i + (size_t)j * n] += pr.dQ[i + (size_t)j * n] + StA[i + (size_t)j * n] + StA[j + (size_t)i * n];
  if (reduced) {
    for (size_t i = 0; i < (size_t)n * n; ++i) dSm[i] += KtG[i];
  } else {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) dSm[i + (size_t)j * n] += KtG[i + (size_t)j * n] + KtG[j + (size_t)i * n];
    gemm_tn(n, n, p, 1.0, Km.data(), p, RmKm.data(), p, 1.0, dSm.data(), n);
  }
  gemm_tn(n, 1, n, 1.0, Sm.data(), n, pr.Hvt, n, 1.0, dSv.data(), n);
  gemm_tn(n, 1, n, 1.0, pr.At, n, Sv.data(), n, 1.0, dSv.data(), n);
  gemm_tn(n, 1, p, 1.0, Gm.data(), p, Lv.data(), p, 1.0, dSv.data(), n);
  if (!reduced) {
    gemm_tn(n, 1, p, 1.0, Km.data(), p, Gv.data(), p, 1.0, dSv.data(), n);
    gemm_tn(n, 1, p, 1.0, RmKm.data(), p, Lv.data(), p, 1.0, dSv.data(), n);
  }
  ds += dot(n, pr.Hvt, Sv.data());
  if (reduced) {
    ds += 0.5 * dot(p, Lv.data(), Gv.data());
  } else {
    ds += dot(p, Lv.data(), Gv.data());
    ds += 0.5 * dot(p, Lv.data(), RmLv.data());
  }
  orc_flatten(n, dSm.data(), dSv.data(), ds, dallSs);
}

StageIn stage_of(const orc_problem& pb, int k) {
  const int n = pb.nx, m = pb.nu, ncm = pb.nc_max;
  StageIn in;
  in.n = n;
  in.m = m;
  in.nc = (ncm > 0) ? (pb.nc ? pb.nc[k] : ncm) : 0;
  in.ldc = std::max(ncm, 1);
  in.A = pb.A + (size_t)k * n * n;
  in.B = pb.B + (size_t)k * n * m;
  in.Hv = pb.Hv + (size_t)k * n;
  in.Q = pb.Q + (size_t)k * n * n;
  in.P = pb.P + (size_t)k * m * n;
  in.R = pb.R + (size_t)k * m * m;
  in.q = pb.q + (size_t)k * n;
  in.r = pb.r + (size_t)k * m;
  in.c = pb.c[k];
  in.C = ncm > 0 ? pb.C + (size_t)k * ncm * n : nullptr;
  in.D = ncm > 0 ? pb.D + (size_t)k * ncm * m : nullptr;
  in.e = ncm > 0 ? pb.e + (size_t)k * ncm : nullptr;
  return in;
}

// ILQR/SLQ::calculateControllerWorker tail (ILQR.cpp:170-180, SLQ.cpp:160-168): un-project K~, L~ into K, bias, dbias
void unproject_controller(const orc_problem& pb, int k, const Projected& pr, const double* Km, const double* Lv, orc_solution* sol) {
  const int n = pb.nx, m = pb.nu, p = pr.p;
  double* K = sol->K + (size_t)k * m * n;
  double* db = sol->dbias + (size_t)k * m;
  double* bias = sol->bias + (size_t)k * m;
  for (size_t i = 0; i < (size_t)m * n; ++i) K[i] = -pr.Cmt[i];
  gemm_nn(m, n, p, 1.0, pr.Pu.data(), m, Km, p, 1.0, K, m);  // gain = -CmProjected + Qu * projectedKm
  for (int i = 0; i < m; ++i) bias[i] = pb.u_nom ? pb.u_nom[(size_t)k * m + i] : 0.0;
  if (pb.x_nom) gemm_nn(m, 1, n, -1.0, K, m, pb.x_nom + (size_t)k * n, n, 1.0, bias, m);  // bias = u_nom - K x_nom
  for (int i = 0; i < m; ++i) db[i] = -pr.Evt[i];
  gemm_nn(m, 1, p, 1.0, pr.Pu.data(), m, Lv, p, 1.0, db, m);  // deltaBias = -EvProjected + Qu * projectedLv
}

void copy_last_controller(const orc_problem& pb, orc_solution* sol) {
  // GaussNewtonDDP::calculateController, GaussNewtonDDP.cpp:609-618: node N := node N-1 (final time is not an event)
  const int n = pb.nx, m = pb.nu, N = pb.N;
  if (N < 1) return;
  std::copy(sol->K + (size_t)(N - 1) * m * n, sol->K + (size_t)N * m * n, sol->K + (size_t)N * m * n);
  std::copy(sol->bias + (size_t)(N - 1) * m, sol->bias + (size_t)N * m, sol->bias + (size_t)N * m);
  std::copy(sol->dbias + (size_t)(N - 1) * m, sol->dbias + (size_t)N * m, sol->dbias + (size_t)N * m);
}

bool all_finite(const double* v, size_t count) {
  for (size_t i = 0; i < count; ++i)
    if (!std::isfinite(v[i])) return false;
  return true;
}

// ILQR::solveSequentialRiccatiEquations + riccatiEquationsWorker (ILQR.cpp:186-299, single partition, no events) and
// ILQR::calculateControllerWorker (ILQR.cpp:162-181)
int backward_ilqr(const orc_settings& st, const orc_problem& pb, orc_solution* sol) {
  const int n = pb.nx, m = pb.nu, N = pb.N;
  int status = ORC_STATUS_OK;
  // valueFunctionTrajectory.back() = finalValueFunction (GaussNewtonDDP.cpp:526)
  std::copy(pb.Qf, pb.Qf + (size_t)n * n, sol->Sm + (size_t)N * n * n);
  std::copy(pb.qf, pb.qf + n, sol->Sv + (size_t)N * n);
  sol->s[N] = pb.cf[0];
  // The final-node projection (ILQR.cpp:200-209) only produces a controller entry that calculateController overwrites
  // with the copy of node N-1 (GaussNewtonDDP.cpp:609-618); it is therefore not restated.
  Projected pr;
  pr.resize(n, m);
  vec Km((size_t)m * n), Lv(m);
  for (int k = N - 1; k >= 0; --k) {
    const StageIn in = stage_of(pb, k);
    const double* SmNext = sol->Sm + (size_t)(k + 1) * n * n;
    if (pb.event && pb.event[k]) {
      // pre-event node (ILQR.cpp:263-295): value function by riccatiTransversalityConditions (RiccatiTransversalityConditions.h:40-56)
      // from the jump model data, controller from the regular model data projected with Sm = 0 and the pre-event value function
      const double* SvNext = sol->Sv + (size_t)(k + 1) * n;
      double* Sm = sol->Sm + (size_t)k * n * n;
      double* Sv = sol->Sv + (size_t)k * n;
      vec SmTransAm((size_t)n * n), SmHv(n), tmp(n);
      gemm_tn(n, n, n, 1.0, SmNext, n, in.A, n, 0.0, SmTransAm.data(), n);  // Sm^T * dfdx
      std::copy(in.Q, in.Q + (size_t)n * n, Sm);
      gemm_tn(n, n, n, 1.0, SmTransAm.data(), n, in.A, n, 1.0, Sm, n);      // += SmTransAm^T * dfdx
      gemm_nn(n, 1, n, 1.0, SmNext, n, in.Hv, n, 0.0, SmHv.data(), n);
      for (int i = 0; i < n; ++i) tmp[i] = SvNext[i] + SmHv[i];
      std::copy(in.q, in.q + n, Sv);
      gemm_tn(n, 1, n, 1.0, in.A, n, tmp.data(), n, 1.0, Sv, n);
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc += in.Hv[i] * (SvNext[i] + 0.5 * SmHv[i]);
      sol->s[k] = sol->s[k + 1] + in.c + acc;
      status |= project_stage(st, in, nullptr, pr);  // SmDummy = 0 (ILQR.cpp:281-282): Hm = R
      const int p = pr.p;
      for (size_t i = 0; i < (size_t)p * n; ++i) Km[i] = -pr.Pt[i] - pr.dGm[i];
      gemm_tn(p, n, n, -1.0, pr.Bt.data(), n, Sm, n, 1.0, Km.data(), p);
      for (int i = 0; i < p; ++i) Lv[i] = -pr.rt[i] - pr.dGv[i];
      gemm_tn(p, 1, n, -1.0, pr.Bt.data(), n, Sv, n, 1.0, Lv.data(), p);
      unproject_controller(pb, k, pr, Km.data(), Lv.data(), sol);
      continue;
    }
    status |= project_stage(st, in, SmNext, pr);
    compute_map(st.reduced_form != 0, n, pr.p, ptrs_of(pr), SmNext, sol->Sv + (size_t)(k + 1) * n, sol->s[k + 1], Km.data(), Lv.data(),
                sol->Sm + (size_t)k * n * n, sol->Sv + (size_t)k * n, &sol->s[k]);
    unproject_controller(pb, k, pr, Km.data(), Lv.data(), sol);
  }
  copy_last_controller(pb, sol);
  if (!all_finite(sol->K, (size_t)(N + 1) * m * n) || !all_finite(sol->dbias, (size_t)(N + 1) * m) ||
      !all_finite(sol->Sm, (size_t)(N + 1) * n * n))
    status |= ORC_STATUS_NONFINITE;
  return status;
}

// lerp of every projected field, LinearInterpolation::interpolate (implementation/LinearInterpolation.h:128-146):
// alpha * lhs + (1 - alpha) * rhs. When adjacent nodes differ in projected input dimension the reference picks the nearer
// node (areSameSize false).
void lerp_vec(double alpha, const vec& a, const vec& b, size_t count, vec& out) {
  out.resize(count);
  for (size_t i = 0; i < count; ++i) out[i] = alpha * a[i] + (1.0 - alpha) * b[i];
}
void interpolate_projected(int index, double alpha, const std::vector<Projected>& traj, Projected& out) {
  const Projected& L = traj[index];
  const Projected& Rr = traj[index + 1];
  const int n = L.n;
  if (L.p != Rr.p) {
    // state-sized fields still interpolate; input-sized fields take the nearer node
    const Projected& pick = (alpha > 0.5) ? L : Rr;
    out = pick;
    lerp_vec(alpha, L.At, Rr.At, (size_t)n * n, out.At);
    lerp_vec(alpha, L.Hvt, Rr.Hvt, n, out.Hvt);
    lerp_vec(alpha, L.Qt, Rr.Qt, (size_t)n * n, out.Qt);
    lerp_vec(alpha, L.qt, Rr.qt, n, out.qt);
    lerp_vec(alpha, L.dQ, Rr.dQ, (size_t)n * n, out.dQ);
    out.ct = alpha * L.ct + (1.0 - alpha) * Rr.ct;
    return;
  }
  const int p = L.p;
  out.n = n;
  out.m = L.m;
  out.p = p;
  out.nc = L.nc;
  lerp_vec(alpha, L.Hvt, Rr.Hvt, n, out.Hvt);
  lerp_vec(alpha, L.At, Rr.At, (size_t)n * n, out.At);
  lerp_vec(alpha, L.Bt, Rr.Bt, (size_t)n * p, out.Bt);
  out.ct = alpha * L.ct + (1.0 - alpha) * Rr.ct;
  lerp_vec(alpha, L.qt, Rr.qt, n, out.qt);
  lerp_vec(alpha, L.Qt, Rr.Qt, (size_t)n * n, out.Qt);
  lerp_vec(alpha, L.rt, Rr.rt, p, out.rt);
  lerp_vec(alpha, L.Pt, Rr.Pt, (size_t)p * n, out.Pt);
  lerp_vec(alpha, L.Rt, Rr.Rt, (size_t)p * p, out.Rt);
  lerp_vec(alpha, L.dQ, Rr.dQ, (size_t)n * n, out.dQ);
  lerp_vec(alpha, L.dGm, Rr.dGm, (size_t)p * n, out.dGm);
  lerp_vec(alpha, L.dGv, Rr.dGv, p, out.dGv);
}

// classic RK4 as boost::numeric::odeint::runge_kutta4 (explicit_generic_rk, c = {0, 1/2, 1/2, 1}, b = {1/6, 1/3, 1/3, 1/6});
// stepper typedef ocs2_core/include/ocs2_core/integration/steppers.h:52-53
template <class F>
void rk4_step(F&& f, vec& y, double t, double h, vec& k1, vec& k2, vec& k3, vec& k4, vec& tmp) {
  const size_t d = y.size();
  f(t, y, k1);
  for (size_t i = 0; i < d; ++i) tmp[i] = y[i] + (h * 0.5) * k1[i];
  f(t + h * 0.5, tmp, k2);
  for (size_t i = 0; i < d; ++i) tmp[i] = y[i] + (h * 0.5) * k2[i];
  f(t + h * 0.5, tmp, k3);
  for (size_t i = 0; i < d; ++i) tmp[i] = y[i] + h * k3[i];
  f(t + h, tmp, k4);
  const double b1 = h * (1.0 / 6.0), b2 = h * (1.0 / 3.0);
  for (size_t i = 0; i < d; ++i) y[i] = y[i] + b1 * k1[i] + b2 * k2[i] + b2 * k3[i] + b1 * k4[i];
}

// boost::numeric::odeint::detail::less_with_sign / less_eq_with_sign for dt > 0
inline bool less_with_sign(double t1, double t2) { return (t2 - t1) > std::numeric_limits<double>::epsilon(); }
inline bool less_eq_with_sign(double t1, double t2) { return (t1 - t2) <= std::numeric_limits<double>::epsilon(); }

// SLQ::solveSequentialRiccatiEquations / riccatiEquationsWorker / integrateRiccatiEquationNominalTime (SLQ.cpp:174-302, events
// included) and SLQ::calculateControllerWorker (SLQ.cpp:127-169)
int backward_slq(const orc_settings& st, const orc_problem& pb, orc_solution* sol) {
  const int n = pb.nx, m = pb.nu, N = pb.N;
  int status = ORC_STATUS_OK;
  // per-node projection with Hm = R (SLQ.cpp:183-199, 206-208)
  std::vector<Projected> traj(N + 1);
  for (int k = 0; k <= N; ++k) {
    traj[k].resize(n, m);
    status |= project_stage(st, stage_of(pb, k), nullptr, traj[k]);
  }
  std::copy(pb.Qf, pb.Qf + (size_t)n * n, sol->Sm + (size_t)N * n * n);
  std::copy(pb.qf, pb.qf + n, sol->Sv + (size_t)N * n);
  sol->s[N] = pb.cf[0];

  // normalised time z = -t, reversed (retrieveActiveNormalizedTime, DDP_HelperFunctions.cpp:309-328)
  vec z(N + 1);
  for (int j = 0; j <= N; ++j) z[j] = -pb.time[N - j];
  const size_t dim = (size_t)n * (n + 1) / 2 + n + 1;
  vec y(dim), k1(dim), k2(dim), k3(dim), k4(dim), tmp(dim);
  orc_flatten(n, pb.Qf, pb.qf, pb.cf[0], y.data());
  Projected interp;
  interp.resize(n, m);
  const bool reduced = st.reduced_form != 0;
  auto flow = [&](double zz, const vec& yy, vec& dy) {
    // ContinuousTimeRiccatiEquations::computeFlowMap, ContinuousTimeRiccatiEquations.cpp:152-170
    int index;
    double alpha;
    orc_time_segment(-zz, pb.time, N + 1, &index, &alpha);
    interpolate_projected(index, alpha, traj, interp);
    flow_map_slq(reduced, n, interp.p, ptrs_of(interp), yy.data(), dy.data());
  };
  // boost::numeric::odeint::integrate_times(stepper, sys, x, times_begin, times_end, dt, observer) with a plain stepper
  // (call site ocs2_core/include/ocs2_core/integration/implementation/Integrator.h:298-311)
  const double dt = st.time_step;
  for (int j = 0;; ++j) {
    double current_time = z[j];
    if (j > 0) {  // observer: allSsTrajectory[j] belongs to node N - j (SLQ.cpp:247-250); node N keeps the terminal value
      orc_unflatten(n, y.data(), sol->Sm + (size_t)(N - j) * n * n, sol->Sv + (size_t)(N - j) * n, &sol->s[N - j]);
    }
    if (j == N) break;
    const int lower = N - 1 - j;  // the interval [node lower, node lower + 1]
    if (pb.event && pb.event[lower]) {
      // node `lower` is a pre-event node, node lower + 1 the post-event node: the segments are integrated separately and joined by
      // ContinuousTimeRiccatiEquations::computeJumpMap (SLQ.cpp:286-296, ContinuousTimeRiccatiEquations.cpp:135-147) =
      // riccatiTransversalityConditions on modelDataEventTimes[i] (RiccatiTransversalityConditions.h:40-56)
      int ord = 0;
      for (int q = 0; q < lower; ++q) ord += pb.event[q] != 0;
      const double* Ae = pb.jA + (size_t)ord * n * n;
      const double* Hve = pb.jHv + (size_t)ord * n;
      vec Sp((size_t)n * n), Svp(n), Sm((size_t)n * n), Sv(n), SmTransAm((size_t)n * n), SmHv(n), tmpv(n);
      double sp;
      orc_unflatten(n, y.data(), Sp.data(), Svp.data(), &sp);
      gemm_tn(n, n, n, 1.0, Sp.data(), n, Ae, n, 0.0, SmTransAm.data(), n);
      std::copy(pb.jQ + (size_t)ord * n * n, pb.jQ + (size_t)(ord + 1) * n * n, Sm.begin());
      gemm_tn(n, n, n, 1.0, SmTransAm.data(), n, Ae, n, 1.0, Sm.data(), n);
      gemm_nn(n, 1, n, 1.0, Sp.data(), n, Hve, n, 0.0, SmHv.data(), n);
      for (int i = 0; i < n; ++i) tmpv[i] = Svp[i] + SmHv[i];
      std::copy(pb.jq + (size_t)ord * n, pb.jq + (size_t)(ord + 1) * n, Sv.begin());
      gemm_tn(n, 1, n, 1.0, Ae, n, tmpv.data(), n, 1.0, Sv.data(), n);
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc += Hve[i] * (Svp[i] + 0.5 * SmHv[i]);
      orc_flatten(n, Sm.data(), Sv.data(), sp + pb.jc[ord] + acc, y.data());
      continue;
    }
    double current_dt = dt;
    while (less_with_sign(current_time, z[j + 1])) {
      current_dt = std::min(dt, z[j + 1] - current_time);
      rk4_step(flow, y, current_time, current_dt, k1, k2, k3, k4, tmp);
      current_time += current_dt;
      current_dt = std::max(dt, current_dt);
    }
  }
  // controller (SLQ.cpp:127-169)
  vec Km((size_t)m * n), Lv(m);
  for (int k = 0; k <= N; ++k) {
    const Projected& pr = traj[k];
    const int p = pr.p;
    for (size_t i = 0; i < (size_t)p * n; ++i) Km[i] = -(pr.dGm[i] + pr.Pt[i]);
    gemm_tn(p, n, n, -1.0, pr.Bt.data(), n, sol->Sm + (size_t)k * n * n, n, 1.0, Km.data(), p);
    for (int i = 0; i < p; ++i) Lv[i] = -(pr.dGv[i] + pr.rt[i]);
    gemm_tn(p, 1, n, -1.0, pr.Bt.data(), n, sol->Sv + (size_t)k * n, n, 1.0, Lv.data(), p);
    unproject_controller(pb, k, pr, Km.data(), Lv.data(), sol);
  }
  copy_last_controller(pb, sol);
  if (!all_finite(sol->K, (size_t)(N + 1) * m * n) || !all_finite(sol->dbias, (size_t)(N + 1) * m) ||
      !all_finite(sol->Sm, (size_t)(N + 1) * n * n))
    status |= ORC_STATUS_NONFINITE;
  return status;
}

// ---- counter-based RNG shared with the CUDA generator (ocs2_b200/csrc/synthetic.cuh) ----
inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline double urand(uint64_t seed, int64_t problem, int node, int field, int idx) {
  const uint64_t ctr = ((((uint64_t)problem * 1024ULL + (uint64_t)node) * 16ULL + (uint64_t)field) << 16) + (uint64_t)idx;
  const uint64_t h = mix64(mix64(seed) ^ ctr);
  const double u01 = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  return 2.0 * u01 - 1.0;  // U[-1, 1)
}
enum { F_A = 0, F_B, F_HV, F_M, F_Q, F_R, F_C, F_CC, F_D, F_E, F_MF, F_QF, F_CF, F_X0 };

}  // namespace

extern "C" {

int orc_inverse_uut(int m, const double* H, double* Ui) {
  vec L(H, H + (size_t)m * m);
  const bool ok = cholesky_lower(m, L.data(), m);
  vec U((size_t)m * m);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) U[i + (size_t)j * m] = L[j + (size_t)i * m];  // matrixU() = L^T
  upper_inverse(m, U.data(), m, Ui, m);
  return ok ? 0 : 1;
}

void orc_constraint_projection(int m, int nc, const double* D, int ldd, const double* Ui, double* Ddagger, double* RcInv, double* Pu) {
  // QR of (RmInvUmUmT^T * Dm^T), m x nc
  vec Mx((size_t)m * nc), Qm((size_t)m * m);
  for (int j = 0; j < nc; ++j)
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int k = 0; k < m; ++k) acc += Ui[k + (size_t)i * m] * D[j + (size_t)k * ldd];
      Mx[i + (size_t)j * m] = acc;
    }
  householder_qr(m, nc, Mx.data(), m, Qm.data());
  vec Rc((size_t)nc * nc, 0.0);
  for (int j = 0; j < nc; ++j)
    for (int i = 0; i <= j; ++i) Rc[i + (size_t)j * nc] = Mx[i + (size_t)j * m];
  // setTriangularMinimumEigenvalues (LinearAlgebra.cpp:38-47) with weakEpsilon = 1e-9 (NumericTraits.h:51-53)
  const double minEig = 1e-9;
  for (int i = 0; i < nc; ++i) {
    double& ev = Rc[i + (size_t)i * nc];
    ev = (ev < 0.0) ? std::min(-minEig, ev) : std::max(minEig, ev);
  }
  upper_inverse(nc, Rc.data(), nc, RcInv, nc);  // DmDaggerTRmDmDaggerUUT = Rc^-1
  // DmDagger = RmInvUmUmT * (Qc * RcInv^T)
  vec QcRt((size_t)m * nc);
  gemm_nt(m, nc, nc, 1.0, Qm.data(), m, RcInv, nc, 0.0, QcRt.data(), m);
  gemm_nn(m, nc, m, 1.0, Ui, m, QcRt.data(), m, 0.0, Ddagger, m);
  // RmInvConstrainedUUT = RmInvUmUmT * Qu
  gemm_nn(m, m - nc, m, 1.0, Ui, m, Qm.data() + (size_t)nc * m, m, 0.0, Pu, m);
}

int orc_shift_hessian(int strategy, int n, double* M, double eps) {
  switch (strategy) {
    case ORC_HC_DIAGONAL_SHIFT:
      for (int i = 0; i < n; ++i) M[i + (size_t)i * n] += eps;  // matrix.diagonal().array() += minEigenvalue
      return 0;
    case ORC_HC_GERSHGORIN_MODIFICATION:
      make_psd_gershgorin(n, M, eps);
      return 0;
    case ORC_HC_EIGENVALUE_MODIFICATION:
      make_psd_eigenvalue(n, M, eps);
      return 0;
    default:
      return 1;  // CHOLESKY_MODIFICATION (Eigen::IncompleteCholesky): out of scope
  }
}

void orc_flatten(int n, const double* Sm, const double* Sv, double s, double* allSs) {
  size_t count = 0;
  for (int col = 0; col < n; ++col)
    for (int row = 0; row <= col; ++row) allSs[count++] = Sm[row + (size_t)col * n];
  for (int i = 0; i < n; ++i) allSs[count++] = Sv[i];
  allSs[count] = s;
}

void orc_unflatten(int n, const double* allSs, double* Sm, double* Sv, double* s) {
  size_t count = 0;
  for (int col = 0; col < n; ++col)
    for (int row = 0; row <= col; ++row) {
      Sm[row + (size_t)col * n] = allSs[count];
      Sm[col + (size_t)row * n] = allSs[count];
      ++count;
    }
  for (int i = 0; i < n; ++i) Sv[i] = allSs[count++];
  *s = allSs[count];
}

void orc_time_segment(double t, const double* time, int count, int* index, double* alpha) {
  if (count <= 1) {
    *index = 0;
    *alpha = 1.0;
    return;
  }
  // lookup::findIntervalInTimeArray = lower_bound - 1 (ocs2_core/include/ocs2_core/misc/Lookup.h:89-116)
  const int idx = (int)(std::lower_bound(time, time + count, t) - time) - 1;
  const int lastInterval = count - 1;
  if (idx >= 0) {
    if (idx < lastInterval) {
      const double intervalLength = time[idx + 1] - time[idx];
      const double timeTillNext = time[idx + 1] - t;
      const double minIntervalTime = 2.0 * 1e-9;
      if (intervalLength > minIntervalTime) {
        *index = idx;
        *alpha = timeTillNext / intervalLength;
        return;
      }
      *index = idx;
      *alpha = (timeTillNext < 0.5 * intervalLength) ? 0.0 : 1.0;
      return;
    }
    *index = std::max(lastInterval - 1, 0);
    *alpha = 0.0;
    return;
  }
  *index = 0;
  *alpha = 1.0;
}

int orc_project_stage(const orc_settings* st, int n, int m, int nc, int ldc, const double* A, const double* B, const double* Hv,
                      const double* Q, const double* P, const double* R, const double* q, const double* r, double c, const double* C,
                      const double* D, const double* e, const double* Sm, orc_projected* out, int* status) {
  Projected pr;
  pr.resize(n, m);
  StageIn in{n, m, nc, ldc, A, B, Hv, Q, P, R, q, r, c, C, D, e};
  const int stt = project_stage(*st, in, Sm, pr);
  if (status) *status = stt;
  const int p = pr.p;
  auto cp = [](const vec& v, size_t cnt, double* dst) {
    if (dst) std::copy(v.begin(), v.begin() + cnt, dst);
  };
  cp(pr.At, (size_t)n * n, out->At);
  cp(pr.Bt, (size_t)n * p, out->Bt);
  cp(pr.Hvt, n, out->Hvt);
  cp(pr.Qt, (size_t)n * n, out->Qt);
  cp(pr.Pt, (size_t)p * n, out->Pt);
  cp(pr.Rt, (size_t)p * p, out->Rt);
  cp(pr.qt, n, out->qt);
  cp(pr.rt, p, out->rt);
  if (out->ct) *out->ct = pr.ct;
  cp(pr.Cmt, (size_t)m * n, out->Cmt);
  cp(pr.Evt, m, out->Evt);
  cp(pr.Pu, (size_t)m * p, out->Pu);
  cp(pr.dQ, (size_t)n * n, out->dQ);
  cp(pr.dGm, (size_t)p * n, out->dGm);
  cp(pr.dGv, p, out->dGv);
  return p;
}

void orc_compute_map(int reduced, int n, int p, const orc_projected* pr, const double* SmNext, const double* SvNext, double sNext,
                     double* Km, double* Lv, double* Sm, double* Sv, double* s) {
  ProjPtrs pp{pr->At, pr->Bt, pr->Hvt, pr->Qt, pr->Pt, pr->Rt, pr->qt, pr->rt, pr->dQ, pr->dGm, pr->dGv, *pr->ct};
  compute_map(reduced != 0, n, p, pp, SmNext, SvNext, sNext, Km, Lv, Sm, Sv, s);
}

void orc_flow_map_slq(int reduced, int n, int p, const orc_projected* pr, const double* allSs, double* dallSs) {
  ProjPtrs pp{pr->At, pr->Bt, pr->Hvt, pr->Qt, pr->Pt, pr->Rt, pr->qt, pr->rt, pr->dQ, pr->dGm, pr->dGv, *pr->ct};
  flow_map_slq(reduced != 0, n, p, pp, allSs, dallSs);
}

int orc_backward(const orc_settings* st, const orc_problem* pb, orc_solution* sol) {
  const int status = (st->algorithm == ORC_ALG_SLQ) ? backward_slq(*st, *pb, sol) : backward_ilqr(*st, *pb, sol);
  sol->status = status;
  return status;
}

int orc_rollout(const orc_settings* st, const orc_problem* pb, const orc_solution* sol, const double* x0, double alpha, double* x,
                double* u, double* t_out, int max_out, int* n_out) {
  const int n = pb->nx, m = pb->nu, N = pb->N;
  int status = ORC_STATUS_OK;
  if (st->algorithm == ORC_ALG_ILQR) {
    // discrete LQ model: x_{k+1} = x_nom_{k+1} + A_k (x_k - x_nom_k) + B_k (u_k - u_nom_k) + Hv_k
    // (discrete-model semantics ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:48-52), with the policy
    // u_k = bias_k + alpha * deltaBias_k + K_k x_k (DDP_HelperFunctions.cpp:296-304, LinearController.cpp:79-87)
    if (max_out < N + 1) return -1;
    std::copy(x0, x0 + n, x);
    vec dx(n), du(m);
    for (int k = 0; k <= N; ++k) {
      const double* xk = x + (size_t)k * n;
      double* uk = u + (size_t)k * m;
      const double* K = sol->K + (size_t)k * m * n;
      for (int i = 0; i < m; ++i) uk[i] = sol->bias[(size_t)k * m + i] + alpha * sol->dbias[(size_t)k * m + i];
      gemm_nn(m, 1, n, 1.0, K, m, xk, n, 1.0, uk, m);
      if (k == N) break;
      for (int i = 0; i < n; ++i) dx[i] = xk[i] - (pb->x_nom ? pb->x_nom[(size_t)k * n + i] : 0.0);
      for (int i = 0; i < m; ++i) du[i] = uk[i] - (pb->u_nom ? pb->u_nom[(size_t)k * m + i] : 0.0);
      double* xn = x + (size_t)(k + 1) * n;
      for (int i = 0; i < n; ++i) xn[i] = pb->Hv[(size_t)k * n + i] + (pb->x_nom ? pb->x_nom[(size_t)(k + 1) * n + i] : 0.0);
      gemm_nn(n, 1, n, 1.0, pb->A + (size_t)k * n * n, n, dx.data(), n, 1.0, xn, n);
      if (!(pb->event && pb->event[k]))  // a pre-event node jumps: x+ = x_nom+ + A_e dx + Hv_e, the input does not act
        gemm_nn(n, 1, m, 1.0, pb->B + (size_t)k * n * m, n, du.data(), m, 1.0, xn, n);
      if (t_out) t_out[k] = pb->time ? pb->time[k] : (double)k;
    }
    if (t_out) t_out[N] = pb->time ? pb->time[N] : (double)N;
    *n_out = N + 1;
    if (!all_finite(x, (size_t)(N + 1) * n)) status |= ORC_STATUS_NONFINITE;
    return status;
  }
  // continuous LQ model: xdot = A(t)(x - x_nom(t)) + B(t)(u - u_nom(t)) + Hv(t), all lerped on the node grid
  // (LinearSystemDynamics::computeFlowMap, ocs2_core/src/dynamics/LinearSystemDynamics.cpp:54-58, is the Hv = 0, nominal = 0 case);
  // TimeTriggeredRollout::run (TimeTriggeredRollout.cpp:46-115): integrateAdaptive with a plain RK4 stepper =
  // boost::odeint integrate_const steps of `timeStep` plus a truncated last step; start nudged by weakEpsilon
  // (RolloutBase.cpp:62-64).
  const double* time = pb->time;
  const double t0 = time[0], tf = time[N];
  const double tStart = std::min(t0 + 1e-9, tf);
  const double dt = st->time_step;
  vec Ai((size_t)n * n), Bi((size_t)n * m), Hvi(n), Ki((size_t)m * n), bi(m), xni(n), uni(m), uu(m), dxv(n);
  auto policy = [&](double t, const double* xx, double* uo) {
    int idx;
    double a;
    orc_time_segment(t, time, N + 1, &idx, &a);
    for (int i = 0; i < m; ++i) {
      const double bl = sol->bias[(size_t)idx * m + i] + alpha * sol->dbias[(size_t)idx * m + i];
      const double br = sol->bias[(size_t)(idx + 1) * m + i] + alpha * sol->dbias[(size_t)(idx + 1) * m + i];
      uo[i] = a * bl + (1.0 - a) * br;
    }
    for (size_t i = 0; i < (size_t)m * n; ++i) Ki[i] = a * sol->K[(size_t)idx * m * n + i] + (1.0 - a) * sol->K[(size_t)(idx + 1) * m * n + i];
    gemm_nn(m, 1, n, 1.0, Ki.data(), m, xx, n, 1.0, uo, m);
  };
  auto flow = [&](double t, const vec& xx, vec& dxdt) {
    policy(t, xx.data(), uu.data());
    int idx;
    double a;
    orc_time_segment(t, time, N + 1, &idx, &a);
    for (size_t i = 0; i < (size_t)n * n; ++i) Ai[i] = a * pb->A[(size_t)idx * n * n + i] + (1.0 - a) * pb->A[(size_t)(idx + 1) * n * n + i];
    for (size_t i = 0; i < (size_t)n * m; ++i) Bi[i] = a * pb->B[(size_t)idx * n * m + i] + (1.0 - a) * pb->B[(size_t)(idx + 1) * n * m + i];
    for (int i = 0; i < n; ++i) {
      dxdt[i] = a * pb->Hv[(size_t)idx * n + i] + (1.0 - a) * pb->Hv[(size_t)(idx + 1) * n + i];
      const double xn = pb->x_nom ? a * pb->x_nom[(size_t)idx * n + i] + (1.0 - a) * pb->x_nom[(size_t)(idx + 1) * n + i] : 0.0;
      dxv[i] = xx[i] - xn;
    }
    for (int i = 0; i < m; ++i) {
      const double un = pb->u_nom ? a * pb->u_nom[(size_t)idx * m + i] + (1.0 - a) * pb->u_nom[(size_t)(idx + 1) * m + i] : 0.0;
      uu[i] -= un;
    }
    gemm_nn(n, 1, n, 1.0, Ai.data(), n, dxv.data(), n, 1.0, dxdt.data(), n);
    gemm_nn(n, 1, m, 1.0, Bi.data(), n, uu.data(), m, 1.0, dxdt.data(), n);
  };
  vec y(x0, x0 + n), k1(n), k2(n), k3(n), k4(n), tmp(n);
  int count = 0;
  auto observe = [&](double t) {
    if (count < max_out) {
      std::copy(y.begin(), y.end(), x + (size_t)count * n);
      policy(t, y.data(), u + (size_t)count * m);  // inputs reconstructed at output nodes (TimeTriggeredRollout.cpp:98-102)
      if (t_out) t_out[count] = t;
    }
    ++count;
  };
  // RolloutBase::findActiveModesTimeInterval (RolloutBase.cpp:43-67): the event times split [t0, tf] into intervals whose start is
  // nudged by weakEpsilon; here the event times are the stamps of the pre-event nodes
  std::vector<double> switching{t0};
  std::vector<int> event_node;
  if (pb->event)
    for (int k = 0; k < N; ++k)
      if (pb->event[k]) {
        switching.push_back(time[k]);
        event_node.push_back(k);
      }
  switching.push_back(tf);
  const int num_intervals = (int)switching.size() - 1;
  for (int iv = 0; iv < num_intervals; ++iv) {
    const double tEnd = switching[iv + 1];
    const double tBegin = std::min(switching[iv] + 1e-9, tEnd);
    if (tBegin < tEnd) {
      // integrate_const (boost/numeric/odeint/integrate/detail/integrate_const.hpp, stepper_tag)
      double t = tBegin;
      int step = 0;
      while (less_eq_with_sign(t + dt, tEnd)) {
        observe(t);
        rk4_step(flow, y, t, dt, k1, k2, k3, k4, tmp);
        ++step;
        t = tBegin + (double)step * dt;
      }
      observe(t);
      // integrate_adaptive's last truncated step (integrate_adaptive.hpp, stepper_tag)
      const double end = tBegin + dt * (double)step;
      if (less_with_sign(end, tEnd)) {
        rk4_step(flow, y, end, tEnd - end, k1, k2, k3, k4, tmp);
        observe(tEnd);
      }
    } else {
      observe(tEnd);
    }
    if (iv + 1 < num_intervals) {
      // jump map of the LQ model (TimeTriggeredRollout.cpp:104-108): x+ = x_nom(post) + A_e (x - x_nom(pre)) + Hv_e
      const int k = event_node[iv];
      const double* Ae = pb->jA + (size_t)iv * n * n;
      for (int i = 0; i < n; ++i) dxv[i] = y[i] - (pb->x_nom ? pb->x_nom[(size_t)k * n + i] : 0.0);
      for (int i = 0; i < n; ++i) tmp[i] = pb->jHv[(size_t)iv * n + i] + (pb->x_nom ? pb->x_nom[(size_t)(k + 1) * n + i] : 0.0);
      gemm_nn(n, 1, n, 1.0, Ae, n, dxv.data(), n, 1.0, tmp.data(), n);
      y = tmp;
    }
  }
  *n_out = count;
  if (count > max_out) return -1;
  if (!all_finite(x, (size_t)count * n)) status |= ORC_STATUS_NONFINITE;
  return status;
}

double orc_discrete_lq_cost(const orc_problem* pb, const double* x, const double* u) {
  // sum_k [ c + q.dx + r.du + 1/2 dx'Q dx + du'P dx + 1/2 du'R du ] + terminal, with dx = x - x_nom, du = u - u_nom
  const int n = pb->nx, m = pb->nu, N = pb->N;
  double J = 0.0;
  vec dx(n), du(m), t1(std::max(n, m));
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < n; ++i) dx[i] = x[(size_t)k * n + i] - (pb->x_nom ? pb->x_nom[(size_t)k * n + i] : 0.0);
    for (int i = 0; i < m; ++i) du[i] = u[(size_t)k * m + i] - (pb->u_nom ? pb->u_nom[(size_t)k * m + i] : 0.0);
    J += pb->c[k] + dot(n, pb->q + (size_t)k * n, dx.data()) + dot(m, pb->r + (size_t)k * m, du.data());
    gemm_nn(n, 1, n, 1.0, pb->Q + (size_t)k * n * n, n, dx.data(), n, 0.0, t1.data(), n);
    J += 0.5 * dot(n, dx.data(), t1.data());
    if (pb->event && pb->event[k]) {  // pre-jump cost: state terms only (undo the r.du added above)
      J -= dot(m, pb->r + (size_t)k * m, du.data());
      continue;
    }
    gemm_nn(m, 1, n, 1.0, pb->P + (size_t)k * m * n, m, dx.data(), n, 0.0, t1.data(), m);
    J += dot(m, du.data(), t1.data());
    gemm_nn(m, 1, m, 1.0, pb->R + (size_t)k * m * m, m, du.data(), m, 0.0, t1.data(), m);
    J += 0.5 * dot(m, du.data(), t1.data());
  }
  for (int i = 0; i < n; ++i) dx[i] = x[(size_t)N * n + i] - (pb->x_nom ? pb->x_nom[(size_t)N * n + i] : 0.0);
  gemm_nn(n, 1, n, 1.0, pb->Qf, n, dx.data(), n, 0.0, t1.data(), n);
  J += pb->cf[0] + dot(n, pb->qf, dx.data()) + 0.5 * dot(n, dx.data(), t1.data());
  return J;
}

void orc_generate_problem(uint64_t seed, int64_t problem, int algorithm, int n, int m, int nc, int N, double dt, double* A, double* B,
                          double* Hv, double* Q, double* P, double* R, double* q, double* r, double* c, double* C, double* D, double* e,
                          double* Qf, double* qf, double* cf, double* x0) {
  // Family of SURVEY.md §8(d): A = I + dt*Ac (ILQR) or Ac (SLQ), Ac ~ U(-1,1)/sqrt(n); B = dt*Bc or Bc, Bc ~ U(-1,1);
  // joint cost W = M^T M/(n+m) + 0.1 I with M ~ U(-1,1)^{(n+m)x(n+m)} (as getRandomCost,
  // ocs2_oc/test/include/ocs2_oc/test/testProblemsGeneration.h:45-58), scaled by dt for the discrete model (ILQR.cpp:149-150);
  // q, r ~ U(-1,1)*scale; c ~ U(0,1)*scale; Hv ~ 0.01 U(-1,1); D = [I | U(-1,1)] (full row rank), C ~ U(-1,1), e ~ 0.1 U(-1,1).
  // Every product is an explicit sequential fma chain so that the CUDA generator reproduces the bits exactly.
  const bool discrete = (algorithm == ORC_ALG_ILQR);
  const int nodes = discrete ? N : N + 1;
  const int nm = n + m;
  const double invSqrtN = 1.0 / std::sqrt((double)n);
  const double scale = discrete ? dt : 1.0;
  const double invNm = 1.0 / (double)nm;
  vec M((size_t)nm * nm);
  for (int k = 0; k < nodes; ++k) {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const double ac = urand(seed, problem, k, F_A, i + j * n) * invSqrtN;
        A[(size_t)k * n * n + i + (size_t)j * n] = discrete ? ((i == j ? 1.0 : 0.0) + dt * ac) : ac;
      }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) B[(size_t)k * n * m + i + (size_t)j * n] = scale * urand(seed, problem, k, F_B, i + j * n);
    for (int i = 0; i < n; ++i) Hv[(size_t)k * n + i] = 0.01 * urand(seed, problem, k, F_HV, i);
    for (int j = 0; j < nm; ++j)
      for (int i = 0; i < nm; ++i) M[i + (size_t)j * nm] = urand(seed, problem, k, F_M, i + j * nm);
    auto W = [&](int i, int j) {  // (M^T M)(i,j)/(n+m) + 0.1 delta_ij
      double acc = 0.0;
      for (int l = 0; l < nm; ++l) acc = std::fma(M[l + (size_t)i * nm], M[l + (size_t)j * nm], acc);
      return std::fma(acc, invNm, (i == j) ? 0.1 : 0.0);
    };
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) Q[(size_t)k * n * n + i + (size_t)j * n] = scale * W(std::min(i, j), std::max(i, j));
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) P[(size_t)k * m * n + i + (size_t)j * m] = scale * W(j, n + i);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) R[(size_t)k * m * m + i + (size_t)j * m] = scale * W(n + std::min(i, j), n + std::max(i, j));
    for (int i = 0; i < n; ++i) q[(size_t)k * n + i] = scale * urand(seed, problem, k, F_Q, i);
    for (int i = 0; i < m; ++i) r[(size_t)k * m + i] = scale * urand(seed, problem, k, F_R, i);
    c[k] = scale * (0.5 * (urand(seed, problem, k, F_C, 0) + 1.0));
    if (nc > 0) {
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < nc; ++i) C[(size_t)k * nc * n + i + (size_t)j * nc] = urand(seed, problem, k, F_CC, i + j * nc);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < nc; ++i)
          D[(size_t)k * nc * m + i + (size_t)j * nc] = (j < nc) ? ((i == j) ? 1.0 : 0.0) : urand(seed, problem, k, F_D, i + j * nc);
      for (int i = 0; i < nc; ++i) e[(size_t)k * nc + i] = 0.1 * urand(seed, problem, k, F_E, i);
    }
  }
  // terminal: Qf = Mf^T Mf / n + 0.1 I, qf ~ U(-1,1), cf ~ U(0,1)
  vec Mf((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) Mf[i + (size_t)j * n] = urand(seed, problem, 1023, F_MF, i + j * n);
  const double invN = 1.0 / (double)n;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const int a = std::min(i, j), b = std::max(i, j);
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc = std::fma(Mf[l + (size_t)a * n], Mf[l + (size_t)b * n], acc);
      Qf[i + (size_t)j * n] = std::fma(acc, invN, (i == j) ? 0.1 : 0.0);
    }
  for (int i = 0; i < n; ++i) qf[i] = urand(seed, problem, 1023, F_QF, i);
  cf[0] = 0.5 * (urand(seed, problem, 1023, F_CF, 0) + 1.0);
  for (int i = 0; i < n; ++i) x0[i] = urand(seed, problem, 1023, F_X0, i);
}

double orc_baseline_run(const orc_settings* st, uint64_t seed, int64_t first, int64_t count, int n, int m, int nc, int N, double dt,
                        int threads, double* checksum) {
  const bool discrete = st->algorithm == ORC_ALG_ILQR;
  const int nodes = discrete ? N : N + 1;
  struct Buffers {
    vec A, B, Hv, Q, P, R, q, r, c, C, D, e, Qf, qf, cf, x0, time;
  };
  std::vector<Buffers> data((size_t)count);
  // generation is not timed (the GPU arm also starts from resident inputs)
  {
    std::atomic<int64_t> next{0};
    auto gen = [&]() {
      int64_t i;
      while ((i = next++) < count) {
        Buffers& b = data[(size_t)i];
        b.A.resize((size_t)nodes * n * n);
        b.B.resize((size_t)nodes * n * m);
        b.Hv.resize((size_t)nodes * n);
        b.Q.resize((size_t)nodes * n * n);
        b.P.resize((size_t)nodes * m * n);
        b.R.resize((size_t)nodes * m * m);
        b.q.resize((size_t)nodes * n);
        b.r.resize((size_t)nodes * m);
        b.c.resize(nodes);
        b.C.resize((size_t)nodes * std::max(nc, 1) * n);
        b.D.resize((size_t)nodes * std::max(nc, 1) * m);
        b.e.resize((size_t)nodes * std::max(nc, 1));
        b.Qf.resize((size_t)n * n);
        b.qf.resize(n);
        b.cf.resize(1);
        b.x0.resize(n);
        b.time.resize(N + 1);
        for (int k = 0; k <= N; ++k) b.time[k] = dt * (double)k;
        orc_generate_problem(seed, first + i, st->algorithm, n, m, nc, N, dt, b.A.data(), b.B.data(), b.Hv.data(), b.Q.data(), b.P.data(),
                             b.R.data(), b.q.data(), b.r.data(), b.c.data(), b.C.data(), b.D.data(), b.e.data(), b.Qf.data(),
                             b.qf.data(), b.cf.data(), b.x0.data());
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(gen);
    for (auto& th : pool) th.join();
  }
  std::atomic<int64_t> next{0};
  std::vector<double> partial(threads, 0.0);
  auto work = [&](int tid) {
    vec K((size_t)(N + 1) * m * n), db((size_t)(N + 1) * m), bias((size_t)(N + 1) * m), Sm((size_t)(N + 1) * n * n), Sv((size_t)(N + 1) * n),
        s(N + 1);
    const int maxOut = N + 8;
    vec x((size_t)maxOut * n), u((size_t)maxOut * m);
    int64_t i;
    while ((i = next++) < count) {
      Buffers& b = data[(size_t)i];
      orc_problem pb{};
      pb.nx = n;
      pb.nu = m;
      pb.nc_max = nc;
      pb.N = N;
      pb.A = b.A.data();
      pb.B = b.B.data();
      pb.Hv = b.Hv.data();
      pb.Q = b.Q.data();
      pb.P = b.P.data();
      pb.R = b.R.data();
      pb.q = b.q.data();
      pb.r = b.r.data();
      pb.c = b.c.data();
      pb.C = b.C.data();
      pb.D = b.D.data();
      pb.e = b.e.data();
      pb.nc = nullptr;
      pb.Qf = b.Qf.data();
      pb.qf = b.qf.data();
      pb.cf = b.cf.data();
      pb.x_nom = nullptr;
      pb.u_nom = nullptr;
      pb.time = b.time.data();
      orc_solution sol{K.data(), db.data(), bias.data(), Sm.data(), Sv.data(), s.data(), 0};
      orc_backward(st, &pb, &sol);
      int nOut = 0;
      orc_rollout(st, &pb, &sol, b.x0.data(), 1.0, x.data(), u.data(), nullptr, maxOut, &nOut);
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += x[(size_t)(nOut - 1) * n + j];
      for (int j = 0; j < m * n; ++j) acc += K[j];
      partial[tid] += acc;
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  double total = 0.0;
  for (double v : partial) total += v;
  if (checksum) *checksum = total;
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_batch_solve(const orc_settings* st, uint64_t seed, int64_t first, int64_t count, int n, int m, int nc, int N, double dt, double alpha,
                    int threads, double* K, double* dbias, double* bias, double* Sm, double* Sv, double* s, double* x, double* u,
                    int max_out, int32_t* status) {
  const bool discrete = st->algorithm == ORC_ALG_ILQR;
  const int nodes = discrete ? N : N + 1;
  std::atomic<int64_t> next{0};
  std::atomic<int> out_nodes{0};
  auto work = [&]() {
    const int ncm = std::max(nc, 1);
    vec A((size_t)nodes * n * n), B((size_t)nodes * n * m), Hv((size_t)nodes * n), Q((size_t)nodes * n * n), P((size_t)nodes * m * n),
        R((size_t)nodes * m * m), q((size_t)nodes * n), r((size_t)nodes * m), c(nodes), C((size_t)nodes * ncm * n), D((size_t)nodes * ncm * m),
        e((size_t)nodes * ncm), Qf((size_t)n * n), qf(n), cf(1), x0(n), time(N + 1);
    vec lK((size_t)(N + 1) * m * n), ldb((size_t)(N + 1) * m), lbias((size_t)(N + 1) * m), lSm((size_t)(N + 1) * n * n), lSv((size_t)(N + 1) * n),
        ls(N + 1), lx((size_t)max_out * n), lu((size_t)max_out * m);
    for (int k = 0; k <= N; ++k) time[k] = dt * (double)k;
    int64_t i;
    while ((i = next++) < count) {
      orc_generate_problem(seed, first + i, st->algorithm, n, m, nc, N, dt, A.data(), B.data(), Hv.data(), Q.data(), P.data(), R.data(), q.data(),
                           r.data(), c.data(), C.data(), D.data(), e.data(), Qf.data(), qf.data(), cf.data(), x0.data());
      orc_problem pb{};
      pb.nx = n;
      pb.nu = m;
      pb.nc_max = nc;
      pb.N = N;
      pb.A = A.data();
      pb.B = B.data();
      pb.Hv = Hv.data();
      pb.Q = Q.data();
      pb.P = P.data();
      pb.R = R.data();
      pb.q = q.data();
      pb.r = r.data();
      pb.c = c.data();
      pb.C = C.data();
      pb.D = D.data();
      pb.e = e.data();
      pb.Qf = Qf.data();
      pb.qf = qf.data();
      pb.cf = cf.data();
      pb.time = time.data();
      orc_solution sol{lK.data(), ldb.data(), lbias.data(), lSm.data(), lSv.data(), ls.data(), 0};
      const int stat = orc_backward(st, &pb, &sol);
      int nOut = 0;
      orc_rollout(st, &pb, &sol, x0.data(), alpha, lx.data(), lu.data(), nullptr, max_out, &nOut);
      out_nodes = nOut;
      auto put = [&](double* dst, const vec& src, size_t per) {
        if (dst) std::copy(src.begin(), src.begin() + per, dst + (size_t)i * per);
      };
      put(K, lK, (size_t)(N + 1) * m * n);
      put(dbias, ldb, (size_t)(N + 1) * m);
      put(bias, lbias, (size_t)(N + 1) * m);
      put(Sm, lSm, (size_t)(N + 1) * n * n);
      put(Sv, lSv, (size_t)(N + 1) * n);
      put(s, ls, (size_t)(N + 1));
      if (x) std::copy(lx.begin(), lx.begin() + (size_t)nOut * n, x + (size_t)i * max_out * n);
      if (u) std::copy(lu.begin(), lu.begin() + (size_t)nOut * m, u + (size_t)i * max_out * m);
      if (status) status[i] = stat;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < std::max(threads, 1); ++t) pool.emplace_back(work);
  for (auto& th : pool) th.join();
  return out_nodes;
}

}  // extern "C"
