/*
 * lq_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A dependency-free restatement, operation for operation, of the reference's (RIVeR-Lab/ocs2) Eigen
 * implementation of the batched-LQ hot path: ILQR discrete Riccati sweep, SLQ continuous Riccati flow
 * map under fixed-step RK4, and the LinearController rollout of the LQ model.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library. The shipped library (ocs2_b200/csrc -> libocs2_ddp_cuda.so) never links or calls it.
 *
 * Parity pinning: the reference itself cannot be compiled here (Eigen3 and Boost.odeint are absent from
 * this image), so this oracle is pinned against the reference's own golden vectors
 * (tests/test_oracle_golden.py): the Riccati flatten order (ocs2_ddp/test/RiccatiTest.cpp:107-131), the
 * MATLAB CARE known answer (ocs2_ddp/test/testContinuousTimeLqr.cpp:41-72), the LLT / constraint
 * projection identities (ocs2_core/test/misc/testLinearAlgebra.cpp), the change-of-input-variables
 * equivalences (ocs2_oc/test/testChangeOfInputVariables.cpp) and the DDP == dense-KKT check
 * (ocs2_ddp/test/CorrectnessTest.cpp:215-223, KKT restated in oracle/kkt_oracle.py).
 *
 * All matrices are column-major (Eigen default), doubles. Citations are relative to /root/reference.
 */
#ifndef LQ_ORACLE_H_
#define LQ_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum values mirror the reference's enum order */
enum { ORC_ALG_ILQR = 0, ORC_ALG_SLQ = 1 };                 /* ocs2_ddp/include/ocs2_ddp/DDP_Settings.h Algorithm */
enum { ORC_STRATEGY_LINE_SEARCH = 0, ORC_STRATEGY_LM = 1 }; /* search_strategy/StrategySettings.h:43 */
enum {                                                      /* ocs2_ddp/include/ocs2_ddp/HessianCorrection.h:44-49 */
  ORC_HC_DIAGONAL_SHIFT = 0,
  ORC_HC_CHOLESKY_MODIFICATION = 1,
  ORC_HC_EIGENVALUE_MODIFICATION = 2,
  ORC_HC_GERSHGORIN_MODIFICATION = 3
};
enum { ORC_STATUS_OK = 0, ORC_STATUS_CHOL_NOT_PD = 1, ORC_STATUS_NONFINITE = 2, ORC_STATUS_CONSTRAINT_RANK = 4 };

typedef struct orc_settings {
  int32_t algorithm;          /* ORC_ALG_* */
  int32_t reduced_form;       /* ddp preComputeRiccatiTerms && LINE_SEARCH (ILQR.cpp:68, SLQ.cpp:65) */
  int32_t strategy;           /* ORC_STRATEGY_* */
  int32_t hessian_correction; /* ORC_HC_* (lineSearch.hessianCorrectionStrategy) */
  double hessian_multiple;    /* lineSearch.hessianCorrectionMultiple */
  double lm_riccati_multiple; /* LM riccatiMultiple */
  double time_step;           /* ddp timeStep (SLQ backward RK4) / rollout timeStep */
} orc_settings;

/* One LQ problem. Per-field arrays, contiguous over nodes: field[node][block]. ILQR uses nodes 0..N-1 (discrete
 * stage data), SLQ uses nodes 0..N (continuous-time model data at every time node). */
typedef struct orc_problem {
  int32_t nx, nu, nc_max, N;
  const double *A, *B, *Hv;         /* dynamics.dfdx n*n, dynamics.dfdu n*m, dynamicsBias n   (ModelData.h:43-60) */
  const double *Q, *P, *R, *q, *r, *c; /* cost.dfdxx n*n, dfdux m*n, dfduu m*m, dfdx n, dfdu m, f 1 */
  const double *C, *D, *e;          /* stateInputEqConstraint dfdx/dfdu/f; blocks nc_max*n, nc_max*m, nc_max; ld = nc_max */
  const int32_t* nc;                /* active constraints per node (NULL => nc_max) */
  const double *Qf, *qf, *cf;       /* terminal value function (already Hessian-shifted, GaussNewtonDDP.cpp:724-727) */
  const double *x_nom, *u_nom;      /* [N+1][n], [N+1][m] nominal trajectories (NULL => 0) */
  const double* time;               /* [N+1] node times (SLQ / continuous rollout) */
  const int32_t* event;             /* ILQR only, [N] or NULL: event[k] != 0 marks node k as a PRE-EVENT node (time[k] == time[k+1],
                                       k+1 in postEventIndices_). Its A, Hv, Q, q, c are the jump ModelData (modelDataEventTimes:
                                       jump map linearisation and pre-jump cost), its B, P, R, r, C, D, e the regular model data of the
                                       node, used only for the controller (ILQR.cpp:263-295).
                                       SLQ: [N+1] or NULL: event[k] != 0 marks node k as a pre-event node and node k + 1 as its post-event
                                       node (time[k+1] = time[k] + weakEpsilon as the reference's rollouts stamp them, RolloutBase.cpp:62-64).
                                       Every node keeps its own continuous-time model data; the jump ModelData of the e-th event (in
                                       node order) are jA, jHv, jQ, jq, jc below. */
  const double *jA, *jHv, *jQ, *jq, *jc; /* SLQ jump model data per event: [e][n*n], [e][n], [e][n*n], [e][n], [e] (modelDataEventTimes) */
} orc_problem;

typedef struct orc_solution {
  double *K, *dbias, *bias; /* LinearController gainArray_/deltaBiasArray_/biasArray_: [N+1][m*n], [N+1][m], [N+1][m] */
  double *Sm, *Sv, *s;      /* valueFunctionTrajectory: [N+1][n*n], [N+1][n], [N+1] */
  int32_t status;
} orc_solution;

/* ---- small dense building blocks (exported so tests can pin them one by one) ---- */
/* LinearAlgebra::computeInverseMatrixUUT, ocs2_core/src/misc/LinearAlgebra.cpp:119-124. Ui = U^-1, H = U^T U. */
int orc_inverse_uut(int m, const double* H, double* Ui);
/* LinearAlgebra::computeConstraintProjection, LinearAlgebra.cpp:129-155 (+ :38-47). D is nc x m with ld ldd. */
void orc_constraint_projection(int m, int nc, const double* D, int ldd, const double* Ui, double* Ddagger, double* RcInv,
                               double* Pu);
/* hessian_correction::shiftHessian, ocs2_ddp/src/HessianCorrection.cpp:53-74 (DIAGONAL_SHIFT, GERSHGORIN only). */
int orc_shift_hessian(int strategy, int n, double* M, double eps);
/* ContinuousTimeRiccatiEquations::convert2Vector / convert2Matrix, ContinuousTimeRiccatiEquations.cpp:55-108 */
void orc_flatten(int n, const double* Sm, const double* Sv, double s, double* allSs);
void orc_unflatten(int n, const double* allSs, double* Sm, double* Sv, double* s);
/* LinearInterpolation::timeSegment, implementation/LinearInterpolation.h:69-107 */
void orc_time_segment(double t, const double* time, int count, int* index, double* alpha);

/* Projected stage (GaussNewtonDDP::computeProjectionAndRiccatiModification, GaussNewtonDDP.cpp:734-750).
 * Sm may be NULL (SLQ: Hm = R, SLQ.cpp:206-208). Outputs are caller-allocated at full (unprojected) sizes; p = m - nc
 * is returned. Layouts: At n*n, Bt n*p, Hvt n, Qt n*n, Pt p*n (ld p), Rt p*p (ld p), qt n, rt p, ct 1, Cmt m*n
 * (=Ddagger*C), Evt m (=Ddagger*e), Pu m*p, dQ n*n, dGm p*n (ld p), dGv p. */
typedef struct orc_projected {
  double *At, *Bt, *Hvt, *Qt, *Pt, *Rt, *qt, *rt, *ct, *Cmt, *Evt, *Pu, *dQ, *dGm, *dGv;
} orc_projected;
int orc_project_stage(const orc_settings* st, int n, int m, int nc, int ldc, const double* A, const double* B, const double* Hv,
                      const double* Q, const double* P, const double* R, const double* q, const double* r, double c,
                      const double* C, const double* D, const double* e, const double* Sm, orc_projected* out, int* status);

/* DiscreteTimeRiccatiEquations::computeMapILQR, DiscreteTimeRiccatiEquations.cpp:65-154 */
void orc_compute_map(int reduced, int n, int p, const orc_projected* pr, const double* SmNext, const double* SvNext,
                     double sNext, double* Km, double* Lv, double* Sm, double* Sv, double* s);
/* ContinuousTimeRiccatiEquations::computeFlowMapSLQ on already-interpolated projected data, :152-292 */
void orc_flow_map_slq(int reduced, int n, int p, const orc_projected* pr, const double* allSs, double* dallSs);

/* ---- whole-path drivers ---- */
/* ILQR: ILQR.cpp:186-299 + GaussNewtonDDP.cpp:516-642; SLQ: SLQ.cpp:127-302. Single partition (exact sweep). */
int orc_backward(const orc_settings* st, const orc_problem* pb, orc_solution* sol);
/* LQ-model rollout with the LinearController (DDP_HelperFunctions.cpp:296-304, LinearController.cpp:79-87,
 * TimeTriggeredRollout.cpp:46-115). Discrete (ILQR data): x[N+1][n], u[N+1][m] (u[N] re-evaluates the copied last policy
 * at x_N). Continuous (SLQ data): RK4 constant steps (integrate_adaptive with a plain stepper); outputs at the
 * step times, count returned through n_out (capacity max_out), t_out optional. */
int orc_rollout(const orc_settings* st, const orc_problem* pb, const orc_solution* sol, const double* x0, double alpha,
                double* x, double* u, double* t_out, int max_out, int* n_out);
/* trajectory cost of the LQ model along a discrete rollout (used by the V(x0)==cost self-consistency test) */
double orc_discrete_lq_cost(const orc_problem* pb, const double* x, const double* u);

/* ---- synthetic, counter-based problem generator (bit-identical to the CUDA generator o2c_generate_synthetic) ---- */
/* Fills one problem (index `problem`) of the seeded family described in SURVEY.md §8(d) / DESIGN.md into caller arrays
 * laid out like orc_problem (field[node][block]); x0 is n doubles. nodes = N (ILQR) or N+1 (SLQ). */
void orc_generate_problem(uint64_t seed, int64_t problem, int algorithm, int n, int m, int nc, int N, double dt, double* A,
                          double* B, double* Hv, double* Q, double* P, double* R, double* q, double* r, double* c, double* C,
                          double* D, double* e, double* Qf, double* qf, double* cf, double* x0);

/* ---- batched CPU baseline: `count` generated problems [first, first+count), backward + one rollout (alpha = 1),
 * one problem per task on `threads` std::threads. Returns elapsed seconds (generation excluded); checksum accumulates
 * sum of x_N and K_0 entries so the work cannot be optimised away. */
double orc_baseline_run(const orc_settings* st, uint64_t seed, int64_t first, int64_t count, int n, int m, int nc, int N, double dt,
                        int threads, double* checksum);

/* ---- batched solve with the results kept: problems [first, first+count) of the seeded family, backward + one rollout with step
 * length alpha, one problem per task on `threads` std::threads. Caller-allocated outputs (any may be NULL), indexed
 * [problem][node][block] with column-major blocks like orc_solution: K [count][N+1][m*n], dbias/bias [count][N+1][m],
 * Sm [count][N+1][n*n], Sv [count][N+1][n], s [count][N+1], x [count][out_nodes][n], u [count][out_nodes][m] (out_nodes = N+1 for
 * ILQR; SLQ: the rollout step schedule, capacity max_out, the count actually written is returned), status [count]. Used by the
 * all-problems GPU parity test. */
int orc_batch_solve(const orc_settings* st, uint64_t seed, int64_t first, int64_t count, int n, int m, int nc, int N, double dt, double alpha,
                    int threads, double* K, double* dbias, double* bias, double* Sm, double* Sv, double* s, double* x, double* u,
                    int max_out, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* LQ_ORACLE_H_ */
