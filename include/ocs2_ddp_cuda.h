/*
 * ocs2_ddp_cuda.h — C ABI of the B200-native batched LQ solver (libocs2_ddp_cuda.so).
 *
 * This is the drop-in boundary for ONE hot path of RIVeR-Lab/ocs2: the LQ sub-problem of the DDP inner loop, applied to
 * `batch` independent optimal control problems at once, FP64, on one CUDA device (sm_100a):
 *
 *   o2c_backward  replaces, per problem,
 *       ILQR::solveSequentialRiccatiEquations / riccatiEquationsWorker      ocs2_ddp/src/ILQR.cpp:186-299
 *       SLQ::solveSequentialRiccatiEquations / riccatiEquationsWorker       ocs2_ddp/src/SLQ.cpp:174-302
 *       GaussNewtonDDP::computeProjectionAndRiccatiModification             ocs2_ddp/src/GaussNewtonDDP.cpp:734-782
 *       DiscreteTimeRiccatiEquations::computeMap                            ocs2_ddp/src/riccati_equations/DiscreteTimeRiccatiEquations.cpp:50-154
 *       ContinuousTimeRiccatiEquations::computeFlowMap (+ RK4 integrate_times) .../ContinuousTimeRiccatiEquations.cpp:152-292
 *       GaussNewtonDDP::calculateController + ILQR/SLQ::calculateControllerWorker  GaussNewtonDDP.cpp:588-642, ILQR.cpp:162-181, SLQ.cpp:127-169
 *   o2c_rollout   replaces incrementController + rolloutTrajectory of the LQ model with the LinearController
 *       ocs2_ddp/src/DDP_HelperFunctions.cpp:125-138, 296-304; ocs2_core/src/control/LinearController.cpp:79-87;
 *       ocs2_oc/src/rollout/TimeTriggeredRollout.cpp:46-115
 *
 * The reference has no FFI for this path (the seam is C++ virtual dispatch inside GaussNewtonDDP, see
 * ocs2_ddp/include/ocs2_ddp/GaussNewtonDDP.h:149-193); INTEGRATION.md shows the ILQR/SLQ subclass a maintainer would add to
 * call this ABI (built: include/ocs2_ddp_cuda/GaussNewtonDDP_CUDA.h, ocs2::ILQR_CUDA / ocs2::SLQ_CUDA). Inputs are the ModelData fields (ocs2_core/include/ocs2_core/model_data/ModelData.h:43-60) of every time node
 * in a struct-of-arrays batch layout; every matrix block is column-major (Eigen default) and contiguous.
 *
 * Conventions: all functions return an o2c_error (0 = success) and never throw; o2c_last_error() gives the message of the
 * last failure on the calling thread. Calls on one handle must be serialised by the caller; different handles are independent:
 * they may live on different devices and be driven from different host threads at the same time (the library keeps no
 * process-wide launch state; include/ocs2_ddp_cuda/ShardedRiccatiSolver.h shards one batch over several devices that way).
 * There is NO CPU fallback: without a CUDA device o2c_create fails with O2C_ERR_CUDA.
 *
 * Riccati form: under LINE_SEARCH the shape-specialised kernels evaluate the REDUCED form for both O2C_FORM_* values. The full form
 * (preComputeRiccatiTerms = false) is the same map written with K~'G~ + G~'K~ + K~'H~K~ in place of -G~'G~ (H~ = Pu'Hm Pu = I); the
 * reference's own RiccatiTest.cpp:87-105 holds the two equal to 1e-9. They differ only in rounding, which matters when Hm is barely
 * positive definite. Configurations served by the generic kernels (see o2c_kernel_variant) evaluate the form that was asked for.
 *
 * Equality constraints: the shape-specialised kernels use the range-space form of the reference's projection (Z = L^-1 D', the
 * Cholesky factor of Z'Z is the R factor of the reference's QR of U^-T D'); K, dbias, bias and the value function do not depend on the
 * choice of null-space basis, so the results equal the Householder-QR projection up to rounding. A dependent constraint row sets
 * O2C_STATUS_CONSTRAINT_RANK (the reference clamps the pivot, LinearAlgebra.cpp:38-47): the results of that problem are then
 * regularised, not the reference's.
 */
#ifndef OCS2_DDP_CUDA_H_
#define OCS2_DDP_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O2C_ABI_VERSION 4

typedef enum o2c_error {
  O2C_OK = 0,
  O2C_ERR_INVALID_ARGUMENT = 1,
  O2C_ERR_UNSUPPORTED = 2, /* e.g. the CHOLESKY_MODIFICATION hessian correction */
  O2C_ERR_CUDA = 3,
  O2C_ERR_OUT_OF_MEMORY = 4,
  O2C_ERR_NOT_READY = 5 /* rollout before backward, download before compute, ... */
} o2c_error;

/* values mirror the reference enums */
enum { O2C_ALG_ILQR = 0, O2C_ALG_SLQ = 1 };                 /* ddp::Algorithm (ocs2_ddp/include/ocs2_ddp/DDP_Settings.h) */
enum { O2C_FORM_FULL = 0, O2C_FORM_REDUCED = 1 };           /* ddp preComputeRiccatiTerms (ILQR.cpp:68, SLQ.cpp:65) */
enum { O2C_STRATEGY_LINE_SEARCH = 0, O2C_STRATEGY_LEVENBERG_MARQUARDT = 1 }; /* search_strategy::Type */
enum {                                                      /* hessian_correction::Strategy (HessianCorrection.h:44-49) */
  O2C_HC_DIAGONAL_SHIFT = 0,
  O2C_HC_CHOLESKY_MODIFICATION = 1,   /* unsupported */
  O2C_HC_EIGENVALUE_MODIFICATION = 2, /* LinearAlgebra::makePsdEigenvalue, LinearAlgebra.cpp:52-72 (generic kernels) */
  O2C_HC_GERSHGORIN_MODIFICATION = 3
};

/* per-problem status bits (replace the reference's exceptions, SURVEY.md §5 "failure detection") */
enum {
  O2C_STATUS_OK = 0,
  O2C_STATUS_CHOL_NOT_PD = 1,    /* Hm = R + B'SB not positive definite (LinearAlgebra.cpp:119-124 would silently continue) */
  O2C_STATUS_NONFINITE = 2,      /* non-finite gains / value function / rollout (GaussNewtonDDP.cpp:621-636, DDP_HelperFunctions.cpp:132-134) */
  O2C_STATUS_CONSTRAINT_RANK = 4, /* |Rc_ii| clamped to 1e-9 in the constraint QR (LinearAlgebra.cpp:38-47) */
  O2C_STATUS_NOT_PSD = 8          /* set by o2c_check_numerical_stability only: some S_k fails checkBeingPSD (GaussNewtonDDP.cpp:555-579) */
};

/* o2c_lq_view.flags */
enum {
  /* Q, R and Qf hold only their upper triangle, packed column by column (element (i, j), i <= j, at j (j + 1) / 2 + i; blocks of
   * n (n + 1) / 2 and m (m + 1) / 2 doubles): the order of ContinuousTimeRiccatiEquations::convert2Vector
   * (ocs2_ddp/src/riccati_equations/ContinuousTimeRiccatiEquations.cpp:55-108) and of Eigen's triangularView<Upper> traversal. The
   * cost Hessians are symmetric, so a host caller moves 19 % fewer bytes over PCIe for the legged shape. Strides are in doubles of
   * the packed blocks. Honoured by o2c_upload, o2c_import_device and o2c_solve_host. */
  O2C_LQ_SYMMETRIC_PACKED = 1
};

typedef struct o2c_config {
  int32_t nx;                 /* state dimension n */
  int32_t nu;                 /* input dimension m */
  int32_t nc_max;             /* max number of state-input equality constraints per node (0 = unconstrained) */
  int32_t num_stages;         /* N: time nodes 0..N; ILQR consumes stage data of nodes 0..N-1, SLQ of nodes 0..N */
  int32_t batch;              /* number of independent problems held by the handle */
  int32_t algorithm;          /* O2C_ALG_* */
  int32_t riccati_form;       /* O2C_FORM_* */
  int32_t strategy;           /* O2C_STRATEGY_* */
  int32_t hessian_correction; /* O2C_HC_* (lineSearch.hessianCorrectionStrategy) */
  int32_t device;             /* CUDA device ordinal */
  int32_t max_alphas;         /* capacity for simultaneous rollout step lengths (line search), >= 1 */
  int32_t has_nominal;        /* 1: x_nom/u_nom are provided; 0: nominal trajectories are zero (deviation coordinates) */
  double hessian_multiple;    /* lineSearch.hessianCorrectionMultiple */
  double lm_riccati_multiple; /* levenbergMarquardt riccatiMultiple */
  double time_step;           /* ddp timeStep (SLQ RK4 integrate_times) and rollout timeStep (continuous rollout) */
} o2c_config;

/* One field of a batch: block(problem, node) starts at ptr + problem*problem_stride + node*node_stride (strides in doubles).
 * ptr == NULL means "absent". */
typedef struct o2c_field {
  double* ptr;
  int64_t problem_stride;
  int64_t node_stride;
} o2c_field;

/* Struct-of-arrays view of the LQ data (host or device memory, as stated by the function taking it). Field names follow
 * ModelData: dynamics.dfdx/dfdu, dynamicsBias, cost.dfdxx/dfdux/dfduu/dfdx/dfdu/f, stateInputEqConstraint.dfdx/dfdu/f. */
typedef struct o2c_lq_view {
  o2c_field A, B, Hv;          /* n*n, n*m, n */
  o2c_field Q, P, R, q, r, c;  /* n*n, m*n, m*m, n, m, 1 */
  o2c_field C, D, e;           /* nc_max*n, nc_max*m, nc_max; column-major with leading dimension nc_max */
  const int32_t* nc;           /* active constraints per (problem, node); NULL => nc_max everywhere */
  int64_t nc_problem_stride, nc_node_stride;
  o2c_field Qf, qf, cf;        /* terminal value function (already Hessian-corrected); node_stride ignored */
  o2c_field x_nom, u_nom;      /* nominal trajectories, N+1 nodes of n / m (only read when has_nominal) */
  o2c_field x0;                /* initial state for the rollout, n per problem; node_stride ignored */
  const double* time;          /* N+1 node times shared by the batch (required for SLQ; ILQR: optional) */
  /* ILQR events (ILQR.cpp:263-295, RiccatiTransversalityConditions.h:40-56): event[problem][node] != 0 marks a PRE-EVENT node
   * (time[node] == time[node+1], node+1 in postEventIndices_). Its A, Hv, Q, q, c blocks hold the jump ModelData
   * (modelDataEventTimes: jump map linearisation x+ = A_e dx + Hv_e and pre-jump cost), its B, P, R, r, C, D, e the regular model
   * data of the node, which only shape the controller entry. NULL = no events. Host memory in o2c_upload / o2c_solve_host, device
   * memory in o2c_import_device. With a pre-event node at the last stage (node N-1) the controller entry of node N is still
   * the copy of node N-1: the reference keeps the entry built from node N's OWN ModelData in that case (GaussNewtonDDP.cpp:609-618,
   * ILQR.cpp:196-209), which an ILQR handle does not carry (stage data of nodes 0..N-1 only). LEVENBERG_MARQUARDT handles
   * do not take ILQR events (O2C_ERR_UNSUPPORTED from o2c_backward): deltaGm / deltaGv of a pre-event node are built from the node's
   * regular dynamics (ILQR.cpp:263-295), which this layout replaces by the jump map. SLQ handles: see jump_* below. */
  const int32_t* event;
  int64_t event_problem_stride, event_node_stride;
  /* SLQ events (SLQ.cpp:256-302, ContinuousTimeRiccatiEquations.cpp:135-147): event[problem][node] != 0 marks a pre-event node k whose
   * successor k+1 is the post-event node (stamped time[k] + weakEpsilon by the reference's rollouts, RolloutBase.cpp:62-64; the stamps
   * must differ). The time grid is shared by the batch, so the flags must be the same for every problem. Every node keeps its own
   * continuous-time model data; the jump ModelData of the e-th event in node order (modelDataEventTimes[e]: dynamics.dfdx,
   * dynamicsBias, cost.dfdxx, cost.dfdx, cost.f) are block(problem, e) = ptr + problem*problem_stride + e*node_stride of the fields
   * below. The backward pass integrates the inter-event segments separately and joins them with riccatiTransversalityConditions; the
   * rollout restarts weakEpsilon after every event from x+ = x_nom(k+1) + A_e (x - x_nom(k)) + Hv_e. Host memory (o2c_upload,
   * o2c_solve_host); o2c_import_device does not take SLQ events. */
  o2c_field jump_A, jump_Hv, jump_Q, jump_q, jump_c; /* n*n, n, n*n, n, 1; jump_A and jump_Q required with SLQ events */
  int32_t flags;                                      /* O2C_LQ_* bits, 0 = dense blocks everywhere */
} o2c_lq_view;

/* Struct-of-arrays view of the solution. Controller and value function have N+1 nodes (node N of the controller is the copy of
 * node N-1, GaussNewtonDDP.cpp:609-618; node N of the value function is the terminal condition, :526). */
typedef struct o2c_solution_view {
  o2c_field K, dbias, bias;  /* LinearController gainArray_ (m*n), deltaBiasArray_ (m), biasArray_ (m) */
  o2c_field Sm, Sv, s;       /* valueFunctionTrajectory dfdxx (n*n), dfdx (n), f (1) */
  o2c_field x, u;            /* rollout: out_nodes nodes of n / m per (alpha, problem); see alpha_stride */
  int64_t x_alpha_stride, u_alpha_stride; /* stride (doubles) between rollouts of consecutive step lengths */
  int32_t* status;           /* per-problem O2C_STATUS_* bits */
} o2c_solution_view;

typedef struct o2c_handle o2c_handle;

#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

/* ---- life cycle ---- */
int o2c_abi_version(void);
const char* o2c_last_error(void);
o2c_error o2c_create(const o2c_config* config, o2c_handle** handle);
o2c_error o2c_destroy(o2c_handle* handle);
o2c_error o2c_get_config(const o2c_handle* handle, o2c_config* config);
/* number of CUDA devices visible to the process (0 and O2C_ERR_CUDA when there is none): the valid range of o2c_config.device */
o2c_error o2c_device_count(int32_t* count);
/* levenbergMarquardt riccatiMultiple of the NEXT backward pass. The reference's strategy adapts it from iteration to iteration
 * (LevenbergMarquardtStrategy.cpp: lmModule_.riccatiMultiple); the handle otherwise keeps o2c_config.lm_riccati_multiple. */
o2c_error o2c_set_lm_riccati_multiple(o2c_handle* handle, double riccati_multiple);
o2c_error o2c_sync(o2c_handle* handle);
/* the CUDA stream (cudaStream_t) all compute of this handle is enqueued on; time it with CUDA events on this stream */
o2c_error o2c_compute_stream(o2c_handle* handle, void** stream);

/* ---- data movement ---- */
/* Library-owned, device-resident views (the layout the kernels consume: one interleaved record per (problem, node)).
 * Producers that already live on the device (an LQ approximator, the synthetic generator) write through these. */
o2c_error o2c_device_lq_view(o2c_handle* handle, o2c_lq_view* view);
o2c_error o2c_device_solution_view(o2c_handle* handle, o2c_solution_view* view);
/* number of output nodes of a rollout (N+1 for ILQR; the RK4 step schedule of the continuous rollout for SLQ) */
o2c_error o2c_rollout_num_nodes(o2c_handle* handle, int32_t* out_nodes);
/* rollout output times (out_nodes doubles, host memory) */
o2c_error o2c_rollout_times(o2c_handle* handle, double* times);

/* host SoA -> device for problems [problem_begin, problem_begin+problem_count); the view is indexed from problem 0 of the
 * host arrays (block of handle problem p is read at host index p - problem_begin). Asynchronous w.r.t. the host only when
 * the host memory is pinned; ordered before subsequent compute calls. */
o2c_error o2c_upload(o2c_handle* handle, const o2c_lq_view* host_view, int32_t problem_begin, int32_t problem_count);
/* same, from device memory in an arbitrary strided SoA layout */
o2c_error o2c_import_device(o2c_handle* handle, const o2c_lq_view* device_view, int32_t problem_begin, int32_t problem_count);
/* device -> host SoA; fields with ptr == NULL are skipped; n_alpha rollouts are copied */
o2c_error o2c_download(o2c_handle* handle, const o2c_solution_view* host_view, int32_t problem_begin, int32_t problem_count,
                       int32_t n_alpha);
o2c_error o2c_set_time(o2c_handle* handle, const double* host_time /* N+1 */);

/* ---- compute (asynchronous, enqueued on the compute stream) ---- */
/* backward pass + controller for problems [begin, begin+count) */
o2c_error o2c_backward(o2c_handle* handle, int32_t problem_begin, int32_t problem_count);
/* forward rollouts of the LQ model for n_alpha step lengths (host array `alphas`) from the resident x0 */
o2c_error o2c_rollout(o2c_handle* handle, const double* alphas, int32_t n_alpha, int32_t problem_begin, int32_t problem_count);
/* backward + one rollout with step length alpha (one "LQ solve" of the benchmark metric) */
o2c_error o2c_solve(o2c_handle* handle, double alpha, int32_t problem_begin, int32_t problem_count);
/* ddp::Settings::checkNumericalStability_ for problems [begin, begin+count) after o2c_backward: every S_k of the value function goes
 * through the reference's checkBeingPSD (ocs2_core/src/Types.cpp:206-236 as called from GaussNewtonDDP.cpp:555-579): finite,
 * self-adjoint to 1e-6 (Eigen isApprox), smallest eigenvalue >= -epsilon. A failing problem gets O2C_STATUS_NOT_PSD (non-finite
 * entries: O2C_STATUS_NONFINITE) or-ed into its status word, where the reference would throw. Asynchronous on the compute stream. */
o2c_error o2c_check_numerical_stability(o2c_handle* handle, int32_t problem_begin, int32_t problem_count);
/* number of kernel launches enqueued by this handle so far (bench.py's gpu_launches) */
o2c_error o2c_launch_count(const o2c_handle* handle, int64_t* launches);
/* name of the sweep kernel variant that o2c_backward dispatches to for this config (diagnostics / profiles) */
const char* o2c_kernel_variant(const o2c_handle* handle);

/* ---- the step after: batched Armijo line search on the LQ model ----
 * LineSearchStrategy::run / lineSearchTask (ocs2_ddp/src/search_strategy/LineSearchStrategy.cpp:125-258): the candidates
 * alpha_e = max_step_length * contraction_rate^e, as long as alpha_e >= min_step_length (numerics::almost_ge), are rolled out on the
 * LQ model (all candidates of all problems in one launch; at most max_alphas of them), the merit of a rollout is its LQ-model cost
 * (computeRolloutPerformanceIndex with no constraint terms: the LQ rollout satisfies the linearised constraints), and per problem the
 * largest alpha with  merit < baseline - armijo_coefficient * alpha * IS(deltaBias)  wins; 0 (index -1) if none does.
 * IS = trapezoidal integral of |deltaBias|^2 over the controller time stamps (computeControllerUpdateIS,
 * DDP_HelperFunctions.cpp:285-291). Discrete (ILQR) model only; O2C_ERR_UNSUPPORTED for SLQ. Requires o2c_backward. */
typedef struct o2c_line_search_settings { /* search_strategy::line_search::Settings, StrategySettings.h:85-105 */
  double min_step_length;    /* 0.05 */
  double max_step_length;    /* 1.0  */
  double contraction_rate;   /* 0.5  */
  double armijo_coefficient; /* 1e-4 */
} o2c_line_search_settings;
/* baseline_merit: host array with one merit per problem of the range (the performance index of the nominal trajectory), or NULL =
 * the LQ cost of the zero-deviation trajectory (sum of the constants c_k + cf). The rollouts of all candidates stay resident (x, u of
 * candidate e are rollout e of o2c_download). */
o2c_error o2c_line_search(o2c_handle* handle, const o2c_line_search_settings* settings, const double* baseline_merit, int32_t problem_begin,
                          int32_t problem_count);
/* results of the last o2c_line_search for problems [begin, begin+count) (host arrays, any may be NULL): chosen step length and
 * candidate index per problem, merits [n_candidates][problem_count], baseline and IS per problem, the candidate step lengths
 * (capacity max_alphas) and their number */
o2c_error o2c_line_search_result(o2c_handle* handle, double* step_length, int32_t* candidate_index, double* merits, double* baseline,
                                 double* update_is, double* candidates, int32_t* n_candidates, int32_t problem_begin, int32_t problem_count);

/* ---- the data format after the path: the flattened policy of ocs2_msgs/mpc_flattened_controller ----
 * LinearController::flatten at the controller's own time stamps (ocs2_core/src/control/LinearController.cpp:87-140, called from
 * MPC_ROS_Interface::createMpcPolicyMsg, ocs2_ros_interfaces/src/mpc/MPC_ROS_Interface.cpp:175): per node one float32 record of
 * m*(n+1) values, row i = [uff_i, K_i,:] (row-major), with uff = bias + step_length * deltaBias, i.e. the controller after
 * incrementController(step_length). host_out: [problem_count][N+1][m*(n+1)] floats; the conversion runs on the device, so a consumer
 * that only publishes the policy moves 4*m*(n+1) bytes per node instead of the FP64 controller arrays. Blocking. */
o2c_error o2c_download_flattened_controller(o2c_handle* handle, float* host_out, double step_length, int32_t problem_begin,
                                            int32_t problem_count);

/* ---- the step before: ILQR discretisation of continuous-time linearisations ----
 * ILQR::discreteLQWorker (ocs2_ddp/src/ILQR.cpp:137-157) with rk4SensitivityDiscretization
 * (ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:130-169): the discrete A, B of the interval after node k are assembled from
 * the four stage linearisations (dfdx, dfdu) of one RK4 step of length dt_k — k1 at (t, x), k2 at (t + dt/2, x + dt/2 f1), k3 at
 * (t + dt/2, x + dt/2 f2), k4 at (t + dt, x + dt f3). The stage evaluations need the caller's system dynamics; this call does the
 * batched dense part on the device: the input sensitivity chain (k2.dfdu += dt/2 k2.dfdx k1.dfdu, ...), the state sensitivity chain
 * (k2.dfdx += dt/2 k2.dfdx k1.dfdx, ...), the assembly A = I + dt/6 k1 + dt/3 k2 + dt/3 k3 + dt/6 k4 (B alike), Hv := 0, and
 * (scale_cost != 0) modelData.cost *= dt_k on the resident Q, P, R, q, r, c of the node. A node with dt_k == 0 keeps the
 * continuous-time data (k1) unscaled (ILQR.cpp:123-130). ILQR handles only; all pointers of the view are DEVICE memory except dt. */
typedef struct o2c_discretization_view {
  o2c_field dfdx[4], dfdu[4]; /* n*n and n*m blocks per (problem, node) of the stages k1..k4; with stages == 1 only [0] is read */
  const double* dt;           /* HOST array of num_stages step lengths, or NULL = time[k+1] - time[k] of the handle's time grid */
  int32_t stages;             /* 4, or 1 for a model that is constant over the step (k1 = k2 = k3 = k4) */
} o2c_discretization_view;
o2c_error o2c_discretize(o2c_handle* handle, const o2c_discretization_view* device_view, int32_t scale_cost, int32_t problem_begin,
                         int32_t problem_count);

/* ---- end-to-end convenience: host buffers in, host buffers out, chunked H2D / compute / D2H pipeline ---- */
o2c_error o2c_solve_host(o2c_handle* handle, const o2c_lq_view* host_lq, const o2c_solution_view* host_solution, double alpha,
                         int32_t problem_count, int32_t chunk);

/* ---- synthetic data (benchmarks / parity tests): seeded counter-based generator, bit-identical to oracle's
 * orc_generate_problem; problem p of the handle gets global problem index first_problem_index + p ---- */
o2c_error o2c_generate_synthetic(o2c_handle* handle, uint64_t seed, int64_t first_problem_index, double dt);

/* pinned host memory helpers for callers without a CUDA runtime of their own */
o2c_error o2c_host_alloc(void** ptr, uint64_t bytes);
o2c_error o2c_host_free(void* ptr);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* OCS2_DDP_CUDA_H_ */
