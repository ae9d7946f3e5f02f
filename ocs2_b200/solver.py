"""Python host side over the C ABI, mirroring the reference's interface for the batched LQ path.

Names follow the reference: ``Settings`` carries the ``ddp::Settings`` fields that change the hot-path arithmetic
(ocs2_ddp/include/ocs2_ddp/DDP_Settings.h:63-120, search_strategy/StrategySettings.h:66-132); ``BatchedLqSolver`` exposes
``solveSequentialRiccatiEquations`` / ``calculateController`` (GaussNewtonDDP.h:167-176) and ``rolloutTrajectory``
(DDP_HelperFunctions.cpp:125-138); ``LinearController`` holds ``timeStamp_/gainArray_/biasArray_/deltaBiasArray_``
(ocs2_core/include/ocs2_core/control/LinearController.h:109-112). Errors raise (the reference throws std::runtime_error).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import lib as _l


@dataclass
class Settings:
    """Subset of ddp::Settings that reaches the LQ arithmetic."""
    algorithm: int = _l.ALG_ILQR                      # ddp.algorithm (ILQR | SLQ)
    preComputeRiccatiTerms: bool = True               # reduced-form Riccati (only with LINE_SEARCH)
    strategy: int = _l.STRATEGY_LINE_SEARCH           # ddp.strategy
    hessianCorrectionStrategy: int = _l.HC_DIAGONAL_SHIFT   # lineSearch.hessianCorrectionStrategy
    hessianCorrectionMultiple: float = 1e-6           # lineSearch.hessianCorrectionMultiple (numeric_traits::limitEpsilon)
    riccatiMultiple: float = 0.0                      # levenbergMarquardt riccatiMultiple
    timeStep: float = 1e-2                            # ddp.timeStep / rollout.timeStep
    backwardPassIntegratorType: str = "RK4"           # only RK4 is provided on the device
    nThreads: int = 1                                 # ignored: parallelism comes from the batch

    @property
    def reduced_form(self) -> bool:
        return bool(self.preComputeRiccatiTerms) and self.strategy == _l.STRATEGY_LINE_SEARCH


def _colmajor(a: np.ndarray) -> np.ndarray:
    """(..., rows, cols) natural -> contiguous buffer whose trailing block is column-major."""
    return np.ascontiguousarray(np.swapaxes(np.asarray(a, dtype=np.float64), -1, -2))


def pack_upper(a: np.ndarray) -> np.ndarray:
    """(..., k, k) symmetric -> (..., k (k + 1) / 2): upper triangle column by column, element (i, j), i <= j, at j (j + 1) / 2 + i
    (O2C_LQ_SYMMETRIC_PACKED; the order of ContinuousTimeRiccatiEquations::convert2Vector)."""
    k = a.shape[-1]
    rows = np.concatenate([np.arange(j + 1) for j in range(k)])
    cols = np.concatenate([np.full(j + 1, j) for j in range(k)])
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)[..., rows, cols])


def _field(arr: Optional[np.ndarray], block: int, nodes: int) -> _l.Field:
    if arr is None:
        return _l.Field(None, 0, 0)
    assert arr.flags["C_CONTIGUOUS"] and arr.dtype == np.float64
    return _l.Field(arr.ctypes.data, nodes * block, block)


@dataclass
class LqBatch:
    """Host LQ data of a batch in natural numpy layout (ModelData fields, ModelData.h:43-60):
    A (B,nodes,n,n) B (B,nodes,n,m) Hv (B,nodes,n) Q (B,nodes,n,n) P (B,nodes,m,n) R (B,nodes,m,m) q (B,nodes,n) r (B,nodes,m)
    c (B,nodes) C (B,nodes,ncmax,n) D (B,nodes,ncmax,m) e (B,nodes,ncmax) nc (B,nodes) int32 Qf (B,n,n) qf (B,n) cf (B,)
    x_nom (B,N+1,n) u_nom (B,N+1,m) x0 (B,n) time (N+1,) event (B,nodes) int. nodes = N for ILQR, N+1 for SLQ."""
    A: np.ndarray
    B: np.ndarray
    Q: np.ndarray
    R: np.ndarray
    Qf: np.ndarray
    Hv: Optional[np.ndarray] = None
    P: Optional[np.ndarray] = None
    q: Optional[np.ndarray] = None
    r: Optional[np.ndarray] = None
    c: Optional[np.ndarray] = None
    C: Optional[np.ndarray] = None
    D: Optional[np.ndarray] = None
    e: Optional[np.ndarray] = None
    nc: Optional[np.ndarray] = None
    qf: Optional[np.ndarray] = None
    cf: Optional[np.ndarray] = None
    x_nom: Optional[np.ndarray] = None
    u_nom: Optional[np.ndarray] = None
    x0: Optional[np.ndarray] = None
    time: Optional[np.ndarray] = None
    event: Optional[np.ndarray] = None   # (B, nodes) int: 1 marks a pre-event node. ILQR: jump data in the node's A, Hv, Q, q, c
    # SLQ: jump ModelData of the e-th event in node order (the same event nodes for every problem: the time grid is shared)
    jump_A: Optional[np.ndarray] = None   # (B, E, n, n)
    jump_Hv: Optional[np.ndarray] = None  # (B, E, n)
    jump_Q: Optional[np.ndarray] = None   # (B, E, n, n)
    jump_q: Optional[np.ndarray] = None   # (B, E, n)
    jump_c: Optional[np.ndarray] = None   # (B, E)
    _keep: dict = field(default_factory=dict, repr=False)

    @property
    def batch(self) -> int:
        return self.A.shape[0]

    def view(self, N: int, symmetric_packed: bool = False) -> _l.LqView:
        """Builds the o2c_lq_view over column-major copies of the arrays (kept alive in self._keep). symmetric_packed: Q, R, Qf travel
        as packed upper triangles (O2C_LQ_SYMMETRIC_PACKED)."""
        k = self._keep
        nodes = self.A.shape[1]
        n, m = self.B.shape[-2], self.B.shape[-1]
        ncm = 0 if self.D is None else self.D.shape[-2]
        for name in ("A", "B", "Q", "P", "R", "C", "D", "Qf"):
            v = getattr(self, name)
            if symmetric_packed and name in ("Q", "R", "Qf"):
                k[name] = pack_upper(v)
            else:
                k[name] = _colmajor(v) if v is not None else None
        for name in ("Hv", "q", "r", "c", "e", "qf", "cf", "x_nom", "u_nom", "x0", "time"):
            v = getattr(self, name)
            k[name] = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
        k["nc"] = np.ascontiguousarray(self.nc, dtype=np.int32) if self.nc is not None else None
        lv = _l.LqView()
        lv.A = _field(k["A"], n * n, nodes)
        lv.B = _field(k["B"], n * m, nodes)
        lv.Hv = _field(k["Hv"], n, nodes)
        lv.Q = _field(k["Q"], n * (n + 1) // 2 if symmetric_packed else n * n, nodes)
        lv.P = _field(k["P"], m * n, nodes)
        lv.R = _field(k["R"], m * (m + 1) // 2 if symmetric_packed else m * m, nodes)
        lv.flags = _l.LQ_SYMMETRIC_PACKED if symmetric_packed else 0
        lv.q = _field(k["q"], n, nodes)
        lv.r = _field(k["r"], m, nodes)
        lv.c = _field(k["c"], 1, nodes)
        if ncm:
            lv.C = _field(k["C"], ncm * n, nodes)
            lv.D = _field(k["D"], ncm * m, nodes)
            lv.e = _field(k["e"], ncm, nodes)
            if k["nc"] is not None:
                lv.nc = k["nc"].ctypes.data
                lv.nc_problem_stride = nodes
                lv.nc_node_stride = 1
        lv.Qf = _field(k["Qf"], n * (n + 1) // 2 if symmetric_packed else n * n, 1)
        lv.qf = _field(k["qf"], n, 1)
        lv.cf = _field(k["cf"], 1, 1)
        lv.x_nom = _field(k["x_nom"], n, N + 1)
        lv.u_nom = _field(k["u_nom"], m, N + 1)
        lv.x0 = _field(k["x0"], n, 1)
        lv.time = k["time"].ctypes.data if k["time"] is not None else None
        k["event"] = np.ascontiguousarray(self.event, dtype=np.int32) if self.event is not None else None
        if k["event"] is not None:
            lv.event = k["event"].ctypes.data
            lv.event_problem_stride = nodes
            lv.event_node_stride = 1
        if self.jump_A is not None:
            E = self.jump_A.shape[1]
            for name in ("jump_A", "jump_Q"):
                k[name] = _colmajor(getattr(self, name))
            for name in ("jump_Hv", "jump_q", "jump_c"):
                v = getattr(self, name)
                k[name] = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
            lv.jump_A, lv.jump_Q = _field(k["jump_A"], n * n, E), _field(k["jump_Q"], n * n, E)
            lv.jump_Hv, lv.jump_q, lv.jump_c = _field(k["jump_Hv"], n, E), _field(k["jump_q"], n, E), _field(k["jump_c"], 1, E)
        return lv


@dataclass
class LineSearchSettings:
    """search_strategy::line_search::Settings (ocs2_ddp/include/ocs2_ddp/search_strategy/StrategySettings.h:85-105)."""
    minStepLength: float = 0.05
    maxStepLength: float = 1.0
    contractionRate: float = 0.5
    armijoCoefficient: float = 1e-4


@dataclass
class LineSearchResult:
    stepLength: np.ndarray      # (count,) chosen step length per problem, 0 when no candidate satisfies the Armijo condition
    candidateIndex: np.ndarray  # (count,) index into `candidates`, -1 when none
    merits: np.ndarray          # (n_candidates, count) LQ-model cost of every candidate rollout
    baselineMerit: np.ndarray   # (count,)
    controllerUpdateIS: np.ndarray  # (count,) trapezoidal integral of |deltaBias|^2 (computeControllerUpdateIS)
    candidates: np.ndarray      # (n_candidates,) step lengths, largest first


@dataclass
class LinearController:
    """LinearController arrays for the batch: timeStamp_ (N+1,), gainArray_ (B,N+1,m,n), biasArray_ (B,N+1,m), deltaBiasArray_ (B,N+1,m)."""
    timeStamp_: np.ndarray
    gainArray_: np.ndarray
    biasArray_: np.ndarray
    deltaBiasArray_: np.ndarray


@dataclass
class Solution:
    controller: LinearController
    Sm: np.ndarray       # valueFunctionTrajectory dfdxx (B,N+1,n,n)
    Sv: np.ndarray       # dfdx (B,N+1,n)
    s: np.ndarray        # f (B,N+1)
    status: np.ndarray   # (B,) int32 O2C_STATUS_* bits
    x: Optional[np.ndarray] = None   # (n_alpha,B,out_nodes,n)
    u: Optional[np.ndarray] = None   # (n_alpha,B,out_nodes,m)
    t: Optional[np.ndarray] = None   # (out_nodes,)

    @property
    def K(self):
        return self.controller.gainArray_

    @property
    def dbias(self):
        return self.controller.deltaBiasArray_

    @property
    def bias(self):
        return self.controller.biasArray_


class BatchedLqSolver:
    """One handle of libocs2_ddp_cuda.so: `batch` independent LQ problems resident on one CUDA device."""

    def __init__(self, settings: Settings, nx: int, nu: int, num_stages: int, batch: int, nc_max: int = 0, device: int = 0,
                 has_nominal: bool = False, max_alphas: int = 1):
        if settings.algorithm == _l.ALG_SLQ and settings.backwardPassIntegratorType != "RK4":
            raise _l.O2cError(2, "only the fixed-step RK4 backward pass is provided (backwardPassIntegratorType must be RK4)")
        self._lib = _l.load_library()
        self.settings = settings
        self.nx, self.nu, self.N, self.batch, self.nc_max = nx, nu, num_stages, batch, nc_max
        self.nodes = num_stages + 1 if settings.algorithm == _l.ALG_SLQ else num_stages
        cfg = _l.Config(nx=nx, nu=nu, nc_max=nc_max, num_stages=num_stages, batch=batch, algorithm=settings.algorithm,
                        riccati_form=_l.FORM_REDUCED if settings.reduced_form else _l.FORM_FULL, strategy=settings.strategy,
                        hessian_correction=settings.hessianCorrectionStrategy, device=device, max_alphas=max_alphas,
                        has_nominal=int(has_nominal), hessian_multiple=settings.hessianCorrectionMultiple,
                        lm_riccati_multiple=settings.riccatiMultiple, time_step=settings.timeStep)
        self._h = C.c_void_p()
        _l.check(self._lib.o2c_create(C.byref(cfg), C.byref(self._h)))
        self.max_alphas = max_alphas
        self._n_alpha = 0

    # ---- life cycle -------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.o2c_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def kernel_variant(self) -> str:
        return self._lib.o2c_kernel_variant(self._h).decode()

    @property
    def launch_count(self) -> int:
        v = C.c_int64(0)
        _l.check(self._lib.o2c_launch_count(self._h, C.byref(v)))
        return v.value

    @property
    def compute_stream(self) -> int:
        s = C.c_void_p()
        _l.check(self._lib.o2c_compute_stream(self._h, C.byref(s)))
        return s.value or 0

    def sync(self):
        _l.check(self._lib.o2c_sync(self._h))

    # ---- data -------------------------------------------------------------------------------------------------------
    def upload(self, lq: LqBatch, problem_begin: int = 0, symmetric_packed: bool = False):
        view = lq.view(self.N, symmetric_packed)
        _l.check(self._lib.o2c_upload(self._h, C.byref(view), problem_begin, lq.batch))
        self.sync()

    def import_device(self, view: _l.LqView, problem_begin: int = 0, problem_count: Optional[int] = None):
        """o2c_import_device: LQ data that already live in device memory, in any strided SoA layout (pointers of `view` are device
        pointers; `view.time`, if set, too)."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_import_device(self._h, C.byref(view), problem_begin, cnt))

    def discretize(self, view: "_l.DiscretizationView", scale_cost: bool = True, problem_begin: int = 0, problem_count: Optional[int] = None):
        """o2c_discretize: ILQR::discreteLQWorker on caller-supplied stage linearisations in device memory (rk4SensitivityDiscretization):
        writes A, B, Hv = 0 of the resident records and scales the resident cost blocks by the step length."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_discretize(self._h, C.byref(view), 1 if scale_cost else 0, problem_begin, cnt))

    def set_time(self, time: Sequence[float]):
        t = np.ascontiguousarray(time, dtype=np.float64)
        assert t.shape == (self.N + 1,)
        _l.check(self._lib.o2c_set_time(self._h, t.ctypes.data_as(C.POINTER(C.c_double))))

    def generate_synthetic(self, seed: int, first_problem_index: int = 0, dt: float = 0.01):
        _l.check(self._lib.o2c_generate_synthetic(self._h, seed, first_problem_index, dt))

    def device_lq_view(self) -> _l.LqView:
        v = _l.LqView()
        _l.check(self._lib.o2c_device_lq_view(self._h, C.byref(v)))
        return v

    def device_solution_view(self) -> _l.SolutionView:
        v = _l.SolutionView()
        _l.check(self._lib.o2c_device_solution_view(self._h, C.byref(v)))
        return v

    @property
    def rollout_num_nodes(self) -> int:
        v = C.c_int32(0)
        _l.check(self._lib.o2c_rollout_num_nodes(self._h, C.byref(v)))
        return v.value

    def rollout_times(self) -> np.ndarray:
        t = np.zeros(self.rollout_num_nodes)
        _l.check(self._lib.o2c_rollout_times(self._h, t.ctypes.data_as(C.POINTER(C.c_double))))
        return t

    # ---- compute (asynchronous on the compute stream) -----------------------------------------------------------------
    def solveSequentialRiccatiEquations(self, problem_begin: int = 0, problem_count: Optional[int] = None):
        """Backward pass of every problem: ILQR::/SLQ::solveSequentialRiccatiEquations + calculateController."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_backward(self._h, problem_begin, cnt))

    def calculateController(self):
        """The controller (K, bias, deltaBias) is produced by the same kernel as the backward pass; kept for API parity."""
        return None

    def rolloutTrajectory(self, alphas: Sequence[float] = (1.0,), problem_begin: int = 0, problem_count: Optional[int] = None):
        """incrementController(alpha) + LQ-model rollout for every step length in `alphas`."""
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_rollout(self._h, a.ctypes.data_as(C.POINTER(C.c_double)), len(a), problem_begin, cnt))
        self._n_alpha = len(a)

    def setRiccatiMultiple(self, riccatiMultiple: float):
        """levenbergMarquardt riccatiMultiple of the next backward pass (the reference's strategy adapts it every iteration)."""
        _l.check(self._lib.o2c_set_lm_riccati_multiple(self._h, float(riccatiMultiple)))
        self.settings.riccatiMultiple = float(riccatiMultiple)

    def checkNumericalStability(self, problem_begin: int = 0, problem_count: Optional[int] = None):
        """ddp::Settings::checkNumericalStability_ for the value function (GaussNewtonDDP.cpp:555-579): checkBeingPSD of every S_k; a
        failing problem gets STATUS_NOT_PSD in its status word (the reference throws). Run after the backward pass."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_check_numerical_stability(self._h, problem_begin, cnt))

    def solve(self, alpha: float = 1.0, problem_begin: int = 0, problem_count: Optional[int] = None):
        """backward pass + one rollout: one 'LQ solve' per problem (the benchmark metric)."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        _l.check(self._lib.o2c_solve(self._h, alpha, problem_begin, cnt))
        self._n_alpha = 1

    def lineSearch(self, settings: Optional[LineSearchSettings] = None, baselineMerit: Optional[np.ndarray] = None, problem_begin: int = 0,
                   problem_count: Optional[int] = None) -> LineSearchResult:
        """LineSearchStrategy::run on the LQ model: every candidate step length of every problem is rolled out in one launch and the
        largest one satisfying the Armijo condition is chosen per problem (o2c_line_search)."""
        ls = settings or LineSearchSettings()
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        cs = _l.LineSearchSettings(ls.minStepLength, ls.maxStepLength, ls.contractionRate, ls.armijoCoefficient)
        base = None if baselineMerit is None else np.ascontiguousarray(baselineMerit, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        _l.check(self._lib.o2c_line_search(self._h, C.byref(cs), base.ctypes.data_as(dp) if base is not None else None, problem_begin, cnt))
        step, idx = np.zeros(cnt), np.zeros(cnt, dtype=np.int32)
        merits, bl, upd, cand = np.zeros((self.max_alphas, cnt)), np.zeros(cnt), np.zeros(cnt), np.zeros(self.max_alphas)
        nc = C.c_int32(0)
        _l.check(self._lib.o2c_line_search_result(self._h, step.ctypes.data_as(dp), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  merits.ctypes.data_as(dp), bl.ctypes.data_as(dp), upd.ctypes.data_as(dp),
                                                  cand.ctypes.data_as(dp), C.byref(nc), problem_begin, cnt))
        self._n_alpha = nc.value
        # o2c_line_search_result packs the merits as [n_candidates][count]
        return LineSearchResult(step, idx, merits.reshape(-1)[:nc.value * cnt].reshape(nc.value, cnt), bl, upd, cand[:nc.value])

    def flatten(self, stepLength: float = 1.0, problem_begin: int = 0, problem_count: Optional[int] = None) -> np.ndarray:
        """LinearController::flatten at the controller's own time stamps, after incrementController(stepLength): float32
        (count, N+1, m*(n+1)), rows [uff_i, K_i,:] — the payload of ocs2_msgs/mpc_flattened_controller. Converted on the device."""
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        out = np.zeros((cnt, self.N + 1, self.nu * (self.nx + 1)), dtype=np.float32)
        _l.check(self._lib.o2c_download_flattened_controller(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), float(stepLength), problem_begin, cnt))
        return out

    # ---- results ----------------------------------------------------------------------------------------------------
    def _solution_buffers(self, count: int, n_alpha: int, want_value: bool = True):
        n, m, N = self.nx, self.nu, self.N
        on = self.rollout_num_nodes
        bufs = dict(K=np.zeros((count, N + 1, n, m)), dbias=np.zeros((count, N + 1, m)), bias=np.zeros((count, N + 1, m)),
                    Sm=np.zeros((count, N + 1, n, n)) if want_value else None, Sv=np.zeros((count, N + 1, n)) if want_value else None,
                    s=np.zeros((count, N + 1)) if want_value else None, status=np.zeros(count, dtype=np.int32),
                    x=np.zeros((max(n_alpha, 1), count, on, n)) if n_alpha else None,
                    u=np.zeros((max(n_alpha, 1), count, on, m)) if n_alpha else None)
        sv = _l.SolutionView()
        sv.K = _field(bufs["K"], m * n, N + 1)
        sv.dbias = _field(bufs["dbias"], m, N + 1)
        sv.bias = _field(bufs["bias"], m, N + 1)
        sv.Sm = _field(bufs["Sm"], n * n, N + 1)
        sv.Sv = _field(bufs["Sv"], n, N + 1)
        sv.s = _field(bufs["s"], 1, N + 1)
        if n_alpha:
            sv.x = _field(bufs["x"], n, on)
            sv.u = _field(bufs["u"], m, on)
            sv.x_alpha_stride = count * on * n
            sv.u_alpha_stride = count * on * m
        sv.status = bufs["status"].ctypes.data
        return bufs, sv

    def _to_solution(self, bufs, n_alpha) -> Solution:
        ctrl = LinearController(timeStamp_=self.rollout_times() if self.settings.algorithm == _l.ALG_ILQR else None,
                                gainArray_=np.swapaxes(bufs["K"], -1, -2), biasArray_=bufs["bias"], deltaBiasArray_=bufs["dbias"])
        return Solution(controller=ctrl, Sm=np.swapaxes(bufs["Sm"], -1, -2) if bufs["Sm"] is not None else None, Sv=bufs["Sv"], s=bufs["s"],
                        status=bufs["status"], x=bufs["x"], u=bufs["u"], t=self.rollout_times() if n_alpha else None)

    def download(self, problem_begin: int = 0, problem_count: Optional[int] = None, n_alpha: Optional[int] = None) -> Solution:
        cnt = self.batch - problem_begin if problem_count is None else problem_count
        na = self._n_alpha if n_alpha is None else n_alpha
        bufs, sv = self._solution_buffers(cnt, na)
        _l.check(self._lib.o2c_download(self._h, C.byref(sv), problem_begin, cnt, na))
        return self._to_solution(bufs, na)

    def solve_host(self, lq: LqBatch, alpha: float = 1.0, chunk: int = 0, symmetric_packed: bool = False, want_value: bool = True) -> Solution:
        """End to end through host buffers: chunked H2D -> sweep + rollout -> D2H pipeline (o2c_solve_host). symmetric_packed: the cost
        Hessians travel as packed upper triangles; want_value = False skips the value function (Sm, Sv, s) on the way back."""
        view = lq.view(self.N, symmetric_packed)
        bufs, sv = self._solution_buffers(lq.batch, 1, want_value)
        _l.check(self._lib.o2c_solve_host(self._h, C.byref(view), C.byref(sv), alpha, lq.batch, chunk))
        return self._to_solution(bufs, 1)
