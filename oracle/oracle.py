"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The oracle restates the reference's Eigen path (see oracle/lq_oracle.h); it is never used by ocs2_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblq_oracle.so")

ALG_ILQR, ALG_SLQ = 0, 1
STRATEGY_LINE_SEARCH, STRATEGY_LM = 0, 1
HC_DIAGONAL_SHIFT, HC_CHOLESKY_MODIFICATION, HC_EIGENVALUE_MODIFICATION, HC_GERSHGORIN_MODIFICATION = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Settings(C.Structure):
    _fields_ = [
        ("algorithm", C.c_int32),
        ("reduced_form", C.c_int32),
        ("strategy", C.c_int32),
        ("hessian_correction", C.c_int32),
        ("hessian_multiple", C.c_double),
        ("lm_riccati_multiple", C.c_double),
        ("time_step", C.c_double),
    ]


class _Problem(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("nc_max", C.c_int32), ("N", C.c_int32)] + [
        (name, _dp) for name in ("A", "B", "Hv", "Q", "P", "R", "q", "r", "c", "C", "D", "e")
    ] + [("nc", _ip)] + [(name, _dp) for name in ("Qf", "qf", "cf", "x_nom", "u_nom", "time")] + [("event", _ip)] + [
        (name, _dp) for name in ("jA", "jHv", "jQ", "jq", "jc")]


class _Solution(C.Structure):
    _fields_ = [(name, _dp) for name in ("K", "dbias", "bias", "Sm", "Sv", "s")] + [("status", C.c_int32)]


class _Projected(C.Structure):
    _fields_ = [(name, _dp) for name in
                ("At", "Bt", "Hvt", "Qt", "Pt", "Rt", "qt", "rt", "ct", "Cmt", "Evt", "Pu", "dQ", "dGm", "dGv")]


def build(force: bool = False) -> str:
    """Compile oracle/liblq_oracle.so with the committed Makefile (g++ only, no reference sources involved)."""
    src = os.path.join(_HERE, "lq_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liblq_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_inverse_uut.restype = C.c_int
        _lib.orc_shift_hessian.restype = C.c_int
        _lib.orc_project_stage.restype = C.c_int
        _lib.orc_backward.restype = C.c_int
        _lib.orc_rollout.restype = C.c_int
        _lib.orc_discrete_lq_cost.restype = C.c_double
        _lib.orc_baseline_run.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else _dp()


def _f(a):
    """Column-major (Fortran) contiguous float64 copy."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def make_settings(algorithm=ALG_ILQR, reduced_form=True, strategy=STRATEGY_LINE_SEARCH, hessian_correction=HC_DIAGONAL_SHIFT,
                  hessian_multiple=1e-6, lm_riccati_multiple=0.0, time_step=1e-2) -> Settings:
    return Settings(algorithm, int(reduced_form), strategy, hessian_correction, hessian_multiple, lm_riccati_multiple, time_step)


# --------------------------------------------------------------------------------------------------------------------
# single building blocks
# --------------------------------------------------------------------------------------------------------------------
def inverse_uut(H):
    H = _f(H)
    m = H.shape[0]
    Ui = np.zeros((m, m), order="F")
    rc = lib().orc_inverse_uut(m, _p(H), _p(Ui))
    return Ui, rc


def constraint_projection(D, Ui):
    D = _f(D)
    Ui = _f(Ui)
    nc, m = D.shape
    Dd = np.zeros((m, nc), order="F")
    RcInv = np.zeros((nc, nc), order="F")
    Pu = np.zeros((m, m - nc), order="F")
    lib().orc_constraint_projection(m, nc, _p(D), nc, _p(Ui), _p(Dd), _p(RcInv), _p(Pu))
    return Dd, RcInv, Pu


def shift_hessian(strategy, M, eps):
    M = _f(M).copy(order="F")
    rc = lib().orc_shift_hessian(strategy, M.shape[0], _p(M), C.c_double(eps))
    return M, rc


def flatten(Sm, Sv, s):
    Sm = _f(Sm)
    n = Sm.shape[0]
    out = np.zeros(n * (n + 1) // 2 + n + 1)
    lib().orc_flatten(n, _p(Sm), _p(np.ascontiguousarray(Sv, dtype=np.float64)), C.c_double(s), _p(out))
    return out


def unflatten(n, allSs):
    allSs = np.ascontiguousarray(allSs, dtype=np.float64)
    Sm = np.zeros((n, n), order="F")
    Sv = np.zeros(n)
    s = C.c_double(0)
    lib().orc_unflatten(n, _p(allSs), _p(Sm), _p(Sv), C.byref(s))
    return Sm, Sv, s.value


def time_segment(t, time):
    time = np.ascontiguousarray(time, dtype=np.float64)
    idx = C.c_int(0)
    alpha = C.c_double(0)
    lib().orc_time_segment(C.c_double(t), _p(time), len(time), C.byref(idx), C.byref(alpha))
    return idx.value, alpha.value


@dataclass
class ProjectedStage:
    p: int
    status: int
    At: np.ndarray
    Bt: np.ndarray
    Hvt: np.ndarray
    Qt: np.ndarray
    Pt: np.ndarray
    Rt: np.ndarray
    qt: np.ndarray
    rt: np.ndarray
    ct: float
    Cmt: np.ndarray
    Evt: np.ndarray
    Pu: np.ndarray
    dQ: np.ndarray
    dGm: np.ndarray
    dGv: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    def c_struct(self):
        ct = np.array([self.ct])
        self._keep = [ct]
        return _Projected(_p(self.At), _p(self.Bt), _p(self.Hvt), _p(self.Qt), _p(self.Pt), _p(self.Rt), _p(self.qt), _p(self.rt),
                          _p(ct), _p(self.Cmt), _p(self.Evt), _p(self.Pu), _p(self.dQ), _p(self.dGm), _p(self.dGv))


def project_stage(st: Settings, A, B, Hv, Q, P, R, q, r, c, Cm=None, Dm=None, e=None, Sm=None) -> ProjectedStage:
    A, B, Q, P, R = map(_f, (A, B, Q, P, R))
    Hv, q, r = (np.ascontiguousarray(v, dtype=np.float64) for v in (Hv, q, r))
    n, m = B.shape
    nc = 0 if Dm is None else np.asarray(Dm).shape[0]
    if nc:
        Cm, Dm = _f(Cm), _f(Dm)
        e = np.ascontiguousarray(e, dtype=np.float64)
    p = m - nc
    bufs = dict(At=np.zeros((n, n), order="F"), Bt=np.zeros((n, p), order="F"), Hvt=np.zeros(n), Qt=np.zeros((n, n), order="F"),
                Pt=np.zeros((p, n), order="F"), Rt=np.zeros((p, p), order="F"), qt=np.zeros(n), rt=np.zeros(p), ct=np.zeros(1),
                Cmt=np.zeros((m, n), order="F"), Evt=np.zeros(m), Pu=np.zeros((m, p), order="F"), dQ=np.zeros((n, n), order="F"),
                dGm=np.zeros((p, n), order="F"), dGv=np.zeros(p))
    out = _Projected(*[_p(bufs[k]) for k in ("At", "Bt", "Hvt", "Qt", "Pt", "Rt", "qt", "rt", "ct", "Cmt", "Evt", "Pu", "dQ", "dGm", "dGv")])
    status = C.c_int(0)
    SmF = _f(Sm) if Sm is not None else None
    pp = lib().orc_project_stage(C.byref(st), n, m, nc, max(nc, 1), _p(A), _p(B), _p(Hv), _p(Q), _p(P), _p(R), _p(q), _p(r),
                                 C.c_double(c), _p(Cm) if nc else _dp(), _p(Dm) if nc else _dp(), _p(e) if nc else _dp(),
                                 _p(SmF), C.byref(out), C.byref(status))
    assert pp == p
    bufs["ct"] = float(bufs["ct"][0])
    return ProjectedStage(p=p, status=status.value, **bufs)


def compute_map(reduced, pr: ProjectedStage, SmNext, SvNext, sNext):
    n = pr.At.shape[0]
    p = pr.p
    SmNext = _f(SmNext)
    SvNext = np.ascontiguousarray(SvNext, dtype=np.float64)
    Km = np.zeros((p, n), order="F")
    Lv = np.zeros(p)
    Sm = np.zeros((n, n), order="F")
    Sv = np.zeros(n)
    s = C.c_double(0)
    cs = pr.c_struct()
    lib().orc_compute_map(int(reduced), n, p, C.byref(cs), _p(SmNext), _p(SvNext), C.c_double(sNext), _p(Km), _p(Lv), _p(Sm), _p(Sv),
                          C.byref(s))
    return Km, Lv, Sm, Sv, s.value


def flow_map_slq(reduced, pr: ProjectedStage, allSs):
    n = pr.At.shape[0]
    allSs = np.ascontiguousarray(allSs, dtype=np.float64)
    out = np.zeros_like(allSs)
    cs = pr.c_struct()
    lib().orc_flow_map_slq(int(reduced), n, pr.p, C.byref(cs), _p(allSs), _p(out))
    return out


# --------------------------------------------------------------------------------------------------------------------
# whole problems. Python-side layout: arrays indexed [node, row, col] (C-order over nodes, each block stored column-major,
# i.e. arr[k] is the transposed view of a Fortran block). To keep this simple every per-node matrix field is held as
# an ndarray of shape (nodes, cols, rows): arr[k].T is the matrix.  Helpers below convert from natural (nodes, rows, cols).
# --------------------------------------------------------------------------------------------------------------------
def to_colmajor_nodes(a):
    """(nodes, rows, cols) natural -> contiguous (nodes, cols, rows) buffer whose per-node block is column-major."""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


def from_colmajor_nodes(a):
    return np.swapaxes(a, -1, -2)


@dataclass
class Problem:
    """One LQ problem in natural numpy layout: A (nodes,n,n), B (nodes,n,m), Hv (nodes,n), Q (nodes,n,n), P (nodes,m,n),
    R (nodes,m,m), q (nodes,n), r (nodes,m), c (nodes,), C (nodes,ncmax,n), D (nodes,ncmax,m), e (nodes,ncmax), nc (nodes,) int32,
    Qf (n,n), qf (n,), cf float, x_nom (N+1,n), u_nom (N+1,m), time (N+1,). nodes = N for ILQR, N+1 for SLQ."""
    N: int
    A: np.ndarray
    B: np.ndarray
    Hv: np.ndarray
    Q: np.ndarray
    P: np.ndarray
    R: np.ndarray
    q: np.ndarray
    r: np.ndarray
    c: np.ndarray
    Qf: np.ndarray
    qf: np.ndarray
    cf: float
    C: np.ndarray | None = None
    D: np.ndarray | None = None
    e: np.ndarray | None = None
    nc: np.ndarray | None = None
    x_nom: np.ndarray | None = None
    u_nom: np.ndarray | None = None
    time: np.ndarray | None = None
    event: np.ndarray | None = None  # (nodes,) int: 1 marks a pre-event node. ILQR: jump data in the node's A, Hv, Q, q, c
    # SLQ: jump model data of the e-th event (node order): jA (E,n,n), jHv (E,n), jQ (E,n,n), jq (E,n), jc (E,)
    jA: np.ndarray | None = None
    jHv: np.ndarray | None = None
    jQ: np.ndarray | None = None
    jq: np.ndarray | None = None
    jc: np.ndarray | None = None

    @property
    def nx(self):
        return self.A.shape[1]

    @property
    def nu(self):
        return self.B.shape[2]

    @property
    def nc_max(self):
        return 0 if self.D is None else self.D.shape[1]

    def c_struct(self):
        keep = {}
        for name in ("A", "B", "Q", "P", "R", "C", "D"):
            v = getattr(self, name)
            keep[name] = to_colmajor_nodes(v) if v is not None else None
        for name in ("Hv", "q", "r", "c", "e", "qf", "x_nom", "u_nom", "time"):
            v = getattr(self, name)
            keep[name] = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
        keep["Qf"] = np.ascontiguousarray(np.asarray(self.Qf, dtype=np.float64).T)
        keep["cf"] = np.array([self.cf], dtype=np.float64)
        keep["nc"] = np.ascontiguousarray(self.nc, dtype=np.int32) if self.nc is not None else None
        ev = getattr(self, "event", None)
        keep["event"] = np.ascontiguousarray(ev, dtype=np.int32) if ev is not None else None
        if keep["time"] is None:
            keep["time"] = np.arange(self.N + 1, dtype=np.float64)
        for name in ("jA", "jQ"):
            v = getattr(self, name)
            keep[name] = to_colmajor_nodes(v) if v is not None else None
        for name in ("jHv", "jq", "jc"):
            v = getattr(self, name)
            keep[name] = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
        pb = _Problem(self.nx, self.nu, self.nc_max, self.N, _p(keep["A"]), _p(keep["B"]), _p(keep["Hv"]), _p(keep["Q"]), _p(keep["P"]),
                      _p(keep["R"]), _p(keep["q"]), _p(keep["r"]), _p(keep["c"]), _p(keep["C"]), _p(keep["D"]), _p(keep["e"]),
                      keep["nc"].ctypes.data_as(_ip) if keep["nc"] is not None else _ip(), _p(keep["Qf"]), _p(keep["qf"]), _p(keep["cf"]),
                      _p(keep["x_nom"]), _p(keep["u_nom"]), _p(keep["time"]),
                      keep["event"].ctypes.data_as(_ip) if keep["event"] is not None else _ip(), _p(keep["jA"]), _p(keep["jHv"]),
                      _p(keep["jQ"]), _p(keep["jq"]), _p(keep["jc"]))
        return pb, keep


@dataclass
class Solution:
    """K (N+1,m,n), dbias (N+1,m), bias (N+1,m), Sm (N+1,n,n), Sv (N+1,n), s (N+1,), status."""
    K: np.ndarray
    dbias: np.ndarray
    bias: np.ndarray
    Sm: np.ndarray
    Sv: np.ndarray
    s: np.ndarray
    status: int = 0


def backward(st: Settings, pb: Problem) -> Solution:
    n, m, N = pb.nx, pb.nu, pb.N
    Kc = np.zeros((N + 1, n, m))  # column-major blocks of m x n
    db = np.zeros((N + 1, m))
    bias = np.zeros((N + 1, m))
    Smc = np.zeros((N + 1, n, n))
    Sv = np.zeros((N + 1, n))
    s = np.zeros(N + 1)
    cp, keep = pb.c_struct()
    sol = _Solution(_p(Kc), _p(db), _p(bias), _p(Smc), _p(Sv), _p(s), 0)
    status = lib().orc_backward(C.byref(st), C.byref(cp), C.byref(sol))
    del keep
    return Solution(K=from_colmajor_nodes(Kc).copy(), dbias=db, bias=bias, Sm=from_colmajor_nodes(Smc).copy(), Sv=Sv, s=s, status=status)


def rollout(st: Settings, pb: Problem, sol: Solution, x0, alpha=1.0, max_out=None):
    n, m, N = pb.nx, pb.nu, pb.N
    if max_out is None:
        n_ev = 0 if pb.event is None else int(np.count_nonzero(pb.event))
        max_out = N + 1 if st.algorithm == ALG_ILQR else int(np.ceil((pb.time[-1] - pb.time[0]) / st.time_step)) + 4 + 3 * n_ev
    x = np.zeros((max_out, n))
    u = np.zeros((max_out, m))
    t = np.zeros(max_out)
    cp, keep = pb.c_struct()
    Kc = to_colmajor_nodes(sol.K)
    Smc = to_colmajor_nodes(sol.Sm)
    csol = _Solution(_p(Kc), _p(np.ascontiguousarray(sol.dbias)), _p(np.ascontiguousarray(sol.bias)), _p(Smc), _p(sol.Sv), _p(sol.s), 0)
    n_out = C.c_int(0)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    status = lib().orc_rollout(C.byref(st), C.byref(cp), C.byref(csol), _p(x0), C.c_double(alpha), _p(x), _p(u), _p(t), max_out,
                               C.byref(n_out))
    del keep
    assert status >= 0, "rollout output capacity too small"
    k = n_out.value
    return x[:k], u[:k], t[:k], status


def discrete_lq_cost(pb: Problem, x, u) -> float:
    cp, keep = pb.c_struct()
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    val = lib().orc_discrete_lq_cost(C.byref(cp), _p(x), _p(u))
    del keep
    return val


def line_search(st: Settings, pb: Problem, sol: Solution, x0, min_step=0.05, max_step=1.0, rate=0.5, armijo=1e-4, baseline=None):
    """LineSearchStrategy::run / lineSearchTask (ocs2_ddp/src/search_strategy/LineSearchStrategy.cpp:125-258) restated for the
    single-thread case ("equivalent to a single core line search") on the discrete LQ model: candidates max*rate^e >= min
    (numerics::almost_ge), merit = LQ-model cost of the rollout, Armijo condition against
    baseline - armijo * alpha * computeControllerUpdateIS (DDP_HelperFunctions.cpp:285-291, trapezoidal rule over the time stamps).
    Returns (step length, candidate index or -1, merits, baseline, IS, candidates)."""
    eps, tiny = np.finfo(np.float64).eps, np.finfo(np.float64).tiny
    cands, e = [], 0
    while True:
        a = max_step * rate ** e
        diff, mag = abs(a - min_step), min(abs(a), abs(min_step))
        if not (a > min_step or diff <= eps * mag or diff < tiny):
            break
        cands.append(a)
        e += 1
    t = np.asarray(pb.time, dtype=np.float64)
    sq = (np.asarray(sol.dbias) ** 2).sum(axis=1)
    IS = 0.0
    for k in range(1, len(t)):
        IS += (sq[k - 1] + sq[k]) * (0.5 * (t[k] - t[k - 1]))
    base = float(baseline) if baseline is not None else float(np.sum(pb.c[:pb.N]) + pb.cf)
    merits = []
    for a in cands:
        x, u, _, _ = rollout(st, pb, sol, x0, alpha=a)
        merits.append(discrete_lq_cost(pb, x, u))
    best, idx = 0.0, -1
    for e, a in enumerate(cands):
        if merits[e] < base - armijo * a * IS:
            best, idx = a, e
            break
    return best, idx, np.array(merits), base, IS, np.array(cands)


def flatten_controller(time, K, bias, dbias=None, alpha=0.0, query_times=None):
    """LinearController::flatten / flattenSingle (ocs2_core/src/control/LinearController.cpp:87-140): per query time the bias and gain
    are linearly interpolated over the time stamps (LinearInterpolation::timeSegment) and serialised as float32 rows
    [uff_i, K_i,:]. K (N+1,m,n), bias (N+1,m); with dbias the bias is first incremented by alpha*dbias (incrementController,
    DDP_HelperFunctions.cpp:296-304). Returns (len(query_times), m*(n+1)) float32; query_times defaults to the time stamps."""
    K, bias = np.asarray(K, dtype=np.float64), np.asarray(bias, dtype=np.float64)
    if dbias is not None:
        bias = bias + alpha * np.asarray(dbias, dtype=np.float64)
    time = np.asarray(time, dtype=np.float64)
    query = time if query_times is None else np.asarray(query_times, dtype=np.float64)
    m, n = K.shape[1], K.shape[2]
    out = np.zeros((len(query), m * (n + 1)), dtype=np.float32)
    for t_i, t in enumerate(query):
        idx, a = time_segment(float(t), time)
        if len(time) > 1 and idx + 1 < len(time):
            uff, k = a * bias[idx] + (1.0 - a) * bias[idx + 1], a * K[idx] + (1.0 - a) * K[idx + 1]
        else:
            uff, k = bias[idx], K[idx]
        for i in range(m):
            out[t_i, i * (n + 1)] = np.float32(uff[i])
            for j in range(n):
                out[t_i, i * (n + 1) + j + 1] = np.float32(k[i, j])
    return out


def unflatten_controller(flat, n, m):
    """LinearController::unFlatten (LinearController.cpp:145-171): float rows [uff_i, K_i,:] -> (bias (T,m), gain (T,m,n)) in double."""
    flat = np.asarray(flat, dtype=np.float32)
    if flat.shape[1] != m + m * n:
        raise RuntimeError("LinearController::unFlatten received array of wrong length.")
    rows = flat.reshape(flat.shape[0], m, n + 1).astype(np.float64)
    return rows[:, :, 0].copy(), rows[:, :, 1:].copy()


def rk4_sensitivity_discretization(dfdx, dfdu, dt):
    """rk4SensitivityDiscretization (ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:130-169) on the four stage linearisations
    k1..k4 (dfdx[s] (n,n), dfdu[s] (n,m)) of one RK4 step: the input sensitivity chain, the state sensitivity chain (one temporary per
    product), and the assembly dfdx = I + dt/6 k1 + dt/3 k2 + dt/3 k3 + dt/6 k4 (dfdu alike). Returns (A, B) of the discrete model."""
    A = [np.array(a, dtype=np.float64) for a in dfdx]
    B = [np.array(b, dtype=np.float64) for b in dfdu]
    h2, h6, h3 = dt / 2.0, dt / 6.0, dt / 3.0
    B[1] = B[1] + h2 * (A[1] @ B[0])
    B[2] = B[2] + h2 * (A[2] @ B[1])
    B[3] = B[3] + dt * (A[3] @ B[2])
    tmp = h2 * (A[1] @ A[0])
    A[1] = A[1] + tmp
    tmp = h2 * (A[2] @ A[1])
    A[2] = A[2] + tmp
    tmp = dt * (A[3] @ A[2])
    A[3] = A[3] + tmp
    Ad = h6 * A[0] + h3 * A[1] + h3 * A[2] + h6 * A[3]
    Ad[np.diag_indices_from(Ad)] += 1.0
    Bd = h6 * B[0] + h3 * B[1] + h3 * B[2] + h6 * B[3]
    return Ad, Bd


def generate_problem(seed, problem, algorithm, n, m, nc, N, dt):
    """One problem of the seeded synthetic family (bit-identical to the CUDA generator). Returns (Problem, x0)."""
    nodes = N if algorithm == ALG_ILQR else N + 1
    A = np.zeros((nodes, n, n))
    B = np.zeros((nodes, m, n))
    Hv = np.zeros((nodes, n))
    Q = np.zeros((nodes, n, n))
    P = np.zeros((nodes, n, m))
    R = np.zeros((nodes, m, m))
    q = np.zeros((nodes, n))
    r = np.zeros((nodes, m))
    c = np.zeros(nodes)
    ncm = max(nc, 1)
    Cm = np.zeros((nodes, n, ncm))
    Dm = np.zeros((nodes, m, ncm))
    e = np.zeros((nodes, ncm))
    Qf = np.zeros((n, n))
    qf = np.zeros(n)
    cf = np.zeros(1)
    x0 = np.zeros(n)
    lib().orc_generate_problem(C.c_uint64(seed), C.c_int64(problem), algorithm, n, m, nc, N, C.c_double(dt), _p(A), _p(B), _p(Hv), _p(Q),
                               _p(P), _p(R), _p(q), _p(r), _p(c), _p(Cm), _p(Dm), _p(e), _p(Qf), _p(qf), _p(cf), _p(x0))
    sw = from_colmajor_nodes
    pb = Problem(N=N, A=sw(A), B=sw(B), Hv=Hv, Q=sw(Q), P=sw(P), R=sw(R), q=q, r=r, c=c, Qf=Qf.T, qf=qf, cf=float(cf[0]),
                 C=sw(Cm) if nc else None, D=sw(Dm) if nc else None, e=e if nc else None, time=dt * np.arange(N + 1))
    return pb, x0


def baseline_run(st: Settings, seed, first, count, n, m, nc, N, dt, threads):
    """Timed CPU baseline on `count` generated problems (backward + rollout). Returns (seconds, checksum)."""
    chk = C.c_double(0)
    secs = lib().orc_baseline_run(C.byref(st), C.c_uint64(seed), C.c_int64(first), C.c_int64(count), n, m, nc, N, C.c_double(dt), threads,
                                  C.byref(chk))
    return secs, chk.value


def batch_solve(st: Settings, seed, first, count, n, m, nc, N, dt, alpha=1.0, threads=None, max_out=None):
    """backward + one rollout of `count` generated problems on `threads` CPU threads, results kept. Returns a dict of arrays in the
    natural layout of Solution: K (count,N+1,m,n), dbias, bias (count,N+1,m), Sm (count,N+1,n,n), Sv (count,N+1,n), s (count,N+1),
    x (count,out_nodes,n), u (count,out_nodes,m), status (count,)."""
    threads = threads or (os.cpu_count() or 1)
    if max_out is None:
        max_out = N + 1 if st.algorithm == ALG_ILQR else int(np.ceil(N * dt / st.time_step)) + 4
    Kc = np.zeros((count, N + 1, n, m))
    db, bias = np.zeros((count, N + 1, m)), np.zeros((count, N + 1, m))
    Smc, Sv, s = np.zeros((count, N + 1, n, n)), np.zeros((count, N + 1, n)), np.zeros((count, N + 1))
    x, u = np.zeros((count, max_out, n)), np.zeros((count, max_out, m))
    status = np.zeros(count, dtype=np.int32)
    fn = lib().orc_batch_solve
    fn.restype = C.c_int
    k = fn(C.byref(st), C.c_uint64(seed), C.c_int64(first), C.c_int64(count), n, m, nc, N, C.c_double(dt), C.c_double(alpha), threads, _p(Kc),
           _p(db), _p(bias), _p(Smc), _p(Sv), _p(s), _p(x), _p(u), max_out, status.ctypes.data_as(_ip))
    return dict(K=np.swapaxes(Kc, -1, -2), dbias=db, bias=bias, Sm=np.swapaxes(Smc, -1, -2), Sv=Sv, s=s, x=x[:, :k], u=u[:, :k], status=status)
