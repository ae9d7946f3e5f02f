"""Pins the CPU oracle against the reference's own golden vectors / known-answer tests / identities (CPU only).

Each test cites the reference test it transfers (paths relative to /root/reference). Eigen's Random() fixtures are not
reproducible without Eigen, so the *properties* and tolerances are transferred with numpy-seeded inputs.
"""
import os

import numpy as np
import pytest

from oracle import kkt_oracle
from oracle import oracle as orc

RNG = np.random.default_rng(0)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def spd(n, rng=RNG):
    """generateSPDmatrix (ocs2_core/include/ocs2_core/misc/randomMatrices.h): symmetric, diagonally dominant."""
    A = rng.uniform(-1, 1, (n, n))
    A = 0.5 * (A + A.T)
    return A + n * np.eye(n)


def random_cost(n, m, rng):
    """getRandomCost (ocs2_oc/test/include/ocs2_oc/test/testProblemsGeneration.h:45-58): joint PSD M^T M."""
    M = rng.uniform(-1, 1, (n + m, n + m))
    W = M.T @ M
    return W[:n, :n], W[n:, :n], W[n:, n:], rng.uniform(-1, 1, n), rng.uniform(-1, 1, m), float(rng.uniform(-1, 1))


# ---------------------------------------------------------------------------------------------------------------------
def test_flatten_golden_order():
    """ocs2_ddp/test/RiccatiTest.cpp:107-131 (testFlattenSMatrix): exact golden vector."""
    Sm = np.array([[1, 2, 4, 7], [2, 3, 5, 8], [4, 5, 6, 9], [7, 8, 9, 10]], dtype=float)
    allSs = orc.flatten(Sm, np.array([11.0, 12, 13, 14]), 15.0)
    assert np.array_equal(allSs, np.arange(1.0, 16.0))


def test_flatten_unflatten_roundtrip():
    """ocs2_ddp/test/RiccatiTest.cpp:133-157 (stateDim = 42)."""
    n = 42
    Sm = RNG.uniform(-1, 1, (n, n))
    Sm = Sm + Sm.T
    Sv = RNG.uniform(-1, 1, n)
    s = float(RNG.uniform(-1, 1))
    Sm2, Sv2, s2 = orc.unflatten(n, orc.flatten(Sm, Sv, s))
    assert s2 == s and np.array_equal(Sv2, Sv) and np.array_equal(Sm2, Sm)


def test_llt_of_inverse():
    """ocs2_core/test/misc/testLinearAlgebra.cpp:111-125 (LLTofInverse, n = 10, tol 1e-9)."""
    A = spd(10)
    Ui, rc = orc.inverse_uut(A)
    assert rc == 0
    assert np.allclose(np.tril(Ui, -1), 0.0)  # upper triangular
    assert np.abs(np.linalg.inv(A) - Ui @ Ui.T).max() < 1e-9


def test_llt_reports_indefinite():
    _, rc = orc.inverse_uut(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert rc == 1


def test_constraint_projection_against_full_computation():
    """ocs2_core/test/misc/testLinearAlgebra.cpp:127-165 (m = 4 constraints, n = 15 inputs, tol 1e-9)."""
    nc, m = 4, 15
    D = RNG.uniform(-1, 1, (nc, m))
    R = spd(m)
    Ui, _ = orc.inverse_uut(R)
    Rinv = Ui @ Ui.T
    Dd, RcInv, Pu = orc.constraint_projection(D, Ui)
    RmProjected = np.linalg.inv(D @ Rinv @ D.T)
    Dd_check = Rinv @ D.T @ RmProjected
    assert np.abs(Dd - Dd_check).max() < 1e-9
    assert np.abs(RcInv @ RcInv.T - Dd_check.T @ R @ Dd_check).max() < 1e-9
    nullProj = np.eye(m) - Dd_check @ D
    assert np.abs(Pu @ Pu.T - Rinv.T @ nullProj.T @ R @ nullProj @ Rinv).max() < 1e-9
    # invariants listed in SURVEY.md appendix A.3
    assert np.abs(D @ Dd - np.eye(nc)).max() < 1e-12
    assert np.abs(D @ Pu).max() < 1e-12
    assert np.abs(Pu.T @ R @ Pu - np.eye(m - nc)).max() < 1e-12


def test_make_psd_gershgorin():
    """ocs2_core/test/misc/testLinearAlgebra.cpp:167-189."""
    n = 10
    dd = spd(n)
    corr, rc = orc.shift_hessian(orc.HC_GERSHGORIN_MODIFICATION, dd, 1e-6)
    assert rc == 0 and np.allclose(dd, corr, rtol=1e-9)
    lam_min = np.linalg.eigvalsh(dd).min()
    nd = dd - (lam_min + 1e-2) * np.eye(n)
    corr, _ = orc.shift_hessian(orc.HC_GERSHGORIN_MODIFICATION, nd, 1e-3)
    assert np.linalg.eigvalsh(corr).min() >= 1e-3


def test_diagonal_shift_and_unsupported():
    M = spd(5)
    out, rc = orc.shift_hessian(orc.HC_DIAGONAL_SHIFT, M, 1e-3)
    assert rc == 0 and np.array_equal(out, M + 1e-3 * np.eye(5))
    _, rc = orc.shift_hessian(orc.HC_CHOLESKY_MODIFICATION, M, 1e-3)
    assert rc == 1


@pytest.mark.parametrize("n", [1, 3, 10, 24])
def test_make_psd_eigenvalue_against_numpy_eigh(n):
    """LinearAlgebra::makePsdEigenvalue (LinearAlgebra.cpp:52-72): eigenvalues below eps are raised to eps (V max(L, eps) V'); a matrix
    whose spectrum is already above eps is only symmetrised. The oracle's Jacobi solver is pinned against numpy's LAPACK eigh."""
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (n, n))
    M = 0.5 * (X + X.T)  # indefinite
    eps = 1e-3
    out, rc = orc.shift_hessian(orc.HC_EIGENVALUE_MODIFICATION, M, eps)
    w, V = np.linalg.eigh(M)
    want = (V * np.maximum(w, eps)) @ V.T if w.min() < eps else 0.5 * (M + M.T)
    assert rc == 0 and np.abs(out - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    assert np.linalg.eigvalsh(0.5 * (out + out.T)).min() >= eps * (1 - 1e-9)
    # benign case: untouched up to symmetrisation, even for a slightly asymmetric input
    P = spd(n) + 1e-13 * rng.uniform(-1, 1, (n, n))
    out, _ = orc.shift_hessian(orc.HC_EIGENVALUE_MODIFICATION, P, eps)
    assert np.array_equal(out, 0.5 * (P + P.T))


@pytest.mark.parametrize("nc", [0, 2])
def test_change_of_input_variables_equivalence(nc):
    """ocs2_oc/test/testChangeOfInputVariables.cpp: projected model evaluated at u~ == original at u = Pu u~ + Px x + u0."""
    rng = np.random.default_rng(5 + nc)
    n, m = 4, 5
    Q, P, R, q, r, c = random_cost(n, m, rng)
    R = R + np.eye(m)
    A, B, Hv = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, m)), rng.uniform(-1, 1, n)
    Cm = rng.uniform(-1, 1, (nc, n)) if nc else None
    Dm = rng.uniform(-1, 1, (nc, m)) if nc else None
    e = rng.uniform(-1, 1, nc) if nc else None
    st = orc.make_settings(hessian_multiple=0.0)
    S = spd(n, rng)
    pr = orc.project_stage(st, A, B, Hv, Q, P, R, q, r, c, Cm, Dm, e, Sm=S)
    assert pr.p == m - nc
    dx = rng.uniform(-1, 1, n)
    ut = rng.uniform(-1, 1, m - nc)
    u = pr.Pu @ ut - pr.Cmt @ dx - pr.Evt

    def cost(Q, P, R, q, r, c, x, u):
        return c + q @ x + r @ u + 0.5 * x @ Q @ x + u @ P @ x + 0.5 * u @ R @ u

    assert cost(pr.Qt, pr.Pt, pr.Rt, pr.qt, pr.rt, pr.ct, dx, ut) == pytest.approx(cost(Q, P, R, q, r, c, dx, u), rel=1e-12)
    assert np.allclose(pr.At @ dx + pr.Bt @ ut + pr.Hvt, A @ dx + B @ u + Hv, rtol=1e-12, atol=1e-13)
    if nc:
        assert np.allclose(Cm @ dx + Dm @ u + e, 0.0, atol=1e-12)  # constraint satisfied for every u~
    # Hm = R + B'SB; Pu' Hm Pu = I (GaussNewtonDDP.cpp:774-781)
    Hm = R + B.T @ S @ B
    assert np.allclose(pr.Pu.T @ Hm @ pr.Pu, np.eye(m - nc), atol=1e-10)


def test_flow_map_reduced_equals_full():
    """ocs2_ddp/test/RiccatiTest.cpp:87-105 (STATE_DIM 48, INPUT_DIM 10, projected data with R~ = I, tol 1e-9)."""
    rng = np.random.default_rng(3)
    n, m = 48, 10
    st = orc.make_settings(algorithm=orc.ALG_SLQ, hessian_multiple=0.0)
    pr = orc.project_stage(st, rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, m)), rng.uniform(-1, 1, n), spd(n, rng),
                           rng.uniform(-1, 1, (m, n)), np.eye(m), rng.uniform(-1, 1, n), rng.uniform(-1, 1, m), 0.3)
    assert np.allclose(pr.Rt, np.eye(m))
    pr.dQ = np.asfortranarray(0.1 * spd(n, rng))
    allSs = rng.uniform(-1, 1, n * (n + 1) // 2 + n + 1)
    d_red = orc.flow_map_slq(True, pr, allSs)
    d_full = orc.flow_map_slq(False, pr, allSs)
    assert np.abs(d_red - d_full).max() < 1e-9


def test_discrete_map_reduced_equals_full():
    """Same identity for DiscreteTimeRiccatiEquations (reduced form is exact when projected Hm = I)."""
    rng = np.random.default_rng(4)
    n, m = 12, 5
    st = orc.make_settings(hessian_multiple=1e-5)
    Q, P, R, q, r, c = random_cost(n, m, rng)
    R = R + np.eye(m)
    S = spd(n, rng)
    pr = orc.project_stage(st, np.eye(n) + 0.1 * rng.uniform(-1, 1, (n, n)), 0.1 * rng.uniform(-1, 1, (n, m)), 0.01 * rng.uniform(-1, 1, n),
                           Q, P, R, q, r, c, Sm=S)
    Sv = rng.uniform(-1, 1, n)
    a = orc.compute_map(True, pr, S, Sv, 0.7)
    b = orc.compute_map(False, pr, S, Sv, 0.7)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-9, atol=1e-9)


def test_time_segment_rules():
    """implementation/LinearInterpolation.h:69-107 and ocs2_core/test/misc/testInterpolation.cpp semantics."""
    t = np.array([0.0, 1.0, 2.0, 3.0])
    assert orc.time_segment(-1.0, t) == (0, 1.0)
    assert orc.time_segment(0.0, t) == (0, 1.0)  # lower_bound: t == t0 -> interval -1 -> clamp
    assert orc.time_segment(0.25, t) == (0, 0.75)
    assert orc.time_segment(1.0, t) == (0, 0.0)  # exact node j >= 1 -> (j-1, alpha = 0)
    assert orc.time_segment(2.5, t) == (2, 0.5)
    assert orc.time_segment(3.0, t) == (2, 0.0)
    assert orc.time_segment(9.0, t) == (2, 0.0)
    short = np.array([0.0, 1.0, 1.0 + 1e-9, 2.0])
    assert orc.time_segment(1.0 + 0.2e-9, short) == (1, 1.0)
    assert orc.time_segment(1.0 + 0.8e-9, short) == (1, 0.0)


def test_slq_rk4_converges_to_matlab_care():
    """ocs2_ddp/test/testContinuousTimeLqr.cpp:41-72: MATLAB golden K and S (tol 1e-9) for A=[1 2;3 4], B=[5;6],
    Q=[3 2;2 4], R=5, P=[0.1 0.2]. The SLQ Riccati flow map integrated backwards with RK4 (h = 1e-3, T = 20) must converge to it."""
    h, T = 1e-3, 20.0
    N = int(round(T / h))
    nodes = N + 1
    rep = lambda M: np.repeat(np.asarray(M, dtype=float)[None], nodes, 0)  # noqa: E731
    pb = orc.Problem(N=N, A=rep([[1.0, 2.0], [3.0, 4.0]]), B=rep([[5.0], [6.0]]), Hv=np.zeros((nodes, 2)), Q=rep([[3.0, 2.0], [2.0, 4.0]]),
                     P=rep([[0.1, 0.2]]), R=rep([[5.0]]), q=np.zeros((nodes, 2)), r=np.zeros((nodes, 1)), c=np.zeros(nodes),
                     Qf=np.zeros((2, 2)), qf=np.zeros(2), cf=0.0, time=h * np.arange(nodes))
    st = orc.make_settings(algorithm=orc.ALG_SLQ, reduced_form=True, hessian_multiple=0.0, time_step=h)
    sol = orc.backward(st, pb)
    assert sol.status == 0
    K_check = np.array([[-0.905054653909129, -1.802904101100247]])
    S_check = np.array([[1.109884545577592, -0.187358243057052], [-0.187358243057052, 1.625218620131083]])
    assert np.allclose(sol.K[0], K_check, rtol=1e-9, atol=0)
    assert np.allclose(sol.Sm[0], S_check, rtol=1e-9, atol=0)
    # full-form Riccati gives the same answer
    st_full = orc.make_settings(algorithm=orc.ALG_SLQ, reduced_form=False, hessian_multiple=0.0, time_step=h)
    sol_full = orc.backward(st_full, pb)
    assert np.allclose(sol_full.Sm[0], S_check, rtol=1e-9, atol=0)


def random_lq(rng, n, m, N, nc=0, dt=0.05):
    A = np.stack([np.eye(n) + dt * rng.uniform(-1, 1, (n, n)) for _ in range(N)])
    B = np.stack([dt * rng.uniform(-1, 1, (n, m)) * 3 for _ in range(N)])
    Q, P, R, q, r, c = (np.zeros((N, n, n)), np.zeros((N, m, n)), np.zeros((N, m, m)), np.zeros((N, n)), np.zeros((N, m)), np.zeros(N))
    for k in range(N):
        Qk, Pk, Rk, qk, rk, ck = random_cost(n, m, rng)
        Q[k], P[k], R[k], q[k], r[k], c[k] = dt * Qk, dt * Pk, dt * (Rk + 0.1 * np.eye(m)), dt * qk, dt * rk, dt * ck
    Qf, _, _, qf, _, cf = random_cost(n, m, rng)
    kw = {}
    if nc:
        D = rng.uniform(-1, 1, (N, nc, m))
        D[:, :, :nc] += 2 * np.eye(nc)
        kw = dict(C=rng.uniform(-1, 1, (N, nc, n)), D=D, e=0.1 * rng.uniform(-1, 1, (N, nc)))
    return orc.Problem(N=N, A=A, B=B, Hv=0.01 * rng.uniform(-1, 1, (N, n)), Q=Q, P=P, R=R, q=q, r=r, c=c, Qf=Qf, qf=qf, cf=cf,
                       time=dt * np.arange(N + 1), **kw)


@pytest.mark.parametrize("n,m,nc,N", [(3, 2, 0, 50), (3, 2, 2, 50), (9, 9, 3, 100), (4, 1, 0, 30)])
@pytest.mark.parametrize("reduced", [True, False])
def test_ilqr_sweep_matches_dense_kkt(n, m, nc, N, reduced):
    """ocs2_ddp/test/CorrectnessTest.cpp:51-306 (random LQ, +-state-input constraints, DDP vs dense KKT; the reference uses
    5e-3/1e-2 after several DDP iterations — a single exact LQ solve must agree to round-off)."""
    rng = np.random.default_rng(100 * n + 10 * m + nc)
    pb = random_lq(rng, n, m, N, nc)
    st = orc.make_settings(algorithm=orc.ALG_ILQR, reduced_form=reduced, hessian_multiple=0.0)
    sol = orc.backward(st, pb)
    assert sol.status == 0
    x0 = rng.uniform(-1, 1, n)
    x, u, _, status = orc.rollout(st, pb, sol, x0, alpha=1.0)
    assert status == 0
    xk, uk, cost = kkt_oracle.solve_discrete_lq(pb, x0)
    assert np.abs(x - xk).max() < 1e-9 * max(1.0, np.abs(xk).max())
    assert np.abs(u[:N] - uk).max() < 1e-9 * max(1.0, np.abs(uk).max())
    # V(x0) = 1/2 x0'S0x0 + Sv0'x0 + s0 equals the rolled-out cost
    V = 0.5 * x0 @ sol.Sm[0] @ x0 + sol.Sv[0] @ x0 + sol.s[0]
    assert V == pytest.approx(cost, rel=1e-9, abs=1e-10)
    assert orc.discrete_lq_cost(pb, x, u) == pytest.approx(cost, rel=1e-9, abs=1e-10)


def test_ilqr_lm_strategy_zero_multiple_equals_line_search():
    rng = np.random.default_rng(77)
    pb = random_lq(rng, 5, 3, 40)
    a = orc.backward(orc.make_settings(strategy=orc.STRATEGY_LINE_SEARCH, hessian_multiple=0.0, reduced_form=False), pb)
    b = orc.backward(orc.make_settings(strategy=orc.STRATEGY_LM, lm_riccati_multiple=0.0, reduced_form=False), pb)
    assert np.allclose(a.K, b.K, rtol=1e-10, atol=1e-12) and np.allclose(a.Sm, b.Sm, rtol=1e-10, atol=1e-12)


def test_slq_matches_ilqr_in_the_small_step_limit():
    """SLQ (continuous flow map, RK4) and ILQR (discrete sweep) solve the same problem as dt -> 0: consistency of the two paths."""
    rng = np.random.default_rng(9)
    n, m, N, dt = 3, 2, 400, 2.5e-3
    nodes = N + 1
    Ac = rng.uniform(-1, 1, (n, n))
    Bc = rng.uniform(-1, 1, (n, m))
    Qc, Pc, Rc, _, _, _ = random_cost(n, m, rng)
    Rc = Rc + np.eye(m)
    Qf = spd(n, rng) / n
    rep = lambda M, k: np.repeat(np.asarray(M, dtype=float)[None], k, 0)  # noqa: E731
    cont = orc.Problem(N=N, A=rep(Ac, nodes), B=rep(Bc, nodes), Hv=np.zeros((nodes, n)), Q=rep(Qc, nodes), P=rep(Pc, nodes), R=rep(Rc, nodes),
                       q=np.zeros((nodes, n)), r=np.zeros((nodes, m)), c=np.zeros(nodes), Qf=Qf, qf=np.zeros(n), cf=0.0,
                       time=dt * np.arange(nodes))
    import scipy.linalg

    Md = scipy.linalg.expm(np.block([[Ac, Bc], [np.zeros((m, n + m))]]) * dt)
    disc = orc.Problem(N=N, A=rep(Md[:n, :n], N), B=rep(Md[:n, n:], N), Hv=np.zeros((N, n)), Q=rep(Qc * dt, N), P=rep(Pc * dt, N),
                       R=rep(Rc * dt, N), q=np.zeros((N, n)), r=np.zeros((N, m)), c=np.zeros(N), Qf=Qf, qf=np.zeros(n), cf=0.0,
                       time=dt * np.arange(nodes))
    s_slq = orc.backward(orc.make_settings(algorithm=orc.ALG_SLQ, hessian_multiple=0.0, time_step=dt), cont)
    s_ilqr = orc.backward(orc.make_settings(algorithm=orc.ALG_ILQR, hessian_multiple=0.0), disc)
    assert np.allclose(s_slq.Sm[0], s_ilqr.Sm[0], rtol=2e-2)
    assert np.allclose(s_slq.K[0], s_ilqr.K[0], rtol=5e-2, atol=5e-2)
    # continuous rollout with node-aligned steps lands on N+1 outputs and ends at tf
    x0 = rng.uniform(-1, 1, n)
    st = orc.make_settings(algorithm=orc.ALG_SLQ, hessian_multiple=0.0, time_step=dt)
    x, u, t, status = orc.rollout(st, cont, s_slq, x0)
    assert status == 0 and len(t) == N + 1 and t[-1] == cont.time[-1] and t[0] == pytest.approx(1e-9)


def test_cartpole_fixture_regenerates_and_matches_kkt():
    """config[0] anchor: the committed fixture equals a fresh oracle run, and the oracle equals the dense KKT solution."""
    g = np.load(os.path.join(GOLDEN, "cartpole_ilqr.npz"))
    N = int(g["N"])
    pb = orc.Problem(N=N, A=g["A"], B=g["B"], Hv=np.zeros((N, 4)), Q=g["Q"], P=g["P"], R=g["R"], q=g["q"], r=g["r"], c=g["c"], Qf=g["Qf"],
                     qf=g["qf"], cf=float(g["cf"]), x_nom=g["x_nom"], u_nom=g["u_nom"], time=float(g["dt"]) * np.arange(N + 1))
    st = orc.make_settings(hessian_multiple=float(g["eps"]), time_step=float(g["dt"]))
    sol = orc.backward(st, pb)
    for name in ("K", "dbias", "bias", "Sm", "Sv", "s"):
        assert np.allclose(getattr(sol, name), g[name], rtol=1e-12, atol=1e-12), name
    x, u, _, _ = orc.rollout(st, pb, sol, g["x0"])
    assert np.allclose(x, g["x"], rtol=1e-12, atol=1e-12) and np.allclose(u, g["u"], rtol=1e-12, atol=1e-12)
    # KKT in deviation coordinates (nominal removed): dx_{k+1} = A dx + B du, cost around the nominal
    dev = orc.Problem(N=N, A=g["A"], B=g["B"], Hv=np.zeros((N, 4)), Q=g["Q"], P=g["P"], R=g["R"], q=g["q"], r=g["r"], c=g["c"], Qf=g["Qf"],
                      qf=g["qf"], cf=float(g["cf"]))
    st0 = orc.make_settings(hessian_multiple=0.0)
    sol0 = orc.backward(st0, dev)
    dx0 = g["x0"] - g["x_nom"][0]
    xk, uk, _ = kkt_oracle.solve_discrete_lq(dev, dx0)
    xd, ud, _, _ = orc.rollout(st0, dev, sol0, dx0)
    assert np.abs(xd - xk).max() < 1e-9 and np.abs(ud[:N] - uk).max() < 1e-9
    # eps = 1e-6 shift perturbs the solution only slightly
    assert np.abs((x - g["x_nom"]) - xk).max() < 1e-3


def test_generator_is_deterministic_and_well_posed():
    pb, x0 = orc.generate_problem(0, 5, orc.ALG_ILQR, 9, 9, 3, 20, 0.01)
    pb2, x02 = orc.generate_problem(0, 5, orc.ALG_ILQR, 9, 9, 3, 20, 0.01)
    assert np.array_equal(pb.A, pb2.A) and np.array_equal(x0, x02)
    pb3, _ = orc.generate_problem(0, 6, orc.ALG_ILQR, 9, 9, 3, 20, 0.01)
    assert not np.array_equal(pb.A, pb3.A)
    assert np.allclose(pb.Q, np.swapaxes(pb.Q, 1, 2)) and np.allclose(pb.R, np.swapaxes(pb.R, 1, 2))
    W = np.block([[pb.Q[3], pb.P[3].T], [pb.P[3], pb.R[3]]])
    assert np.linalg.eigvalsh(W).min() > 0
    assert np.allclose(pb.D[0][:, :3], np.eye(3))
    st = orc.make_settings(hessian_multiple=1e-3)
    sol = orc.backward(st, pb)
    assert sol.status == 0
    x, u, _, status = orc.rollout(st, pb, sol, x0)
    assert status == 0 and np.isfinite(x).all()
    # constraints hold along the rollout: C x + D u + e = 0
    for k in range(pb.N):
        assert np.abs(pb.C[k] @ x[k] + pb.D[k] @ u[k] + pb.e[k]).max() < 1e-10


@pytest.mark.parametrize("n,m,nc", [(4, 2, 0), (6, 3, 2)])
def test_ilqr_events_match_dense_kkt(n, m, nc):
    """Event (pre-jump) nodes, ILQR.cpp:263-295 + riccatiTransversalityConditions (RiccatiTransversalityConditions.h:40-56): the value
    function passes through the jump map x+ = A_e x + Hv_e with the pre-jump cost (Q_e, q_e, c_e), no input acts. Equivalent dense KKT
    problem: the same stages with B = 0, P = 0, r = 0 (and no constraint) at the event nodes."""
    rng = np.random.default_rng(9 + n)
    N = 14
    pb = random_lq(rng, n, m, N, nc)
    event = np.zeros(N, dtype=np.int32)
    event[[3, 9, N - 1]] = 1  # includes an event at the last stage
    for k in np.nonzero(event)[0]:
        pb.A[k] = np.eye(n) + 0.3 * rng.uniform(-1, 1, (n, n))  # jump map
        pb.Hv[k] = 0.1 * rng.uniform(-1, 1, n)
    pb.event = event
    st = orc.make_settings(algorithm=orc.ALG_ILQR, reduced_form=True, hessian_multiple=0.0)
    sol = orc.backward(st, pb)
    assert sol.status == 0
    x0 = rng.uniform(-1, 1, n)
    x, u, _, status = orc.rollout(st, pb, sol, x0, alpha=1.0)
    assert status == 0
    import copy
    kk = copy.deepcopy(pb)
    kk.event = None
    for k in np.nonzero(event)[0]:
        kk.B[k] = 0.0
        kk.P[k] = 0.0
        kk.r[k] = 0.0
        kk.R[k] = np.eye(m)
        if nc:
            kk.nc = (np.full(N, nc, np.int32) if kk.nc is None else kk.nc.copy())
            kk.nc[k] = 0
    xk, uk, cost = kkt_oracle.solve_discrete_lq(kk, x0)
    assert np.abs(x - xk).max() < 1e-9 * max(1.0, np.abs(xk).max())
    reg = event == 0
    assert np.abs(u[:N][reg] - uk[reg]).max() < 1e-9 * max(1.0, np.abs(uk).max())
    V = 0.5 * x0 @ sol.Sm[0] @ x0 + sol.Sv[0] @ x0 + sol.s[0]
    assert V == pytest.approx(cost, rel=1e-9, abs=1e-10)
    assert orc.discrete_lq_cost(pb, x, u) == pytest.approx(cost, rel=1e-9, abs=1e-10)


def test_flatten_controller_layout_and_round_trip():
    """ocs2_core/test/control/testLinearController.cpp:7-26: flatten at the controller's time stamps, unFlatten, compare to 1e-6;
    plus the serialisation order of LinearController.cpp:107-140 (row i = [uff_i, K_i,:], float32) and an interpolated query."""
    rng = np.random.default_rng(5)
    time = np.array([0.0, 1.0])
    bias = rng.uniform(-1, 1, (2, 2))
    gain = rng.uniform(-1, 1, (2, 2, 3))
    flat = orc.flatten_controller(time, gain, bias)
    assert flat.dtype == np.float32 and flat.shape == (2, 2 + 2 * 3)
    for k in range(2):
        for i in range(2):
            assert flat[k, i * 4] == np.float32(bias[k, i])
            assert (flat[k, i * 4 + 1:i * 4 + 4] == gain[k, i].astype(np.float32)).all()
    bias_out, gain_out = orc.unflatten_controller(flat, 3, 2)
    assert np.allclose(bias_out, bias, rtol=1e-6, atol=1e-6) and np.allclose(gain_out, gain, rtol=1e-6, atol=1e-6)
    # a query between the stamps interpolates bias and gain linearly (LinearInterpolation::interpolate)
    mid = orc.flatten_controller(time, gain, bias, query_times=[0.25])
    b_mid, k_mid = orc.unflatten_controller(mid, 3, 2)
    assert np.allclose(b_mid[0], 0.75 * bias[0] + 0.25 * bias[1], atol=1e-6) and np.allclose(k_mid[0], 0.75 * gain[0] + 0.25 * gain[1], atol=1e-6)
    # incrementController: uff = bias + alpha * deltaBias
    dbias = rng.uniform(-1, 1, (2, 2))
    inc = orc.flatten_controller(time, gain, bias, dbias=dbias, alpha=0.5)
    assert np.allclose(orc.unflatten_controller(inc, 3, 2)[0], bias + 0.5 * dbias, atol=1e-6)
    with pytest.raises(RuntimeError):
        orc.unflatten_controller(flat[:, :-1], 3, 2)


def _slq_random_problem(rng, n, m, N, dt, event_nodes=()):
    """Random continuous-time LQ data on N+1 nodes; event_nodes are PRE-event nodes k: node k+1 is stamped weakEpsilon later
    (RolloutBase.cpp:62-64) and every event carries its own jump model data."""
    sym = lambda M: 0.5 * (M + M.T)
    time = np.zeros(N + 1)
    for k in range(1, N + 1):
        time[k] = time[k - 1] + (1e-9 if (k - 1) in event_nodes else dt)
    A = 0.5 * rng.uniform(-1, 1, (N + 1, n, n))
    B = rng.uniform(-1, 1, (N + 1, n, m))
    Q = np.stack([sym(rng.uniform(-1, 1, (n, n))) + n * np.eye(n) for _ in range(N + 1)])
    R = np.stack([sym(rng.uniform(-1, 1, (m, m))) + m * np.eye(m) for _ in range(N + 1)])
    pb = orc.Problem(N=N, A=A, B=B, Hv=0.1 * rng.uniform(-1, 1, (N + 1, n)), Q=Q, P=0.1 * rng.uniform(-1, 1, (N + 1, m, n)), R=R,
                     q=rng.uniform(-1, 1, (N + 1, n)), r=rng.uniform(-1, 1, (N + 1, m)), c=rng.uniform(-1, 1, N + 1),
                     Qf=sym(rng.uniform(-1, 1, (n, n))) + n * np.eye(n), qf=rng.uniform(-1, 1, n), cf=0.3, time=time,
                     x_nom=0.2 * rng.uniform(-1, 1, (N + 1, n)), u_nom=0.2 * rng.uniform(-1, 1, (N + 1, m)))
    E = len(event_nodes)
    if E:
        ev = np.zeros(N + 1, dtype=np.int32)
        ev[list(event_nodes)] = 1
        pb.event = ev
        pb.jA = np.stack([np.eye(n) + 0.3 * rng.uniform(-1, 1, (n, n)) for _ in range(E)])
        pb.jHv = 0.1 * rng.uniform(-1, 1, (E, n))
        pb.jQ = np.stack([sym(rng.uniform(-1, 1, (n, n))) + np.eye(n) for _ in range(E)])
        pb.jq = 0.2 * rng.uniform(-1, 1, (E, n))
        pb.jc = rng.uniform(-1, 1, E)
    return pb


def _slice_problem(pb, lo, hi, Qf, qf, cf):
    """nodes lo..hi of an SLQ problem as a problem of its own with the given terminal value function"""
    s = slice(lo, hi + 1)
    return orc.Problem(N=hi - lo, A=pb.A[s], B=pb.B[s], Hv=pb.Hv[s], Q=pb.Q[s], P=pb.P[s], R=pb.R[s], q=pb.q[s], r=pb.r[s], c=pb.c[s],
                       Qf=Qf, qf=qf, cf=cf, time=pb.time[s], x_nom=pb.x_nom[s], u_nom=pb.u_nom[s])


@pytest.mark.parametrize("reduced", [True, False])
def test_slq_events_compose_from_event_free_segments(reduced):
    """SLQ with events (SLQ.cpp:256-302): the backward pass integrates the inter-event segments separately and joins them with
    computeJumpMap = riccatiTransversalityConditions (ContinuousTimeRiccatiEquations.cpp:135-147, RiccatiTransversalityConditions.h:
    40-56). Pinned structurally: the solution with events equals the already pinned event-free pass on each segment, chained through
    the transversality conditions evaluated here in numpy."""
    rng = np.random.default_rng(77)
    n, m, N, dt = 5, 2, 14, 0.02
    events = (4, 9)
    pb = _slq_random_problem(rng, n, m, N, dt, events)
    st = orc.make_settings(algorithm=orc.ALG_SLQ, reduced_form=reduced, hessian_multiple=1e-6, time_step=0.007)
    full = orc.backward(st, pb)
    assert full.status == 0
    Qf, qf, cf = pb.Qf, pb.qf, pb.cf
    hi = N
    for ord_, k in reversed(list(enumerate(events))):
        seg = orc.backward(st, _slice_problem(pb, k + 1, hi, Qf, qf, cf))
        for name in ("Sm", "Sv", "s"):
            assert np.allclose(getattr(full, name)[k + 1:hi + 1], getattr(seg, name), rtol=1e-12, atol=1e-12), (name, k)
        assert np.allclose(full.K[k + 1:hi], seg.K[:hi - k - 1], rtol=1e-12, atol=1e-12)
        Sm, Sv, s = seg.Sm[0], seg.Sv[0], seg.s[0]
        Ae, Hve = pb.jA[ord_], pb.jHv[ord_]
        SmHv = Sm @ Hve
        Qf = pb.jQ[ord_] + (Sm.T @ Ae).T @ Ae
        qf = pb.jq[ord_] + Ae.T @ (Sv + SmHv)
        cf = float(s + pb.jc[ord_] + Hve @ (Sv + 0.5 * SmHv))
        Qf = np.triu(Qf) + np.triu(Qf, 1).T  # convert2Vector keeps the upper triangle
        hi = k
    seg = orc.backward(st, _slice_problem(pb, 0, hi, Qf, qf, cf))
    for name in ("Sm", "Sv", "s"):
        assert np.allclose(getattr(full, name)[:hi + 1], getattr(seg, name), rtol=1e-12, atol=1e-12), name
    assert np.allclose(full.K[:hi], seg.K[:hi], rtol=1e-12, atol=1e-12)

    # rollout: the segments chained through the jump map of the LQ model x+ = x_nom(post) + A_e (x - x_nom(pre)) + Hv_e
    x0 = rng.uniform(-1, 1, n)
    x, u, t, status = orc.rollout(st, pb, full, x0, alpha=0.7)
    assert status == 0
    pre = [int(np.argmin(np.abs(t - pb.time[k]))) for k in events]
    for ord_, (k, i) in enumerate(zip(events, pre)):
        assert t[i] == pb.time[k] and t[i + 1] == pb.time[k] + 1e-9  # the segment ends on the event, the next starts weakEpsilon later
        want = pb.x_nom[k + 1] + pb.jA[ord_] @ (x[i] - pb.x_nom[k]) + pb.jHv[ord_]
        assert np.allclose(x[i + 1], want, rtol=1e-13, atol=1e-13)
    # first segment: identical to the event-free rollout of nodes 0..k0 (same start nudge, same steps)
    k0 = events[0]
    segp = _slice_problem(pb, 0, k0, Qf, qf, cf)
    segsol = orc.Solution(K=full.K[:k0 + 1], dbias=full.dbias[:k0 + 1], bias=full.bias[:k0 + 1], Sm=full.Sm[:k0 + 1], Sv=full.Sv[:k0 + 1],
                          s=full.s[:k0 + 1])
    xs, us, ts, _ = orc.rollout(st, segp, segsol, x0, alpha=0.7)
    assert np.array_equal(ts, t[:pre[0] + 1]) and np.allclose(xs, x[:pre[0] + 1], rtol=1e-13, atol=1e-13)
    assert np.allclose(us, u[:pre[0] + 1], rtol=1e-12, atol=1e-12)


def test_rk4_sensitivity_discretization_reproduces_the_cartpole_fixture():
    """ILQR::discreteLQWorker / rk4SensitivityDiscretization (ILQR.cpp:137-157, SensitivityIntegratorImpl.cpp:130-169): the oracle's
    restatement on the four stage linearisations of the analytic cartpole model reproduces the discrete A, B stored in the committed
    fixture (tests/golden/cartpole_ilqr.npz, generated by make_cartpole_fixture.py)."""
    sympy = pytest.importorskip("sympy")  # the analytic model of the generator
    assert sympy is not None
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_cartpole_fixture", os.path.join(os.path.dirname(__file__), "golden", "make_cartpole_fixture.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "cartpole_ilqr.npz"))
    dt = float(fx["dt"])
    for k in (0, 17, 63, 99):
        x, u = fx["x_nom"][k], float(fx["u_nom"][k, 0])
        f1, A1, B1 = gen.lin(x, u)
        f2, A2, B2 = gen.lin(x + dt / 2 * f1, u)
        f3, A3, B3 = gen.lin(x + dt / 2 * f2, u)
        f4, A4, B4 = gen.lin(x + dt * f3, u)
        Ad, Bd = orc.rk4_sensitivity_discretization([A1, A2, A3, A4], [B1, B2, B3, B4], dt)
        assert np.allclose(Ad, fx["A"][k], rtol=1e-13, atol=1e-13) and np.allclose(Bd, fx["B"][k], rtol=1e-13, atol=1e-13)
    # a model that is constant over the step: the chain collapses to the 4th-order Taylor polynomial of the matrix exponential
    rng = np.random.default_rng(2)
    A, B, h = rng.uniform(-1, 1, (3, 3)), rng.uniform(-1, 1, (3, 2)), 0.05
    Ad, Bd = orc.rk4_sensitivity_discretization([A] * 4, [B] * 4, h)
    hA = h * A
    taylor = np.eye(3) + hA + hA @ hA / 2 + hA @ hA @ hA / 6 + hA @ hA @ hA @ hA / 24
    assert np.allclose(Ad, taylor, rtol=1e-13, atol=1e-14)
    assert np.allclose(Bd, h * (np.eye(3) + hA / 2 + hA @ hA / 6 + hA @ hA @ hA / 24) @ B, rtol=1e-13, atol=1e-14)


def test_batch_solve_is_the_single_problem_path_run_in_threads():
    """orc_batch_solve (the all-problems checker of the GPU parity tests) returns bit for bit what backward + rollout return per problem."""
    for alg, (n, m, nc) in ((orc.ALG_ILQR, (24, 24, 0)), (orc.ALG_ILQR, (9, 9, 3)), (orc.ALG_SLQ, (12, 4, 0))):
        N, dt, seed, first, count = 20, 0.01, 5, 1771, 9
        st = orc.make_settings(algorithm=alg, hessian_multiple=1e-4, time_step=dt)
        got = orc.batch_solve(st, seed, first, count, n, m, nc, N, dt, alpha=0.7, threads=4)
        assert (got["status"] == 0).all()
        for i in (0, 4, count - 1):
            pb, x0 = orc.generate_problem(seed, first + i, alg, n, m, nc, N, dt)
            ref = orc.backward(st, pb)
            x, u, _, _ = orc.rollout(st, pb, ref, x0, alpha=0.7)
            for name in ("K", "dbias", "bias", "Sm", "Sv", "s"):
                assert np.array_equal(got[name][i], getattr(ref, name)), name
            assert got["x"].shape[1] == len(x)
            assert np.array_equal(got["x"][i], x) and np.array_equal(got["u"][i], u)
