"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerance: the north star asks for <= 1e-9 relative error on K, k (deltaBias), S, Sv, s and the rollout trajectories; the
tests measure the TRUE relative error max|gpu - oracle| / max|oracle| per field and problem (no floor of 1 on the denominator: the
legged costs are scaled by dt = 0.01, so most fields are far below 1 in magnitude) and require it <= REL_TOL. A field whose
reference is identically zero (bias in deviation coordinates) must be reproduced exactly.
"""
import os

import numpy as np
import pytest

import ocs2_b200 as o2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
REL_TOL = 1e-9
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not (np.isfinite(a).all() and np.isfinite(b).all()):
        return 0.0 if np.array_equal(np.isnan(a), np.isnan(b)) else np.inf
    if a.size == 0:
        return 0.0
    scale = float(np.abs(b).max())
    return float(np.abs(a - b).max() / scale) if scale > 0.0 else (0.0 if not np.abs(a).max() else np.inf)


def orc_settings(st: o2.Settings):
    return orc.make_settings(algorithm=st.algorithm, reduced_form=st.reduced_form, strategy=st.strategy,
                             hessian_correction=st.hessianCorrectionStrategy, hessian_multiple=st.hessianCorrectionMultiple,
                             lm_riccati_multiple=st.riccatiMultiple, time_step=st.timeStep)


def check_against_oracle(st, pb, x0, sol, i, alphas=(1.0,), what=""):
    """sol: ocs2_b200.Solution of a batch; i: index inside it; pb/x0: the oracle problem."""
    ost = orc_settings(st)
    ref = orc.backward(ost, pb)
    assert int(sol.status[i]) & 1 == ref.status & 1, f"{what}: CHOL_NOT_PD flag differs"
    for name, got, want in (("K", sol.K[i], ref.K), ("dbias", sol.dbias[i], ref.dbias), ("bias", sol.bias[i], ref.bias),
                            ("Sm", sol.Sm[i], ref.Sm), ("Sv", sol.Sv[i], ref.Sv), ("s", sol.s[i], ref.s)):
        err = rel_err(got, want)
        assert err <= REL_TOL, f"{what} problem {i} field {name}: rel err {err:.3e}"
    if sol.x is not None:
        for ia, alpha in enumerate(alphas):
            x, u, t, _ = orc.rollout(ost, pb, ref, x0, alpha=alpha)
            assert sol.x.shape[2] == len(x), f"{what}: rollout node count {sol.x.shape[2]} vs oracle {len(x)}"
            ex, eu = rel_err(sol.x[ia, i], x), rel_err(sol.u[ia, i], u)
            assert ex <= REL_TOL and eu <= REL_TOL, f"{what} problem {i} alpha {alpha}: rollout rel err x {ex:.3e} u {eu:.3e}"
            assert rel_err(sol.t, t) <= 1e-15
    return ref


SHAPES = {
    "cartpole": (4, 1, 0),
    "ballbot": (10, 3, 0),
    "quadrotor": (12, 4, 0),
    "manipulator": (9, 9, 3),
    "legged": (24, 24, 0),
    "legged_c6": (24, 24, 6),      # one constraint tile of the legged DMMA kernel
    "legged_c14": (24, 24, 14),    # two tiles (a trot: 2 stance feet x 3 + 2 swing feet x (3 + 1))
    "test32c": (3, 2, 2),
}


@pytest.mark.parametrize("shape", ["cartpole", "ballbot", "manipulator", "legged", "legged_c6", "legged_c14", "test32c"])
@pytest.mark.parametrize("variant", ["ls_reduced_diag", "ls_full_diag", "ls_reduced_gershgorin", "lm_full"])
def test_ilqr_generated_batch_matches_oracle(shape, variant):
    n, m, nc = SHAPES[shape]
    N, batch, dt, seed = 100, 24, 0.01, 0
    st = o2.Settings(algorithm=o2.ALG_ILQR, hessianCorrectionMultiple=1e-3 if not shape.startswith("legged") else 1e-5, timeStep=dt)
    if variant == "ls_full_diag":
        st.preComputeRiccatiTerms = False
    elif variant == "ls_reduced_gershgorin":
        st.hessianCorrectionStrategy = o2.HC_GERSHGORIN_MODIFICATION
    elif variant == "lm_full":
        st.strategy = o2.STRATEGY_LEVENBERG_MARQUARDT
        st.riccatiMultiple = 0.37
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, max_alphas=2) as solver:
        solver.generate_synthetic(seed, first_problem_index=1000, dt=dt)
        # LINE_SEARCH + DIAGONAL_SHIFT in either Riccati form runs the shape's specialised kernel (the forms are the same map, the oracle
        # below is evaluated in the form asked for); every other variant goes through the generic kernel
        # (the legged shape: its DMMA kernel also carries the Gershgorin and Levenberg-Marquardt modifications as template variants)
        fast = (variant in ("ls_reduced_diag", "ls_full_diag") and shape != "test32c") or shape == "legged"
        assert ("generic" in solver.kernel_variant) != fast
        solver.solveSequentialRiccatiEquations()
        alphas = (1.0, 0.35)
        solver.rolloutTrajectory(alphas)
        sol = solver.download()
        assert (sol.status == 0).all()
        for i in (0, 7, batch - 1):
            pb, x0 = orc.generate_problem(seed, 1000 + i, orc.ALG_ILQR, n, m, nc, N, dt)
            check_against_oracle(st, pb, x0, sol, i, alphas, what=f"{shape}/{variant}")


@pytest.mark.parametrize("shape,substeps", [("quadrotor", 1.0), ("test32c", 2.5), ("cartpole", 1.0), ("legged", 1.0), ("legged", 2.5)])
@pytest.mark.parametrize("variant", ["ls_reduced_diag", "ls_full_gershgorin", "lm_full"])
def test_slq_rk4_generated_batch_matches_oracle(shape, substeps, variant):
    n, m, nc = SHAPES[shape]
    N, batch, dt, seed = 100, 12, 0.01, 3
    st = o2.Settings(algorithm=o2.ALG_SLQ, hessianCorrectionMultiple=1e-3, timeStep=dt / substeps)
    if variant == "ls_full_gershgorin":
        st.preComputeRiccatiTerms = False
        st.hessianCorrectionStrategy = o2.HC_GERSHGORIN_MODIFICATION
    elif variant == "lm_full":
        st.strategy = o2.STRATEGY_LEVENBERG_MARQUARDT
        st.riccatiMultiple = 0.2
    if shape == "quadrotor" and variant == "ls_reduced_diag":  # the full form of the same settings takes the same kernel
        st_full = o2.Settings(algorithm=o2.ALG_SLQ, hessianCorrectionMultiple=1e-3, timeStep=dt / substeps, preComputeRiccatiTerms=False)
        with o2.BatchedLqSolver(st_full, n, m, N, batch, nc_max=nc) as solver:
            assert solver.kernel_variant == "slq_rpl_kernel"
            solver.generate_synthetic(seed, first_problem_index=50, dt=dt)
            solver.solve(alpha=1.0)
            sol = solver.download()
            pb, x0 = orc.generate_problem(seed, 50, orc.ALG_SLQ, n, m, nc, N, dt)
            check_against_oracle(st_full, pb, x0, sol, 0, what="slq quadrotor full form")
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        if shape == "legged":  # projection / RK4 flow map / controller on the FP64 tensor pipe for the line-search configuration
            assert solver.kernel_variant == ("slq_wpp_kernel" if variant == "ls_reduced_diag" else "slq_generic_kernel")
        solver.generate_synthetic(seed, first_problem_index=50, dt=dt)
        solver.solve(alpha=1.0)
        sol = solver.download()
        assert (sol.status == 0).all()
        for i in (0, batch - 1):
            pb, x0 = orc.generate_problem(seed, 50 + i, orc.ALG_SLQ, n, m, nc, N, dt)
            check_against_oracle(st, pb, x0, sol, i, what=f"slq {shape}/{variant}")


def test_generator_is_bit_identical_to_oracle():
    """The device generator must reproduce the oracle generator bit for bit (so sampled full-size parity is meaningful):
    checked through a problem whose solution is exactly linear in the data: K, S of a 1-stage problem, and via x0."""
    n, m, nc, N, dt = 9, 9, 3, 4, 0.01
    st = o2.Settings(hessianCorrectionMultiple=0.0)
    with o2.BatchedLqSolver(st, n, m, N, 3, nc_max=nc) as solver:
        solver.generate_synthetic(11, first_problem_index=5, dt=dt)
        solver.solve()
        sol = solver.download()
        for i in range(3):
            pb, x0 = orc.generate_problem(11, 5 + i, orc.ALG_ILQR, n, m, nc, N, dt)
            assert np.array_equal(sol.x[0, i, 0], x0)          # x0 passes through the rollout untouched
            assert np.array_equal(sol.Sm[i, N], pb.Qf)         # terminal value function is a straight copy
            assert np.array_equal(sol.Sv[i, N], pb.qf)


def test_cartpole_anchor_fixture_batch1():
    """config[0]: cartpole ILQR (nx=4, nu=1, N=100) single problem with non-zero nominal trajectories, uploaded from host."""
    g = np.load(os.path.join(GOLDEN, "cartpole_ilqr.npz"))
    N = int(g["N"])
    st = o2.Settings(hessianCorrectionMultiple=float(g["eps"]), timeStep=float(g["dt"]))
    lq = o2.LqBatch(A=g["A"][None], B=g["B"][None], Q=g["Q"][None], R=g["R"][None], Qf=g["Qf"][None], P=g["P"][None], q=g["q"][None],
                    r=g["r"][None], c=g["c"][None], qf=g["qf"][None], cf=np.array([float(g["cf"])]), x_nom=g["x_nom"][None],
                    u_nom=g["u_nom"][None], x0=g["x0"][None])
    with o2.BatchedLqSolver(st, 4, 1, N, 1, has_nominal=True) as solver:
        solver.upload(lq)
        solver.solve()
        sol = solver.download()
    assert sol.status[0] == 0
    for name in ("K", "dbias", "bias", "Sm", "Sv", "s"):
        assert rel_err(getattr(sol, name)[0], g[name]) <= REL_TOL, name
    assert rel_err(sol.x[0, 0], g["x"]) <= REL_TOL and rel_err(sol.u[0, 0], g["u"]) <= REL_TOL


def _random_batch(rng, batch, n, m, N, ncmax, algorithm, ragged_nc=True, dt=0.02):
    nodes = N + 1 if algorithm == o2.ALG_SLQ else N
    disc = algorithm == o2.ALG_ILQR
    A = rng.uniform(-1, 1, (batch, nodes, n, n)) * (dt if disc else 1.0) + (np.eye(n) if disc else 0.0)
    Bm = rng.uniform(-1, 1, (batch, nodes, n, m)) * (3 * dt if disc else 1.0)
    M = rng.uniform(-1, 1, (batch, nodes, n + m, n + m))
    W = np.einsum("bkij,bkil->bkjl", M, M) / (n + m) + 0.1 * np.eye(n + m)
    sc = dt if disc else 1.0
    Mf = rng.uniform(-1, 1, (batch, n, n))
    kw = {}
    if ncmax:
        D = rng.uniform(-1, 1, (batch, nodes, ncmax, m))
        D[..., :ncmax] += 2 * np.eye(ncmax)
        nc = rng.integers(0, ncmax + 1, (batch, nodes)).astype(np.int32) if ragged_nc else np.full((batch, nodes), ncmax, np.int32)
        kw = dict(C=rng.uniform(-1, 1, (batch, nodes, ncmax, n)), D=D, e=0.1 * rng.uniform(-1, 1, (batch, nodes, ncmax)), nc=nc)
    return o2.LqBatch(A=A, B=Bm, Q=sc * W[..., :n, :n], R=sc * W[..., n:, n:], P=sc * W[..., n:, :n], Hv=0.01 * rng.uniform(-1, 1, (batch, nodes, n)),
                      q=sc * rng.uniform(-1, 1, (batch, nodes, n)), r=sc * rng.uniform(-1, 1, (batch, nodes, m)),
                      c=sc * rng.uniform(0, 1, (batch, nodes)), Qf=np.einsum("bij,bil->bjl", Mf, Mf) / n + 0.1 * np.eye(n),
                      qf=rng.uniform(-1, 1, (batch, n)), cf=rng.uniform(0, 1, batch), x_nom=0.1 * rng.uniform(-1, 1, (batch, N + 1, n)),
                      u_nom=0.1 * rng.uniform(-1, 1, (batch, N + 1, m)), x0=rng.uniform(-1, 1, (batch, n)), time=dt * np.arange(N + 1), **kw)


def _oracle_problem(lq, i, N):
    g = lambda a: None if a is None else a[i]  # noqa: E731
    return orc.Problem(N=N, A=lq.A[i], B=lq.B[i], Hv=lq.Hv[i], Q=lq.Q[i], P=lq.P[i], R=lq.R[i], q=lq.q[i], r=lq.r[i], c=lq.c[i], Qf=lq.Qf[i],
                       qf=lq.qf[i], cf=float(lq.cf[i]), C=g(lq.C), D=g(lq.D), e=g(lq.e), nc=g(lq.nc), x_nom=g(lq.x_nom), u_nom=g(lq.u_nom),
                       time=lq.time, event=g(lq.event), jA=g(lq.jump_A), jHv=g(lq.jump_Hv), jQ=g(lq.jump_Q), jq=g(lq.jump_q), jc=g(lq.jump_c))


@pytest.mark.parametrize("algorithm", [o2.ALG_ILQR, o2.ALG_SLQ])
def test_host_upload_with_ragged_constraints_and_nominal(algorithm):
    """Uploaded host SoA data, per-stage varying number of active constraints (incl. nc = 0 and nc = nc_max), nominal trajectories."""
    rng = np.random.default_rng(21)
    batch, n, m, N, ncmax = 6, 5, 4, 30, 3
    lq = _random_batch(rng, batch, n, m, N, ncmax, algorithm)
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=1e-4, timeStep=0.02, preComputeRiccatiTerms=(algorithm == o2.ALG_ILQR))
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=ncmax, has_nominal=True) as solver:
        solver.upload(lq)
        solver.solve(alpha=0.8)
        sol = solver.download()
        for i in range(batch):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, alphas=(0.8,), what="ragged")


def test_solve_host_pipeline_equals_resident_path():
    rng = np.random.default_rng(5)
    batch, n, m, N = 37, 6, 3, 25
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR)
    lq.x_nom = lq.u_nom = None
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        solver.upload(lq)
        solver.solve()
        a = solver.download()
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        b = solver.solve_host(lq, alpha=1.0, chunk=5)   # 8 chunks over 3 lanes
    for name in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    check_against_oracle(st, _oracle_problem(lq, 36, N), lq.x0[36], b, 36, what="solve_host")


def test_indefinite_hamiltonian_sets_status_like_oracle():
    rng = np.random.default_rng(8)
    batch, n, m, N = 4, 4, 2, 10
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR)
    lq.x_nom = lq.u_nom = None
    lq.R[2, 5] = -np.eye(m)  # Hm = R + B'SB indefinite at stage 5 of problem 2
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        solver.upload(lq)
        solver.solve()
        sol = solver.download()
    assert sol.status[2] & o2.STATUS_CHOL_NOT_PD and sol.status[2] & o2.STATUS_NONFINITE
    assert (sol.status[[0, 1, 3]] == 0).all()
    ref = orc.backward(orc_settings(st), _oracle_problem(lq, 2, N))
    assert ref.status & 1
    assert np.array_equal(np.isnan(sol.K[2]), np.isnan(ref.K))
    check_against_oracle(st, _oracle_problem(lq, 1, N), lq.x0[1], sol, 1, what="neighbour of the failing problem")


def test_abi_rejects_bad_arguments():
    with pytest.raises(o2.O2cError):
        o2.BatchedLqSolver(o2.Settings(hessianCorrectionStrategy=o2.HC_CHOLESKY_MODIFICATION), 4, 2, 10, 1)
    with pytest.raises(o2.O2cError):
        o2.BatchedLqSolver(o2.Settings(), 4, 2, 10, 1, nc_max=3)
    with pytest.raises(o2.O2cError):
        o2.BatchedLqSolver(o2.Settings(algorithm=o2.ALG_SLQ, backwardPassIntegratorType="ODE45"), 4, 2, 10, 1)
    with o2.BatchedLqSolver(o2.Settings(), 4, 2, 10, 2) as s:
        with pytest.raises(o2.O2cError):
            s.solve(problem_begin=1, problem_count=5)


# ---- BASELINE.json's full sizes: sampled parity (the generator is counter based, so the oracle can regenerate any sampled
# ---- problem) plus an all-problems status check ----
FULL = [
    ("ballbot", o2.ALG_ILQR, 65536, 1e-3),
    ("quadrotor", o2.ALG_SLQ, 32768, 1e-3),
    ("manipulator", o2.ALG_ILQR, 16384, 1e-3),
    ("legged", o2.ALG_ILQR, 16384, 1e-5),
]


@pytest.mark.parametrize("shape,algorithm,batch,eps", FULL)
def test_full_size_configs_sampled_parity(shape, algorithm, batch, eps):
    n, m, nc = SHAPES[shape]
    N, dt, seed = 100, 0.01, 0
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=eps, timeStep=dt)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.generate_synthetic(seed, 0, dt)
        solver.solve()
        status = solver.download(n_alpha=0).status if batch <= 4096 else None
        # >= 64 sampled problems: both ends, the boundaries of the resident waves of the persistent kernels (12 problems per SM x 148
        # SMs = 1776 for the legged kernel; the first / last problem a warp or CTA picks up there), and a seeded random spread
        rng = np.random.default_rng(11)
        sample = {0, 1, 2, batch // 3, batch // 2 + 1, batch - 2, batch - 1}
        for wave in (148, 296, 1036, 1480, 1775, 1776, 1777, 2 * 1776 - 1, 2 * 1776, 2 * 1776 + 1, 2071, 2072, 2073, 9 * 1776 - 1, 9 * 1776):
            if wave < batch:
                sample.add(wave)
        sample |= {int(v) for v in rng.integers(0, batch, size=64 - len(sample) + 8)}
        sample = sorted(sample)
        assert len(sample) >= 64
        ost = orc_settings(st)
        for i in sample:
            sol = solver.download(problem_begin=i, problem_count=1)
            assert sol.status[0] == 0
            pb, x0 = orc.generate_problem(seed, i, algorithm, n, m, nc, N, dt)
            ref = check_against_oracle(st, pb, x0, sol, 0, what=f"full {shape} #{i}")
            if algorithm == o2.ALG_ILQR and i in sample[:8]:
                # size-independent property: V(x0) equals the rolled-out cost of the LQ model (alpha = 1, exact LQ data)
                V = 0.5 * x0 @ sol.Sm[0, 0] @ x0 + sol.Sv[0, 0] @ x0 + sol.s[0, 0]
                J = orc.discrete_lq_cost(pb, sol.x[0, 0], sol.u[0, 0])
                # the eps*I shift makes V an upper bound that is tight to O(eps)
                assert abs(V - J) <= 10 * eps * max(1.0, abs(J)) * N
                del ref
        # every problem finished without a status flag (download statuses only, in slabs)
        import ctypes as C

        from ocs2_b200 import lib as _l
        stat = np.zeros(batch, dtype=np.int32)
        sv = _l.SolutionView()
        sv.status = stat.ctypes.data
        _l.check(solver._lib.o2c_download(solver.handle, C.byref(sv), 0, batch, 0))
        assert (stat == 0).all()
        del status, ost


@pytest.mark.parametrize("shape,algorithm,eps,batch", [("legged", o2.ALG_ILQR, 1e-5, 2048), ("manipulator", o2.ALG_ILQR, 1e-3, 2048),
                                                       ("ballbot", o2.ALG_ILQR, 1e-3, 4096), ("quadrotor", o2.ALG_SLQ, 1e-3, 2048)])
def test_all_problems_of_a_batch_match_the_oracle(shape, algorithm, eps, batch):
    """EVERY problem of a batch (2048 legged problems = the per-GPU share of BASELINE config 5) against the oracle run on all host
    threads: true relative error per field and problem <= 1e-9."""
    n, m, nc = SHAPES[shape]
    N, dt, seed, first = 100, 0.01, 4, 7000
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=eps, timeStep=dt)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.generate_synthetic(seed, first, dt)
        solver.solve()
        worst = {}
        slab = 256
        for b0 in range(0, batch, slab):
            cnt = min(slab, batch - b0)
            sol = solver.download(problem_begin=b0, problem_count=cnt)
            ref = orc.batch_solve(orc_settings(st), seed, first + b0, cnt, n, m, nc, N, dt, alpha=1.0)
            assert (sol.status == 0).all() and (ref["status"] == 0).all()
            assert sol.x.shape[2] == ref["x"].shape[1]
            for name, got in (("K", sol.K), ("dbias", sol.dbias), ("bias", sol.bias), ("Sm", sol.Sm), ("Sv", sol.Sv), ("s", sol.s),
                              ("x", sol.x[0]), ("u", sol.u[0])):
                want = ref[name]
                axes = tuple(range(1, want.ndim))
                scale = np.abs(want).max(axis=axes)
                err = np.abs(got - want).max(axis=axes)
                rel = np.where(scale > 0, err / np.where(scale > 0, scale, 1.0), np.where(err > 0, np.inf, 0.0))
                worst[name] = max(worst.get(name, 0.0), float(rel.max()))
        print(f"all-problems parity {shape} x{batch}: worst true relative error per field {worst}")
        assert max(worst.values()) <= REL_TOL, worst


# ---- the shape-specialised DMMA/TMA kernel (nx = nu = 24, unconstrained, LINE_SEARCH, reduced, DIAGONAL_SHIFT) ----
@pytest.mark.parametrize("N", [1, 2, 3, 7, 100])
def test_legged_dmma_kernel_fused_solve_matches_oracle(N):
    """o2c_solve on the legged shape runs ONE fused kernel (sweep + rollout); odd/even/minimal horizons exercise the TMA ring parities."""
    n = m = 24
    batch, dt, seed, alpha = 40, 0.01, 2, 0.6
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=dt)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        assert solver.kernel_variant in ("ilqr_wpp_kernel", "ilqr_dmma_kernel")
        solver.generate_synthetic(seed, first_problem_index=300, dt=dt)
        l0 = solver.launch_count
        solver.solve(alpha=alpha)
        assert solver.launch_count - l0 == 1
        sol = solver.download()
        assert (sol.status == 0).all()
        for i in (0, 1, 17, batch - 1):
            pb, x0 = orc.generate_problem(seed, 300 + i, orc.ALG_ILQR, n, m, 0, N, dt)
            check_against_oracle(st, pb, x0, sol, i, (alpha,), what=f"dmma N={N}")
        # backward alone + generic multi-alpha rollout give the same controller / trajectories
        solver.solveSequentialRiccatiEquations()
        solver.rolloutTrajectory((alpha,))
        sol2 = solver.download()
        for name in ("K", "dbias", "bias", "Sm", "Sv", "s"):
            assert np.array_equal(getattr(sol, name), getattr(sol2, name)), name
        assert rel_err(sol2.x, sol.x) <= 1e-12 and rel_err(sol2.u, sol.u) <= 1e-12


def test_legged_dmma_kernel_uploaded_data_subrange_and_status():
    """Host-uploaded random 24x24 data through the DMMA kernel: sub-range solve, an indefinite Hm in one problem, neighbours unaffected."""
    rng = np.random.default_rng(77)
    batch, n, m, N = 9, 24, 24, 12
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR)
    lq.x_nom = lq.u_nom = None
    lq.R[4, 6] = -np.eye(m)
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        assert solver.kernel_variant in ("ilqr_wpp_kernel", "ilqr_dmma_kernel")
        solver.upload(lq)
        solver.solve(alpha=1.0, problem_begin=2, problem_count=6)
        sol = solver.download(problem_begin=2, problem_count=6)
    assert sol.status[2] & o2.STATUS_CHOL_NOT_PD and sol.status[2] & o2.STATUS_NONFINITE
    assert (np.delete(sol.status, 2) == 0).all()
    ref = orc.backward(orc_settings(st), _oracle_problem(lq, 4, N))
    assert ref.status & 1
    assert np.array_equal(np.isnan(sol.K[2]), np.isnan(ref.K))
    for i in (0, 1, 3, 5):
        check_against_oracle(st, _oracle_problem(lq, 2 + i, N), lq.x0[2 + i], sol, i, what="dmma uploaded")


# ---- the row-per-lane kernels (small shapes: several problems per warp) ----
@pytest.mark.parametrize("shape,algorithm,variant", [("ballbot", o2.ALG_ILQR, "ilqr_rpl_kernel"), ("manipulator", o2.ALG_ILQR, "ilqr_rpl_kernel"),
                                                     ("cartpole", o2.ALG_ILQR, "ilqr_rpl_kernel"), ("quadrotor", o2.ALG_SLQ, "slq_rpl_kernel"),
                                                     ("quadrotor", o2.ALG_ILQR, "ilqr_rpl_kernel"), ("ballbot", o2.ALG_SLQ, "slq_rpl_kernel"),
                                                     ("cartpole", o2.ALG_SLQ, "slq_rpl_kernel")])
@pytest.mark.parametrize("batch,begin,count", [(1, 0, 1), (2, 1, 1), (7, 0, 7), (11, 3, 5)])
def test_rpl_kernels_partial_warps_and_subranges(shape, algorithm, variant, batch, begin, count):
    """A warp carries 32 // nx problems: batches that do not fill the last warp, and sub-range solves, must not touch neighbours."""
    n, m, nc = SHAPES[shape]
    N, dt, seed = 9, 0.01, 5
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=1e-3, timeStep=dt)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, max_alphas=2) as solver:
        assert solver.kernel_variant == variant
        solver.generate_synthetic(seed, first_problem_index=20, dt=dt)
        solver.solve(alpha=0.7, problem_begin=begin, problem_count=count)
        sol = solver.download(problem_begin=begin, problem_count=count)
        assert (sol.status == 0).all()
        for i in sorted({0, count - 1}):
            pb, x0 = orc.generate_problem(seed, 20 + begin + i, algorithm, n, m, nc, N, dt)
            check_against_oracle(st, pb, x0, sol, i, (0.7,), what=f"rpl {shape} batch {batch}")
        # several step lengths through o2c_rollout (the continuous rollout of SLQ has its own row-per-lane kernel)
        solver.solveSequentialRiccatiEquations(begin, count)
        solver.rolloutTrajectory((1.0, 0.25), begin, count)
        sol2 = solver.download(problem_begin=begin, problem_count=count)
        pb, x0 = orc.generate_problem(seed, 20 + begin, algorithm, n, m, nc, N, dt)
        check_against_oracle(st, pb, x0, sol2, 0, (1.0, 0.25), what=f"rpl {shape} multi-alpha")


def test_rpl_constraints_rank_deficient_flag_and_ragged_fallback():
    """Manipulator shape through the range-space constraint path: a duplicated constraint row sets CONSTRAINT_RANK on that problem
    only; per-node (ragged) constraint counts stay on the row-per-lane kernel (inactive rows are masked), with the oracle's results."""
    rng = np.random.default_rng(5)
    batch, n, m, nc, N = 6, 9, 9, 3, 10
    lq = _random_batch(rng, batch, n, m, N, nc, o2.ALG_ILQR, ragged_nc=False)
    lq.x_nom = lq.u_nom = None
    lq.nc = None
    lq.D[2, 4, 1] = lq.D[2, 4, 0]  # rows 0 and 1 of D identical at node 4 of problem 2
    lq.C[2, 4, 1] = lq.C[2, 4, 0]
    lq.e[2, 4, 1] = lq.e[2, 4, 0]
    st = o2.Settings(hessianCorrectionMultiple=1e-3)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.upload(lq)
        assert solver.kernel_variant == "ilqr_rpl_kernel"
        solver.solve()
        sol = solver.download()
    assert sol.status[2] & o2.STATUS_CONSTRAINT_RANK
    assert (np.delete(sol.status, 2) == 0).all()
    lq.nc = np.full((batch, N), nc, np.int32)
    for i in (0, 1, 3, 5):
        check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, what="rpl constraints uploaded")
    # ragged counts: the same kernel (rows beyond a node's active count are masked)
    lq2 = _random_batch(rng, batch, n, m, N, nc, o2.ALG_ILQR, ragged_nc=True)
    lq2.x_nom = lq2.u_nom = None
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.upload(lq2)
        assert solver.kernel_variant == "ilqr_rpl_kernel"
        solver.solve()
        sol2 = solver.download()
    assert (sol2.status == 0).all()
    for i in range(batch):
        check_against_oracle(st, _oracle_problem(lq2, i, N), lq2.x0[i], sol2, i, what="ragged counts on the row-per-lane kernel")


# ---- the step after the backward pass: batched Armijo line search on the LQ model (o2c_line_search) ----
@pytest.mark.parametrize("shape", ["ballbot", "manipulator", "legged", "test32c"])
def test_line_search_matches_oracle(shape):
    n, m, nc = SHAPES[shape]
    N, batch, dt, seed = 20, 9, 0.01, 4
    st = o2.Settings(hessianCorrectionMultiple=1e-3 if shape != "legged" else 1e-5, timeStep=dt)
    ls = o2.LineSearchSettings(minStepLength=0.05, maxStepLength=1.0, contractionRate=0.5, armijoCoefficient=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, max_alphas=6) as solver:
        solver.generate_synthetic(seed, first_problem_index=0, dt=dt)
        solver.solveSequentialRiccatiEquations()
        sol = solver.download(n_alpha=0)
        res0 = solver.lineSearch(ls)
        assert len(res0.candidates) == 5 and np.allclose(res0.candidates, [1.0, 0.5, 0.25, 0.125, 0.0625])
        refs = []
        for i in range(batch):
            pb, x0 = orc.generate_problem(seed, i, orc.ALG_ILQR, n, m, nc, N, dt)
            refs.append((pb, x0, orc.backward(orc_settings(st), pb)))
        # On the exact LQ model alpha = 1 minimises the merit, so with the default coefficient candidate 0 always wins; larger Armijo
        # coefficients (the penalty scales with alpha) move the choice down the candidate list: thresholds from problem 0's merits
        base_all = res0.merits.max(axis=0) + 0.05 * (res0.merits.max(axis=0) - res0.merits.min(axis=0)) + 1e-12
        ce = (base_all[0] - res0.merits[:, 0]) / (res0.candidates * res0.controllerUpdateIS[0])
        coefficients = [1e-4, 0.5 * (ce[0] + ce[1]), 0.5 * (ce[2] + ce[3]), 10.0 * ce.max()]
        picked = set()
        for coef in coefficients:
            ls.armijoCoefficient = float(coef)
            base_in = None if coef == coefficients[0] else base_all
            res = solver.lineSearch(ls, baselineMerit=base_in)
            rolled = solver.download()
            for i, (pb, x0, ref) in enumerate(refs):
                step, idx, merits, b, IS, cands = orc.line_search(orc_settings(st), pb, ref, x0, ls.minStepLength, ls.maxStepLength,
                                                                  ls.contractionRate, ls.armijoCoefficient,
                                                                  None if base_in is None else base_in[i])
                assert rel_err(res.merits[:, i], merits) <= REL_TOL, f"{shape} #{i}: merits"
                assert abs(res.controllerUpdateIS[i] - IS) <= REL_TOL * max(1.0, abs(IS))
                assert abs(res.baselineMerit[i] - b) <= REL_TOL * max(1.0, abs(b))
                # the Armijo comparison is only decidable when it is not a tie to rounding
                margins = np.abs(merits - (b - ls.armijoCoefficient * cands * IS)) / max(1.0, abs(b))
                if margins.min() > 1e-9:
                    assert res.candidateIndex[i] == idx and res.stepLength[i] == step, f"{shape} #{i}: picked {res.candidateIndex[i]} vs {idx}"
                picked.add(int(res.candidateIndex[i]))
                # the rollouts of all candidates stay resident: candidate e is rollout e
                e = max(int(res.candidateIndex[i]), 0)
                x, u, _, _ = orc.rollout(orc_settings(st), pb, ref, x0, alpha=float(res.candidates[e]))
                assert rel_err(rolled.x[e, i], x) <= REL_TOL and rel_err(rolled.u[e, i], u) <= REL_TOL
        assert len(picked) >= 3, f"the coefficients should exercise several outcomes of the Armijo rule, got {picked}"
        del sol


def test_line_search_argument_checks():
    with o2.BatchedLqSolver(o2.Settings(), 4, 2, 10, 3, max_alphas=2) as s:
        s.generate_synthetic(0, 0, 0.01)
        with pytest.raises(o2.O2cError) as e:
            s.lineSearch()  # before the backward pass
        assert e.value.code == 5
        s.solveSequentialRiccatiEquations()
        with pytest.raises(o2.O2cError):
            s.lineSearch(o2.LineSearchSettings())  # 5 candidates > max_alphas = 2
        r = s.lineSearch(o2.LineSearchSettings(minStepLength=0.5))
        assert list(r.candidates) == [1.0, 0.5]
    with o2.BatchedLqSolver(o2.Settings(algorithm=o2.ALG_SLQ), 4, 2, 10, 3, max_alphas=6) as s:
        s.generate_synthetic(0, 0, 0.01)
        s.solveSequentialRiccatiEquations()
        with pytest.raises(o2.O2cError) as e:
            s.lineSearch()
        assert e.value.code == 2  # discrete model only


@pytest.mark.parametrize("shape,algorithm,variant", [("legged", o2.ALG_ILQR, "ilqr_wpp_kernel"), ("ballbot", o2.ALG_ILQR, "ilqr_rpl_kernel"),
                                                     ("manipulator", o2.ALG_ILQR, "ilqr_rpl_kernel"), ("cartpole", o2.ALG_ILQR, "ilqr_rpl_kernel"),
                                                     ("quadrotor", o2.ALG_SLQ, "slq_rpl_kernel"), ("quadrotor", o2.ALG_ILQR, "ilqr_rpl_kernel"),
                                                     ("ballbot", o2.ALG_SLQ, "slq_rpl_kernel"), ("cartpole", o2.ALG_SLQ, "slq_rpl_kernel"),
                                                     ("legged", o2.ALG_SLQ, "slq_wpp_kernel")])
def test_fast_kernels_with_nominal_trajectories(shape, algorithm, variant):
    """The real DDP iteration linearises about a nominal trajectory: bias = u_nom - K x_nom and the rollout is
    x_{k+1} = x_nom_{k+1} + A dx + B du + Hv (xdot = A (x - x_nom(t)) + B (u - u_nom(t)) + Hv for SLQ). The specialised kernels keep
    serving that case (has_nominal = 1)."""
    n, m, nc = SHAPES[shape]
    rng = np.random.default_rng(21)
    batch, N = 7, 11
    lq = _random_batch(rng, batch, n, m, N, nc, algorithm, ragged_nc=False)
    lq.nc = None
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=1e-4, timeStep=0.02)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True, max_alphas=2) as solver:
        solver.upload(lq)
        assert solver.kernel_variant == variant
        solver.solve(alpha=0.6)
        sol = solver.download()
        assert (sol.status == 0).all()
        if nc:
            lq.nc = np.full((batch, N), nc, np.int32)
        for i in (0, 3, batch - 1):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, (0.6,), what=f"{shape} with nominal")
        # the generic multi-alpha rollout on the same controller agrees with the fused one
        solver.solveSequentialRiccatiEquations()
        solver.rolloutTrajectory((0.6,))
        sol2 = solver.download()
        assert np.array_equal(sol.bias, sol2.bias)
        assert rel_err(sol2.x, sol.x) <= 1e-12 and rel_err(sol2.u, sol.u) <= 1e-12


@pytest.mark.parametrize("algorithm", [o2.ALG_ILQR, o2.ALG_SLQ])
@pytest.mark.parametrize("n,m,nc", [(5, 3, 0), (6, 4, 2), (24, 24, 0)])
def test_eigenvalue_modification_matches_oracle(algorithm, n, m, nc):
    """hessian_correction::EIGENVALUE_MODIFICATION (LinearAlgebra::makePsdEigenvalue): state costs made indefinite at some nodes so
    that the eigenvalue clamp really acts (other nodes take the symmetrise-only branch)."""
    rng = np.random.default_rng(3)
    batch, N = 4, 6
    lq = _random_batch(rng, batch, n, m, N, nc, algorithm, ragged_nc=False)
    lq.x_nom = lq.u_nom = None
    lq.nc = None
    nodes = lq.Q.shape[1]
    scale = np.abs(lq.Q).max()
    for b in range(batch):
        for k in range(0, nodes, 2):  # every other node: push part of the spectrum of Q below zero
            v = rng.uniform(-1, 1, (n, 2))
            lq.Q[b, k] -= 3.0 * scale * (v @ v.T) / n
    st = o2.Settings(algorithm=algorithm, hessianCorrectionStrategy=o2.HC_EIGENVALUE_MODIFICATION, hessianCorrectionMultiple=1e-3, timeStep=0.02)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        assert "generic" in solver.kernel_variant
        solver.upload(lq)
        solver.solve(alpha=1.0)
        sol = solver.download()
    if nc:
        lq.nc = np.full((batch, nodes), nc, np.int32)
    clamped = 0
    for i in range(batch):
        pb = _oracle_problem(lq, i, N)
        check_against_oracle(st, pb, lq.x0[i], sol, i, what=f"eigenvalue modification n={n}")
        clamped += int(np.linalg.eigvalsh(lq.Q[i, 0]).min() < 0)
    assert clamped > 0, "the test data should contain indefinite state costs"


@pytest.mark.parametrize("n,m,nc,nominal", [(5, 3, 0, False), (6, 4, 2, True), (24, 24, 0, False), (24, 24, 0, True), (10, 3, 0, True), (10, 3, 0, False),
                                               (9, 9, 3, True), (9, 9, 3, False), (4, 1, 0, False), (24, 24, 6, True), (24, 24, 13, False)])
def test_ilqr_events_match_oracle(n, m, nc, nominal):
    """Pre-event nodes (ILQR.cpp:263-295): value function through riccatiTransversalityConditions on the jump model data, controller
    entry from the regular data with Sm = 0, rollout through the jump map."""
    rng = np.random.default_rng(17 + n)
    batch, N = 5, 12
    lq = _random_batch(rng, batch, n, m, N, nc, o2.ALG_ILQR, ragged_nc=False)
    if not nominal:
        lq.x_nom = lq.u_nom = None
    event = np.zeros((batch, N), dtype=np.int32)
    for b in range(batch):
        for k in rng.choice(N, size=b % 3, replace=False):  # 0, 1 or 2 events per problem, different places (problem 0: none)
            event[b, k] = 1
            lq.A[b, k] = np.eye(n) + 0.3 * rng.uniform(-1, 1, (n, n))
            lq.Hv[b, k] = 0.1 * rng.uniform(-1, 1, n)
    event[1, N - 1] = 1  # an event at the last stage
    lq.event = event
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    ls = o2.LineSearchSettings()
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=nominal, max_alphas=6) as solver:
        solver.upload(lq)
        # the shape-specialised kernels carry the event branch themselves; every other shape goes through the generic kernels
        assert {24: "ilqr_wpp", 10: "ilqr_rpl", 9: "ilqr_rpl", 4: "ilqr_rpl"}.get(n, "generic") in solver.kernel_variant
        solver.solve(alpha=0.7)
        sol = solver.download()
        assert (sol.status == 0).all()
        res = solver.lineSearch(ls)
        for i in range(batch):
            pb = _oracle_problem(lq, i, N)
            ref = check_against_oracle(st, pb, lq.x0[i], sol, i, (0.7,), what=f"events n={n} #{i}")
            _, _, merits, _, _, _ = orc.line_search(orc_settings(st), pb, ref, lq.x0[i], ls.minStepLength, ls.maxStepLength, ls.contractionRate,
                                                    ls.armijoCoefficient)
            assert rel_err(res.merits[:, i], merits) <= REL_TOL
        # the host pipeline uploads in chunks: event flags travel with their chunk
        sol_h = solver.solve_host(lq, alpha=0.7, chunk=2)
        for name in ("K", "dbias", "Sm", "Sv", "s", "x", "u"):
            assert np.array_equal(getattr(sol_h, name), getattr(sol, name)), f"solve_host with events: {name}"
        # a later upload without events (or with an all-zero flag array) clears them: the specialised kernels come back where they exist
        lq.event = np.zeros_like(event)
        solver.upload(lq)
        assert ("generic" in solver.kernel_variant) == (n not in (24, 10, 9, 4))
        lq.event = None
        solver.upload(lq)
        assert ("generic" in solver.kernel_variant) == (n not in (24, 10, 9, 4))
        solver.solve(alpha=0.7)
        sol = solver.download()
        check_against_oracle(st, _oracle_problem(lq, 2, N), lq.x0[2], sol, 2, (0.7,), what=f"events cleared n={n}")
        lq.event = event
        solver.upload(lq)
        lq.event = None
        sol_h = solver.solve_host(lq, alpha=0.7, chunk=2)
        assert ("generic" in solver.kernel_variant) == (n not in (24, 10, 9, 4))
        for name in ("K", "dbias", "Sm", "Sv", "s", "x", "u"):
            assert rel_err(getattr(sol_h, name), getattr(sol, name)) <= 1e-12, f"solve_host after events were cleared: {name}"


@pytest.mark.parametrize("n,m,nominal", [(24, 24, False), (10, 3, True), (5, 2, True)])
def test_flattened_controller_matches_oracle(n, m, nominal):
    """o2c_download_flattened_controller = LinearController::flatten (LinearController.cpp:87-140) of the incremented controller at
    its own time stamps, converted to float32 on the device: bit-identical to the restated serialisation of the downloaded FP64
    arrays, and within float precision of the oracle's controller (the reference's own round-trip tolerance is 1e-6)."""
    rng = np.random.default_rng(23 + n)
    batch, N = 6, 15
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR, ragged_nc=False)
    if not nominal:
        lq.x_nom = lq.u_nom = None
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch, has_nominal=nominal) as solver:
        solver.upload(lq)
        with pytest.raises(o2.O2cError):
            solver.flatten(1.0)  # NOT_READY before the backward pass
        solver.solveSequentialRiccatiEquations()
        sol = solver.download()
        flat = solver.flatten(0.5)
        assert flat.shape == (batch, N + 1, m * (n + 1)) and flat.dtype == np.float32
        time = st.timeStep * np.arange(N + 1)
        for i in range(batch):
            want = orc.flatten_controller(time, sol.K[i], sol.bias[i], dbias=sol.dbias[i], alpha=0.5)
            assert np.array_equal(flat[i], want), f"problem {i}: serialisation differs"
            ref = orc.backward(orc_settings(st), _oracle_problem(lq, i, N))
            want_ref = orc.flatten_controller(time, ref.K, ref.bias, dbias=ref.dbias, alpha=0.5)
            assert rel_err(flat[i], want_ref) <= 1e-6
        part = solver.flatten(0.5, problem_begin=2, problem_count=3)
        assert np.array_equal(part, flat[2:5])


@pytest.mark.parametrize("alg,n,m,nc,events", [(o2.ALG_ILQR, 24, 24, 0, True), (o2.ALG_ILQR, 9, 9, 3, False), (o2.ALG_ILQR, 6, 3, 2, True),
                                                (o2.ALG_SLQ, 12, 4, 0, False)])
def test_import_device_matches_host_upload(alg, n, m, nc, events):
    """o2c_import_device: a producer that already lives on the device hands over strided SoA arrays (here: torch tensors with padded
    node and problem strides, per-node constraint counts and event flags in device memory). Same records as the host upload => the
    same solution, bit for bit."""
    import ctypes as C
    import torch
    from ocs2_b200 import lib as o2lib
    rng = np.random.default_rng(31 + n)
    batch, N = 7, 11
    lq = _random_batch(rng, batch, n, m, N, nc, alg, ragged_nc=nc > 0 and n == 6)
    nodes = lq.A.shape[1]
    if events:
        ev = np.zeros((batch, nodes), dtype=np.int32)
        ev[1, 3] = ev[4, 0] = ev[4, 7] = ev[6, nodes - 1] = 1
        lq.event = ev
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=1e-4)
    dev = torch.device("cuda:0")
    keep = []

    def dfield(arr, block, count):
        """(B, count, block) host buffer -> device tensor padded to (B + 1, count + 2, block + 3); returns the strided o2c_field"""
        if arr is None:
            return o2lib.Field(None, 0, 0)
        host = np.zeros((batch + 1, count + 2, block + 3))
        host[:batch, :count, :block] = np.asarray(arr, dtype=np.float64).reshape(batch, count, block)
        t = torch.from_numpy(host).to(dev)
        keep.append(t)
        return o2lib.Field(t.data_ptr(), (count + 2) * (block + 3), block + 3)

    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True) as a, \
            o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True) as b:
        a.upload(lq)
        a.solve(alpha=0.8)
        want = a.download()
        hv = lq.view(N)  # column-major host copies live in lq._keep
        k = lq._keep
        dv = o2lib.LqView()
        dv.A, dv.B, dv.Hv = dfield(k["A"], n * n, nodes), dfield(k["B"], n * m, nodes), dfield(k["Hv"], n, nodes)
        dv.Q, dv.P, dv.R = dfield(k["Q"], n * n, nodes), dfield(k["P"], m * n, nodes), dfield(k["R"], m * m, nodes)
        dv.q, dv.r, dv.c = dfield(k["q"], n, nodes), dfield(k["r"], m, nodes), dfield(k["c"], 1, nodes)
        if nc:
            dv.C, dv.D, dv.e = dfield(k["C"], nc * n, nodes), dfield(k["D"], nc * m, nodes), dfield(k["e"], nc, nodes)
            if k["nc"] is not None:
                t = torch.from_numpy(k["nc"]).to(dev)
                keep.append(t)
                dv.nc, dv.nc_problem_stride, dv.nc_node_stride = t.data_ptr(), nodes, 1
        dv.Qf, dv.qf, dv.cf = dfield(k["Qf"], n * n, 1), dfield(k["qf"], n, 1), dfield(k["cf"], 1, 1)
        dv.x_nom, dv.u_nom, dv.x0 = dfield(k["x_nom"], n, N + 1), dfield(k["u_nom"], m, N + 1), dfield(k["x0"], n, 1)
        if k["time"] is not None:
            t = torch.from_numpy(k["time"]).to(dev)
            keep.append(t)
            dv.time = t.data_ptr()
        if events:
            t = torch.zeros((batch, nodes + 5), dtype=torch.int32, device=dev)
            t[:, :nodes] = torch.from_numpy(k["event"]).to(dev)
            keep.append(t)
            dv.event, dv.event_problem_stride, dv.event_node_stride = t.data_ptr(), nodes + 5, 1
        torch.cuda.synchronize()
        b.import_device(dv)
        same_kernel = b.kernel_variant == a.kernel_variant
        assert same_kernel or nc  # device-side counts cannot be inspected: they are treated as ragged (generic kernel)
        b.solve(alpha=0.8)
        got = b.download()
        for name in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u", "status"):
            if same_kernel:
                assert np.array_equal(getattr(got, name), getattr(want, name)), f"{name} differs between import_device and upload"
            else:
                assert rel_err(getattr(got, name), getattr(want, name)) <= 1e-10, f"{name} differs between import_device and upload"
        assert hv is not None


def test_device_views_and_config_round_trip():
    """o2c_get_config returns what o2c_create took; o2c_device_lq_view describes the resident records well enough that another handle can
    import them device-to-device; o2c_device_solution_view points at the resident controller (checked with a raw cudaMemcpy)."""
    import ctypes as C
    from ocs2_b200 import lib as o2lib
    n, m, N, batch = 10, 3, 20, 9
    st = o2.Settings(hessianCorrectionMultiple=1e-3)
    with o2.BatchedLqSolver(st, n, m, N, batch) as a, o2.BatchedLqSolver(st, n, m, N, batch) as b:
        cfg = o2lib.Config()
        o2lib.check(a._lib.o2c_get_config(a.handle, C.byref(cfg)))
        assert (cfg.nx, cfg.nu, cfg.nc_max, cfg.num_stages, cfg.batch, cfg.algorithm) == (n, m, 0, N, batch, o2.ALG_ILQR)
        assert cfg.hessian_multiple == 1e-3 and cfg.time_step == st.timeStep
        assert a.compute_stream != 0
        a.generate_synthetic(seed=99, first_problem_index=1000)
        a.solve()
        want = a.download()
        b.import_device(a.device_lq_view())
        b.solve()
        got = b.download()
        for name in ("K", "dbias", "Sm", "Sv", "s", "x", "u"):
            assert np.array_equal(getattr(got, name), getattr(want, name)), name
        sv = a.device_solution_view()
        try:
            cudart = C.CDLL("libcudart.so")
        except OSError:
            cudart = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
        for p in (2, batch - 1):
            for node in (0, N):
                host = np.zeros(m * n)
                src = sv.K.ptr + 8 * (p * sv.K.problem_stride + node * sv.K.node_stride)
                assert cudart.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(src), C.c_size_t(host.nbytes), 2) == 0
                assert np.array_equal(host.reshape(n, m).T, want.K[p, node])


@pytest.mark.parametrize("alg,n,m,nc", [(o2.ALG_ILQR, 24, 24, 0), (o2.ALG_ILQR, 10, 3, 0), (o2.ALG_ILQR, 5, 2, 1), (o2.ALG_SLQ, 12, 4, 0), (o2.ALG_SLQ, 5, 2, 1)])
def test_empty_problem_range_is_a_no_op(alg, n, m, nc):
    """An empty problem range (count = 0) is accepted by every entry point and leaves the resident solution untouched."""
    rng = np.random.default_rng(3)
    batch, N = 4, 6
    lq = _random_batch(rng, batch, n, m, N, nc, alg, ragged_nc=False)
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True, max_alphas=6) as solver:
        solver.upload(lq)
        solver.solve()
        want = solver.download()
        solver.solveSequentialRiccatiEquations(problem_begin=2, problem_count=0)
        solver.rolloutTrajectory((1.0,), problem_begin=batch, problem_count=0)
        solver.solve(problem_begin=1, problem_count=0)
        if alg == o2.ALG_ILQR:
            assert solver.lineSearch(problem_begin=1, problem_count=0).stepLength.shape == (0,)
            solver.solve()
        assert solver.flatten(1.0, problem_begin=3, problem_count=0).shape[0] == 0
        assert solver.download(problem_begin=0, problem_count=0).K.shape[0] == 0
        got = solver.download()
        for name in ("K", "dbias", "Sm", "Sv", "s", "x", "u"):
            assert np.array_equal(getattr(got, name), getattr(want, name)), name


@pytest.mark.parametrize("n,m,nc", [(24, 24, 0), (10, 3, 0), (9, 9, 3), (12, 4, 0), (7, 2, 0)])
@pytest.mark.parametrize("N", [1, 2, 9])
def test_ilqr_events_dense_and_short_horizons(n, m, nc, N):
    """Event nodes everywhere (problem 0), at the first and last node only (problem 1), alternating (problem 2), none (problem 3):
    every kernel family against the oracle, horizons down to a single stage."""
    rng = np.random.default_rng(41 + n + N)
    batch = 4
    lq = _random_batch(rng, batch, n, m, N, nc, o2.ALG_ILQR, ragged_nc=False)
    event = np.zeros((batch, N), dtype=np.int32)
    event[0, :] = 1
    event[1, 0] = event[1, N - 1] = 1
    event[2, ::2] = 1
    for b in range(batch):
        for k in np.nonzero(event[b])[0]:
            lq.A[b, k] = np.eye(n) + 0.2 * rng.uniform(-1, 1, (n, n))
            lq.Hv[b, k] = 0.1 * rng.uniform(-1, 1, n)
    lq.event = event
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True) as solver:
        solver.upload(lq)
        solver.solve(alpha=0.6)
        sol = solver.download()
        for i in range(batch):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, (0.6,), what=f"dense events n={n} N={N} #{i}")


@pytest.mark.parametrize("n,m,nc,variant", [(12, 4, 0, "ls_reduced"), (10, 3, 0, "ls_reduced"), (4, 1, 0, "ls_reduced"), (5, 3, 2, "ls_full"), (6, 2, 0, "lm_full"),
                                            (24, 24, 0, "ls_reduced")])
def test_slq_events_match_oracle(n, m, nc, variant):
    """SLQ with events (SLQ.cpp:256-302): inter-event segments integrated separately, joined by computeJumpMap =
    riccatiTransversalityConditions on the event's jump model data; the continuous rollout restarts weakEpsilon after every event from
    the LQ jump map (TimeTriggeredRollout.cpp:46-115). Event nodes are shared by the batch, jump data are per problem."""
    rng = np.random.default_rng(53 + n)
    batch, N, dt = 5, 13, 0.02
    events = (3, 8)
    lq = _random_batch(rng, batch, n, m, N, nc, o2.ALG_SLQ, ragged_nc=False, dt=dt)
    time = np.zeros(N + 1)
    for k in range(1, N + 1):
        time[k] = time[k - 1] + (1e-9 if (k - 1) in events else dt)
    lq.time = time
    ev = np.zeros((batch, N + 1), dtype=np.int32)
    ev[:, list(events)] = 1
    E = len(events)
    lq.event = ev
    lq.jump_A = np.eye(n) + 0.3 * rng.uniform(-1, 1, (batch, E, n, n))
    lq.jump_Hv = 0.1 * rng.uniform(-1, 1, (batch, E, n))
    Mq = rng.uniform(-1, 1, (batch, E, n, n))
    lq.jump_Q = np.einsum("beij,beil->bejl", Mq, Mq) / n + 0.1 * np.eye(n)
    lq.jump_q = 0.2 * rng.uniform(-1, 1, (batch, E, n))
    lq.jump_c = rng.uniform(-1, 1, (batch, E))
    st = o2.Settings(algorithm=o2.ALG_SLQ, hessianCorrectionMultiple=1e-4, timeStep=0.007, preComputeRiccatiTerms=variant == "ls_reduced",
                     strategy=o2.STRATEGY_LEVENBERG_MARQUARDT if variant == "lm_full" else o2.STRATEGY_LINE_SEARCH,
                     riccatiMultiple=0.3 if variant == "lm_full" else 0.0)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=True, max_alphas=2) as solver:
        solver.upload(lq)
        # the quadrotor shape keeps its row-per-lane kernels, the legged shape its tensor-pipe kernels (jump map on tiles)
        assert ("slq_rpl" if n in (12, 10, 4) else ("slq_wpp" if (n == 24 and variant == "ls_reduced") else "generic")) in solver.kernel_variant
        solver.solveSequentialRiccatiEquations()
        solver.rolloutTrajectory((1.0, 0.4))
        sol = solver.download()
        assert (sol.status == 0).all()
        for i in range(batch):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, (1.0, 0.4), what=f"slq events n={n} #{i}")
        # the host pipeline (chunked uploads) takes the same path
        sol_h = solver.solve_host(lq, alpha=1.0, chunk=2)
        for name in ("K", "dbias", "Sm", "Sv", "s"):
            assert np.array_equal(getattr(sol_h, name), getattr(sol, name)), name
        assert np.array_equal(sol_h.x[0], sol.x[0])
        # event flags that differ between problems are rejected: the time grid (and with it the step schedule) is shared
        bad = ev.copy()
        bad[2, events[0]] = 0
        bad[2, events[0] + 1] = 1
        lq.event = bad
        with pytest.raises(o2.O2cError):
            solver.upload(lq)
        # without events the schedules are rebuilt and the shape-specialised kernels return where they exist
        lq.event = None
        lq.jump_A = lq.jump_Hv = lq.jump_Q = lq.jump_q = lq.jump_c = None
        lq.time = dt * np.arange(N + 1)
        solver.upload(lq)
        assert ("generic" in solver.kernel_variant) == (n not in (12, 10, 4, 24))
        solver.solve()
        sol = solver.download()
        check_against_oracle(st, _oracle_problem(lq, 1, N), lq.x0[1], sol, 1, (1.0,), what=f"slq events cleared n={n}")


@pytest.mark.parametrize("n,m,stages", [(4, 1, 4), (24, 24, 4), (10, 3, 1), (7, 5, 4)])
def test_discretize_matches_oracle(n, m, stages):
    """o2c_discretize = ILQR::discreteLQWorker on caller-supplied stage linearisations (rk4SensitivityDiscretization,
    SensitivityIntegratorImpl.cpp:130-169; cost *= dt, Hv := 0, ILQR.cpp:137-157): A, B of the resident records against the oracle's
    restatement, cost blocks scaled by the step length, a zero-length interval keeps the continuous-time data; then the discretised
    batch solves like the same data uploaded from the host."""
    import torch
    from ocs2_b200 import lib as o2lib
    rng = np.random.default_rng(61 + n)
    batch, N = 4, 8
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR, ragged_nc=False)
    dts = rng.uniform(0.005, 0.03, N)
    dts[5] = 0.0
    Ac = rng.uniform(-1, 1, (4, batch, N, n, n))
    Bc = rng.uniform(-1, 1, (4, batch, N, n, m))
    dev = torch.device("cuda:0")
    keep = []

    def dfield(arr, block):  # (batch, N, rows, cols) natural -> column-major blocks on the device, padded strides
        host = np.zeros((batch, N + 1, block + 2))
        host[:, :N, :block] = np.swapaxes(arr, -1, -2).reshape(batch, N, block)
        t = torch.from_numpy(host).to(dev)
        keep.append(t)
        return o2lib.Field(t.data_ptr(), (N + 1) * (block + 2), block + 2)

    dv = o2lib.DiscretizationView()
    for s in range(stages):
        dv.dfdx[s], dv.dfdu[s] = dfield(Ac[s], n * n), dfield(Bc[s], n * m)
    dv.dt = dts.ctypes.data
    dv.stages = stages
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch, has_nominal=True) as a, o2.BatchedLqSolver(st, n, m, N, batch, has_nominal=True) as b:
        a.upload(lq)
        torch.cuda.synchronize()
        a.discretize(dv)
        # the same thing on the host: oracle discretisation + cost * dt, uploaded
        want = o2.LqBatch(**{f: getattr(lq, f) for f in ("A", "B", "Q", "R", "Qf", "Hv", "P", "q", "r", "c", "qf", "cf", "x_nom", "u_nom", "x0", "time")})
        want.A, want.B, want.Hv = lq.A.copy(), lq.B.copy(), np.zeros_like(lq.Hv)
        want.Q, want.P, want.R, want.q, want.r, want.c = (getattr(lq, f).copy() for f in ("Q", "P", "R", "q", "r", "c"))
        for p in range(batch):
            for k in range(N):
                src = [s if stages == 4 else 0 for s in range(4)]
                if dts[k] == 0.0:
                    want.A[p, k], want.B[p, k], want.Hv[p, k] = Ac[0, p, k], Bc[0, p, k], lq.Hv[p, k]
                    continue
                want.A[p, k], want.B[p, k] = orc.rk4_sensitivity_discretization([Ac[s, p, k] for s in src], [Bc[s, p, k] for s in src], dts[k])
                for f in ("Q", "P", "R", "q", "r", "c"):
                    getattr(want, f)[p, k] *= dts[k]
        b.upload(want)
        a.solve(alpha=0.9)
        b.solve(alpha=0.9)
        sa, sb = a.download(), b.download()
        for name in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u"):
            assert rel_err(getattr(sa, name), getattr(sb, name)) <= 1e-10, name
        ref = orc.backward(orc_settings(st), _oracle_problem(want, 1, N))
        assert rel_err(sa.K[1], ref.K) <= REL_TOL and rel_err(sa.Sm[1], ref.Sm) <= REL_TOL
    with o2.BatchedLqSolver(o2.Settings(algorithm=o2.ALG_SLQ), n, m, N, batch) as slq:
        with pytest.raises(o2.O2cError):
            slq.discretize(dv)


def test_flattened_controller_in_slices():
    """The float conversion runs through a bounded device scratch: a batch larger than one slice comes back identical to the per-problem
    conversion."""
    n, m, N, batch = 24, 24, 100, 1300  # 1300 * 101 * 600 floats = 315 MB > one 256 MiB slice
    st = o2.Settings(hessianCorrectionMultiple=1e-5)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        solver.generate_synthetic(4, 0)
        solver.solveSequentialRiccatiEquations()
        flat = solver.flatten(1.0)
        assert flat.shape == (batch, N + 1, m * (n + 1)) and np.isfinite(flat).all()
        for p in (0, 1107, 1108, batch - 1):  # both sides of the slice boundary
            assert np.array_equal(flat[p], solver.flatten(1.0, problem_begin=p, problem_count=1)[0])


# ---- round 2: ABI v4 additions and the boundary fixes of ADVICE.md ----
@pytest.mark.parametrize("alg,n,m,nc", [(o2.ALG_ILQR, 24, 24, 0), (o2.ALG_ILQR, 9, 9, 3), (o2.ALG_SLQ, 12, 4, 0), (o2.ALG_ILQR, 5, 2, 0)])
def test_symmetric_packed_upload_is_bit_identical_to_dense(alg, n, m, nc):
    """O2C_LQ_SYMMETRIC_PACKED: Q, R, Qf as packed upper triangles (column by column) expand to the same records as the dense upload,
    through o2c_upload and through the chunked o2c_solve_host pipeline."""
    rng = np.random.default_rng(3 + n)
    batch, N = 7, 15
    lq = _random_batch(rng, batch, n, m, N, nc, alg, ragged_nc=False)
    lq.x_nom = lq.u_nom = None
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=1e-4, timeStep=0.02)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.upload(lq)
        solver.solve(0.9)
        dense = solver.download()
        solver.upload(lq, symmetric_packed=True)
        solver.solve(0.9)
        packed = solver.download()
        host = solver.solve_host(lq, alpha=0.9, chunk=3, symmetric_packed=True)
        slim = solver.solve_host(lq, alpha=0.9, chunk=3, symmetric_packed=True, want_value=False)
        for name in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u"):
            assert np.array_equal(getattr(dense, name), getattr(packed, name)), name
            assert np.array_equal(getattr(dense, name), getattr(host, name)), f"solve_host packed: {name}"
        assert slim.Sm is None and slim.Sv is None and slim.s is None  # the value function never crossed the bus
        for name in ("K", "dbias", "bias", "x", "u"):
            assert np.array_equal(getattr(dense, name), getattr(slim, name)), f"solve_host policy only: {name}"
        check_against_oracle(st, _oracle_problem(lq, 3, N), lq.x0[3], packed, 3, (0.9,), what="packed upload")


def test_pack_upper_order_is_the_reference_flatten_order():
    """pack_upper = upper triangle column by column = the Sm part of convert2Vector (golden order of RiccatiTest.cpp:116-126)."""
    S = np.arange(16, dtype=np.float64).reshape(4, 4)
    S = S + S.T
    assert np.array_equal(o2.pack_upper(S), orc.flatten(S, np.zeros(4), 0.0)[:10])


@pytest.mark.parametrize("n,m", [(24, 24), (6, 3)])
def test_check_numerical_stability_flags_non_psd_value_functions(n, m):
    """ddp checkNumericalStability_ (GaussNewtonDDP.cpp:555-579): a terminal cost with a negative eigenvalue makes S_N (and some S_k
    below it) fail checkBeingPSD; the check ors NOT_PSD into exactly those problems' status words and leaves the others clean."""
    rng = np.random.default_rng(9)
    batch, N = 6, 8
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR)
    lq.x_nom = lq.u_nom = None
    bad = (1, 4)
    for b in bad:
        v = rng.uniform(-1, 1, n)
        lq.Qf[b] = lq.Qf[b] - 3.0 * np.outer(v, v)  # indefinite but symmetric and finite
    lq.Qf[5] = lq.Qf[5] + np.triu(1e-3 * rng.uniform(-1, 1, (n, n)), 1)  # not self-adjoint (1e-3 relative >> 1e-6)
    st = o2.Settings(hessianCorrectionMultiple=1e-4)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        solver.upload(lq)
        solver.solveSequentialRiccatiEquations()
        before = solver.download(n_alpha=0).status.copy()
        solver.checkNumericalStability()
        after = solver.download(n_alpha=0).status
        for b in range(batch):
            want = b in bad or b == 5
            assert bool(after[b] & o2.STATUS_NOT_PSD) == want, (b, before, after)
            assert (after[b] & ~o2.STATUS_NOT_PSD) == before[b]
        # the reference's own criterion, evaluated on the host for every S_k of the flagged and unflagged problems
        sol = solver.download(n_alpha=0)
        for b in range(batch):
            finite = [S for S in sol.Sm[b] if np.isfinite(S).all()]  # non-finite entries are reported as NONFINITE instead
            if len(finite) < N + 1:
                assert after[b] & o2.STATUS_NONFINITE
            psd = all(np.linalg.eigvalsh(np.tril(S) + np.tril(S, -1).T).min() >= -np.finfo(float).eps and
                      np.linalg.norm(S - S.T) <= 1e-6 * np.linalg.norm(S) for S in finite)
            assert psd == (not (after[b] & o2.STATUS_NOT_PSD)), b


def test_ilqr_events_under_levenberg_marquardt_are_refused():
    """deltaGm / deltaGv of a pre-event node are built from the node's regular dynamics (ILQR.cpp:263-295), which the record layout
    replaces by the jump map: the library refuses instead of returning a different controller."""
    rng = np.random.default_rng(2)
    batch, n, m, N = 3, 5, 2, 6
    lq = _random_batch(rng, batch, n, m, N, 0, o2.ALG_ILQR)
    lq.x_nom = lq.u_nom = None
    lq.event = np.zeros((batch, N), dtype=np.int32)
    lq.event[1, 2] = 1
    st = o2.Settings(strategy=o2.STRATEGY_LEVENBERG_MARQUARDT, riccatiMultiple=0.3, preComputeRiccatiTerms=False)
    with o2.BatchedLqSolver(st, n, m, N, batch) as solver:
        solver.upload(lq)
        with pytest.raises(o2.O2cError) as e:
            solver.solveSequentialRiccatiEquations()
        assert e.value.code == 2  # O2C_ERR_UNSUPPORTED
        lq.event = None
        solver.upload(lq)  # without events the LM handle works
        solver.solve()
        check_against_oracle(st, _oracle_problem(lq, 1, N), lq.x0[1], solver.download(), 1, what="LM after events were cleared")


def test_rollout_before_backward_is_not_ready():
    with o2.BatchedLqSolver(o2.Settings(), 4, 2, 10, 2) as solver:
        solver.generate_synthetic(0, 0, 0.01)
        with pytest.raises(o2.O2cError) as e:
            solver.rolloutTrajectory((1.0,))
        assert e.value.code == 5  # O2C_ERR_NOT_READY


@pytest.mark.parametrize("alg", [o2.ALG_ILQR, o2.ALG_SLQ])
def test_chunked_whole_batch_upload_does_the_whole_range_bookkeeping(alg, monkeypatch):
    """o2c_upload splits uploads above 1 GiB (the legged batch is 38.8 GB); every chunk is a partial upload, so the whole-range
    bookkeeping happens once up front: (1) a first whole-batch SLQ upload WITH events works, (2) a whole-batch upload
    WITHOUT events / with uniform constraint counts brings the specialised kernel back after an upload that had events / ragged counts.
    O2C_UPLOAD_CHUNK_BYTES forces the split at test sizes."""
    rng = np.random.default_rng(13)
    n, m, nc = (9, 9, 3) if alg == o2.ALG_ILQR else (12, 4, 0)
    batch, N = 6, 10
    fast = "ilqr_rpl_kernel" if alg == o2.ALG_ILQR else "slq_rpl_kernel"
    lq = _random_batch(rng, batch, n, m, N, nc, alg, ragged_nc=True)
    if alg == o2.ALG_ILQR:
        lq.x_nom = lq.u_nom = None
    nodes = N + 1 if alg == o2.ALG_SLQ else N
    event = np.zeros((batch, nodes), dtype=np.int32)
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=1e-4, timeStep=0.02)
    nominal = alg == o2.ALG_SLQ  # (the SLQ rollout restarts from x_nom after an event)
    if alg == o2.ALG_SLQ:
        event[:, 4] = 1
        time = np.zeros(N + 1)
        for k in range(1, N + 1):  # node 5 is the post-event node, stamped weakEpsilon after node 4
            time[k] = time[k - 1] + (1e-9 if k - 1 == 4 else 0.02)
        lq.time = time
        lq.jump_A = np.eye(n) + 0.1 * rng.uniform(-1, 1, (batch, 1, n, n))
        lq.jump_Hv = 0.1 * rng.uniform(-1, 1, (batch, 1, n))
        Mq = rng.uniform(-1, 1, (batch, 1, n, n))
        lq.jump_Q = np.einsum("beij,beil->bejl", Mq, Mq) / n
        lq.jump_q = rng.uniform(-1, 1, (batch, 1, n))
        lq.jump_c = rng.uniform(0, 1, (batch, 1))
    else:
        event[2, 3] = event[4, 7] = 1
    lq.event = event
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc, has_nominal=nominal) as solver:
        assert solver.kernel_variant == fast
        per_problem = 8 * (nodes * 400 + 600)
        monkeypatch.setenv("O2C_UPLOAD_CHUNK_BYTES", str(2 * per_problem))  # 2-3 problems per chunk
        solver.upload(lq)  # (1): whole batch, split, with events (SLQ: used to fail with "a partial upload must carry the same events")
        solver.solve()
        sol = solver.download()
        for i in (0, 2, batch - 1):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, what=f"chunked upload with events #{i}")
        # (2): the same batch without events and with uniform counts, again split: back on the specialised kernel
        lq.event = None
        lq.jump_A = lq.jump_Hv = lq.jump_Q = lq.jump_q = lq.jump_c = None
        if alg == o2.ALG_SLQ:
            lq.time = 0.02 * np.arange(N + 1)
        if nc:
            lq.nc = np.full((batch, nodes), nc, np.int32)
        solver.upload(lq)
        assert solver.kernel_variant == fast
        solver.solve()
        sol = solver.download()
        for i in (1, 4):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, what=f"chunked upload, events cleared #{i}")


@pytest.mark.parametrize("shape,alg", [("legged", "ilqr"), ("ballbot", "ilqr"), ("quadrotor", "slq")])
def test_riccati_multiple_can_change_between_backward_passes(shape, alg):
    """LevenbergMarquardtStrategy adapts riccatiMultiple after every iteration (LevenbergMarquardtStrategy.cpp:131-147): the handle takes
    the new value without being rebuilt and the next pass equals the oracle evaluated with it."""
    n, m, nc = SHAPES[shape]
    algorithm = o2.ALG_ILQR if alg == "ilqr" else o2.ALG_SLQ
    N, batch, dt, seed = 30, 6, 0.01, 11
    st = o2.Settings(algorithm=algorithm, hessianCorrectionMultiple=1e-4, timeStep=dt, strategy=o2.STRATEGY_LEVENBERG_MARQUARDT, riccatiMultiple=0.05)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.generate_synthetic(seed, first_problem_index=3, dt=dt)
        for mu in (0.05, 0.8, 0.0):
            solver.setRiccatiMultiple(mu)
            assert solver.settings.riccatiMultiple == mu
            solver.solve(alpha=1.0)
            sol = solver.download()
            for i in (0, batch - 1):
                pb, x0 = orc.generate_problem(seed, 3 + i, algorithm, n, m, nc, N, dt)
                check_against_oracle(solver.settings, pb, x0, sol, i, what=f"{shape} mu={mu}")
        with pytest.raises(o2.O2cError):
            solver.setRiccatiMultiple(-1.0)


def test_device_count_matches_torch():
    import ctypes as C

    import torch

    from ocs2_b200 import lib as o2lib

    count = C.c_int32(-1)
    o2lib.check(o2lib.load_library().o2c_device_count(C.byref(count)))
    assert count.value == torch.cuda.device_count() >= 1


@pytest.mark.parametrize("ncmax", [5, 8, 11, 16])
@pytest.mark.parametrize("nominal", [False, True])
def test_legged_dmma_kernel_with_ragged_equality_constraints(ncmax, nominal):
    """The legged DMMA kernel carries up to 16 state-input equality constraints (contact constraints of the reference's legged robot,
    LeggedRobotInterface.cpp:186-190) as one or two 8-row tiles; the active count changes from node to node (contact switches), incl.
    nodes without any constraint and nodes with nc_max. Against the oracle's Householder-QR projection."""
    rng = np.random.default_rng(100 + ncmax)
    batch, n, m, N = 5, 24, 24, 12
    lq = _random_batch(rng, batch, n, m, N, ncmax, o2.ALG_ILQR, ragged_nc=True)
    lq.nc[0, :] = ncmax
    lq.nc[1, :] = 0
    if not nominal:
        lq.x_nom = lq.u_nom = None
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.02)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=ncmax, has_nominal=nominal, max_alphas=2) as solver:
        solver.upload(lq)
        assert solver.kernel_variant == "ilqr_wpp_kernel"
        solver.solve(alpha=0.7)
        sol = solver.download()
        assert (sol.status == 0).all()
        for i in range(batch):
            check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, alphas=(0.7,), what=f"legged nc<={ncmax}")
        # backward pass and rollout as separate calls
        solver.solveSequentialRiccatiEquations()
        solver.rolloutTrajectory((1.0, 0.25))
        sol = solver.download()
        check_against_oracle(st, _oracle_problem(lq, 3, N), lq.x0[3], sol, 3, alphas=(1.0, 0.25), what="legged constraints, two calls")
        # the chunked host pipeline (three stream lanes, constraint counts travel with their chunk) equals the resident path bit for bit
        solver.solve(alpha=0.7)
        resident = solver.download()
        piped = solver.solve_host(lq, alpha=0.7, chunk=2)
        for name in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u"):
            assert np.array_equal(getattr(piped, name), getattr(resident, name)), f"solve_host with constraints: {name}"


def test_legged_dmma_kernel_constraint_rank_flag():
    rng = np.random.default_rng(7)
    batch, n, m, nc, N = 4, 24, 24, 9, 8
    lq = _random_batch(rng, batch, n, m, N, nc, o2.ALG_ILQR, ragged_nc=False)
    lq.x_nom = lq.u_nom = None
    lq.nc = None
    lq.D[1, 3, 8] = lq.D[1, 3, 2]  # a duplicated constraint row (second tile against the first) at node 3 of problem 1
    lq.C[1, 3, 8] = lq.C[1, 3, 2]
    lq.e[1, 3, 8] = lq.e[1, 3, 2]
    st = o2.Settings(hessianCorrectionMultiple=1e-5)
    with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
        solver.upload(lq)
        assert solver.kernel_variant == "ilqr_wpp_kernel"
        solver.solve()
        sol = solver.download()
    assert sol.status[1] & o2.STATUS_CONSTRAINT_RANK
    assert (np.delete(sol.status, 1) == 0).all()
    lq.nc = np.full((batch, N), nc, np.int32)
    for i in (0, 2, 3):
        check_against_oracle(st, _oracle_problem(lq, i, N), lq.x0[i], sol, i, what="legged constraints, rank flag")


@pytest.mark.parametrize("nc", [0, 10])
def test_legged_dmma_kernel_schedule_does_not_change_a_bit(nc, monkeypatch):
    """Sweeper / roller warp roles, ring depth, dynamic problem fetch and the roller's own first sweep are scheduling only: every
    combination must reproduce the fused schedule (no rollers, static stride) bit for bit, for a batch of several rounds and for one
    smaller than the machine."""
    n = m = 24
    N, dt = 6, 0.01
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=dt)
    knobs = ("O2C_WPP_ROLLERS", "O2C_WPP_RESIDENT", "O2C_WPP_RING", "O2C_WPP_DYNAMIC", "O2C_WPP_ROLLER_SWEEPS", "O2C_WPP_WIDE")
    configs = [dict(O2C_WPP_ROLLERS="0", O2C_WPP_DYNAMIC="0", O2C_WPP_WIDE="0"), dict(), dict(O2C_WPP_ROLLERS="2"), dict(O2C_WPP_ROLLERS="3", O2C_WPP_RESIDENT="2"),
               dict(O2C_WPP_RESIDENT="1"), dict(O2C_WPP_RING="1"), dict(O2C_WPP_ROLLER_SWEEPS="0", O2C_WPP_RESIDENT="5"),
               dict(O2C_WPP_ROLLERS="0", O2C_WPP_RESIDENT="3"), dict(O2C_WPP_WIDE="0")]
    # 2000 problems: one round of 14 sweeps per SM on the 14-warp instantiation (the default there) against the two-round schedules
    for batch in (4000, 37, 2000):
        with o2.BatchedLqSolver(st, n, m, N, batch, nc_max=nc) as solver:
            solver.generate_synthetic(9, 0, dt)
            ref = None
            for cfg in configs:
                for k in knobs:
                    monkeypatch.delenv(k, raising=False)
                for k, v in cfg.items():
                    monkeypatch.setenv(k, v)
                solver.solve(alpha=0.9)
                sol = solver.download()
                assert (sol.status == 0).all()
                got = (sol.K, sol.dbias, sol.Sm, sol.Sv, sol.s, sol.x, sol.u)
                if ref is None:
                    ref = got
                    pb, x0 = orc.generate_problem(9, batch - 1, orc.ALG_ILQR, n, m, nc, N, dt)
                    check_against_oracle(st, pb, x0, sol, batch - 1, alphas=(0.9,), what="schedule reference")
                else:
                    for a, b in zip(ref, got):
                        assert np.array_equal(a, b), f"batch {batch}, {cfg}"
