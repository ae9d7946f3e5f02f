"""Generates tests/golden/cartpole_ilqr.npz — the config[0] correctness anchor (cartpole ILQR, nx=4, nu=1, N=100).

The reference builds this problem through CppADCodeGen JIT (not runnable here: no Eigen/Boost/CppAD toolchain), so the LQ data
is restated analytically from the reference sources:
  * dynamics      ocs2_robotic_examples/ocs2_cartpole/include/ocs2_cartpole/dynamics/CartPoleSystemDynamics.h:57-76
  * parameters    .../include/ocs2_cartpole/CartPoleParameters.h:76-92 with config/mpc/task.info:2-9
  * cost          task.info:89-112 (Q = 0, R = 0.1, Q_final = diag(5,1,1,1)), x0 task.info:80-86
  * discretise    ILQR::discreteLQWorker ocs2_ddp/src/ILQR.cpp:137-157 (RK4 sensitivity
                  ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:130-169, cost * dt, Hv := 0)
The nominal trajectory is an open-loop RK4 rollout under a small deterministic input so that x_nom, u_nom are non-trivial.
The stored outputs come from the CPU oracle (oracle/lq_oracle.cpp); the KKT oracle cross-checks them in tests/test_oracle_golden.py.

Run:  python tests/golden/make_cartpole_fixture.py
"""
import os
import sys

import numpy as np
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

cartMass, poleMass, poleLength, gravity = 2.0, 0.2, 1.0, 9.81
poleHalfLength = poleLength / 2.0
poleMoi = 1.0 / 12.0 * poleMass * poleLength**2
poleSteinerMoi = poleMoi + poleMass * poleHalfLength**2

th, px, thd, pxd, uu = sp.symbols("th px thd pxd u")
I = sp.Matrix([[poleSteinerMoi, poleMass * poleHalfLength * sp.cos(th)], [poleMass * poleHalfLength * sp.cos(th), cartMass + poleMass]])
rhs = sp.Matrix([poleMass * poleHalfLength * gravity * sp.sin(th), uu + poleMass * poleHalfLength * thd**2 * sp.sin(th)])
acc = I.inv() * rhs
f_sym = sp.Matrix([thd, pxd, acc[0], acc[1]])
state = sp.Matrix([th, px, thd, pxd])
f_num = sp.lambdify((th, px, thd, pxd, uu), f_sym, "numpy")
dfdx_num = sp.lambdify((th, px, thd, pxd, uu), f_sym.jacobian(state), "numpy")
dfdu_num = sp.lambdify((th, px, thd, pxd, uu), f_sym.jacobian(sp.Matrix([uu])), "numpy")


def lin(x, u):
    return (np.asarray(f_num(*x, u), dtype=float).reshape(4), np.asarray(dfdx_num(*x, u), dtype=float).reshape(4, 4),
            np.asarray(dfdu_num(*x, u), dtype=float).reshape(4, 1))


def rk4_sensitivity(x, u, dt):
    """rk4SensitivityDiscretization, SensitivityIntegratorImpl.cpp:130-169 (time-invariant system)."""
    f1, A1, B1 = lin(x, u)
    f2, A2, B2 = lin(x + dt / 2 * f1, u)
    f3, A3, B3 = lin(x + dt / 2 * f2, u)
    f4, A4, B4 = lin(x + dt * f3, u)
    B2 = B2 + dt / 2 * A2 @ B1
    B3 = B3 + dt / 2 * A3 @ B2
    B4 = B4 + dt * A4 @ B3
    A2 = A2 + dt / 2 * A2 @ A1
    A3 = A3 + dt / 2 * A3 @ A2
    A4 = A4 + dt * A4 @ A3
    A = dt / 6 * A1 + dt / 3 * A2 + dt / 3 * A3 + dt / 6 * A4 + np.eye(4)
    B = dt / 6 * B1 + dt / 3 * B2 + dt / 3 * B3 + dt / 6 * B4
    xn = x + dt / 6 * f1 + dt / 3 * f2 + dt / 3 * f3 + dt / 6 * f4
    return A, B, xn


def main():
    N, T = 100, 5.0
    dt = T / N
    n, m = 4, 1
    Qc = np.zeros((n, n))
    Rc = np.array([[0.1]])
    Qfinal = np.diag([5.0, 1.0, 1.0, 1.0])
    x = np.array([3.14, 0.0, 0.0, 0.0])
    x_nom = np.zeros((N + 1, n))
    u_nom = np.zeros((N + 1, m))
    A = np.zeros((N, n, n))
    B = np.zeros((N, n, m))
    for k in range(N):
        u = 0.5 * np.sin(0.3 * k)
        x_nom[k] = x
        u_nom[k, 0] = u
        A[k], B[k], x = rk4_sensitivity(x, u, dt)
    x_nom[N] = x
    u_nom[N] = u_nom[N - 1]
    # quadratic approximation of the cost around the nominal, times dt (ILQR.cpp:149-150); x_ref = 0, u_ref = 0
    Q = np.repeat((Qc * dt)[None], N, 0)
    R = np.repeat((Rc * dt)[None], N, 0)
    P = np.zeros((N, m, n))
    q = np.stack([dt * Qc @ x_nom[k] for k in range(N)])
    r = np.stack([dt * Rc @ u_nom[k] for k in range(N)])
    c = np.array([dt * 0.5 * (x_nom[k] @ Qc @ x_nom[k] + u_nom[k] @ Rc @ u_nom[k]) for k in range(N)])
    eps = 1e-6
    Qf = Qfinal + eps * np.eye(n)  # hessian-corrected final cost (GaussNewtonDDP.cpp:724-727, DIAGONAL_SHIFT)
    qf = Qfinal @ x_nom[N]
    cf = 0.5 * x_nom[N] @ Qfinal @ x_nom[N]
    pb = orc.Problem(N=N, A=A, B=B, Hv=np.zeros((N, n)), Q=Q, P=P, R=R, q=q, r=r, c=c, Qf=Qf, qf=qf, cf=float(cf), x_nom=x_nom,
                     u_nom=u_nom, time=dt * np.arange(N + 1))
    st = orc.make_settings(algorithm=orc.ALG_ILQR, reduced_form=True, strategy=orc.STRATEGY_LINE_SEARCH,
                           hessian_correction=orc.HC_DIAGONAL_SHIFT, hessian_multiple=eps, time_step=dt)
    sol = orc.backward(st, pb)
    assert sol.status == 0
    x0 = x_nom[0] + np.array([0.05, -0.02, 0.01, 0.03])
    xs, us, _, status = orc.rollout(st, pb, sol, x0, alpha=1.0)
    assert status == 0
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cartpole_ilqr.npz")
    np.savez_compressed(out, N=N, dt=dt, eps=eps, A=A, B=B, Q=Q, P=P, R=R, q=q, r=r, c=c, Qf=Qf, qf=qf, cf=cf, x_nom=x_nom, u_nom=u_nom,
                        x0=x0, K=sol.K, dbias=sol.dbias, bias=sol.bias, Sm=sol.Sm, Sv=sol.Sv, s=sol.s, x=xs, u=us)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
