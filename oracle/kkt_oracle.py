"""Second, independent oracle (TEST INFRASTRUCTURE): dense KKT solve of a discrete LQ optimal control problem.

Restates what the reference's own DDP correctness test compares against: ocs2_test_tools/ocs2_qp_solver
(getConstraintMatrices QpSolver.cpp:103-160, getCostMatrices :162-203, solveDenseQp :222-239), i.e.
min_z 1/2 z'Hz + g'z + c0  s.t.  Gz = b with z = [x_0, u_0, x_1, u_1, ..., x_N], and the dynamics, initial-state and
state-input equality constraints stacked in G. Used by ocs2_ddp/test/CorrectnessTest.cpp:215-223.
"""
from __future__ import annotations

import numpy as np


def solve_discrete_lq(pb, x0):
    """pb: oracle.Problem (ILQR layout, nodes = N, nominal trajectories must be None/zero). Returns x (N+1,n), u (N,m), cost."""
    n, m, N = pb.nx, pb.nu, pb.N
    nz = (N + 1) * n + N * m

    def xi(k):
        return slice(k * (n + m), k * (n + m) + n)

    def ui(k):
        return slice(k * (n + m) + n, (k + 1) * (n + m))

    H = np.zeros((nz, nz))
    g = np.zeros(nz)
    c0 = 0.0
    for k in range(N):
        H[xi(k), xi(k)] += pb.Q[k]
        H[ui(k), ui(k)] += pb.R[k]
        H[ui(k), xi(k)] += pb.P[k]
        H[xi(k), ui(k)] += pb.P[k].T
        g[xi(k)] += pb.q[k]
        g[ui(k)] += pb.r[k]
        c0 += pb.c[k]
    H[xi(N), xi(N)] += pb.Qf
    g[xi(N)] += pb.qf
    c0 += pb.cf

    rows = []
    rhs = []
    # initial state
    G0 = np.zeros((n, nz))
    G0[:, xi(0)] = np.eye(n)
    rows.append(G0)
    rhs.append(np.asarray(x0, dtype=float))
    for k in range(N):
        Gd = np.zeros((n, nz))
        Gd[:, xi(k)] = pb.A[k]
        Gd[:, ui(k)] = pb.B[k]
        Gd[:, xi(k + 1)] = -np.eye(n)
        rows.append(Gd)
        rhs.append(-pb.Hv[k])
        nc = 0 if pb.D is None else (int(pb.nc[k]) if pb.nc is not None else pb.D.shape[1])
        if nc > 0:
            Gc = np.zeros((nc, nz))
            Gc[:, xi(k)] = pb.C[k][:nc]
            Gc[:, ui(k)] = pb.D[k][:nc]
            rows.append(Gc)
            rhs.append(-pb.e[k][:nc])
    G = np.vstack(rows)
    b = np.concatenate(rhs)
    ncon = G.shape[0]
    KKT = np.block([[H, G.T], [G, np.zeros((ncon, ncon))]])
    sol = np.linalg.solve(KKT, np.concatenate([-g, b]))
    z = sol[:nz]
    x = np.stack([z[xi(k)] for k in range(N + 1)])
    u = np.stack([z[ui(k)] for k in range(N)])
    cost = 0.5 * z @ H @ z + g @ z + c0
    return x, u, cost
