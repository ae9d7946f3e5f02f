"""Quick start (Python mirror of the C ABI): solve a batch of random legged-size LQ problems on cuda:0, run the batched line search,
fetch the policy in the ocs2_msgs wire format. Needs a CUDA device (the library has no CPU fallback).

    python -m ocs2_b200.build && python examples/quickstart.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2  # noqa: E402

n, m, N, batch, dt = 24, 24, 100, 2048, 0.01
rng = np.random.default_rng(0)
W = rng.uniform(-1, 1, (batch, N, n + m, n + m))
W = np.einsum("bkij,bkil->bkjl", W, W) / (n + m) + 0.1 * np.eye(n + m)          # [[Q, P'], [P, R]] positive definite
Mf = rng.uniform(-1, 1, (batch, n, n))
lq = o2.LqBatch(
    A=np.eye(n) + dt * rng.uniform(-1, 1, (batch, N, n, n)), B=dt * rng.uniform(-1, 1, (batch, N, n, m)),  # ModelData.dynamics
    Hv=np.zeros((batch, N, n)),                                                                          # dynamicsBias
    Q=dt * W[..., :n, :n], P=dt * W[..., n:, :n], R=dt * W[..., n:, n:],                                   # cost.dfdxx, dfdux, dfduu
    q=dt * rng.uniform(-1, 1, (batch, N, n)), r=dt * rng.uniform(-1, 1, (batch, N, m)), c=np.zeros((batch, N)),
    Qf=np.einsum("bij,bil->bjl", Mf, Mf) / n + 0.1 * np.eye(n), qf=np.zeros((batch, n)), cf=np.zeros(batch),
    x0=rng.uniform(-1, 1, (batch, n)), time=dt * np.arange(N + 1))

settings = o2.Settings(algorithm=o2.ALG_ILQR, hessianCorrectionMultiple=1e-5)    # the ddp::Settings fields that reach the arithmetic
with o2.BatchedLqSolver(settings, n, m, N, batch, max_alphas=6) as solver:
    solver.upload(lq)
    solver.solveSequentialRiccatiEquations()                # backward pass + calculateController, all problems in one launch
    ls = solver.lineSearch(o2.LineSearchSettings())         # Armijo line search on the LQ model, all candidates in one launch
    sol = solver.download(n_alpha=1)                        # LinearController arrays, value function, rollout of candidate 0
    policy = solver.flatten(stepLength=1.0)                 # float32 [uff_i, K_i,:] rows (mpc_flattened_controller payload)
    print("kernel:", solver.kernel_variant)
    print("status ok:", bool((sol.status == 0).all()), " chosen step lengths:", np.unique(ls.stepLength))
    print("K", sol.K.shape, "Sm", sol.Sm.shape, "x", sol.x.shape, "flattened policy", policy.shape, policy.dtype)
