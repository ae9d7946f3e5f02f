"""Stress: repeated create/solve/destroy of handles of every kernel family; device memory must return to its starting level and two
interleaved handles must not disturb each other. Usage: stress_handles.py [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ocs2_b200 as o2

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch.cuda.init()
free0, _ = torch.cuda.mem_get_info()
shapes = [(o2.ALG_ILQR, 24, 24, 0), (o2.ALG_ILQR, 10, 3, 0), (o2.ALG_ILQR, 9, 9, 3), (o2.ALG_SLQ, 12, 4, 0), (o2.ALG_ILQR, 7, 5, 2), (o2.ALG_SLQ, 5, 2, 1)]
ref = {}
for r in range(rounds):
    for alg, n, m, nc in shapes:
        st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=1e-4)
        with o2.BatchedLqSolver(st, n, m, 30, 257, nc_max=nc, max_alphas=3) as a, o2.BatchedLqSolver(st, n, m, 30, 100, nc_max=nc) as b:
            a.generate_synthetic(7, 0)
            b.generate_synthetic(7, 0)
            a.solve(); b.solve(); a.rolloutTrajectory((1.0, 0.5, 0.25)); b.solve()
            sa, sb = a.download(problem_count=100, n_alpha=1), b.download()
            assert np.array_equal(sa.K, sb.K) and (sa.status == 0).all()
            # a's rollout went through the multi-alpha rollout kernel, b's through the fused one: same numbers up to summation order
            assert np.abs(sa.x[0] - sb.x[0]).max() <= 1e-11 * max(1.0, np.abs(sb.x[0]).max())
            key = (alg, n, m, nc)
            if key in ref:
                assert np.array_equal(ref[key], sb.K), "results changed between rounds"
            ref[key] = sb.K
free1, _ = torch.cuda.mem_get_info()
print(f"rounds {rounds}: device memory free before {free0 >> 20} MiB, after {free1 >> 20} MiB, leak {(free0 - free1) >> 20} MiB")
assert free0 - free1 < 64 << 20, "device memory leak"
print("stress ok")
