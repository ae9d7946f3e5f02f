"""small end-to-end runs of every specialised kernel, for compute-sanitizer (memcheck / racecheck): tiny batches, short horizons"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ocs2_b200 as o2
cases = [("legged", 24, 24, 0, o2.ALG_ILQR, 1e-5, 9), ("ballbot", 10, 3, 0, o2.ALG_ILQR, 1e-3, 7), ("manipulator", 9, 9, 3, o2.ALG_ILQR, 1e-3, 7),
         ("cartpole", 4, 1, 0, o2.ALG_ILQR, 1e-6, 17), ("quadrotor", 12, 4, 0, o2.ALG_SLQ, 1e-3, 5)]
for name, n, m, nc, alg, eps, batch in cases:
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=0.01)
    with o2.BatchedLqSolver(st, n, m, 6, batch, nc_max=nc, max_alphas=6) as s:
        s.generate_synthetic(1, 0, 0.01)
        s.solve(0.8)
        s.solveSequentialRiccatiEquations()
        s.rolloutTrajectory((1.0, 0.5))
        if alg == o2.ALG_ILQR:
            s.lineSearch()
        sol = s.download()
        assert (sol.status == 0).all() and np.isfinite(sol.x).all()
        print(name, s.kernel_variant, "ok", flush=True)
