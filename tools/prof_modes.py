"""solve (sweep + fused rollout) against backward-only time of a BASELINE shape. Usage: prof_modes.py workload batch [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2

W = {"legged_slq": (24, 24, 0, 1, 1e-5), "legged": (24, 24, 0, 0, 1e-5), "quadrotor_slq": (12, 4, 0, 1, 1e-3), "manipulator": (9, 9, 3, 0, 1e-3),
     "ballbot": (10, 3, 0, 0, 1e-3), "quadrotor": (12, 4, 0, 0, 1e-3), "cartpole": (4, 1, 0, 0, 1e-3)}
name, batch = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n, m, nc, alg, eps = W[name]
st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=0.01)
with o2.BatchedLqSolver(st, n, m, 100, batch, nc_max=nc) as s:
    s.generate_synthetic(1, 0, 0.01)
    s.sync()
    for mode, fn in (("solve", lambda: s.solve(1.0)), ("backward", s.solveSequentialRiccatiEquations), ("rollout", lambda: s.rolloutTrajectory((1.0,)))):
        for _ in range(2):
            fn()
        s.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        s.sync()
        ms = (time.perf_counter() - t0) / steps * 1e3
        print(json.dumps({"workload": name, "kernel": s.kernel_variant, "mode": mode, "batch": batch, "ms": round(ms, 4), "per_s": round(batch / ms * 1e3)}), flush=True)
