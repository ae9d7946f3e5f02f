"""profiling driver (no torch): `reps` solves of a named workload on `batch` synthetic problems. Usage: prof_shape.py workload batch reps"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2
W = {"legged": (24, 24, 0, 0, 1e-5), "ballbot": (10, 3, 0, 0, 1e-3), "quadrotor_slq": (12, 4, 0, 1, 1e-3), "manipulator": (9, 9, 3, 0, 1e-3), "cartpole": (4, 1, 0, 0, 1e-6)}
name = sys.argv[1]; batch = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, m, nc, alg, eps = W[name]
st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=0.01)
with o2.BatchedLqSolver(st, n, m, 100, batch, nc_max=nc) as s:
    s.generate_synthetic(1, 0, 0.01); s.sync()
    for _ in range(reps):
        t0 = time.perf_counter(); s.solve(1.0); s.sync(); dt = time.perf_counter() - t0
        print(f"{name} {s.kernel_variant} batch {batch}: {dt*1e3:.2f} ms -> {batch/dt:.0f} solves/s", flush=True)
    assert (s.download(problem_begin=0, problem_count=min(batch, 64), n_alpha=0).status == 0).all()
