"""profiling driver (no torch): legged ILQR solve on `batch` synthetic problems, `reps` launches. Used under ncu."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)
with o2.BatchedLqSolver(st, 24, 24, 100, batch) as s:
    s.generate_synthetic(1, 0, 0.01)
    s.sync()
    for _ in range(reps):
        t0 = time.perf_counter()
        s.solve(1.0)
        s.sync()
        dt = time.perf_counter() - t0
        print(f"{s.kernel_variant} batch {batch}: {dt*1e3:.2f} ms -> {batch/dt:.0f} solves/s", flush=True)
    st_ = s.download(problem_begin=0, problem_count=min(batch, 64), n_alpha=0).status
    assert (st_ == 0).all()
