"""backward pass and rollout timed separately (CUDA events on the compute stream). Usage: prof_split.py workload batch [reps]
workload: legged_slq | legged | quadrotor_slq | manipulator | ballbot"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ocs2_b200 as o2

W = {"legged_slq": (24, 24, 0, 1, 1e-5), "legged": (24, 24, 0, 0, 1e-5), "quadrotor_slq": (12, 4, 0, 1, 1e-3), "manipulator": (9, 9, 3, 0, 1e-3),
     "ballbot": (10, 3, 0, 0, 1e-3)}
name = sys.argv[1]
batch = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, m, nc, alg, eps = W[name]
st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=0.01)
with o2.BatchedLqSolver(st, n, m, 100, batch, nc_max=nc) as s:
    s.generate_synthetic(1, 0, 0.01)
    s.sync()
    stream = torch.cuda.ExternalStream(s.compute_stream)
    for _ in range(2):
        s.solveSequentialRiccatiEquations()
        s.rolloutTrajectory((1.0,))
    s.sync()
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        s.solveSequentialRiccatiEquations()
        ev[1].record(stream)
        s.rolloutTrajectory((1.0,))
        ev[2].record(stream)
        s.sync()
        b, r = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        print(f"{name} batch {batch} {s.kernel_variant}: backward {b:.2f} ms, rollout {r:.2f} ms -> {batch / (b + r) * 1e3:.0f} solves/s (backward alone {batch / b * 1e3:.0f})", flush=True)
    assert (s.download(problem_count=min(batch, 64), n_alpha=0).status == 0).all()
