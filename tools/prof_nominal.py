"""legged ILQR with and without nominal trajectories, and with events: time per solve of 16384 problems. Usage: prof_nominal.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ocs2_b200 as o2
st = o2.Settings(hessianCorrectionMultiple=1e-5)
B, N = 16384, 100
for nominal in (False, True):
    with o2.BatchedLqSolver(st, 24, 24, N, B, has_nominal=nominal) as s:
        s.generate_synthetic(1, 0, 0.01); s.sync()
        for r in range(4):
            t0 = time.perf_counter(); s.solve(1.0); s.sync(); dt = time.perf_counter() - t0
        print(f"legged nominal={nominal} {s.kernel_variant}: {dt*1e3:.2f} ms -> {B/dt:.0f} solves/s", flush=True)
