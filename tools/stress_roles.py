"""Stress of the sweeper / roller scheduling of the legged kernel: many launches over random problem ranges and horizons of a few stages
(queue hand-over, tail draining and the roller's own sweep happen thousands of times per second), every result compared bit for bit
with the fused schedule. Usage: stress_roles.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ocs2_b200 as o2

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(0)
st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)
launches = 0
for N, batch, nc in ((3, 6000, 0), (7, 3000, 0), (2, 5000, 9)):
    with o2.BatchedLqSolver(st, 24, 24, N, batch, nc_max=nc) as s:
        s.generate_synthetic(5, 0, 0.01)
        for k in ("O2C_WPP_ROLLERS", "O2C_WPP_RESIDENT", "O2C_WPP_WIDE", "O2C_WPP_DYNAMIC"):
            os.environ.pop(k, None)
        os.environ.update(O2C_WPP_ROLLERS="0", O2C_WPP_DYNAMIC="0", O2C_WPP_WIDE="0")
        s.solve(1.0)
        ref = s.download()
        t_end = time.time() + seconds / 3
        while time.time() < t_end:
            for k in ("O2C_WPP_ROLLERS", "O2C_WPP_RESIDENT", "O2C_WPP_WIDE", "O2C_WPP_DYNAMIC"):
                os.environ.pop(k, None)
            if rng.random() < 0.5:
                os.environ["O2C_WPP_ROLLERS"] = str(rng.integers(1, 4))
            if rng.random() < 0.5:
                os.environ["O2C_WPP_RESIDENT"] = str(rng.integers(1, 12))
            begin = int(rng.integers(0, batch - 1))
            count = int(rng.integers(1, batch - begin + 1))
            for _ in range(20):
                s.solve(1.0, problem_begin=begin, problem_count=count)
            launches += 20
            got = s.download(problem_begin=begin, problem_count=count)
            for name in ("K", "Sm", "x", "u", "status"):
                a, b = getattr(got, name), getattr(ref, name)
                b = b[:, begin:begin + count] if name in ("x", "u") else b[begin:begin + count]
                assert np.array_equal(a, b), (name, N, begin, count, dict(os.environ))
    print(f"N={N} batch={batch} nc={nc}: ok", flush=True)
print(f"{launches} launches, every result bit-identical to the fused schedule")
