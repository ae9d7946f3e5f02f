"""times the parts of the end-to-end path separately: o2c_upload (H2D + pack), o2c_solve, o2c_download (unpack + D2H), o2c_solve_host"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ocs2_b200 as o2
from ocs2_b200 import lib as _l
import bench
eb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, m, nc, alg, eps, _ = bench.WORKLOADS["legged"]
st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=bench.DT)
solver = o2.BatchedLqSolver(st, n, m, bench.N_STAGES, eb, nc_max=nc)
captured = {}
orig = solver._lib
class Wrap:
    def __getattr__(self, k):
        f = getattr(orig, k)
        if k == "o2c_solve_host":
            def g(h, lv, sv, a, cnt, c):
                if not captured:  # time the parts once, while the caller's host buffers are alive
                    captured["done"] = True
                    def t(fn, reps=3):
                        fn(); solver.sync(); t0 = time.perf_counter()
                        for _ in range(reps): fn()
                        solver.sync(); return (time.perf_counter() - t0) / reps * 1e3
                    captured["up"] = t(lambda: _l.check(orig.o2c_upload(h, lv, 0, cnt)))
                    captured["so"] = t(lambda: _l.check(orig.o2c_solve(h, 1.0, 0, cnt)))
                    captured["dn"] = t(lambda: _l.check(orig.o2c_download(h, sv, 0, cnt, 1)))
                return f(h, lv, sv, a, cnt, c)
            return g
        return f
solver._lib = Wrap()
class A: steps = 3
r = bench.run_e2e(o2, np, torch, solver, st, n, m, nc, alg, eb, A, None, 1, lambda: torch.cuda.synchronize())
print("solve_host:", f"{r['ms_per_step']:.1f} ms", f"{r['value']:.0f} solves/s")
up, so, dn = captured["up"], captured["so"], captured["dn"]
print(f"upload {up:.1f} ms ({r['h2d_bytes_per_step']/up/1e6:.1f} GB/s)  solve {so:.2f} ms  download {dn:.1f} ms ({r['d2h_bytes_per_step']/dn/1e6:.1f} GB/s)")
