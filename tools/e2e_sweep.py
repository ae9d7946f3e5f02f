"""e2e (host buffers in/out through o2c_solve_host) throughput vs pipeline chunk size. Usage: e2e_sweep.py [problems] [chunk ...]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ocs2_b200 as o2
from ocs2_b200 import lib as _l
import bench

eb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunks = [int(a) for a in sys.argv[2:]] or [0, 64, 128, 256]
n, m, nc, alg, eps, _ = bench.WORKLOADS["legged"]
st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=bench.DT)
solver = o2.BatchedLqSolver(st, n, m, bench.N_STAGES, eb, nc_max=nc)
class A: steps = 3
for ch in chunks:
    orig = solver._lib.o2c_solve_host
    class Wrap:
        def __init__(self, lib): self.lib = lib
        def __getattr__(self, k):
            f = getattr(self.lib, k)
            if k == "o2c_solve_host":
                return lambda h, lv, sv, a, cnt, c: f(h, lv, sv, a, cnt, ch)
            return f
    solver._lib = Wrap(_l.load_library())
    r = bench.run_e2e(o2, np, torch, solver, st, n, m, nc, alg, eb, A, None, 1, lambda: torch.cuda.synchronize())
    solver._lib = _l.load_library()
    gb = (r["h2d_bytes_per_step"] + 0.0) / 1e9
    print(f"problems {eb} chunk {ch}: {r['value']:.0f} solves/s, {r['ms_per_step']:.1f} ms/step, H2D {gb/ (r['ms_per_step']*1e-3):.1f} GB/s", flush=True)
