"""Throughput of configurations that run through the generic kernels on the legged shape (what the next specialised kernels would buy).
Usage: prof_generic.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cases = [("ILQR LINE_SEARCH (specialised)", dict(algorithm=o2.ALG_ILQR), 0),
         ("ILQR LEVENBERG_MARQUARDT", dict(algorithm=o2.ALG_ILQR, strategy=o2.STRATEGY_LEVENBERG_MARQUARDT, riccatiMultiple=0.1), 0),
         ("ILQR GERSHGORIN", dict(algorithm=o2.ALG_ILQR, hessianCorrectionStrategy=o2.HC_GERSHGORIN_MODIFICATION), 0),
         ("ILQR nc=6 constraints", dict(algorithm=o2.ALG_ILQR), 6),
         ("SLQ-RK4", dict(algorithm=o2.ALG_SLQ), 0)]
for name, kw, nc in cases:
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01, **kw)
    with o2.BatchedLqSolver(st, 24, 24, 100, B, nc_max=nc) as s:
        s.generate_synthetic(1, 0, 0.01); s.sync()
        for r in range(3):
            t0 = time.perf_counter(); s.solve(1.0); s.sync(); dt = time.perf_counter() - t0
        ok = (s.download(problem_count=16, n_alpha=0).status == 0).all()
        print(f"legged {name:34s} {s.kernel_variant:28s} batch {B}: {dt*1e3:8.2f} ms -> {B/dt:9.0f} solves/s  ok={ok}", flush=True)
