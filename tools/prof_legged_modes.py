"""legged ILQR at batch 16384 under LINE_SEARCH + DIAGONAL_SHIFT, LINE_SEARCH + GERSHGORIN and LEVENBERG_MARQUARDT (the three MODE variants of
the DMMA kernel). Usage: prof_legged_modes.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2
for name, kw in (("ls_diag", {}), ("gershgorin", dict(hessianCorrectionStrategy=o2.HC_GERSHGORIN_MODIFICATION)), ("lm", dict(strategy=o2.STRATEGY_LEVENBERG_MARQUARDT, riccatiMultiple=0.1))):
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01, **kw)
    with o2.BatchedLqSolver(st, 24, 24, 100, 16384) as s:
        s.generate_synthetic(1, 0, 0.01); s.sync()
        for _ in range(2): s.solve(1.0)
        s.sync(); t0 = time.perf_counter()
        for _ in range(5): s.solve(1.0)
        s.sync(); ms = (time.perf_counter() - t0) / 5 * 1e3
        print(json.dumps({"workload": "legged " + name, "kernel": s.kernel_variant, "batch": 16384, "ms": round(ms, 3), "per_s": round(16384 / ms * 1e3)}), flush=True)
