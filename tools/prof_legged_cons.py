"""legged ILQR with state-input equality constraints (uniform nc per node): solves/s of the DMMA kernel. Usage: prof_legged_cons.py nc batch"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2
nc, batch = int(sys.argv[1]), int(sys.argv[2])
st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)
with o2.BatchedLqSolver(st, 24, 24, 100, batch, nc_max=nc) as s:
    s.generate_synthetic(1, 0, 0.01)
    s.sync()
    for mode, fn in (("solve", lambda: s.solve(1.0)), ("backward", s.solveSequentialRiccatiEquations)):
        for _ in range(2):
            fn()
        s.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        s.sync()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        ok = bool((s.download(problem_count=64, n_alpha=0).status == 0).all())
        print(json.dumps({"workload": f"legged nc={nc}", "kernel": s.kernel_variant, "mode": mode, "batch": batch, "ms": round(ms, 3), "per_s": round(batch / ms * 1e3), "status_ok": ok, **{k: v for k, v in os.environ.items() if k.startswith("O2C_WPP")}}), flush=True)
