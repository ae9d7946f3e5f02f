"""legged-shape throughput of the settings the headline line does not cover (LM, Gershgorin, constraints, SLQ): which kernel serves
them and how fast. Usage: prof_generic.py [batch] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocs2_b200 as o2

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cases = {
    "ls_diag": dict(st=o2.Settings(hessianCorrectionMultiple=1e-5), nc=0),
    "lm": dict(st=o2.Settings(strategy=o2.STRATEGY_LEVENBERG_MARQUARDT, riccatiMultiple=0.1, preComputeRiccatiTerms=False), nc=0),
    "gershgorin": dict(st=o2.Settings(hessianCorrectionStrategy=o2.HC_GERSHGORIN_MODIFICATION, hessianCorrectionMultiple=1e-5), nc=0),
    "nc6": dict(st=o2.Settings(hessianCorrectionMultiple=1e-5), nc=6),
    "slq": dict(st=o2.Settings(algorithm=o2.ALG_SLQ, hessianCorrectionMultiple=1e-5, timeStep=0.01), nc=0),
}
only = sys.argv[3].split(",") if len(sys.argv) > 3 else list(cases)
for name in only:
    cfg = cases[name]
    with o2.BatchedLqSolver(cfg["st"], 24, 24, 100, batch, nc_max=cfg["nc"]) as s:
        s.generate_synthetic(1, 0, 0.01)
        s.solve(1.0)
        s.sync()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            s.solve(1.0)
            s.sync()
            best = min(best, time.perf_counter() - t0)
        st_ = s.download(problem_begin=0, problem_count=min(batch, 64), n_alpha=0).status
        print(f"legged {name:11s} batch {batch}: {s.kernel_variant:28s} {best*1e3:9.2f} ms -> {batch/best:10.0f} solves/s  status ok {(st_ == 0).all()}", flush=True)
