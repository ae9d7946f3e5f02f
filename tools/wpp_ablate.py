"""Profiling builds of the legged ILQR kernel (ocs2_b200/csrc/riccati_wpp.cu) with extra -D flags, and their timing on the GPU.

  python tools/wpp_ablate.py build NAME=FLAGS [NAME=FLAGS ...]   (here, no GPU)  e.g.  base=  nofactor=-DO2C_WPP_ABLATE=1
      compiles riccati_wpp.cu with -DO2C_WPP_ONLY_BASE + FLAGS and links ocs2_b200/build/exp/libo2c_NAME.so from it and the product objects
  python tools/wpp_ablate.py run [batch ...]                     (on the GPU box) times every variant found, one process each
"""
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXP = os.path.join(ROOT, "ocs2_b200", "build", "exp")


def build(specs):
    from ocs2_b200 import build as b
    b.build()
    os.makedirs(EXP, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        obj = os.path.join(EXP, f"wpp_{name}.o")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, "-DO2C_WPP_ONLY_BASE", *flags.split(), "-Xptxas", "-v", "-c", os.path.join(b.CSRC, "riccati_wpp.cu"), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise SystemExit(out)
        print(name, [l.strip() for l in out.splitlines() if "registers" in l or "spill" in l])
        objs = [os.path.join(ROOT, "ocs2_b200", "build", s.replace(".cu", ".o")) for s in b.SOURCES if s != "riccati_wpp.cu"] + [obj]
        subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(EXP, f"libo2c_{name}.so"), *objs])


def child(path, batches, steps=6):
    import numpy as np
    import ocs2_b200.lib as L
    L.library_path = lambda: path
    import ocs2_b200 as o2
    envs = [dict(kv.split("=") for kv in e.split(",") if kv) for e in os.environ.get("O2C_ABLATE_ENVS", "").split(";")]
    st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)
    with o2.BatchedLqSolver(st, 24, 24, 100, max(batches)) as s:
        s.generate_synthetic(1, 0, 0.01)
        s.sync()
        ref = None
        for env in envs:
            for k in ("O2C_WPP_RESIDENT", "O2C_WPP_ROLLERS", "O2C_WPP_DYNAMIC"):
                os.environ.pop(k, None)
            os.environ.update({"O2C_WPP_" + k: v for k, v in env.items()})
            for batch in batches:
                for mode in ("solve", "backward"):
                    fn = (lambda: s.solve(1.0, problem_count=batch)) if mode == "solve" else (lambda: s.solveSequentialRiccatiEquations(problem_count=batch))
                    for _ in range(2):
                        fn()
                    s.sync()
                    t0 = time.perf_counter()
                    for _ in range(steps):
                        fn()
                    s.sync()
                    ms = (time.perf_counter() - t0) / steps * 1e3
                    same = None
                    if mode == "solve":  # the schedule must not change a single bit of the result
                        sol = s.download(problem_begin=0, problem_count=min(batch, 256), n_alpha=1)
                        got = (sol.x.copy(), sol.u.copy(), sol.K.copy(), sol.status.copy())
                        key = batch
                        if ref is None:
                            ref = {}
                        if key not in ref:
                            ref[key] = got
                        same = all(np.array_equal(a, b) for a, b in zip(ref[key], got)) and bool((got[3] == 0).all())
                    print(json.dumps({"variant": os.path.basename(path)[7:-3], **env, "mode": mode, "batch": batch, "ms": round(ms, 4),
                                      "solves_per_s": round(batch / ms * 1e3), "bit_identical_to_first": same}), flush=True)
                    time.sleep(0.4)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "child":
        child(sys.argv[2], [int(b) for b in sys.argv[3:]])
    else:
        batches = sys.argv[2:] or ["16384", "1776"]
        for path in sorted(glob.glob(os.path.join(EXP, "libo2c_*.so"))):
            subprocess.run([sys.executable, __file__, "child", path, *batches])
