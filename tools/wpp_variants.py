"""legged ILQR kernel variants, selected per launch through environment knobs (O2C_WPP_RESIDENT, O2C_WPP_QPD, O2C_WPP_DYNAMIC).
CUDA events on the library's compute stream, back-to-back steps after warm-up; one JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ocs2_b200 as o2

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)


def run(s, stream, batch, env):
    for k in ("O2C_WPP_RESIDENT", "O2C_WPP_QPD", "O2C_WPP_DYNAMIC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        s.solve(1.0, problem_count=batch)
    s.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for k in range(steps):
        s.solve(1.0, problem_count=batch)
        ev[k + 1].record(stream)
    s.sync()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    status = s.download(problem_begin=0, problem_count=min(batch, 64), n_alpha=0).status
    assert (status == 0).all()
    tot = ev[0].elapsed_time(ev[steps]) / steps
    print(json.dumps({**env, "batch": batch, "ms_mean": round(tot, 4), "ms_min": round(min(ms), 4), "solves_per_s_mean": round(batch / tot * 1e3)}), flush=True)
    time.sleep(0.7)


with o2.BatchedLqSolver(st, 24, 24, 100, 16384) as s:
    s.generate_synthetic(1, 0, 0.01)
    s.sync()
    stream = torch.cuda.ExternalStream(s.compute_stream)
    for rep in range(2):
        for qpd in ("0", "1"):
            for batch in (16384, 2048, 1776):
                run(s, stream, batch, {"O2C_WPP_QPD": qpd})
    for batch in (2048, 4096, 8192):
        for w in ("7", "10", "12"):
            run(s, stream, batch, {"O2C_WPP_RESIDENT": w})
    for w in ("1", "4", "8"):
        for qpd in ("0", "1"):
            run(s, stream, 148 * int(w), {"O2C_WPP_RESIDENT": w, "O2C_WPP_QPD": qpd})
