"""back-to-back legged solves with 20 ms clock sampling: isolates clock / power effects on the sustained number"""
import os, subprocess, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ocs2_b200 as o2
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
st = o2.Settings(hessianCorrectionMultiple=1e-5, timeStep=0.01)
s = o2.BatchedLqSolver(st, 24, 24, 100, 16384)
s.generate_synthetic(1, 0, 0.01); s.sync()
stream = torch.cuda.ExternalStream(s.compute_stream)
for _ in range(3): s.solve(1.0)
s.sync()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
rows = []
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in p.stdout], daemon=True).start()
time.sleep(0.3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
t0 = time.time()
ev[0].record(stream)
for k in range(steps):
    s.solve(1.0); ev[k + 1].record(stream)
s.sync(); t1 = time.time()
time.sleep(0.1); p.terminate()
ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
print("per-step ms:", " ".join(f"{m:.2f}" for m in ms))
print("clock samples in region:", [r for t, r in rows if t0 <= t <= t1][:40])
