"""DRAM bytes per solve of every bench workload, counted by ncu (dram__bytes_read.sum + dram__bytes_write.sum over all kernels of ONE
o2c_solve call at the BASELINE batch size), written to profiles/dram_traffic.json — the source of `roofline.traffic` in bench.py.

  python tools/dram_traffic.py              on the GPU box: runs ncu per workload, writes profiles/r02_dram_<workload>.csv and the json
  python tools/dram_traffic.py child NAME   the profiled process: two solves (the second one is the one counted)
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

NAMES = ("legged", "ballbot", "quadrotor_slq", "manipulator", "legged_constrained", "legged_slq")


def child(name):
    import ocs2_b200 as o2
    n, m, nc, alg, eps, batch = bench.WORKLOADS[name]
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=bench.DT)
    with o2.BatchedLqSolver(st, n, m, bench.N_STAGES, batch, nc_max=nc) as s:
        s.generate_synthetic(1, 0, bench.DT)
        for _ in range(2):
            s.solve(1.0)
            s.sync()
        print("KERNEL", s.kernel_variant, flush=True)


def main():
    out = {}
    for name in NAMES:
        log = os.path.join(ROOT, "profiles", f"r02_dram_{name}.csv")
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", log,
               sys.executable, __file__, "child", name]
        if os.environ.get("O2C_REUSE_CSV") and os.path.exists(log):  # re-summarise committed launch lists without a GPU
            kernel = {"legged": "ilqr_wpp_kernel", "legged_constrained": "ilqr_wpp_kernel", "legged_slq": "slq_wpp_kernel",
                      "quadrotor_slq": "slq_rpl_kernel"}.get(name, "ilqr_rpl_kernel")
        else:
            done = subprocess.run(cmd, capture_output=True, text=True)
            kernel = [l.split()[1] for l in done.stdout.splitlines() if l.startswith("KERNEL")][0]
        rows = [r for r in csv.reader(open(log)) if len(r) > 8]
        hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        h = rows[hdr]
        ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
        launches = {}
        for r in rows[hdr + 1:]:
            launches.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
        ids = sorted(launches)
        names = [launches[i]["name"] for i in ids]
        solve_ids = [i for i in ids if not any(t in launches[i]["name"] for t in ("generate_kernel", "fill_int_kernel"))]
        per = len(solve_ids) // 2  # two identical solves: count the second
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        total, ms, kernels = 0.0, 0.0, []
        for i in solve_ids[per:]:
            L = launches[i]
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += L[k][0] * scale[L[k][1]]
            t, u = L["gpu__time_duration.sum"]
            ms += t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
            kernels.append(L["name"].split("(")[0].split("::")[-1])
        batch = bench.WORKLOADS[name][5]
        out[f"{name}:{kernel}"] = {"dram_bytes_per_solve": total / batch, "kernels_of_one_solve": kernels, "ncu_ms_of_one_solve": round(ms, 3),
                                   "source": f"profiles/r02_dram_{name}.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of "
                                             f"tools/dram_traffic.py child {name}: all kernels of one o2c_solve at batch {batch})",
                                   "file": f"profiles/r02_dram_{name}.csv"}
        print(name, kernel, f"{total / batch / 1e6:.3f} MB/solve", kernels, f"{ms:.2f} ms", flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    child(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[1] == "child" else main()
