"""Summarises an ncu report (raw page) into the handful of numbers DESIGN.md / profiles/ quote. Usage: ncu_summary.py file.ncu-rep [kernel-substr]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if len(sys.argv) > 2 and sys.argv[2] not in d.get("Kernel Name", ""):
        continue
    for k in keys:
        for h in hdr:
            if h == k:
                print(f"{h:75s} {d[h]:>18s} {units[hdr.index(h)]}")
    print("-- warp stall reasons (warps per issue-active cycle) --")
    for h in sorted(stall, key=lambda h: -float(d[h] or 0)):
        v = float(d[h] or 0)
        if v >= 0.05:
            print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
