"""Aggregates an `ncu --page source --csv --print-source sass` dump: samples and executed instructions per opcode, and the hottest SASS lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0])
lines = []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_stall = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "MUFU")) and "." in op else "")
    s = int(r[ix["# Samples"]] or 0)
    e = int(r[ix["Instructions Executed"]] or 0)
    ops[op][0] += s
    ops[op][1] += e
    lines.append((s, e, src, {h: int(r[ix[h]] or 0) for h in stall_cols}))
    for h in stall_cols:
        tot_stall[h] += int(r[ix[h]] or 0)
ts = sum(v[0] for v in ops.values()); te = sum(v[1] for v in ops.values())
print(f"total samples {ts}, executed warp-instr {te}, SASS lines {len(lines)}")
print("stall totals:", {k: v for k, v in tot_stall.most_common(9)})
print(f"{'opcode':14s} {'samples%':>9s} {'exec%':>8s} {'exec':>12s}")
for op, (s, e) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:14s} {100*s/ts:9.2f} {100*e/te:8.2f} {e:12d}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("hottest lines:")
for i, (s, e, src, st) in sorted(enumerate(lines), key=lambda t: -t[1][0])[:n]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    print(f"  #{i:5d} {s:6d} {src[:70]:70s} {top}")
