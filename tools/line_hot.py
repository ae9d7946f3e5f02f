"""Per-CUDA-line sample attribution from `ncu --page source --csv --print-source sass,cuda`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
out = []; tot = 0; fname = ""; hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    s = int(r[si]) if r[si].isdigit() else 0; e = int(r[ei]) if r[ei].isdigit() else 0
    tot += s; out.append((fname, int(r[0]), s, e, r[1]))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
print("total samples", tot)
for fn, ln, s, e, src in out:
    if 100.0 * s / tot >= thr:
        print(f"{fn[:18]:18s}{ln:5d} {100*s/tot:6.2f}% {e:11d}  {src.strip()[:110]}")
