"""Stall samples of an ncu source page (--print-source sass --csv) grouped into the stretches between synchronisation markers
(WARPSYNC / SYNCS = mbarrier / UBLKCP = TMA / branches): where along the instruction stream a warp spends its time.
Usage: sass_segments.py page.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = next(r for r in rows if r and r[0] == "Address")
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = [(r[1].strip(), int(r[si] or 0), [int(r[i] or 0) for i in stall_cols]) for r in rows if len(r) >= len(hdr) and r[0].startswith("0x")]
tot = sum(x[1] for x in ins)
print("instructions", len(ins), "samples", tot)
names = [hdr[i][6:] for i in stall_cols]
start, acc, dm, st = 0, 0, 0, [0] * len(stall_cols)
for i, (sass, s, sv) in enumerate(ins):
    acc += s
    st = [a + b for a, b in zip(st, sv)]
    dm += "DMMA" in sass
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    if op.startswith(("WARPSYNC", "SYNCS", "UBLKCP", "BRA", "BAR", "EXIT")):
        if 100.0 * acc / tot >= thr:
            top = sorted(zip(st, names), reverse=True)[:3]
            print(f"{start:5d}-{i:5d} dmma={dm:3d} {100.0 * acc / tot:6.2f}%  " + " ".join(f"{n}:{100.0 * v / tot:.1f}" for v, n in top) + f"   | {sass[:50]}")
        start, acc, dm, st = i + 1, 0, 0, [0] * len(stall_cols)
