"""Per-kernel counts of the SASS mnemonics that prove the instruction mix (cuobjdump -sass of the built objects):
DMMA (FP64 tensor pipe), UBLKCP (TMA bulk copy), UBLKPF (TMA L2 prefetch), SYNCS (mbarrier), HMMA/UTC*MMA (must be absent: FP64 path).
Usage: python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = ["riccati_wpp.o", "slq_wpp.o", "riccati_rpl.o", "riccati_generic.o", "rollout.o", "pack.o", "line_search.o"]
keys = ["DMMA", "DFMA", "UBLKCP", "UBLKPF", "SYNCS", "HMMA", "UTCMMA", "UTMALDG", "MUFU.RSQ64H", "LDS", "STS", "SHFL"]
print("kernel".ljust(90), " ".join(k.rjust(11) for k in keys), "   instr")
for o in objs:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ocs2_b200", "build", o)], capture_output=True, text=True).stdout
    for part in sass.split("Function : ")[1:]:
        name = subprocess.run(["c++filt", part.split("\n", 1)[0].strip()], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"o2c::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        lines = [l for l in part.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        print((o[:-2] + ": " + name)[:90].ljust(90), " ".join(str(sum(1 for l in lines if k in l)).rjust(11) for k in keys), str(len(lines)).rjust(8))
