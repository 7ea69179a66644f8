"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md), from the built
library.  usage: python scripts/sass_table.py > profiles/r2_sass_tensor_ops.txt     (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "nerf_simple_b200", "libnerf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "FADD2", "F2FP", "MUFU", "RED", "ATOMG", "LDC"]
counts, order, cur = {}, [], None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
        if ".2CTA" in op:
            counts[cur]["2CTA"] += 1
demangled = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a): instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = tensor-map TMA load, UBLKCP = bulk TMA copy,")
print("# UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, 2CTA = instructions carrying the .2CTA modifier (cta_group::2)")
cols = ["_total", "UTCHMMA", "2CTA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FADD2", "F2FP", "MUFU", "RED", "LDC"]
print("# " + " ".join(f"{c.strip('_'):>8s}" for c in cols) + "  kernel")
only_tc = "--all" not in sys.argv
for name, dm in zip(order, demangled or order):
    c = counts[name]
    if c["_total"] == 0 or (only_tc and c["UTCHMMA"] + c["UBLKCP"] + c["UTMALDG"] == 0):
        continue
    short = dm[:dm.rfind(">(") + 1] if ">(" in dm else dm.split("(")[0]
    short = short.replace("(int)", "").replace("(bool)", "").replace("void ", "")
    print("  " + " ".join(f"{c[k]:8d}" for k in cols) + "  " + short[:110])
