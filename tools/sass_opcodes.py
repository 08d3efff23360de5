#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (profiles/rNN_sass_opcodes.txt).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "surf_b200", "csrc", "libsurf_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDG.E.128", "LDG.E.64", "LDG.E.CON", "LDG.E",
        "STG.E.128", "STG.E", "LDS", "STS", "MUFU", "FFMA", "FFMA2", "FMUL2", "FADD2", "FHFMA", "HFMA2", "F2FP", "BAR.SYNC",
        "ATOM", "RED", "LDL", "STL", "SHFL", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts, order, cur, k = {}, [], None, -1
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            k += 1
            full = names[k] if k < len(names) and names[k] else m.group(1)
            mm = re.match(r"^(.*?>)\(", full) if "<" in full.split("(")[0] + "<"[:0] or re.match(r"^[^(]*<", full) else None
            cur = (mm.group(1) if mm else full.split("(")[0]).replace("(bool)", "").replace("(int)", "")
            while cur in counts:
                cur += "'"
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + ".") or (c.endswith(".CON") and op.startswith(c)):
                    counts[cur][c] += 1
    print("# per-kernel SASS opcode counts of surf_b200/csrc/libsurf_b200.so (cuobjdump -sass, sm_100a); tools/sass_opcodes.py")
    print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, "
          "mbarrier -> SYNCS, packed fp32x2 -> FFMA2/FMUL2/FADD2, fp16 x fp16 + fp32 -> FHFMA")
    print("%-44s" % "kernel" + "".join("%10s" % c for c in ["_total"] + COLS))
    for name in order:
        print("%-44s" % name[:44] + "".join("%10d" % counts[name][c] for c in ["_total"] + COLS))


if __name__ == "__main__":
    sys.exit(main())
