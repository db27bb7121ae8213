"""Static SASS opcode census of every kernel in the shipped library (cuobjdump -sass; needs no GPU) -> profiles/r02_sass_opcode_census.csv.
Shows which hardware paths the kernels use: UBLKCP (bulk TMA copies), SYNCS (mbarriers), UCGABAR (cluster barriers), MAPA (DSMEM),
REDUX (warp reductions), IMAD.WIDE / IADD3.X (the 64-bit field arithmetic); no HMMA / UTC* (no tensor-core path).
usage: python tools/sass_census.py [out.csv]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ceno_b200", "lib", "libceno_b200.so")
COLS = ["UBLKCP.S", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "LDG.E", "STG.E", "LDS.128", "IMAD.WIDE", "IADD3.X", "REDUX.SUM", "UCGABAR_ARV", "UCGABAR_WAIT",
        "ST.E", "STS.128", "MAPA", "CCTL.IVALL", "HMMA", "UTCMMA"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcode_census.csv")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_\.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    counts[cur][c] += 1
    with open(out, "w") as f:
        f.write("static SASS opcode counts per kernel (cuobjdump -sass of the shipped libceno_b200.so); columns: total," + ",".join(COLS) + "\n")
        for k in sorted(counts):
            f.write(",".join([k, str(counts[k]["total"])] + [str(counts[k][c]) for c in COLS]) + "\n")
    print(f"{out}: {len(counts)} kernels")


if __name__ == "__main__":
    main()
