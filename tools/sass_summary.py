"""Static evidence from the built library (no GPU needed): registers / stack / static shared memory per kernel
(cuobjdump -res-usage) and, per kernel, the SASS mnemonics that show which hardware paths the code uses --
UBLKCP (TMA bulk copies), SYNCS (mbarriers), LDGSTS (cp.async), ATOMS / REDS (shared-memory atomics), RED / ATOMG
(global reductions), DFMA (FP64 FMA), UCGABAR (cluster barriers), SHFL.  Writes a CSV to stdout:

    python tools/sass_summary.py > profiles/r5c_sass_static.csv
"""
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "getdist_b200", "lib", "libgdk.so")
MNEMONICS = ["UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "REDS", "RED", "ATOMG", "DFMA", "DMUL", "DADD", "UCGABAR", "SHFL", "LDG", "STG", "LDS", "STS", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(MNEMONICS, 0)
            counts[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["instructions"] += 1
        head = op.split(".")[0]
        if head in counts[cur]:
            counts[cur][head] += 1
    names = demangle(sorted(usage))
    cols = ["kernel", "registers", "stack_bytes", "static_shared_bytes", "local_bytes", "instructions"] + MNEMONICS
    print(",".join(cols))
    for f in sorted(usage, key=lambda k: names[k]):
        u, c = usage[f], counts.get(f, {})
        short = re.sub(r"\(.*", "", names[f]).replace("void ", "").replace(", ", ";")
        row = [short, u.get("REG", 0), u.get("STACK", 0), u.get("SHARED", 0), u.get("LOCAL", 0), c.get("instructions", 0)] + [c.get(k, 0) for k in MNEMONICS]
        print(",".join(str(x) for x in row))


if __name__ == "__main__":
    sys.exit(main())
