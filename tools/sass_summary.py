#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): the instructions that show what a kernel is made
of -- packed FP32 (FFMA2 / FMUL2 / FADD2), MUFU, cp.async (LDGSTS), TMA bulk tensor loads (UTMALDG) and mbarrier waits
(SYNCS), global reductions (RED / REDG / ATOMG), shuffles, grid-barrier acquire loads -- plus registers per thread.

    python tools/sass_summary.py [path/to/libslicednb_b200.so] > profiles/rN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL", "MUFU", "LDGSTS", "UTMALDG", "UTMAREDG", "SYNCS",
         "RED", "REDG", "ATOMG", "ATOMS", "ATOM", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "F2I", "I2F", "F2F"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openmm-lab_b200", "csrc", "libslicednb_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    demangled = {}
    names = list(regs)
    if names:
        out = subprocess.run(["cu++filt"]+names, capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(names, out))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            if op in WATCH:
                counts[cur][op] += 1
    print("# SASS of %s (sm_100a): instructions per kernel, registers per thread, watched mnemonics" % os.path.relpath(lib, ROOT))
    for fn, c in counts.items():
        if not c["total"]:
            continue
        pretty = demangled.get(fn, fn)
        pretty = re.sub(r"\((DirectArgs|SortArgs|PmeArgs|ListArgs|BondedArgs|FinishArgs|BeginArgs|EwaldArgs|ParamArgs).*", r"(\1)", pretty)
        watched = "  ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
        print("%-72s %6d instr  %3s regs   %s" % (pretty[:72], c["total"], regs.get(fn, "?"), watched))


if __name__ == "__main__":
    main()
