#!/usr/bin/env python
"""Per-kernel SASS opcode table of the shipped library (runs without a GPU: cuobjdump reads the sm_100a cubin).

    python tools/sass_table.py > profiles/sass_r02.md

Counts the mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md): DMMA (FP64 tensor core),
UBLKCP (cp.async.bulk = TMA 1-D), SYNCS (mbarrier), LDGSTS (cp.async), DFMA (FP64 FMA pipe), LDS/STS, REDUX/SHFL,
USETMAXREG (setmaxnreg), UCGABAR/ LDS via DSMEM (clusters)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nbed_b200", "libnbed_b200.so")
OPS = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "USETMAXREG",
       "UCGABAR", "MAPA", "ATOM", "RED", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    tables = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            tables[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            tables[cur][op] += 1
            tables[cur]["_total"] += 1
    names = list(tables)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for a, b in zip(names, dm):
        b = re.sub(r"\((int|bool|unsigned int)\)", "", b).replace("nbd::", "")
        demangle[a] = re.sub(r"^void ", "", b[: b.rfind("(")] if "(" in b else b)
    print("# SASS opcode counts per kernel (`cuobjdump -sass nbed_b200/libnbed_b200.so`, sm_100a), round 2\n")
    print("Static instruction counts (not executed counts).  DMMA = FP64 tensor-core `DMMA.8x8x4` (what `mma.sync.m8n8k4.f64`\n"
          "lowers to on sm_100a; tcgen05 has no f64 kind), UBLKCP = `cp.async.bulk` (TMA 1-D), SYNCS = mbarrier ops,\n"
          "LDGSTS = `cp.async`, USETMAXREG = `setmaxnreg`.\n")
    print("| kernel | total | " + " | ".join(OPS) + " |")
    print("|---|---|" + "---|" * len(OPS))
    tot = collections.Counter()
    for k in sorted(names, key=lambda x: demangle[x]):
        t = tables[k]
        row = [str(sum(v for o, v in t.items() if o == op or (op in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR") and o.startswith(op) and o != "LDGSTS" and not (op == "LDG" and o.startswith("LDGSTS"))))) for op in OPS]
        for op, v in zip(OPS, row):
            tot[op] += int(v)
        print(f"| `{demangle[k]}` | {t['_total']} | " + " | ".join(row) + " |")
    print("| **all kernels** | | " + " | ".join(str(tot[o]) for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
