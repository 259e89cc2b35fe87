#!/usr/bin/env python
"""SASS census of libopkb200.so (sm_100a cubins): per kernel, how many of the instructions that
characterise its design it holds -- bulk-async copies of the TMA engine (UBLKCP), mbarrier
operations (SYNCS), 128-bit shared / global accesses, Float64 FMAs, warp shuffles, and (to show
their absence) tensor-core opcodes.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "optimpacknextgen.jl_b200", "libopkb200.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDS.128", r"\bLDS(\.U)?\.128"), ("LDS", r"\bLDS\b"),
            ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("STG.128", r"\bSTG\.E(\.\w+)*\.128"), ("DFMA", r"\bDFMA"),
            ("DADD", r"\bDADD"), ("DMUL", r"\bDMUL"), ("FFMA", r"\bFFMA"), ("FADD", r"\bFADD"), ("FMUL", r"\bFMUL"),
            ("F2F", r"\bF2F"), ("SHFL", r"\bSHFL"), ("MUFU.RCP64H", r"\bMUFU\.RCP64H"),
            ("tensor (HMMA|UTC*MMA|QGMMA)", r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCQMMA|UTCIMMA|QGMMA|HGMMA)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        ins = line.split("*/", 1)[-1] if line.strip().startswith("/*") else line
        if not re.search(r"[A-Z]{3,}", ins):
            continue
        kernels[cur]["total"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, ins):
                kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    names = [n for n, _ in PATTERNS]
    print("# SASS census of optimpacknextgen.jl_b200/libopkb200.so (cuobjdump -sass, sm_100a), %d kernels" % len(kernels))
    print("# columns: total instructions | " + " | ".join(names))
    tot = collections.Counter()
    for (k, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*$", "", dn)
        short = re.sub(r"^void ", "", short)
        print("%-110s %6d | %s" % (short[:110], c["total"], " ".join("%5d" % c[n] for n in names)))
        tot.update(c)
    print("%-110s %6d | %s" % ("TOTAL", tot["total"], " ".join("%5d" % tot[n] for n in names)))
    print("# tensor-core opcodes: %d (none expected: every kernel is an HBM-bound stream, <= 1.5 flop/byte)"
          % tot[names[-1]])


if __name__ == "__main__":
    sys.exit(main())
