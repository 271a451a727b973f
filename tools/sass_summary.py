#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the built library (cuobjdump -sass): instruction count, global load / store widths,
atomics / reductions, shared-memory and local-memory traffic, warp shuffles / votes. Written to profiles/ as evidence of
what the kernels compile to (nothing here is a contraction, so no UTCMMA / TMEM instructions are expected).

usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "microphaser_b200", "_lib", "libmicrophaser_gpu.so")
GROUPS = [("LDG.128", r"^LDG\S*\.128"), ("LDG.64", r"^LDG\S*\.64"), ("LDG.32", r"^LDG(?!\S*\.(U8|S8|U16|S16|64|128))"), ("LDG.8/16", r"^LDG\S*\.(U8|S8|U16|S16)"),
          ("STG.128", r"^STG\S*\.128"), ("STG.64", r"^STG\S*\.64"), ("STG.32", r"^STG(?!\S*\.(U8|S8|U16|S16|64|128))"), ("STG.8/16", r"^STG\S*\.(U8|S8|U16|S16)"),
          ("ATOMG/RED", r"^(ATOMG|RED|ATOM)\b"), ("LDS/STS", r"^(LDS|STS)"), ("ATOMS", r"^ATOMS"), ("LDL/STL", r"^(LDL|STL)"),
          ("SHFL", r"^SHFL"), ("VOTE/MATCH/REDUX", r"^(VOTE|MATCH|REDUX)"), ("BAR", r"^BAR"), ("UTCMMA/TMEM", r"^(UTC|TCGEN|LDTM|STTM)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, stats = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            k = re.search(r"(k_\w+(<[^>]*>)?)", name)
            kern = k.group(1) if k else name
            stats[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            stats[kern]["total"] += 1
            for g, pat in GROUPS:
                if re.match(pat, op):
                    stats[kern][g] += 1
    cols = ["total"] + [g for g, _ in GROUPS]
    print("SASS of %s (sm_100a), static instruction counts per kernel" % os.path.relpath(LIB, ROOT))
    print("%-34s" % "kernel" + "".join("%10s" % c[:10] for c in cols))
    for k, c in stats.items():
        print("%-34s" % k[:34] + "".join("%10d" % c[x] for x in cols))


if __name__ == "__main__":
    main()
