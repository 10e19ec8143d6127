#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libgdbgpu.so the counts of the Blackwell data-movement instructions (UBLKCP = bulk
copy engine / TMA, LDGSTS = asynchronous global -> shared copies, SYNCS = mbarrier operations) beside the plain vector loads and
stores, and the instructions themselves for the dominant kernel.   usage: python tools/sass_summary.py > profiles/r02_sass_k5.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "genomicsdb_b200", "lib", "libgdbgpu.so")
PAT = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|LDGSTS|SYNCS|LDGDEPBAR|STG\.E\.128|LDG\.E\.128|LDS\.128|STS\.128|REDUX|BAR\.SYNC|SHFL|MATCH|VOTE)\b[\.\w]*")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    print("# cuobjdump -sass genomicsdb_b200/lib/libgdbgpu.so (sm_100a), %d kernels; instruction counts per kernel" % (len(funcs) - 1))
    total = collections.Counter()
    rows = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        c = collections.Counter(m.group(1) for m in PAT.finditer(f))
        total.update(c)
        rows.append((demangled, c, f))
    print("# whole library:", dict(total))
    print()
    for name, c, _ in sorted(rows, key=lambda r: r[0]):
        if c:
            print("%-70s %s" % (name[:70], " ".join("%s=%d" % kv for kv in sorted(c.items()))))
    print()
    for name, c, f in rows:
        if "k_remap_mix<2, true>" in name or "k_remap_mix<(int)2, (bool)1>" in name:
            print("# ---- %s: its bulk-copy / async-copy / mbarrier instructions and the 16-byte store stream" % name)
            for line in f.split("\n"):
                if re.search(r"UBLKCP|LDGSTS|SYNCS|LDGDEPBAR|DEPBAR|STG\.E\.128", line):
                    print(re.sub(r"\s+", " ", line.strip()))
            break


if __name__ == "__main__":
    main()
