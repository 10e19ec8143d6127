#!/usr/bin/env python3
"""Per-source-line hot spots of the first kernel in an .ncu-rep (needs -lineinfo at compile time).
usage: ncu_source_hot.py report.ncu-rep [top_n]  -> source lines sorted by warp instructions executed and by stall samples."""
import csv
import io
import subprocess
import sys


def main(rep, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    data, fname, hdr, kernels = [], "?", None, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or not r[0].isdigit():
            continue
        try:
            data.append((fname, int(r[0]), r[1], int(float(r[ci] or 0)), int(float(r[si] or 0))))
        except (ValueError, IndexError):
            pass
    # several launches of the same kernel repeat the table: merge by (file, line)
    agg = {}
    for f, ln, src, ni, ns in data:
        k = (f, ln)
        a = agg.setdefault(k, [src, 0, 0])
        a[1] += ni
        a[2] += ns
    tot_i = sum(a[1] for a in agg.values()) or 1
    tot_s = sum(a[2] for a in agg.values()) or 1
    print("== %d warp-instructions, %d samples (all captured launches)" % (tot_i, tot_s))
    for title, key in (("by instructions", 1), ("by stall samples", 2)):
        print("-- " + title)
        for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
            print("%-18s %5d %6.2f%% inst %6.2f%% smp | %s" % (f[:18], ln, 100.0 * a[1] / tot_i, 100.0 * a[2] / tot_s, a[0].strip()[:140]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
