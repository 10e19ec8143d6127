#!/usr/bin/env python3
"""Fragment read: time to open a GenomicsDB array from its on-disk fragment -- files -> HBM columns
+ indexes -- for a GZIP-compressed and an uncompressed copy of the same config-#4 column slice, next to the CPU cost of
inflating the same tiles with zlib on one host core (what ReadState::decompress_tile_gzip does per tile)."""
import argparse
import glob
import json
import os
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--sites", type=int, default=156)
    args = ap.parse_args()
    from genomicsdb_b200 import capi
    from tools import gdbsyn, tiledb_writer
    d = gdbsyn.generate(**dict(gdbsyn.CONFIGS["config4"], n_samples=args.samples), site_lo=0, site_hi=args.sites, with_fixed=True)
    tmp = tempfile.mkdtemp(prefix="gdbopen")
    res = dict(cells=int(len(d["end"])), samples=args.samples, candidate_sites=args.sites)
    ctx = capi.Context(0)
    for comp in (False, True):
        ws = os.path.join(tmp, "ws%d" % comp)
        t0 = time.time()
        winfo = tiledb_writer.write_array(ws, "syn", d, capacity=1000, compress=comp)
        key = "gzip" if comp else "plain"
        res[key + "_write_s"] = time.time() - t0
        frag = glob.glob(os.path.join(ws, "syn", "__*"))[0]
        res[key + "_file_bytes"] = sum(os.path.getsize(f) for f in glob.glob(os.path.join(frag, "*.tdb")))
        times = []
        ctx.set_profiling(True)
        ctx.reset_kernel_stats()
        for _ in range(3):
            t0 = time.perf_counter()
            arr = capi.Array.open(ctx, ws, "syn", [("AD", "R", "int"), ("PL", "G", "int")])
            ctx.synchronize()
            times.append(time.perf_counter() - t0)
            info = arr.info()
            arr.close()
        res[key + "_open_s"] = min(times)
        if comp:
            ms, nl = ctx.kernel_stats("k_inflate")
            res["k_inflate_ms"] = ms / max(1, nl)
        ctx.set_profiling(False)
        res[key + "_staged_bytes"] = int(info.staged_bytes)
        if comp:   # the CPU comparator: the same tiles through zlib on one core (hot-path attributes only)
            tabs = winfo["tile_tables"]
            names = [a[0] for a in tiledb_writer.ATTRS] + ["__coords"]
            t0 = time.perf_counter()
            total = 0
            for ai, name in enumerate(names):
                if name not in ("__coords", "END", "REF", "ALT", "GT", "AD", "PL"):
                    continue
                for fname, offs in ((name + ".tdb", tabs["tile_offsets"][ai]),) + (((name + "_var.tdb", tabs["var_offsets"][ai]),) if ai < len(tabs["var_offsets"]) and tabs["var_offsets"][ai] else ()):
                    raw = open(os.path.join(frag, fname), "rb").read()
                    ends = list(offs[1:]) + [len(raw)]
                    for lo, hi in zip(offs, ends):
                        if hi > lo:
                            total += len(zlib.decompress(raw[lo:hi]))
            res["cpu_zlib_1core_s"] = time.perf_counter() - t0
            res["cpu_zlib_bytes"] = total
    print(json.dumps(res))


if __name__ == "__main__":
    main()
