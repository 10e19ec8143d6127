#!/usr/bin/env python3
"""The columnar-iterator row at BASELINE config #4's width: cells/s of the device's cell list + the two
operators that run as kernels (ColumnHistogramOperator, AlleleCountOperator), shard resident in HBM, next to the
unmodified reference's `gt_mpi_gather --produce-histogram` / `--print-AC` (oracle/_ref, one process = one core) and our
`gt_gpu_gather` twins on the same on-disk fragment (tool wall clocks, outputs compared byte for byte)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--sites", type=int, default=300)
    ap.add_argument("--skip-reference", action="store_true")
    args = ap.parse_args()
    import torch
    from genomicsdb_b200 import capi
    from tools import gdbsyn, tiledb_writer
    d = gdbsyn.generate(**dict(gdbsyn.CONFIGS["config4"], n_samples=args.samples), site_lo=0, site_hi=args.sites)
    lo, hi = d["col_range"]
    n_cells = int(len(d["end"]))
    ctx = capi.Context(0)
    arr = gdbsyn.to_gpu_array(ctx, d)
    res = dict(samples=args.samples, candidate_sites=args.sites, array_cells=n_cells)

    def timed(fn, reps=5):
        fn()
        ctx.synchronize()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            ctx.synchronize()
            t.append(time.perf_counter() - t0)
        return min(t), out

    # whole interval; a mid-array interval exercises the left (END copy) phase over all samples
    for name, a, b in (("whole", lo, hi), ("mid", lo + (hi - lo) // 2 + 137, hi)):
        def list_only():
            sc = capi.Scan(arr, a, b, flags=capi.SCAN_CELLS_ONLY)
            c = sc.cells_histogram(lo, hi, 1000)     # builds the list on the device; only the bins cross PCIe
            sc.close()
            return int(c.sum())

        def ac():
            sc = capi.Scan(arr, a, b, flags=capi.SCAN_CELLS_ONLY)
            r = sc.cells_allele_counts()
            sc.close()
            return len(r), sum(x[3] for x in r)
        t_h, listed = timed(list_only)
        t_a, (n_rec, n_alt) = timed(ac)
        res[name] = dict(interval=[int(a), int(b)], listed_cells=listed, list_plus_histogram_s=t_h, listed_cells_per_s=listed / t_h,
                         allele_count_s=t_a, allele_count_cells_per_s=listed / t_a, allele_count_records=n_rec, alt_alleles_counted=int(n_alt))
    arr.close()
    ctx.close()
    if not args.skip_reference:
        tmp = tempfile.mkdtemp(prefix="gdbcells")
        ws = os.path.join(tmp, "ws")
        info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
        q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT"))
        ref_bin = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
        ours_bin = os.path.join(REPO, "genomicsdb_b200", "bin", "gt_gpu_gather")
        for mode, flag in (("histogram", "--produce-histogram"), ("allele_counts", "--print-AC"),
                           ("interesting_positions", "--produce-interesting-positions")):   # the last one: scan_and_operate + a per-site summary operator
            t0 = time.time()
            p = subprocess.run(["stdbuf", "-o0", ref_bin, "-j", q, flag], capture_output=True)
            t_ref = time.time() - t0
            if p.returncode != 0:
                res[mode] = dict(reference_error=p.stderr.decode()[-400:])
                continue
            out = os.path.join(tmp, mode + ".txt")
            subprocess.run([ours_bin, "-j", q, "-o", out, flag], capture_output=True)    # first CUDA process on the box pays the driver start
            t0 = time.time()
            p2 = subprocess.run([ours_bin, "-j", q, "-o", out, flag], capture_output=True, text=True)
            t_ours = time.time() - t0
            res[mode] = dict(reference_wall_s=t_ref, ours_wall_s=t_ours, byte_identical=p2.returncode == 0 and open(out, "rb").read() == p.stdout,
                             output_bytes=len(p.stdout), ours_error=p2.stderr[-300:] if p2.returncode else None)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
