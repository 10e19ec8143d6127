#!/usr/bin/env python3
"""Tool against tool at BASELINE config #4's width: the unmodified reference's
    gt_mpi_gather --produce-Broad-GVCF -O bu          (oracle/_ref, one process = one core)
and
    gt_gpu_gather --produce-Broad-GVCF                 (host mirror over libgdbgpu, one B200)
on the same on-disk fragment of a 100 k-sample column slice.  Checks that the two BCF files are byte-identical and prints
both wall times (end-to-end wall of the query tool, fragment files in -> BCF bytes out)."""
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
    ap.add_argument("--sites", type=int, default=40)
    ap.add_argument("--compress", action="store_true")
    args = ap.parse_args()
    from tools import gdbsyn, tiledb_writer
    d = gdbsyn.generate(**dict(gdbsyn.CONFIGS["config4"], n_samples=args.samples), site_lo=0, site_hi=args.sites, with_fixed=True)
    tmp = tempfile.mkdtemp(prefix="gdbtool")
    ws = os.path.join(tmp, "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000, compress=args.compress)
    lo, hi = d["col_range"]
    q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP"))
    ref_bin = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    ours_bin = os.path.join(REPO, "genomicsdb_b200", "bin", "gt_gpu_gather")
    ref_out, our_out = os.path.join(tmp, "ref.bcf"), os.path.join(tmp, "ours.bcf")
    t0 = time.time()
    with open(ref_out, "wb") as f:
        p = subprocess.run([ref_bin, "-j", q, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    t_ref = time.time() - t0
    if p.returncode != 0:
        sys.exit("reference failed: " + p.stderr[-1000:])
    ref_timer = [l for l in p.stderr.splitlines() if "Total scan_and_produce_Broad_GVCF" in l]
    # the synthetic reference genome is ACGT repeated (tools/tiledb_writer.py)
    bases = os.path.join(tmp, "bases.tsv")
    with open(bases, "w") as f:
        for pos in range(max(0, lo - 2), hi + 3):
            f.write("1\t%d\t%s\n" % (pos, "ACGT"[pos % 4]))
    t0 = time.time()
    p2 = subprocess.run([ours_bin, "-j", q, "--produce-Broad-GVCF", "--bcf-header", ref_out, "--reference-bases", bases, "-o", our_out],
                        capture_output=True, text=True)
    t_ours = time.time() - t0
    if p2.returncode != 0:
        sys.exit("gt_gpu_gather failed: " + p2.stderr[-1000:])
    t0 = time.time()
    p3 = subprocess.run([ours_bin, "-j", q, "--produce-Broad-GVCF", "--bcf-header", ref_out, "--reference-bases", bases, "-o", "/dev/null"],
                        capture_output=True, text=True)
    t_ours_null = time.time() - t0
    same = subprocess.run(["cmp", "-s", ref_out, our_out]).returncode == 0
    rec = json.loads(p2.stdout.strip().splitlines()[-1])
    print(json.dumps(dict(samples=args.samples, candidate_sites=args.sites, cells=int(len(d["end"])), compressed_fragment=args.compress,
                          records=rec["records"], bcf_bytes=os.path.getsize(ref_out), byte_identical=same,
                          reference_wall_s=t_ref, reference_timer=ref_timer[-1] if ref_timer else None,
                          ours_wall_s=t_ours, ours_wall_s_to_dev_null=t_ours_null, ours_phases_s=rec.get("seconds"),
                          ours_phases_s_to_dev_null=json.loads(p3.stdout.strip().splitlines()[-1]).get("seconds"),
                          note="reference: one process (one core), output to a file; ours: process start, CUDA context, fragment read, "
                               "scan, records written with writev; both walls include everything")))


if __name__ == "__main__":
    main()
