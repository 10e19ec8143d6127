#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/bcf/<case>.bcf from the UNMODIFIED reference (oracle/_ref).

For every case of tests/golden/manifest.json it runs the reference's own
    gt_mpi_gather --produce-Broad-GVCF -O bu          (BroadCombinedGVCFOperator -> htslib, uncompressed BCF2)
over the committed golden array and keeps the BCF bytes.  The FORMAT blocks of those records (GT and the allele-dependent
fields: typed width, values per sample, bytes) pin the FORMAT collation + BCF typed encoding row of SURVEY 8f-1
(variant_field_handler.cc:804-871, broad_combined_gvcf.cc:53-138,603-727, htslib vcf.c:1926-1972).

Runs in THIS container only (needs /root/reference/tests for the VCF header template and the reference genome).
Usage: python oracle/make_golden_bcf.py          (after bash oracle/build_ref.sh and oracle/make_golden.py)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_TESTS = os.environ.get("GDB_REFERENCE", "/root/reference") + "/tests"
BIN = os.path.join(HERE, "_ref", "bin")
GOLD = os.path.join(REPO, "tests", "golden")


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("reference tests not present; BCF fixtures can only be regenerated where /root/reference exists")
    cases = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
    out_dir, java_dir, sites_dir = os.path.join(GOLD, "bcf"), os.path.join(GOLD, "bcf_java"), os.path.join(GOLD, "bcf_sites_only")
    for d in (out_dir, java_dir, sites_dir):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    tmp = tempfile.mkdtemp(prefix="gdbbcf")
    for c in cases:
        ws = os.path.join(tmp, "ws_" + c["id"])
        shutil.copytree(os.path.join(GOLD, c["workspace"]), ws)
        open(os.path.join(ws, c["array"], ".__consolidation_lock"), "wb").close()
        q = {
            "workspace": ws, "array_name": c["array"],
            "vcf_header_filename": ["inputs/template_vcf_header.vcf"],
            "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_row_ranges"]]}],
            "reference_genome": "inputs/chr1_10MB.fasta.gz",
            "attributes": c["attributes"], "segment_size": 10000,
            "vid_mapping_file": os.path.join(GOLD, c["vid_mapping_file"]),
            "callset_mapping_file": os.path.join(GOLD, c["callset_mapping_file"]),
            "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_column_ranges"]]}],
            "produce_GT_field": c["produce_GT_field"], "produce_FILTER_field": c.get("produce_FILTER_field", False),
            "produce_GT_with_min_PL_value_for_spanning_deletions": c["produce_GT_with_min_PL_value_for_spanning_deletions"],
        }
        qf = os.path.join(tmp, c["id"] + ".json")
        json.dump(q, open(qf, "w"))
        p = subprocess.run([BIN + "/gt_mpi_gather", "-j", qf, "--produce-Broad-GVCF", "-O", "bu"], capture_output=True, cwd=REF_TESTS)
        if p.returncode != 0:
            sys.exit("reference failed on %s: %s" % (c["id"], p.stderr.decode()[-1500:]))
        open(os.path.join(out_dir, c["id"] + ".bcf"), "wb").write(p.stdout)
        # the same query with the conventions of the JNI path (htsjdk's BCF2 codec: missing instead of vector_end, no-call GT)
        pj = subprocess.run([BIN + "/gdb_ref_dump", "-j", qf, "--mode", "bcf-java"], capture_output=True, cwd=REF_TESTS)
        if pj.returncode != 0:
            sys.exit("reference (java conventions) failed on %s: %s" % (c["id"], pj.stderr.decode()[-1500:]))
        open(os.path.join(java_dir, c["id"] + ".bcf"), "wb").write(pj.stdout)
        # and as a sites-only query (no samples, no FORMAT fields; broad_combined_gvcf.cc:178-182,302-318)
        qs = os.path.join(tmp, c["id"] + "_sites_only.json")
        json.dump(dict(q, sites_only_query=True), open(qs, "w"))
        ps = subprocess.run([BIN + "/gt_mpi_gather", "-j", qs, "--produce-Broad-GVCF", "-O", "bu"], capture_output=True, cwd=REF_TESTS)
        if ps.returncode != 0:
            sys.exit("reference (sites only) failed on %s: %s" % (c["id"], ps.stderr.decode()[-1500:]))
        open(os.path.join(sites_dir, c["id"] + ".bcf"), "wb").write(ps.stdout)
        print("%-50s %6d bytes, java conventions %6d bytes, sites only %6d bytes" % (c["id"], len(p.stdout), len(pj.stdout), len(ps.stdout)))
    shutil.rmtree(tmp, ignore_errors=True)
    write_reference_bases(out_dir)


def write_reference_bases(bcf_dir):
    """tests/golden/inputs/ref_bases.tsv: the reference-genome base at every record position of the BCF fixtures (what
    VCFAdapter::get_reference_base_at_position returns when a merged REF is 'N', broad_combined_gvcf.cc:811-816) -- the
    few bytes of tests/inputs/chr1_10MB.fasta.gz the record assembly needs, so the 2.8 MB FASTA does not have to travel."""
    import gzip
    sys.path.insert(0, REPO)
    from oracle import bcf_reader
    seq = {}
    name = None
    with gzip.open(os.path.join(REF_TESTS, "inputs", "chr1_10MB.fasta.gz"), "rt") as f:
        for line in f:
            if line.startswith(">"):
                name = line[1:].split()[0]
                seq[name] = []
            else:
                seq[name].append(line.strip())
    seq = {k: "".join(v) for k, v in seq.items()}
    need = set()
    for fn in sorted(os.listdir(bcf_dir)):
        text, _, recs = bcf_reader.read_bcf(os.path.join(bcf_dir, fn))
        contigs = [l.split("ID=")[1].split(",")[0].split(">")[0] for l in text.split("\n") if l.startswith("##contig=<")]
        for r in recs:
            need.add((contigs[r["chrom"]], r["pos"]))
    with open(os.path.join(GOLD, "inputs", "ref_bases.tsv"), "w") as f:
        for contig, pos in sorted(need):
            if contig in seq and pos < len(seq[contig]):
                f.write("%s\t%d\t%s\n" % (contig, pos, seq[contig][pos].upper()))
    print("reference bases: %d positions" % len(need))


if __name__ == "__main__":
    main()
