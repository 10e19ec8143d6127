#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/cells/<case>.{calls,csv,ac,hist} from the UNMODIFIED reference.

For every case of tests/golden/manifest.json it runs the reference's own query tool in the four modes that sit on the
columnar iterator (SingleCellTileDBIterator -> VariantQueryProcessor::iterate_over_cells):
    gt_mpi_gather --print-calls          VariantCallPrintOperator     (variant_operations.cc:808-846)
    gt_mpi_gather --print-csv            VariantCallPrintCSVOperator  (variant_operations.cc:894-898)
    gt_mpi_gather --print-AC             AlleleCountOperator          (variant_operations.cc:900-1089)
    gt_mpi_gather --produce-histogram    ColumnHistogramOperator      (variant_operations.cc:732-790)
over the committed golden array and keeps stdout.  Where the reference's CI holds a golden for the same query
(tests/golden_outputs/*_calls_*), the --print-calls output is checked against it first (md5).

Runs in THIS container only.  Usage: python oracle/make_golden_cells.py   (after oracle/build_ref.sh, oracle/make_golden.py)
"""
import hashlib
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

CI_ATTRIBUTES = ["REF", "ALT", "BaseQRankSum", "MQ", "RAW_MQ", "MQ0", "ClippingRankSum", "MQRankSum", "ReadPosRankSum", "DP", "GT", "GQ", "SB", "AD", "PL",
                 "DP_FORMAT", "MIN_DP", "PID", "PGT"]   # tests/run.py:52
# the reference CI's own goldens for --print-calls (tests/run.py:185-870), by manifest case
CI_CALLS = {
    "t0_1_2_at_0": "t0_1_2_calls_at_0", "t0_1_2_at_12150": "t0_1_2_calls_at_12150", "t0_1_2_multiple_positions": "t0_1_2_calls_at_multiple_positions",
    "t0_1_2_at_12100_12141_to_12150_12300_17384": "t0_1_2_calls_at_12100_12141_to_12150_12300_17384",
    "t6_7_8_at_0": "t6_7_8_calls_at_0", "t6_7_8_at_8029500": "t6_7_8_calls_at_8029500",
    "test_new_fields_at_0": "test_new_fields_MLEAC_only.json", "t0_with_missing_PL_SB_fields_at_0": "t0_with_missing_PL_SB_fields_t1_calls.json",
    "t0_1_2_phased_GT_at_0": "t0_1_2_calls_at_0_phased_GT", "t0_1_2_phased_GT_at_12150": "t0_1_2_calls_at_12150_phased_GT",
    "t0_1_2_DS_ID_at_0": "t0_1_2_DS_ID_calls_at_0",
    "t0_1_2_uncompressed_at_0": "t0_1_2_calls_at_0", "t0_1_2_two_fragments_at_0": "t0_1_2_calls_at_0", "t0_1_2_two_fragments_at_12150": "t0_1_2_calls_at_12150",
}


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("reference tests not present; fixtures can only be regenerated where /root/reference exists")
    cases = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
    out_dir = os.path.join(GOLD, "cells")
    shutil.rmtree(out_dir, ignore_errors=True)
    os.makedirs(out_dir)
    tmp = tempfile.mkdtemp(prefix="gdbcells")
    ci_ok = ci_n = 0
    ci_manifest = {}
    for c in cases:
        ws = os.path.join(tmp, "ws_" + c["id"])
        shutil.copytree(os.path.join(GOLD, c["workspace"]), ws)
        open(os.path.join(ws, c["array"], ".__consolidation_lock"), "wb").close()
        q = {
            "workspace": ws, "array_name": c["array"],
            "vcf_header_filename": ["inputs/template_vcf_header.vcf"], "reference_genome": "inputs/chr1_10MB.fasta.gz",
            "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_row_ranges"]]}],
            "attributes": c["attributes"], "segment_size": 10000,
            "vid_mapping_file": os.path.join(GOLD, c["vid_mapping_file"]),
            "callset_mapping_file": os.path.join(GOLD, c["callset_mapping_file"]),
            "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_column_ranges"]]}],
        }
        qf = os.path.join(tmp, c["id"] + ".json")
        json.dump(q, open(qf, "w"))
        sizes = []
        for ext, args in (("calls", ["--print-calls"]), ("csv", ["--print-csv"]), ("ac", ["--print-AC"]), ("hist", ["--produce-histogram"]),
                          ("ipos", ["--produce-interesting-positions"])):   # InterestingLocationsPrinter through scan_and_operate (hot-path seam)
            # stdbuf: this build of the tool leaves without flushing a block-buffered stdout
            p = subprocess.run(["stdbuf", "-o0", BIN + "/gt_mpi_gather", "-j", qf] + args, capture_output=True, cwd=REF_TESTS)
            if p.returncode != 0:
                print("reference failed on %s %s: %s" % (c["id"], ext, p.stderr.decode()[-600:]))
                continue
            open(os.path.join(out_dir, c["id"] + "." + ext), "wb").write(p.stdout)
            sizes.append("%s %d" % (ext, len(p.stdout)))
        if c["id"] in CI_CALLS:   # the CI's query for this array: its attribute list, its golden
            attrs = ["MLEAC"] if c["id"] == "test_new_fields_at_0" else (CI_ATTRIBUTES + ["DS", "ID"] if "DS_ID" in c["id"] else CI_ATTRIBUTES)
            qc = os.path.join(tmp, c["id"] + "_ci.json")
            json.dump(dict(q, attributes=attrs), open(qc, "w"))
            p = subprocess.run(["stdbuf", "-o0", BIN + "/gt_mpi_gather", "-j", qc, "--print-calls"], capture_output=True, cwd=REF_TESTS)
            gold = open(os.path.join(REF_TESTS, "golden_outputs", CI_CALLS[c["id"]]), "rb").read()
            same = p.returncode == 0 and (hashlib.md5(gold).hexdigest() == hashlib.md5(p.stdout).hexdigest() or json.loads(gold) == json.loads(p.stdout))
            ci_n += 1
            ci_ok += same
            sizes.append("CI golden %s" % ("ok" if same else "DIFFERS"))
            if same:
                open(os.path.join(out_dir, c["id"] + ".ci.calls"), "wb").write(p.stdout)
                ci_manifest[c["id"]] = dict(attributes=attrs, reference_ci_golden=CI_CALLS[c["id"]])
        print("%-50s %s" % (c["id"], ", ".join(sizes)))
    json.dump(ci_manifest, open(os.path.join(out_dir, "ci_manifest.json"), "w"), indent=1)
    print("reference CI calls goldens matched: %d/%d" % (ci_ok, ci_n))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
