#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/variants/<case>.variants from the UNMODIFIED reference: its query tool in its
default mode (gt_mpi_gather -j <query>: run_range_query -> VariantQueryProcessor::gt_get_column_interval + print_variants,
tools/src/gt_mpi_gather.cc:102-319, query_variants.cc:687-843 -- SURVEY 8f-4, the GA4GH range query) over every committed golden
array; where the reference's CI holds a golden for the query (tests/golden_outputs/*_variants_*) the output is checked against it.
Runs in THIS container only.  Usage: python oracle/make_golden_variants.py   (after oracle/build_ref.sh)"""
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
CI = {"t0_1_2_at_0": "t0_1_2_variants_at_0", "t0_1_2_at_12150": "t0_1_2_variants_at_12150", "t6_7_8_at_0": "t6_7_8_variants_at_0",
      "t6_7_8_at_8029500": "t6_7_8_variants_at_8029500", "t0_1_2_phased_GT_at_0": "t0_1_2_variants_at_0_phased_GT",
      "t0_1_2_phased_GT_at_12150": "t0_1_2_variants_at_12150_phased_GT", "t0_1_2_DS_ID_at_0": "t0_1_2_DS_ID_variants_at_0"}
CI_ATTRIBUTES = ["REF", "ALT", "BaseQRankSum", "MQ", "RAW_MQ", "MQ0", "ClippingRankSum", "MQRankSum", "ReadPosRankSum", "DP", "GT", "GQ", "SB", "AD", "PL",
                 "DP_FORMAT", "MIN_DP", "PID", "PGT"]   # tests/run.py:52


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("reference tests not present; fixtures can only be regenerated where /root/reference exists")
    cases = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
    out_dir = os.path.join(GOLD, "variants")
    shutil.rmtree(out_dir, ignore_errors=True)
    os.makedirs(out_dir)
    tmp = tempfile.mkdtemp(prefix="gdbvariants")
    ci_ok = ci_n = 0
    for c in cases:
        ws = os.path.join(tmp, "ws_" + c["id"])
        shutil.copytree(os.path.join(GOLD, c["workspace"]), ws)
        open(os.path.join(ws, c["array"], ".__consolidation_lock"), "wb").close()

        def run(attrs):
            q = {"workspace": ws, "array_name": c["array"], "vcf_header_filename": ["inputs/template_vcf_header.vcf"],
                 "reference_genome": "inputs/chr1_10MB.fasta.gz",
                 "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_row_ranges"]]}],
                 "attributes": attrs, "segment_size": 10000, "vid_mapping_file": os.path.join(GOLD, c["vid_mapping_file"]),
                 "callset_mapping_file": os.path.join(GOLD, c["callset_mapping_file"]),
                 "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in c["query_column_ranges"]]}]}
            qf = os.path.join(tmp, c["id"] + ".json")
            json.dump(q, open(qf, "w"))
            return subprocess.run(["stdbuf", "-o0", BIN + "/gt_mpi_gather", "-j", qf], capture_output=True, cwd=REF_TESTS)
        p = run(c["attributes"])
        if p.returncode != 0:
            print("reference failed on %s: %s" % (c["id"], p.stderr.decode()[-400:]))
            continue
        open(os.path.join(out_dir, c["id"] + ".variants"), "wb").write(p.stdout)
        note = ""
        if c["id"] in CI:
            attrs = CI_ATTRIBUTES + ["DS", "ID"] if "DS_ID" in c["id"] else CI_ATTRIBUTES
            pc = run(attrs)
            gold = open(os.path.join(REF_TESTS, "golden_outputs", CI[c["id"]]), "rb").read()
            ci_n += 1
            same = pc.returncode == 0 and (hashlib.md5(pc.stdout).hexdigest() == hashlib.md5(gold).hexdigest() or json.loads(pc.stdout) == json.loads(gold))
            ci_ok += same
            open(os.path.join(out_dir, c["id"] + ".ci.variants"), "wb").write(pc.stdout)
            note = "  CI golden %s: %s" % (CI[c["id"]], "identical" if same else "DIFFERENT")
        print("%-50s %8d bytes%s" % (c["id"], len(p.stdout), note))
    json.dump({"ci_attributes": CI_ATTRIBUTES, "ci_cases": sorted(CI)}, open(os.path.join(out_dir, "ci.json"), "w"))
    print("reference CI goldens reproduced: %d / %d" % (ci_ok, ci_n))


if __name__ == "__main__":
    main()
