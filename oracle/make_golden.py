#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ from the UNMODIFIED reference built in oracle/_ref.

Runs in THIS container only (needs /root/reference/tests inputs).  For every non-Java scenario of the reference's
own CI driver (tests/run.py:185-870, restated here as data) it
  1. imports the scenario's gVCFs with the reference's vcf2tiledb            -> tests/golden/arrays/<scenario>/<array>
  2. re-runs the reference's golden-output check (gt_mpi_gather --produce-Broad-GVCF, md5 against
     /root/reference/tests/golden_outputs/<file>, exactly as tests/run.py:935-990 does) to pin oracle/_ref
  3. runs oracle/_ref/bin/gdb_ref_dump (the reference engine with a dump operator plugged into
     SingleVariantOperatorBase::operate)                                       -> tests/golden/dumps/<id>.bin
  4. copies the reference's golden VCF text next to it                        -> tests/golden/vcf/<file>
and writes tests/golden/manifest.json, which the parity tests iterate over.

Usage: python oracle/make_golden.py            (after bash oracle/build_ref.sh)
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

VCF_ATTRS = ["END", "REF", "ALT", "BaseQRankSum", "ClippingRankSum", "MQRankSum", "ReadPosRankSum", "MQ", "RAW_MQ", "MQ0",
             "DP", "GT", "GQ", "SB", "AD", "PL", "PGT", "PID", "MIN_DP", "DP_FORMAT", "FILTER"]
ASA_ATTRS = VCF_ATTRS + ["AS_RAW_MQ", "AS_RAW_MQRankSum"]
FULL = [[0, 1000000000]]

# (scenario, loader overrides, [queries]); a query = dict(id, ranges | positions, golden vcf file or None, extra query keys)
SCENARIOS = [
    dict(name="t0_1_2", callsets="t0_1_2.json", compress=True, queries=[
        dict(id="t0_1_2_at_0", ranges=FULL, vcf="t0_1_2_vcf_at_0"),
        dict(id="t0_1_2_multiple_positions", positions=[12000, 12142, 12144, 12160, 12290, 12294, 14000, 17384, 18000],
             vcf="t0_1_2_vcf_at_multiple_positions"),
        dict(id="t0_1_2_at_12150", ranges=[[12150, 1000000000]], vcf="t0_1_2_vcf_at_12150"),
        dict(id="t0_1_2_at_0_FILTER", ranges=FULL, vcf="t0_1_2_vcf_at_0_with_FILTER", extra={"produce_FILTER_field": True}),
        dict(id="t0_1_2_at_12100_12141_to_12150_12300_17384",
             ranges=[[12100, 12100], [12141, 12150], [12300, 12300], [17384, 17384]], vcf=None),
    ]),
    dict(name="t0_1_2_uncompressed", callsets="t0_1_2.json", queries=[
        dict(id="t0_1_2_uncompressed_at_0", ranges=FULL, vcf="t0_1_2_vcf_at_0"),
    ]),
    dict(name="t0_overlapping", callsets="t0_overlapping.json", queries=[
        dict(id="t0_overlapping_at_12202", ranges=[[12202, 1000000000]], vcf="t0_overlapping_at_12202"),
        dict(id="t0_overlapping_at_0", ranges=FULL, vcf=None),
    ]),
    dict(name="t6_7_8", callsets="t6_7_8.json", queries=[
        dict(id="t6_7_8_at_0", ranges=FULL, vcf="t6_7_8_vcf_at_0"),
        dict(id="t6_7_8_at_8029500", ranges=[[8029500, 1000000000]], vcf="t6_7_8_vcf_at_8029500"),
        dict(id="t6_7_8_at_8029500_8029500", ranges=[[8029500, 8029500]], vcf="t6_7_8_vcf_at_8029500-8029500"),
    ]),
    dict(name="test_new_fields", callsets="t6_7_8.json", vid="vid_MLEAC_MLEAF.json", queries=[
        dict(id="test_new_fields_at_0", ranges=FULL, vcf=None, attributes=VCF_ATTRS + ["MLEAC", "MLEAF"],
             extra={"vid_mapping_file": "inputs/vid_MLEAC_MLEAF.json"}),
    ]),
    dict(name="t0_1_2_combined", callsets="t0_1_2_combined.json", queries=[
        dict(id="t0_1_2_combined_at_0", ranges=FULL, vcf="t0_1_2_combined"),
    ]),
    dict(name="t0_with_missing_PL_SB_fields", callsets="t0_with_missing_PL_SB_fields_t1.json", queries=[
        dict(id="t0_with_missing_PL_SB_fields_at_0", ranges=FULL, vcf=None),
    ]),
    dict(name="t0_haploid_triploid_1_2_3_triploid_deletion", callsets="t0_haploid_triploid_1_2_3_triploid_deletion.json",
         vid="vid_DS_ID_phased_GT.json", size_per_column_partition=1200, segment_size=100, queries=[
        dict(id="haploid_triploid_deletion", ranges=FULL, vcf="t0_haploid_triploid_1_2_3_triploid_deletion_vcf",
             extra={"vid_mapping_file": "inputs/vid_DS_ID_phased_GT.json",
                    "callset_mapping_file": "inputs/callsets/t0_haploid_triploid_1_2_3_triploid_deletion.json"}),
        dict(id="haploid_triploid_deletion_produce_GT", ranges=FULL,
             vcf="t0_haploid_triploid_1_2_3_triploid_deletion_vcf_produce_GT",
             extra={"vid_mapping_file": "inputs/vid_DS_ID_phased_GT.json", "produce_GT_field": True,
                    "callset_mapping_file": "inputs/callsets/t0_haploid_triploid_1_2_3_triploid_deletion.json"}),
        dict(id="haploid_triploid_deletion_min_PL", ranges=FULL,
             vcf="t0_haploid_triploid_1_2_3_triploid_deletion_vcf_produce_GT_for_min_value_PL",
             extra={"vid_mapping_file": "inputs/vid_DS_ID_phased_GT.json", "produce_GT_field": True,
                    "produce_GT_with_min_PL_value_for_spanning_deletions": True,
                    "callset_mapping_file": "inputs/callsets/t0_haploid_triploid_1_2_3_triploid_deletion.json"}),
    ]),
    dict(name="t0_1_2_all_asa", callsets="t0_1_2_all_asa.json", vid="vid_all_asa.json", size_per_column_partition=3000,
         queries=[
        dict(id="t0_1_2_all_asa_at_0", ranges=FULL, vcf="t0_1_2_all_asa_loading", attributes=ASA_ATTRS, segment_size=100,
             extra={"vid_mapping_file": "inputs/vid_all_asa.json"}),
    ]),
    dict(name="min_PL_spanning_deletion", callsets="min_PL_spanning_deletion.json", vid="vid_phased_GT.json", queries=[
        dict(id="min_PL_spanning_deletion_no_min_PL", ranges=FULL, vcf="min_PL_spanning_deletion_vcf_no_min_PL",
             extra={"vid_mapping_file": "inputs/vid_phased_GT.json", "produce_GT_field": True}),
        dict(id="min_PL_spanning_deletion_min_PL", ranges=FULL, vcf="min_PL_spanning_deletion_vcf",
             extra={"vid_mapping_file": "inputs/vid_phased_GT.json", "produce_GT_field": True,
                    "produce_GT_with_min_PL_value_for_spanning_deletions": True}),
    ]),
    # two import batches -> two TileDB fragments (exercises the fragment k-way merge, array_read_state.cc:1836-2132)
    dict(name="t0_1_2_two_fragments", callsets="t0_1_2.json", batches=[(0, 1), (2, 2)], queries=[
        dict(id="t0_1_2_two_fragments_at_0", ranges=FULL, vcf="t0_1_2_vcf_at_0"),
        dict(id="t0_1_2_two_fragments_at_12150", ranges=[[12150, 1000000000]], vcf="t0_1_2_vcf_at_12150"),
    ]),
    dict(name="t6_7_8_three_fragments_gz", callsets="t6_7_8.json", compress=True, batches=[(1, 1), (0, 0), (2, 2)], queries=[
        dict(id="t6_7_8_three_fragments_at_0", ranges=FULL, vcf="t6_7_8_vcf_at_0"),
    ]),
    # INFO concatenate (tests/run.py:602-605, vid_info_ops1.json)
    dict(name="info_ops1", callsets="info_ops.json", vid="vid_info_ops1.json", queries=[
        dict(id="info_ops1_at_0", ranges=FULL, vcf=None,
             attributes=["REF", "ALT", "QUAL", "BaseQRankSum", "ClippingRankSum", "GT", "MLEAC", "MLEAF", "var_field"],
             extra={"vid_mapping_file": "inputs/vid_info_ops1.json"}),
    ]),
    # the same gVCFs imported with GT stored phased (tests/run.py:493-520) and with the DS flag / ID string fields (:687-704)
    dict(name="t0_1_2_phased_GT", callsets="t0_1_2.json", vid="vid_phased_GT.json", queries=[
        dict(id="t0_1_2_phased_GT_at_0", ranges=FULL, vcf="t0_1_2_vcf_at_0", extra={"vid_mapping_file": "inputs/vid_phased_GT.json"}),
        dict(id="t0_1_2_phased_GT_at_12150", ranges=[[12150, 1000000000]], vcf="t0_1_2_vcf_at_12150",
             extra={"vid_mapping_file": "inputs/vid_phased_GT.json"}),
    ]),
    dict(name="t0_1_2_DS_ID", callsets="t0_1_2.json", vid="vid_DS_ID.json", queries=[
        dict(id="t0_1_2_DS_ID_at_0", ranges=FULL, vcf=None,
             attributes=["REF", "ALT", "BaseQRankSum", "MQ", "RAW_MQ", "MQ0", "ClippingRankSum", "MQRankSum", "ReadPosRankSum", "DP", "GT", "GQ",
                         "SB", "AD", "PL", "DP_FORMAT", "MIN_DP", "PID", "PGT", "DS", "ID"],
             extra={"vid_mapping_file": "inputs/vid_DS_ID.json"}),
    ]),
    # INFO combine operations (tests/run.py:598-601, vid_info_ops0.json): QUAL median, mean, element-wise sums of A-length and
    # variable-length fields
    dict(name="info_ops0", callsets="info_ops.json", vid="vid_info_ops0.json", queries=[
        dict(id="info_ops0_at_0", ranges=FULL, vcf=None,
             attributes=["REF", "ALT", "QUAL", "BaseQRankSum", "ClippingRankSum", "GT", "MLEAC", "MLEAF", "var_field"],
             extra={"vid_mapping_file": "inputs/vid_info_ops0.json"}),
    ]),
]


def md5(b):
    return hashlib.md5(b).hexdigest()


def run(cmd, **kw):
    p = subprocess.run(cmd, shell=True, capture_output=True, cwd=REF_TESTS, **kw)
    if p.returncode != 0:
        sys.stderr.write("FAILED: %s\n%s\n" % (cmd, p.stderr.decode()[-2000:]))
        sys.exit(1)
    return p.stdout


def loader_json(ws, sc, lb=None, ub=None, first=True):
    d = {
        "row_based_partitioning": False,
        "column_partitions": [{"begin": 0, "workspace": ws, "array_name": sc["name"]}],
        "callset_mapping_file": "inputs/callsets/" + sc["callsets"],
        "vid_mapping_file": "inputs/" + sc.get("vid", "vid.json"),
        "size_per_column_partition": sc.get("size_per_column_partition", 700),
        "treat_deletions_as_intervals": True,
        "vcf_header_filename": "inputs/template_vcf_header.vcf",
        "reference_genome": "inputs/chr1_10MB.fasta.gz",
        "num_parallel_vcf_files": 1, "do_ping_pong_buffering": False, "offload_vcf_output_processing": False,
        "discard_vcf_index": True, "produce_combined_vcf": True, "produce_tiledb_array": True,
        "delete_and_create_tiledb_array": first, "compress_tiledb_array": bool(sc.get("compress", False)),
        "segment_size": sc.get("segment_size", 40), "num_cells_per_tile": 3,
    }
    if lb is not None:
        d["lb_callset_row_idx"] = lb
        d["ub_callset_row_idx"] = ub
        d["produce_combined_vcf"] = False
    return d


def query_json(ws, sc, q):
    d = {
        "workspace": ws, "array_name": sc["name"],
        "vcf_header_filename": ["inputs/template_vcf_header.vcf"],
        "query_row_ranges": [{"range_list": [{"low": 0, "high": 3}]}],
        "reference_genome": "inputs/chr1_10MB.fasta.gz",
        "attributes": q.get("attributes", VCF_ATTRS),
        "segment_size": q.get("segment_size", sc.get("segment_size", 40)),
        "vid_mapping_file": "inputs/" + sc.get("vid", "vid.json"),
        "callset_mapping_file": "inputs/callsets/" + sc["callsets"],
    }
    if "positions" in q:
        d["query_column_ranges"] = [q["positions"]]
    else:
        d["query_column_ranges"] = [{"range_list": [{"low": lo, "high": hi} for lo, hi in q["ranges"]]}]
    d.update(q.get("extra", {}))
    return d


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("reference tests not present; golden fixtures can only be regenerated where /root/reference exists")
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None   # add scenarios, keep the others
    manifest = []
    if only is None:
        for sub in ("arrays", "dumps", "vcf", "inputs/callsets"):
            shutil.rmtree(os.path.join(GOLD, sub), ignore_errors=True)
            os.makedirs(os.path.join(GOLD, sub))
    else:
        manifest = [c for c in json.load(open(os.path.join(GOLD, "manifest.json")))["cases"] if c["scenario"] not in only]
        for name in only:
            shutil.rmtree(os.path.join(GOLD, "arrays", name), ignore_errors=True)
    tmp = tempfile.mkdtemp(prefix="gdbgold")
    used_inputs = set()
    n_vcf_ok = n_vcf = 0
    for sc in SCENARIOS:
        if only is not None and sc["name"] not in only:
            continue
        ws = os.path.join(tmp, "ws_" + sc["name"])
        batches = sc.get("batches", [None])
        for bi, b in enumerate(batches):
            lj = loader_json(ws, sc, *(b if b else (None, None)), first=(bi == 0))
            ljf = os.path.join(tmp, sc["name"] + "_loader%d.json" % bi)
            json.dump(lj, open(ljf, "w"), indent=1)
            run(BIN + "/vcf2tiledb " + ljf)
        ljf = os.path.join(tmp, sc["name"] + "_loader0.json")
        used_inputs.add("callsets/" + sc["callsets"])
        used_inputs.add(sc.get("vid", "vid.json"))
        # copy the array (skip lock files)
        dst = os.path.join(GOLD, "arrays", sc["name"])
        shutil.copytree(ws, dst, ignore=shutil.ignore_patterns(".__*"))
        nfrag = len([d for d in os.listdir(os.path.join(dst, sc["name"])) if d.startswith("__") and
                     os.path.isdir(os.path.join(dst, sc["name"], d))])
        for q in sc["queries"]:
            qj = query_json(ws, sc, q)
            qjf = os.path.join(tmp, q["id"] + "_query.json")
            json.dump(qj, open(qjf, "w"), indent=1)
            vcf_ok = None
            if q.get("vcf"):
                out = run("%s/gt_mpi_gather -s %d -l %s -j %s --produce-Broad-GVCF -p 128" % (BIN, qj["segment_size"], ljf, qjf))
                gold = open(os.path.join(REF_TESTS, "golden_outputs", q["vcf"]), "rb").read()
                vcf_ok = (md5(out) == md5(gold))
                n_vcf += 1
                n_vcf_ok += int(vcf_ok)
                if not vcf_ok:
                    sys.stderr.write("WARNING: reference golden mismatch for %s\n" % q["id"])
                shutil.copy(os.path.join(REF_TESTS, "golden_outputs", q["vcf"]), os.path.join(GOLD, "vcf", q["vcf"]))
            dump = os.path.join(GOLD, "dumps", q["id"] + ".bin")
            info = json.loads(run("%s/gdb_ref_dump -l %s -j %s -o %s --mode dump" % (BIN, ljf, qjf, dump)).decode().strip().splitlines()[-1])
            entry = {
                "id": q["id"], "scenario": sc["name"], "workspace": "arrays/" + sc["name"], "array": sc["name"],
                "num_fragments": nfrag, "compressed": bool(sc.get("compress", False)),
                "vid_mapping_file": "inputs/" + sc.get("vid", "vid.json"),
                "callset_mapping_file": "inputs/callsets/" + sc["callsets"],
                "attributes": qj["attributes"],
                "query_column_ranges": ([[p, p] for p in q["positions"]] if "positions" in q else q["ranges"]),
                "query_row_ranges": [[0, 3]],
                "produce_GT_field": bool(qj.get("produce_GT_field", False)),
                "produce_FILTER_field": bool(qj.get("produce_FILTER_field", False)),
                "produce_GT_with_min_PL_value_for_spanning_deletions":
                    bool(qj.get("produce_GT_with_min_PL_value_for_spanning_deletions", False)),
                "dump": "dumps/" + q["id"] + ".bin", "dump_md5": md5(open(dump, "rb").read()),
                "sites": info["sites"], "calls": info["calls"],
                "vcf_golden": ("vcf/" + q["vcf"]) if q.get("vcf") else None,
                "reference_vcf_md5_matches_golden": vcf_ok,
            }
            manifest.append(entry)
            print("%-50s sites=%-4d calls=%-5d frags=%d vcf_md5=%s" % (q["id"], info["sites"], info["calls"], nfrag, vcf_ok))
    for f in sorted(used_inputs):
        shutil.copy(os.path.join(REF_TESTS, "inputs", f), os.path.join(GOLD, "inputs", f))
    json.dump({"generator": "oracle/make_golden.py", "reference": "Intel-HLS/GenomicsDB v0.10.2 (oracle/_ref, unmodified)",
               "cases": manifest}, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    shutil.rmtree(tmp, ignore_errors=True)
    print("reference VCF goldens reproduced: %d / %d" % (n_vcf_ok, n_vcf))


if __name__ == "__main__":
    main()
