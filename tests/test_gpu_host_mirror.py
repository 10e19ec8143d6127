"""GPU: the host-side C++ mirror of the reference's plugin surface (VariantQueryProcessor::scan_and_operate +
SingleVariantOperatorBase::operate, genomicsdb_b200/csrc/gdb_host_api.h) driven by tools/gt_gpu_gather.cc -- the
counterpart of oracle/ref_harness/gdb_ref_dump.cc -- against the reference's dumps."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, golden_cases

pytestmark = pytest.mark.gpu
CASES = [c for c in golden_cases() if c["id"] in (
    "t0_1_2_at_0", "t0_1_2_multiple_positions", "t0_1_2_at_12150", "t6_7_8_at_8029500", "haploid_triploid_deletion_min_PL",
    "min_PL_spanning_deletion_no_min_PL", "t0_1_2_two_fragments_at_0", "test_new_fields_at_0", "t0_overlapping_at_0")]


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    g.build_gpu_lib()
    return g.build_tools()


def run_case(exe, case, tmp_path, page):
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}],
         "produce_GT_with_min_PL_value_for_spanning_deletions": case["produce_GT_with_min_PL_value_for_spanning_deletions"]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "dump.bin"
    cmd = [exe, "-j", str(qf), "-o", str(out)] + (["-p", str(page)] if page else [])
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return out.read_bytes(), json.loads(p.stdout.strip().splitlines()[-1])["sites"]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_scan_and_operate_mirror_matches_reference(exe, case, tmp_path):
    gold = open(os.path.join(GOLDEN, case["dump"]), "rb").read()
    got, n = run_case(exe, case, tmp_path, 0)
    assert n == case["sites"]
    assert got == gold


@pytest.mark.parametrize("case", CASES[:3], ids=[c["id"] for c in CASES[:3]])
def test_overflow_resume_protocol(exe, case, tmp_path):
    """an operator that reports overflow() after every site forces scan_and_operate to return and be re-entered"""
    gold = open(os.path.join(GOLDEN, case["dump"]), "rb").read()
    got, n = run_case(exe, case, tmp_path, 1)
    assert got == gold


# ---------------------------------------------------------------------------------------------- whole BCF records
ALL = golden_cases()


def run_broad(exe, case, tmp_path, page=0, java=False, sites_only=False, env=None):
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}],
         "produce_GT_field": case["produce_GT_field"], "sites_only_query": sites_only, "produce_FILTER_field": case.get("produce_FILTER_field", False),
         "produce_GT_with_min_PL_value_for_spanning_deletions": case["produce_GT_with_min_PL_value_for_spanning_deletions"]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "out.bcf"
    cmd = [exe, "-j", str(qf), "-o", str(out), "--produce-Broad-GVCF", "--bcf-header",
           os.path.join(GOLDEN, "bcf_sites_only" if sites_only else ("bcf_java" if java else "bcf"), case["id"] + ".bcf"),
           "--reference-bases", os.path.join(GOLDEN, "inputs", "ref_bases.tsv")] + (["-p", str(page)] if page else []) + \
          (["--htsjdk-conventions"] if java else [])
    return subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **env) if env else None), out


@pytest.mark.parametrize("java", [False, True], ids=["htslib", "java"])
@pytest.mark.parametrize("case", ALL, ids=[c["id"] for c in ALL])
def test_broad_gvcf_bcf_file_equals_the_reference(exe, case, tmp_path, java):
    """gt_gpu_gather --produce-Broad-GVCF (BroadCombinedGVCFOperator of the host mirror over the kernels' FORMAT blocks, INFO
    scalars and DP sums) writes, behind the reference's header, the very bytes the unmodified reference's
    `gt_mpi_gather --produce-Broad-GVCF -O bu` wrote for the same query; with --htsjdk-conventions the bytes the reference's
    operator writes when constructed the way the JNI path (Java GenomicsDBFeatureReader) constructs it"""
    gold = open(os.path.join(GOLDEN, "bcf_java" if java else "bcf", case["id"] + ".bcf"), "rb").read()
    p, out = run_broad(exe, case, tmp_path, java=java)
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout.strip().splitlines()[-1])["records"] == case["sites"]
    got = out.read_bytes()
    assert len(got) == len(gold)
    assert got == gold


@pytest.mark.parametrize("case", [c for c in ALL if c["id"] in ("t0_1_2_at_0", "t6_7_8_at_0", "haploid_triploid_deletion_produce_GT")],
                         ids=lambda c: c["id"])
def test_broad_gvcf_overflow_resume(exe, case, tmp_path):
    """a 1-byte output buffer: overflow() after every record, scan_and_operate re-entered with the scan state"""
    gold = open(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"), "rb").read()
    p, out = run_broad(exe, case, tmp_path, page=1)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold


@pytest.mark.parametrize("case", [c for c in ALL if c["id"] in ("t0_1_2_at_0", "t6_7_8_at_0", "haploid_triploid_deletion", "info_ops0_at_0",
                                                                 "min_PL_spanning_deletion_min_PL", "t0_overlapping_at_0", "t0_1_2_DS_ID_at_0",
                                                                 "t0_1_2_at_0_FILTER", "t0_1_2_phased_GT_at_12150")], ids=lambda c: c["id"])
def test_broad_gvcf_sites_only(exe, case, tmp_path):
    """"sites_only_query": true -- records without samples and FORMAT fields, INFO (incl. the DP rule over DP_FORMAT / MIN_DP) as before"""
    gold = open(os.path.join(GOLDEN, "bcf_sites_only", case["id"] + ".bcf"), "rb").read()
    p, out = run_broad(exe, case, tmp_path, sites_only=True)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold


# ---------------------------------------------------------------------------------------------- the JNI path's pump
STREAM_CASES = ["t0_1_2_at_0", "t0_1_2_multiple_positions", "haploid_triploid_deletion_produce_GT", "t0_1_2_DS_ID_at_0", "t0_1_2_at_12150"]


@pytest.mark.parametrize("chunk,capacity", [(7, 64), (100000, 0), (4096, 3000)], ids=["7B-reads-64B-buffer", "100kB-reads", "4kB-reads-3kB-buffer"])
@pytest.mark.parametrize("cid", STREAM_CASES)
def test_bcf_generator_stream_equals_the_reference(exe, cid, tmp_path, chunk, capacity):
    """GenomicsDBBCFGenerator of the mirror (what GenomicsDBQueryStream's JNI methods sit on) pulled with read_and_advance /
    read_next_byte in small and large reads, with buffers that overflow after every record: the stream is the reference's
    BCF (JNI conventions), header first"""
    case = [c for c in ALL if c["id"] == cid][0]
    gold = open(os.path.join(GOLDEN, "bcf_java", cid + ".bcf"), "rb").read()
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}],
         "produce_GT_field": case["produce_GT_field"], "produce_FILTER_field": case.get("produce_FILTER_field", False),
         "produce_GT_with_min_PL_value_for_spanning_deletions": case["produce_GT_with_min_PL_value_for_spanning_deletions"]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "stream.bcf"
    cmd = [exe, "-j", str(qf), "-o", str(out), "--bcf-generator", str(chunk), "--htsjdk-conventions", "--bcf-header",
           os.path.join(GOLDEN, "bcf_java", cid + ".bcf"), "--reference-bases", os.path.join(GOLDEN, "inputs", "ref_bases.tsv")]
    if capacity:
        cmd += ["-p", str(capacity)]
    if cid == "t0_1_2_at_12150":       # the interval given the way the Java reader gives it: contig, 1-based begin and end
        cmd += ["--chromosome", "1", "--begin", "12151", "--end", "1000000001"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout.strip().splitlines()[-1])["bytes"] == len(gold)
    assert out.read_bytes() == gold


def test_bcf_generator_header_only(exe, tmp_path):
    """produce_header_only (GenomicsDBQueryStream's way to fetch the header): the stream is the header, then end()"""
    case = [c for c in ALL if c["id"] == "t0_1_2_at_0"][0]
    gold = open(os.path.join(GOLDEN, "bcf_java", "t0_1_2_at_0.bcf"), "rb").read()
    import struct
    l_text = struct.unpack_from("<I", gold, 5)[0]
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": 0, "high": 1000000000}]}]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "header.bcf"
    p = subprocess.run([exe, "-j", str(qf), "-o", str(out), "--bcf-generator", "1000", "--header-only", "--bcf-header",
                        os.path.join(GOLDEN, "bcf_java", "t0_1_2_at_0.bcf")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold[:9 + l_text]


@pytest.mark.parametrize("cid", ["t0_1_2_multiple_positions", "t6_7_8_at_0", "haploid_triploid_deletion"])
def test_max_alleles_count_operator(exe, cid, tmp_path):
    """MaxAllelesCountOperator over scan_and_operate's per-site summaries: with top_count above the number of sites every site is
    reported -- column, merged REF, number of merged ALT alleles, the ALT alleles -- as the reference's dump has them"""
    from oracle import oracle
    case = [c for c in ALL if c["id"] == cid][0]
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "top.txt"
    p = subprocess.run([exe, "-j", str(qf), "-o", str(out), "--max-alleles", "1000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = out.read_text().splitlines()
    _, sites = oracle.parse_dump(open(os.path.join(GOLDEN, case["dump"]), "rb").read())
    assert lines[0] == "TOTAL %d" % len(sites)
    want = sorted(",".join([str(s["begin"]), s["alleles"][0], str(len(s["alleles"]) - 1)] + list(s["alleles"][1:])) for s in sites)
    assert sorted(lines[1:]) == want
    counts = [int(l.split(",")[2]) for l in lines[1:]]
    assert counts == sorted(counts)          # the reference pops its min-heap: fewest ALT alleles first


@pytest.mark.parametrize("cid", ["t0_1_2_combined_at_0", "haploid_triploid_deletion_produce_GT", "t0_1_2_all_asa_at_0", "info_ops1_at_0"])
def test_broad_gvcf_across_front_windows(exe, cid, tmp_path):
    """one site per front window (GDBGPU_FRONT_BYTES = 1): K6's collation, the INFO combines and the records read every batch's
    rows of a window that is recomputed for the next site -- the file is the reference's all the same"""
    case = [c for c in ALL if c["id"] == cid][0]
    gold = open(os.path.join(GOLDEN, "bcf", cid + ".bcf"), "rb").read()
    p, out = run_broad(exe, case, tmp_path, env={"GDBGPU_FRONT_BYTES": "1"})
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold


