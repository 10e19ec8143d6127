"""GPU: the boundary without anything handed over from a reference run.  The BCF header is generated from the configuration
(template header + queried fields + contigs of the vid mapping + samples of the callset mapping), reference bases come from an
indexed (bgzip) FASTA, workspace / array / mappings / the rank's column partition from the LOADER's JSON, and the Java
reader's query stream is pulled through the C ABI (gdbgpu_stream_*, one call per JNI method of GenomicsDBQueryStream).
Everything is compared with the files the unmodified reference wrote (tests/golden/bcf*)."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, golden_cases

pytestmark = pytest.mark.gpu
ALL = golden_cases()
INPUTS = os.path.join(GOLDEN, "inputs")


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    g.build_gpu_lib()
    return g.build_tools()


def query_json(case, tmp_path, sites_only=False, with_storage=True, ranges=True):
    q = {"attributes": case["attributes"],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}],
         "produce_GT_field": case["produce_GT_field"], "sites_only_query": sites_only, "produce_FILTER_field": case.get("produce_FILTER_field", False),
         "produce_GT_with_min_PL_value_for_spanning_deletions": case["produce_GT_with_min_PL_value_for_spanning_deletions"],
         "vcf_header_filename": os.path.join(INPUTS, "template_vcf_header.vcf"),
         "reference_genome": os.path.join(INPUTS, "ref_sparse.fasta.gz")}
    if ranges:
        q["query_column_ranges"] = [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}]
    if with_storage:
        q.update({"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
                  "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]),
                  "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"])})
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    return str(qf)


def loader_json(case, tmp_path, partitions=None):
    ld = {"row_based_partitioning": False, "size_per_column_partition": 3000,
          "column_partitions": partitions or [{"begin": 0, "workspace": os.path.join(GOLDEN, case["workspace"]), "array": case["array"]}],
          "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]),
          "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"])}
    lf = tmp_path / "loader.json"
    lf.write_text(json.dumps(ld))
    return str(lf)


@pytest.mark.parametrize("case", ALL, ids=lambda c: c["id"])
def test_whole_bcf_with_generated_header(exe, case, tmp_path):
    """no --bcf-header, no --reference-bases: header lines, dictionary indices, contigs, samples and the N -> reference base
    substitutions all come from this library; the file equals the reference's byte for byte"""
    gold = open(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"), "rb").read()
    out = tmp_path / "out.bcf"
    p = subprocess.run([exe, "-j", query_json(case, tmp_path), "-o", str(out), "--produce-Broad-GVCF"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    got = out.read_bytes()
    assert got[:9 + int.from_bytes(gold[5:9], "little")] == gold[:9 + int.from_bytes(gold[5:9], "little")], "header differs"
    assert got == gold


@pytest.mark.parametrize("cid", ["t0_1_2_at_0", "info_ops0_at_0", "t0_1_2_at_0_FILTER"])
def test_sites_only_with_generated_header(exe, cid, tmp_path):
    case = [c for c in ALL if c["id"] == cid][0]
    gold = open(os.path.join(GOLDEN, "bcf_sites_only", cid + ".bcf"), "rb").read()
    out = tmp_path / "out.bcf"
    p = subprocess.run([exe, "-j", query_json(case, tmp_path, sites_only=True), "-o", str(out), "--produce-Broad-GVCF"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold


@pytest.mark.parametrize("cid", ["t0_1_2_at_0", "t6_7_8_at_0"])
def test_loader_json_supplies_storage_and_mappings(exe, cid, tmp_path):
    """gt_mpi_gather -l: the query file names no workspace / array / vid / callsets; the loader's column partition does"""
    case = [c for c in ALL if c["id"] == cid][0]
    gold = open(os.path.join(GOLDEN, "bcf", cid + ".bcf"), "rb").read()
    out = tmp_path / "out.bcf"
    p = subprocess.run([exe, "-l", loader_json(case, tmp_path), "-j", query_json(case, tmp_path, with_storage=False), "-o", str(out), "--produce-Broad-GVCF"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == gold


def records_of(bcf):
    """(header bytes, [record bytes]) of an uncompressed BCF2 stream"""
    l_text = int.from_bytes(bcf[5:9], "little")
    pos, recs = 9 + l_text, []
    while pos < len(bcf):
        n = 8 + int.from_bytes(bcf[pos:pos + 4], "little") + int.from_bytes(bcf[pos + 4:pos + 8], "little")
        recs.append(bcf[pos:pos + n])
        pos += n
    return bcf[:9 + l_text], recs


def test_rank_partitions_leave_the_query_whole(exe, tmp_path):
    """two ranks, two column partitions of one array.  subset_query_column_ranges_based_on_partition edits the base
    configuration's list of ranges, and both of its callers (gt_mpi_gather.cc:554-557, genomicsdb_bcf_generator.cc:51-53) run it
    after read_from_file has handed every range to the query -- in the unmodified reference each rank therefore produces the
    WHOLE query (tests/test_query_config_cpu.py holds the two lists side by side for both), and so does each rank here"""
    case = [c for c in ALL if c["id"] == "t0_1_2_multiple_positions"][0]
    gold = open(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"), "rb").read()
    intervals = sorted(case["query_column_ranges"])
    assert len(intervals) >= 2
    cut = intervals[len(intervals) // 2][0]
    parts = [{"begin": 0, "workspace": os.path.join(GOLDEN, case["workspace"]), "array": case["array"]},
             {"begin": cut, "workspace": os.path.join(GOLDEN, case["workspace"]), "array": case["array"]}]
    lf = loader_json(case, tmp_path, parts)
    qf = query_json(case, tmp_path, with_storage=False)
    for rank in (0, 1):
        out = tmp_path / ("out%d.bcf" % rank)
        p = subprocess.run([exe, "-l", lf, "-j", qf, "-r", str(rank), "-o", str(out), "--produce-Broad-GVCF"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert out.read_bytes() == gold


def test_reference_command_line_and_output_destination(exe, tmp_path):
    """the reference tool's own spellings (--json-config, -O bu) and its destinations: the Broad GVCF goes to stdout, or to the
    configuration's "vcf_output_filename" (json_config.cc:577-617), when no -o is given"""
    case = [c for c in ALL if c["id"] == "t0_1_2_at_12150"][0]
    gold = open(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"), "rb").read()
    qf = query_json(case, tmp_path)
    p = subprocess.run([exe, "--json-config", qf, "--produce-Broad-GVCF", "-O", "bu", "-s", "1048576"], capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")
    assert p.stdout == gold
    q = json.load(open(qf))
    q["vcf_output_filename"] = str(tmp_path / "named_in_the_query.bcf")
    json.dump(q, open(qf, "w"))
    p = subprocess.run([exe, "--json-config", qf, "--produce-Broad-GVCF"], capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")
    assert open(q["vcf_output_filename"], "rb").read() == gold   # (stdout then carries the tool's own one-line run summary)
    p = subprocess.run([exe, "--json-config", qf, "--print-calls"], capture_output=True)
    assert p.returncode == 0 and p.stdout.startswith(b"{")


# ---------------------------------------------------------------------------------------------- the C ABI of the query stream
@pytest.mark.parametrize("cid", ["t0_1_2_at_0", "t0_1_2_multiple_positions", "haploid_triploid_deletion_produce_GT"])
@pytest.mark.parametrize("read_size,capacity", [(7, 64), (1 << 16, 1 << 20)], ids=["7B-reads", "64kB-reads"])
def test_query_stream_c_abi(cid, read_size, capacity, tmp_path):
    """gdbgpu_stream_open / read / read_next_byte / skip / close over ctypes -- the calls the JNI layer makes for
    GenomicsDBQueryStream -- with the loader's JSON: the stream is the reference's BCF as its JNI path writes it"""
    from genomicsdb_b200 import capi
    case = [c for c in ALL if c["id"] == cid][0]
    gold = open(os.path.join(GOLDEN, "bcf_java", cid + ".bcf"), "rb").read()
    s = capi.QueryStream(loader_json(case, tmp_path), query_json(case, tmp_path, with_storage=False), buffer_capacity=capacity,
                         use_missing_values_only_not_vector_end=True)
    assert s.num_bytes_available() == capacity
    got = bytearray()
    k = 0
    while True:
        k += 1
        if k % 5 == 0:
            b = s.read_next_byte()
            if b == -1 and len(got) >= len(gold):
                break
            got.append(b & 0xFF)
            continue
        chunk = s.read(read_size)
        if not chunk:
            break
        got += chunk
    s.close()
    assert bytes(got) == gold


def test_query_stream_header_only_and_skip(tmp_path):
    from genomicsdb_b200 import capi
    case = [c for c in ALL if c["id"] == "t0_1_2_at_0"][0]
    gold = open(os.path.join(GOLDEN, "bcf_java", "t0_1_2_at_0.bcf"), "rb").read()
    hdr_len = 9 + int.from_bytes(gold[5:9], "little")
    s = capi.QueryStream(loader_json(case, tmp_path), query_json(case, tmp_path, with_storage=False), produce_header_only=True,
                         use_missing_values_only_not_vector_end=True)
    got = b""
    while True:
        c = s.read(4096)
        if not c:
            break
        got += c
    s.close()
    assert got == gold[:hdr_len]
    s = capi.QueryStream(loader_json(case, tmp_path), query_json(case, tmp_path, with_storage=False), use_missing_values_only_not_vector_end=True)
    assert s.skip(hdr_len) == hdr_len
    rest = b""
    while True:
        c = s.read(1 << 16)
        if not c:
            break
        rest += c
    s.close()
    assert rest == gold[hdr_len:]


def test_query_stream_contig_interval(tmp_path):
    """chr / start / end (1-based VCF positions) replace the query file's column ranges (genomicsdb_bcf_generator.cc:57-66);
    the file still has to name some: read_from_file refuses a query without (variant_query_config.cc:253)"""
    from genomicsdb_b200 import capi
    case = [c for c in ALL if c["id"] == "t0_1_2_at_12150"][0]
    whole = [c for c in ALL if c["id"] == "t0_1_2_at_0"][0]
    gold = open(os.path.join(GOLDEN, "bcf_java", "t0_1_2_at_12150.bcf"), "rb").read()
    lo, hi = case["query_column_ranges"][0]
    s = capi.QueryStream(loader_json(whole, tmp_path), query_json(whole, tmp_path, with_storage=False), chr="1", start=lo + 1,
                         end=min(hi, 249250620) + 1, use_missing_values_only_not_vector_end=True)
    got = b""
    while True:
        c = s.read(1 << 16)
        if not c:
            break
        got += c
    s.close()
    assert got == gold


def test_query_stream_reports_errors(tmp_path):
    from genomicsdb_b200 import capi
    with pytest.raises(capi.GdbGpuError):
        capi.QueryStream("", str(tmp_path / "missing.json"))
