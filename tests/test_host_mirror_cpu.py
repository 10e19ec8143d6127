"""CPU: the host mirror's parsing logic (BCF header dictionary, vid mapping contigs / combine operations, query JSON) through a
small C++ probe linked against libgdbgpu.so -- no CUDA call is made -- against the same facts read by the test-side readers."""
import json
import os
import subprocess

from conftest import GOLDEN, REPO, golden_cases
from oracle import bcf_reader


def build_probe(tmp_path):
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path / "host_api_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "host_api_probe.cc"), "-o", exe,
                           "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_header_dictionary_vid_mapping_and_query_json(tmp_path):
    exe = build_probe(tmp_path)
    case = [c for c in golden_cases() if c["id"] == "t0_1_2_at_12100_12141_to_12150_12300_17384"][0]
    bcf = os.path.join(GOLDEN, "bcf", case["id"] + ".bcf")
    q = {"workspace": "/nowhere", "array_name": "x", "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]),
         "attributes": case["attributes"], "produce_GT_field": True,
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in reversed(case["query_column_ranges"])]}],
         "query_row_ranges": [{"range_list": [{"low": 0, "high": 2}, 7]}]}
    qf = tmp_path / "q.json"
    qf.write_text(json.dumps(q))
    out = subprocess.run([exe, bcf, os.path.join(GOLDEN, case["vid_mapping_file"]), str(qf)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout)
    text, _, _ = bcf_reader.read_bcf(bcf)
    ids = bcf_reader.header_dict(text)
    for name, v in got["ids"].items():
        assert v == ids.get(name, -1), name
    assert got["ids"]["PASS"] == 0 and got["ids"]["no_such_field"] == -1
    assert (got["contig_1"], got["contig_2"]) == (0, 1)
    assert (got["contig_of_col"], got["contig_offset"]) == ("2", 249250621)
    assert (got["PL_length"], got["MQ_op"], got["DP_FORMAT_is_format"]) == ("G", "median", 1)
    assert got["intervals"] == sorted(case["query_column_ranges"])          # sorted like the reference (variant_query_config.cc:69-73)
    assert got["n_rows"] == 4 and got["produce_GT"] == 1 and got["n_attributes"] == len(case["attributes"])


def test_vid_mapping_of_multidimensional_and_tuple_fields(tmp_path):
    """length / type / vcf_delimiter given as lists (vid_mapper.cc, genomicsdb_multid_vector_field.h): a 2-D field keeps its
    per-dimension descriptors and delimiters; a field of tuples is composite and comes with one flattened attribute per tuple
    element (<name>_tuple_element_<j>, as the array's schema names them) that goes by its parent's VCF name"""
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path / "vid_2d_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "vid_2d_probe.cc"), "-o", exe,
                           "-I" + os.path.join(REPO, "include"), "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    names = ["AS_RAW_MQ", "AS_RAW_MQRankSum", "AS_RAW_MQRankSum_tuple_element_0", "AS_RAW_MQRankSum_tuple_element_1", "PL", "no_such"]
    out = subprocess.run([exe, os.path.join(GOLDEN, "inputs", "vid_all_asa.json")] + names, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout)
    mq = got["AS_RAW_MQ"]
    assert (mq["multid"], mq["composite"], mq["type"], mq["dims"], mq["delims"], mq["op"], mq["dim0_allele"], mq["allele_dependent"]) == \
           (1, 0, "float", ["R", "VAR"], "|,", "element_wise_sum", 1, 0)
    rs = got["AS_RAW_MQRankSum"]
    assert (rs["multid"], rs["composite"], rs["op"], rs["vcf_name"]) == (1, 1, "histogram_sum", "AS_RAW_MQRankSum")
    e0, e1 = got["AS_RAW_MQRankSum_tuple_element_0"], got["AS_RAW_MQRankSum_tuple_element_1"]
    assert (e0["type"], e0["parent"], e0["tuple_index"], e0["vcf_name"], e0["composite"], e0["multid"]) == ("float", "AS_RAW_MQRankSum", 0, "AS_RAW_MQRankSum", 0, 1)
    assert (e1["type"], e1["parent"], e1["tuple_index"], e1["dims"], e1["delims"]) == ("int", "AS_RAW_MQRankSum", 1, ["R", "VAR"], "|,")
    assert got["PL"]["multid"] == 0 and got["PL"]["allele_dependent"] == 1 and got["no_such"] is None
