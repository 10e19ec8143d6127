"""CPU: the vid mapping and the callset mapping as the mirror's VidMapper holds them, against the unmodified reference's
FileBasedVidMapper (oracle/_ref/bin/gdb_ref_dump --mode vid; vid_mapper.cc:686-711,841-1055,1209-1528, known_field_info.cc:239-308):
every spelling of a type and a length (multi-dimensional fields with their "vcf_delimiter" given: without one the reference
separates with NUL bytes, the mirror with '|' and ','), lists and dictionaries, the lengths of known fields the file leaves out, the field
indices (FORMAT twins and flattened tuple elements count -- FILTER cells hold these indices), contigs, callsets, the contig a
column falls into -- and the files the reference refuses."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, REPO

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")
CONTIGS = {"1": {"length": 1000, "tiledb_column_offset": 0}, "2": {"length": 500, "tiledb_column_offset": 1000}}


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path_factory.mktemp("probe") / "vid_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "vid_probe.cc"), "-o", exe,
                           "-I" + os.path.join(REPO, "include"), "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def run_both(probe, tmp_path, vid, callsets=None):
    vf = tmp_path / "vid.json"
    vf.write_text(vid if isinstance(vid, str) else json.dumps(vid))
    ref_args, our_args = [REF, "--mode", "vid", "-j", str(vf)], [probe, str(vf)]
    if callsets is not None:
        cf = tmp_path / "callsets.json"
        cf.write_text(json.dumps(callsets))
        ref_args += ["-l", str(cf)]
        our_args.append(str(cf))
    return subprocess.run(ref_args, capture_output=True, text=True), subprocess.run(our_args, capture_output=True, text=True)


def same(ref, ours):
    assert ref.returncode == 0, ref.stderr[-600:]
    assert ours.returncode == 0, ours.stderr[-600:]
    want, got = json.loads(ref.stdout), json.loads(ours.stdout)
    names = [f["name"] for f in got["fields"]]
    # the mirror lists the fields the file names (with their twins and flattened elements); the mandatory fields the reference
    # appends (END, REF, ALT, QUAL, FILTER) are counted, not listed
    assert [f for f in want["fields"] if f["name"] in names] == got["fields"]
    assert set(f["name"] for f in want["fields"]) - set(names) <= {"END", "REF", "ALT", "QUAL", "FILTER"}
    assert len(want["fields"]) == got["num_fields"]
    for k in ("contigs", "callsets", "locate"):
        assert want[k] == got[k], k
    return got


@pytest.mark.parametrize("vid", sorted(f for f in os.listdir(os.path.join(GOLDEN, "inputs")) if f.startswith("vid") and f.endswith(".json")))
def test_golden_vid_mappings(probe, tmp_path, vid):
    same(*run_both(probe, tmp_path, open(os.path.join(GOLDEN, "inputs", vid)).read(), json.load(open(os.path.join(GOLDEN, "inputs", "callsets", "t0_1_2.json")))))


FIELDS_SPELLINGS = {
    "a": {"type": "Integer", "length": "r", "vcf_field_class": ["FORMAT"]},
    "b": {"type": "Float", "length": "BCF_VL_G", "vcf_field_class": ["FORMAT"]},
    "c": {"type": "String", "length": "var", "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "concatenate"},
    "d": {"type": "Boolean", "vcf_field_class": ["INFO"]},
    "e": {"type": "Flag", "length": 1, "vcf_field_class": ["INFO"]},
    "f": {"type": "Char", "length": "Phased_Ploidy", "vcf_field_class": ["FORMAT"]},
    "g": {"type": "int", "length": "pp", "vcf_field_class": ["FORMAT"]},
    "h": {"type": "integer", "length": "3", "vcf_field_class": ["FORMAT"]},
    "i": {"type": "Int", "length": "0x10", "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "sum"},
    "j": {"type": "float", "length": {"variable_length_descriptor": "A"}, "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "elementwise_sum"},
    "k": {"type": "int", "length": {"fixed_length": 2}, "vcf_field_class": ["FORMAT"]},
    "l": {"type": "int", "length": ["R"], "vcf_field_class": ["FORMAT"]},
    "m": {"type": "int", "length": "no_such_descriptor", "vcf_field_class": ["FORMAT"]},
    "n": {"type": "int", "length": "BCF_VL_FIXED", "vcf_field_class": ["FORMAT"]},
    "o": {"type": "int", "length": "P", "vcf_field_class": ["FORMAT"]},
    "p": {"type": "bool", "length": "VAR"},
    "q": {"type": "float", "length": ["R", {"fixed_length": 2}], "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "element_wise_sum", "vcf_delimiter": ["|", ","]},
    "r": {"type": ["float", "int"], "length": ["A", "var"], "vcf_type": "String", "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "histogram_sum", "vcf_delimiter": [";", ":"]},
}
KNOWN_WITHOUT_LENGTH = {n: {"type": t, "vcf_field_class": c} for n, t, c in [
    ("PL", "int", ["FORMAT"]), ("AD", "int", ["FORMAT"]), ("GT", "int", ["FORMAT"]), ("SB", "int", ["FORMAT"]), ("AF", "float", ["INFO"]), ("AC", "int", ["INFO"]),
    ("PGT", "char", ["FORMAT"]), ("PID", "char", ["FORMAT"]), ("RAW_MQ", "float", ["INFO"]), ("MQ", "float", ["INFO"]), ("MQ0", "int", ["INFO"]), ("ExcessHet", "float", ["INFO"]),
    ("GQ", "int", ["FORMAT"]), ("MIN_DP", "int", ["FORMAT"]), ("DP", "int", ["INFO", "FORMAT"]), ("END", "int", ["INFO"]), ("REF", "char", []), ("ALT", "char", []),
    ("FILTER", "int", []), ("QUAL", "float", []), ("ID", "char", []), ("not_a_known_field", "int", ["INFO"])]}
ACCEPTED = {
    "type_and_length_spellings": {"contigs": CONTIGS, "fields": FIELDS_SPELLINGS},
    "known_fields_without_length": {"contigs": CONTIGS, "fields": KNOWN_WITHOUT_LENGTH},
    "fields_as_a_list": {"contigs": CONTIGS, "fields": [dict(v, name=n) for n, v in FIELDS_SPELLINGS.items()]},
    "fields_as_a_list_of_field_name": {"contigs": CONTIGS, "fields": [dict(v, field_name=n) for n, v in KNOWN_WITHOUT_LENGTH.items()]},
    "contigs_as_lists_in_any_order": {"contigs": [{"contig_name": "z", "length": 10, "tiledb_column_offset": 5000}, {"name": "y", "length": 100, "tiledb_column_offset": 0},
                                                  {"chromosome_name": "x", "length": 900, "tiledb_column_offset": 100}, {"name": "empty", "length": 0, "tiledb_column_offset": 3000}],
                                      "fields": {"GT": {"type": "int", "vcf_field_class": ["FORMAT"]}}},
    # FILTER cells hold field indices: the twin of DP and the flattened elements of a tuple field move everything behind them
    "filters_behind_a_twin_and_a_tuple_field": {"contigs": CONTIGS, "fields": {
        "DP": {"type": "int", "vcf_field_class": ["INFO", "FORMAT"]},
        "PASS": {"type": "int", "vcf_field_class": ["FILTER"]},
        "H": {"type": ["float", "int"], "length": ["R", "var"], "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "histogram_sum", "vcf_delimiter": ["|", ","]},
        "LowQual": {"type": "int", "vcf_field_class": ["FILTER"]},
        "HF": {"type": ["float", "int"], "length": ["R", "var"], "vcf_field_class": ["INFO", "FORMAT"], "VCF_field_combine_operation": "histogram_sum", "vcf_delimiter": ["|", ","]},
        "q10": {"type": "int", "vcf_field_class": ["FILTER"]}}},
    "callsets_in_the_vid_mapping": {"contigs": CONTIGS, "fields": {"GT": {"type": "int", "vcf_field_class": ["FORMAT"]}},
                                    "callsets": {"s1": {"row_idx": 1}, "s0": {"row_idx": 0, "idx_in_file": 0, "filename": "a.vcf"}, "s2": {"row_idx": 2}}},
    "protobuf_callset_mapping_in_the_vid_mapping": {"contigs": CONTIGS, "fields": {"GT": {"type": "int", "vcf_field_class": ["FORMAT"]}},
                                                    "callset_mapping": {"callsets": [{"sample_name": "s1", "row_idx": 1}, {"callset_name": "s0", "row_idx": 0}, {"name": "s2", "row_idx": 2}]}},
}
GT_ONLY = {"GT": {"type": "int", "vcf_field_class": ["FORMAT"]}}
REFUSED = {
    "no_contigs": {"fields": GT_ONLY},
    "no_fields": {"contigs": CONTIGS},
    "field_without_type": {"contigs": CONTIGS, "fields": {"X": {"length": 1}}},
    "unknown_type": {"contigs": CONTIGS, "fields": {"X": {"type": "double"}}},
    "type_of_another_kind": {"contigs": CONTIGS, "fields": {"X": {"type": 5}}},
    "both_name_keys": {"contigs": CONTIGS, "fields": [{"name": "X", "field_name": "X", "type": "int"}]},
    "no_name_key": {"contigs": CONTIGS, "fields": [{"type": "int"}]},
    "the_same_field_twice": {"contigs": CONTIGS, "fields": [{"name": "X", "type": "int"}, {"name": "X", "type": "int"}]},
    "the_same_contig_twice": {"contigs": [{"name": "1", "length": 5, "tiledb_column_offset": 0}, {"name": "1", "length": 5, "tiledb_column_offset": 10}], "fields": GT_ONLY},
    "two_contig_name_keys": {"contigs": [{"name": "1", "contig_name": "1", "length": 5, "tiledb_column_offset": 0}], "fields": GT_ONLY},
    "overlapping_contigs": {"contigs": {"1": {"length": 1000, "tiledb_column_offset": 0}, "2": {"length": 500, "tiledb_column_offset": 999}}, "fields": GT_ONLY},
    "negative_offset": {"contigs": {"1": {"length": 1000, "tiledb_column_offset": -1}}, "fields": GT_ONLY},
    "contig_without_length": {"contigs": {"1": {"tiledb_column_offset": 0}}, "fields": GT_ONLY},
    "unknown_combine_operation": {"contigs": CONTIGS, "fields": {"X": {"type": "int", "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "maximum"}}},
    "concatenate_of_a_fixed_length_field": {"contigs": CONTIGS, "fields": {"X": {"type": "char", "length": 1, "vcf_field_class": ["INFO"], "VCF_field_combine_operation": "concatenate"}}},
    "multi_dimensional_field_of_characters": {"contigs": CONTIGS, "fields": {"X": {"type": "char", "length": ["R", "var"], "vcf_field_class": ["INFO"]}}},
    "length_of_another_kind": {"contigs": CONTIGS, "fields": {"X": {"type": "int", "length": {"neither": 1}}}},
    "syntax_error": "{ \"contigs\": ",
}
CALLSETS_ACCEPTED = {
    "dictionary": {"callsets": {"s2": {"row_idx": 2}, "s0": {"row_idx": 0}, "s1": {"row_idx": 1, "idx_in_file": 1, "filename": "x.vcf.gz"}}},
    "list_with_three_name_keys": {"callsets": [{"sample_name": "s1", "row_idx": 1}, {"callset_name": "s0", "row_idx": 0}, {"name": "s2", "row_idx": 2, "stream_name": "st"}]},
    "the_same_entry_twice": {"callsets": [{"name": "s0", "row_idx": 0}, {"name": "s0", "row_idx": 0}, {"name": "s1", "row_idx": 1}]},
}
CALLSETS_REFUSED = {
    "no_callsets": {"samples": {}},
    "no_row_idx": {"callsets": {"s0": {"idx_in_file": 0}}},
    "a_row_without_a_name": {"callsets": {"s0": {"row_idx": 0}, "s2": {"row_idx": 2}}},
    "one_row_two_names": {"callsets": {"s0": {"row_idx": 0}, "t0": {"row_idx": 0}}},
    "one_name_two_rows": {"callsets": [{"name": "s0", "row_idx": 0}, {"name": "s0", "row_idx": 1}]},
    "two_name_keys": {"callsets": [{"name": "s0", "sample_name": "s0", "row_idx": 0}]},
    "no_name_key": {"callsets": [{"row_idx": 0}]},
    "filename_and_stream_name": {"callsets": {"s0": {"row_idx": 0, "filename": "a", "stream_name": "b"}}},
}


@pytest.mark.parametrize("name", list(ACCEPTED))
def test_vid_mapping_as_the_reference_holds_it(probe, tmp_path, name):
    got = same(*run_both(probe, tmp_path, ACCEPTED[name]))
    if name == "filters_behind_a_twin_and_a_tuple_field":
        idx = {f["name"]: f["idx"] for f in got["fields"]}
        assert (idx["DP"], idx["DP_FORMAT"], idx["PASS"], idx["H"], idx["H_tuple_element_1"], idx["LowQual"], idx["HF"], idx["HF_FORMAT"],
                idx["HF_tuple_element_0"], idx["HF_FORMAT_tuple_element_1"], idx["q10"]) == (0, 1, 2, 3, 5, 6, 7, 8, 9, 12, 13)


@pytest.mark.parametrize("name", list(REFUSED))
def test_vid_mappings_the_reference_refuses(probe, tmp_path, name):
    ref, ours = run_both(probe, tmp_path, REFUSED[name])
    assert ref.returncode != 0, ref.stdout
    assert ours.returncode == 1 and ours.stderr.strip(), (ours.returncode, ours.stdout, ours.stderr)


@pytest.mark.parametrize("name", list(CALLSETS_ACCEPTED))
def test_callset_mapping_as_the_reference_holds_it(probe, tmp_path, name):
    got = same(*run_both(probe, tmp_path, {"contigs": CONTIGS, "fields": GT_ONLY}, CALLSETS_ACCEPTED[name]))
    assert got["callsets"][:2] == ["s0", "s1"]


@pytest.mark.parametrize("name", list(CALLSETS_REFUSED))
def test_callset_mappings_the_reference_refuses(probe, tmp_path, name):
    ref, ours = run_both(probe, tmp_path, {"contigs": CONTIGS, "fields": GT_ONLY}, CALLSETS_REFUSED[name])
    assert ref.returncode != 0, ref.stdout
    assert ours.returncode == 1 and ours.stderr.strip(), (ours.returncode, ours.stdout, ours.stderr)
