"""CPU: the query configuration as the mirror's VariantQueryConfig holds it, against the unmodified reference's own classes
(oracle/_ref/bin/gdb_ref_dump --mode config: JSONConfigBase / VariantQueryConfig::read_from_file, json_config.cc:195-656,
variant_query_config.cc:234-278, after the calls gt_mpi_gather.cc:547-557 makes).  Every accepted form of a column range (list,
position, whole contig, contig position / interval, the protobuf spellings), per-rank lists, partitions given in the query
file, the loader's file beside it -- and the configurations the reference refuses, which the mirror refuses too."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, REPO

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")
VID = os.path.join(GOLDEN, "inputs", "vid.json")
CALLSETS = os.path.join(GOLDEN, "inputs", "callsets", "t0_1_2.json")
C2 = 249250621   # tiledb_column_offset of contig "2" in tests/golden/inputs/vid.json


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path_factory.mktemp("probe") / "config_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "config_probe.cc"), "-o", exe,
                           "-I" + os.path.join(REPO, "include"), "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


BASE = {"workspace": "/ws", "array": "arr", "vid_mapping_file": VID, "callset_mapping_file": CALLSETS,
        "attributes": ["REF", "ALT", "GT", "PL"]}


def q(**kw):
    d = dict(BASE)
    for k, v in kw.items():
        if v is None:
            d.pop(k, None)
        else:
            d[k] = v
    return d


# name -> (query JSON, loader JSON or None, ranks to try)
ACCEPTED = {
    "intervals_positions_unsorted": (q(query_column_ranges=[[[700, 1500], 350, [100, 900], [100, 100], 5]]), None, (0,)),
    "reversed_interval_is_swapped": (q(query_column_ranges=[[[1500, 700], [9, 3]]], query_row_ranges=[[[5, 2], 0]]), None, (0,)),
    "one_list_serves_every_rank": (q(query_column_ranges=[[[0, 5], 45]]), None, (0, 3)),
    "a_list_per_rank": (q(query_column_ranges=[[[0, 5], 45], [76, 87], [[1000, 2000]]], query_row_ranges=[[0], [1, 2], [[0, 2]]]), None, (0, 1, 2)),
    "whole_contig_by_name": (q(query_column_ranges=[["2", [10, 20], "1"]]), None, (0,)),
    "contig_position_and_interval": (q(query_column_ranges=[[{"2": 100}, {"1": [500, 400]}, {"2": [243199000, 243199373]}]]), None, (0,)),
    "contig_interval_cut_at_the_contig_end": (q(query_column_ranges=[[{"2": [243199000, 243299999]}]]), None, (0,)),
    "protobuf_low_high": (q(query_column_ranges=[{"range_list": [{"low": 12, "high": 900}, {"low": 7, "high": 7}]}],
                            query_row_ranges=[{"range_list": [{"low": 0, "high": 1}, {"low": 2, "high": 2}]}]), None, (0,)),
    "protobuf_column_or_interval_list": (q(query_column_ranges=[{"column_or_interval_list": [
        {"column_interval": {"column_interval": {"begin": 50, "end": 70}}},
        {"column_interval": {"contig_interval": {"contig": "2", "begin": 11, "end": 20}}},
        {"column": {"tiledb_column": 4}},
        {"column": {"contig_position": {"contig": "2", "position": 1}}},
        {"contig_position": {"contig": "1", "position": 33}}]}]), None, (0,)),
    "empty_protobuf_entry_for_a_rank": (q(query_column_ranges=[{}, {"range_list": [{"low": 1, "high": 2}]}]), None, (0, 1)),
    "scan_full_beside_ranges": (q(scan_full=True, query_column_ranges=[[[0, 5]]], query_row_ranges=[[1]]), None, (0,)),
    "scan_full_false_is_still_scan_full": (q(scan_full=False), None, (0,)),
    "rows_only": (q(query_column_ranges=None, query_row_ranges=[[[0, 1]]], scan_full=None), None, ()),   # refused: see REFUSED
    "rows_overlapping_and_past_the_array": (q(query_column_ranges=[[[0, 100]]], query_row_ranges=[[[1, 3], 2, [7, 9]]]), None, (0,)),
    "column_partitions_in_the_query": (q(column_partitions=[{"begin": 5000}, {"begin": 0}, {"begin": 9000, "end": 9500}]), None, (0, 1, 2)),
    "column_partitions_contig_bounds": (q(column_partitions=[{"begin": {"1": 100}, "end": {"2": 7}}, {"begin": {"2": [50, 60]}}]), None, (0, 1)),
    "column_partitions_name_their_storage": (q(workspace=None, array=None, column_partitions=[{"begin": 0, "workspace": "/w0", "array": "a0"}, {"begin": 100, "workspace": "/w1", "array_name": "a1"}]), None, (0, 1)),
    "column_partition_overrides_one_rank": (q(column_partitions=[{"begin": 0}, {"begin": 100, "workspace": "/w1", "array": "a1"}]), None, (0, 1)),
    "row_partitions_in_the_query": (q(row_partitions=[{"begin": 2}, {"begin": 0}]), None, ()),   # refused (no column ranges)
    "row_partitions_with_columns": (q(query_column_ranges=[[[0, 100]]], row_partitions=[{"begin": 2, "end": 3}, {"begin": 0, "end": 5}, {"begin": 9, "end": 7}]), None, (0, 1, 2)),
    "per_rank_workspace_and_array": (q(workspace=["/w0", "/w1"], array=None, array_name=["a0", "a1"], query_column_ranges=[[1]]), None, (0, 1)),
    "per_rank_files": (q(query_column_ranges=[[1]], vcf_header_filename=["h0.vcf", "h1.vcf"], reference_genome=["r0.fa", "r1.fa"], vcf_output_filename=["o0.bcf", "o1.bcf"]), None, (0, 1)),
    "output_file": (q(query_column_ranges=[[1]], vcf_output_filename="out.bcf"), None, (0,)),
    "output_file_of_the_rank_s_partition": (q(vcf_output_filename="all.bcf", column_partitions=[{"begin": 0, "vcf_output_filename": "p0.bcf"}, {"begin": 100}]), None, (0, 1)),
    "options": (q(query_column_ranges=[[1]], segment_size=4096, max_diploid_alt_alleles_that_can_be_genotyped=3, produce_GT_field=True, produce_FILTER_field=True,
                  sites_only_query=True, produce_GT_with_min_PL_value_for_spanning_deletions=True, vcf_header_filename="h.vcf", reference_genome="r.fa"), None, (0,)),
    "query_attributes_spelling_and_a_repeated_name": (q(attributes=None, query_attributes=["PL", "REF", "PL", "GT", "REF"], query_column_ranges=[[1]]), None, (0,)),
    "no_attributes": (q(attributes=None, query_column_ranges=[[1]]), None, (0,)),
    "loader_supplies_storage_and_mappings": (q(workspace=None, array=None, vid_mapping_file=None, callset_mapping_file=None, query_column_ranges=[[[0, 50], [4000, 4100], 7000]]),
                                             {"row_based_partitioning": False, "column_partitions": [{"begin": 0, "workspace": "/lw0", "array": "la0"}, {"begin": 4050, "workspace": "/lw1", "array": "la1"}],
                                              "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, (0, 1)),
    "one_loader_partition_serves_every_rank": (q(workspace=None, array=None, query_column_ranges=[[[0, 50], 9000]]),
                                               {"row_based_partitioning": False, "column_partitions": [{"begin": 100, "workspace": "/lw0", "array": "la0"}],
                                                "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, (0,)),
    "one_loader_partition_and_storage_named_outside_it": (q(workspace=None, array=None, query_column_ranges=[[[0, 50], 9000]]),
                                               {"row_based_partitioning": False, "column_partitions": [{"begin": 100}], "workspace": "/lw", "array": "la",
                                                "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, (0,)),
    "loader_partitions_unsorted_with_ends_and_contig_bounds": (q(query_column_ranges=[[[0, 50]]]),
        {"row_based_partitioning": False, "size_per_column_partition": 1000, "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "workspace": "/lw", "array": "la",
         "column_partitions": [{"begin": 5000, "end": 100000}, {"begin": 0}, {"begin": {"2": 100}, "end": {"2": [500, 600]}}, {"begin": 700, "end": 650}, {"begin": 100000, "end": 99999999999}]}, (0, 1, 2, 3, 4)),
    "query_overrides_the_loader": (q(query_column_ranges=[[[0, 50], 9000]]),
                                   {"row_based_partitioning": False, "column_partitions": [{"begin": 0, "workspace": "/lw0", "array": "la0"}, {"begin": 4050, "workspace": "/lw1", "array": "la1"}],
                                    "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, (0, 1)),
}
REFUSED = {
    "no_ranges_at_all": q(),
    "rows_only": ACCEPTED["rows_only"][0],
    "open_ended_row_partition": q(query_column_ranges=[[[0, 100]]], row_partitions=[{"begin": 2}, {"begin": 0, "end": 5}]),   # every row of a range is listed
    "row_partitions_only": ACCEPTED["row_partitions_in_the_query"][0],
    "ranges_and_partitions": q(query_column_ranges=[[1]], column_partitions=[{"begin": 0}]),
    "row_ranges_and_row_partitions": q(query_column_ranges=[[1]], query_row_ranges=[[1]], row_partitions=[{"begin": 0}]),
    "row_and_column_partitions": q(column_partitions=[{"begin": 0}], row_partitions=[{"begin": 0}]),
    "array_and_array_name": q(array_name="other", query_column_ranges=[[1]]),
    "attributes_and_query_attributes": q(query_attributes=["GT"], query_column_ranges=[[1]]),
    "flat_list_of_positions": q(query_column_ranges=[5, 7]),
    "interval_of_three": q(query_column_ranges=[[[1, 2, 3]]]),
    "unknown_contig": q(query_column_ranges=[["no_such_contig"]]),
    "unknown_contig_position": q(query_column_ranges=[[{"no_such_contig": 5}]]),
    "position_past_the_contig": q(query_column_ranges=[[{"2": 243199374}]]),
    "two_partitions_with_one_begin": q(column_partitions=[{"begin": 10}, {"begin": 10}]),
    "partition_without_begin": q(column_partitions=[{"end": 10}]),
    "no_workspace": q(workspace=None, query_column_ranges=[[1]]),
    "no_array": q(array=None, query_column_ranges=[[1]]),
    "no_vid_mapping": q(vid_mapping_file=None, query_column_ranges=[[1]]),
    "no_callset_mapping": q(callset_mapping_file=None, query_column_ranges=[[1]]),
    "syntax_error": "{ \"workspace\": ",
}
LOADER2 = {"row_based_partitioning": False, "column_partitions": [{"begin": 0, "workspace": "/lw0", "array": "la0"}, {"begin": 4050, "workspace": "/lw1", "array": "la1"}],
           "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}
LOADER_REFUSED = {   # (query, loader, rank)
    "rank_past_the_loader_s_partitions": (q(query_column_ranges=[[[0, 50]]]), LOADER2, 2),
    "rank_past_the_loader_s_storage": (q(workspace=None, array=None, query_column_ranges=[[[0, 50]]]), LOADER2, 2),
    "storage_named_inside_the_only_partition_serves_rank_0_alone": (q(workspace=None, array=None, query_column_ranges=[[[0, 50]]]),
        {"row_based_partitioning": False, "column_partitions": [{"begin": 100, "workspace": "/lw0", "array": "la0"}], "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, 5),
    "one_loader_partition_serves_rank_0_alone": (q(query_column_ranges=[[[0, 50]]]),
        {"row_based_partitioning": False, "column_partitions": [{"begin": 100}], "workspace": "/lw", "array": "la", "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, 1),
    "loader_without_partitions": (q(query_column_ranges=[[[0, 50]]]), {"row_based_partitioning": False, "vid_mapping_file": VID, "callset_mapping_file": CALLSETS, "size_per_column_partition": 1000}, 0),
    "loader_without_size_per_column_partition": (q(query_column_ranges=[[[0, 50]]]), {k: v for k, v in LOADER2.items() if k != "size_per_column_partition"}, 0),
}
RANK_REFUSED = {   # accepted for the ranks the lists cover, refused past them
    "a_list_per_rank": (ACCEPTED["a_list_per_rank"][0], 3),
    "per_rank_workspace_and_array": (ACCEPTED["per_rank_workspace_and_array"][0], 2),
    "per_rank_files": (ACCEPTED["per_rank_files"][0], 2),
    "column_partitions_in_the_query": (ACCEPTED["column_partitions_in_the_query"][0], 3),
    "one_column_partition_in_the_query": (q(column_partitions=[{"begin": 5}]), 1),
    "one_row_partition_in_the_query": (q(query_column_ranges=[[1]], row_partitions=[{"begin": 0, "end": 2}]), 1),
}


def run_both(probe, tmp_path, query, loader, rank):
    qf = tmp_path / "q.json"
    qf.write_text(query if isinstance(query, str) else json.dumps(query))
    args = ["-j", str(qf), "-r", str(rank)]
    if loader is not None:
        lf = tmp_path / "l.json"
        lf.write_text(json.dumps(loader))
        args = ["-l", str(lf)] + args
    ref = subprocess.run([REF] + args + ["--mode", "config"], capture_output=True, text=True)
    ours = subprocess.run([probe] + args, capture_output=True, text=True)
    return ref, ours


@pytest.mark.parametrize("name", [n for n, v in ACCEPTED.items() if v[2]])
def test_configuration_as_the_reference_holds_it(probe, tmp_path, name):
    query, loader, ranks = ACCEPTED[name]
    for rank in ranks:
        ref, ours = run_both(probe, tmp_path, query, loader, rank)
        assert ref.returncode == 0, (rank, ref.stderr[-600:])
        assert ours.returncode == 0, (rank, ours.stderr[-600:])
        want = json.loads(ref.stdout)
        want.pop("vcf_output_format")   # the tool sets it after the file is read (gt_mpi_gather.cc:561-562)
        got = json.loads(ours.stdout)
        assert got == want, (rank, got, want)


@pytest.mark.parametrize("name", list(REFUSED))
def test_configurations_the_reference_refuses(probe, tmp_path, name):
    ref, ours = run_both(probe, tmp_path, REFUSED[name], None, 0)
    assert ref.returncode != 0, ref.stdout
    assert ours.returncode == 1 and ours.stderr.strip(), (ours.returncode, ours.stdout, ours.stderr)


@pytest.mark.parametrize("name", list(RANK_REFUSED))
def test_rank_past_the_per_rank_lists(probe, tmp_path, name):
    query, rank = RANK_REFUSED[name]
    ref, ours = run_both(probe, tmp_path, query, None, rank)
    assert ref.returncode != 0, ref.stdout
    assert ours.returncode == 1 and ours.stderr.strip(), (ours.returncode, ours.stdout, ours.stderr)


@pytest.mark.parametrize("name", list(LOADER_REFUSED))
def test_loader_configurations_the_reference_refuses(probe, tmp_path, name):
    query, loader, rank = LOADER_REFUSED[name]
    ref, ours = run_both(probe, tmp_path, query, loader, rank)
    assert ref.returncode != 0, ref.stdout
    assert ours.returncode == 1 and ours.stderr.strip(), (ours.returncode, ours.stdout, ours.stderr)
