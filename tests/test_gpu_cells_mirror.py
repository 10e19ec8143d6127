"""GPU: the columnar-iterator surface of the host mirror (VariantQueryProcessor::iterate_over_cells + the
SingleCellOperatorBase operators, genomicsdb_b200/csrc/gdb_host_api.h) driven by tools/gt_gpu_gather.cc in the reference
tool's four columnar commands: every byte of stdout equals what the unmodified reference's gt_mpi_gather printed for the
same query (tests/golden/cells, oracle/make_golden_cells.py) -- including the queries of the reference's own CI whose
*_calls_* goldens that generator reproduced."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, golden_cases

pytestmark = pytest.mark.gpu
ALL = golden_cases()
CELLS = os.path.join(GOLDEN, "cells")
CI = json.load(open(os.path.join(CELLS, "ci_manifest.json")))
MODES = (("calls", "--print-calls"), ("csv", "--print-csv"), ("ac", "--print-AC"), ("hist", "--produce-histogram"),
         ("ipos", "--produce-interesting-positions"))     # the last one sits on scan_and_operate (InterestingLocationsPrinter), not on the iterator


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    g.build_gpu_lib()
    return g.build_tools()


def run_tool(exe, case, tmp_path, flag, attributes=None, window=None):
    q = {"workspace": os.path.join(GOLDEN, case["workspace"]), "array_name": case["array"],
         "vid_mapping_file": os.path.join(GOLDEN, case["vid_mapping_file"]), "callset_mapping_file": os.path.join(GOLDEN, case["callset_mapping_file"]), "attributes": attributes or case["attributes"],
         "query_column_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_column_ranges"]]}],
         "query_row_ranges": [{"range_list": [{"low": lo, "high": hi} for lo, hi in case["query_row_ranges"]]}]}
    qf = tmp_path / "query.json"
    qf.write_text(json.dumps(q))
    out = tmp_path / "out.txt"
    env = dict(os.environ, GDBGPU_CELLS_WINDOW=str(window)) if window else None
    return subprocess.run([exe, "-j", str(qf), "-o", str(out), flag], capture_output=True, text=True, env=env), out


# the two print commands on every array; the two device-side operators through the tool on a few (tests/test_gpu_cells.py
# already holds their output against the same reference files on every array, through the C ABI)
DEVICE_OP_CASES = ("t0_1_2_multiple_positions", "t0_overlapping_at_0", "haploid_triploid_deletion", "t0_1_2_all_asa_at_0", "t6_7_8_three_fragments_at_0")
RUNS = [(c, m) for c in ALL for m in MODES if m[0] in ("calls", "csv") or c["id"] in DEVICE_OP_CASES]


@pytest.mark.parametrize("case,mode", RUNS, ids=["%s-%s" % (c["id"], m[0]) for c, m in RUNS])
def test_columnar_command_output_equals_the_reference(exe, case, tmp_path, mode):
    ext, flag = mode
    p, out = run_tool(exe, case, tmp_path, flag)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == open(os.path.join(CELLS, case["id"] + "." + ext), "rb").read()


@pytest.mark.parametrize("cid", sorted(CI), ids=sorted(CI))
def test_print_calls_equals_the_reference_ci_golden_queries(exe, cid, tmp_path):
    """the reference CI's own --print-calls queries (tests/run.py attribute lists; outputs md5/JSON-equal to its golden files)"""
    case = [c for c in ALL if c["id"] == cid][0]
    p, out = run_tool(exe, case, tmp_path, "--print-calls", attributes=CI[cid]["attributes"])
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == open(os.path.join(CELLS, cid + ".ci.calls"), "rb").read()


@pytest.mark.parametrize("cid", ["t0_1_2_multiple_positions", "t0_1_2_combined_at_0"])
def test_paging_through_the_cell_list_is_transparent(exe, cid, tmp_path):
    """iterate_over_cells with a two-cell window (gdbgpu_scan_cells_window): same bytes"""
    case = [c for c in ALL if c["id"] == cid][0]
    for ext, flag in MODES[:2]:
        p, out = run_tool(exe, case, tmp_path, flag, window=2)
        assert p.returncode == 0, p.stderr
        assert out.read_bytes() == open(os.path.join(CELLS, cid + "." + ext), "rb").read()


# ---------------------------------------------------------------------------------------------- the GA4GH range query
VARIANTS = os.path.join(GOLDEN, "variants")


@pytest.mark.parametrize("case", ALL, ids=[c["id"] for c in ALL])
def test_range_query_equals_the_reference(exe, case, tmp_path):
    """gt_gpu_gather --range-query (VariantQueryProcessor::gt_get_column_interval of the mirror: the columnar iterator's cell list
    grouped into GA4GH variants, variants of several calls re-indexed by the combine kernels, print_variants) against the stdout
    of the unmodified reference's `gt_mpi_gather -j <query>` (run_range_query, its default command) on the same array"""
    p, out = run_tool(exe, case, tmp_path, "--range-query")
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == open(os.path.join(VARIANTS, case["id"] + ".variants"), "rb").read()


@pytest.mark.parametrize("cid", sorted(json.load(open(os.path.join(GOLDEN, "variants", "ci.json")))["ci_cases"]))
def test_range_query_equals_the_reference_ci_goldens(exe, cid, tmp_path):
    """the reference CI's own range queries (tests/run.py attribute lists; our copies are md5-identical to its *_variants_* goldens)"""
    ci = json.load(open(os.path.join(VARIANTS, "ci.json")))
    case = [c for c in ALL if c["id"] == cid][0]
    attrs = ci["ci_attributes"] + (["DS", "ID"] if "DS_ID" in cid else [])
    p, out = run_tool(exe, case, tmp_path, "--range-query", attributes=attrs)
    assert p.returncode == 0, p.stderr
    assert out.read_bytes() == open(os.path.join(VARIANTS, cid + ".ci.variants"), "rb").read()
