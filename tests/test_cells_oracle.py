"""CPU: the restatement of the columnar iterator + its operators (oracle/cells_oracle.py) against the UNMODIFIED
reference's own --print-csv / --print-AC / --produce-histogram output over the golden arrays (tests/golden/cells)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import cells_oracle, oracle, tiledb_fragment
from util import case_query_rows

CASES = golden_cases()
CELLS = os.path.join(GOLDEN, "cells")


def load(case):
    arr = tiledb_fragment.Array(os.path.join(GOLDEN, case["workspace"]), case["array"])
    end = arr.fixed("END").reshape(-1).astype(np.int64)
    return arr, end


def case_cells(case, arr, end):
    """the iterator's cells over all query intervals of the case -> [(interval, [(cell,row,begin,end)], n_left)]"""
    out = []
    for lo, hi in case["query_column_ranges"]:
        cells, n_left = cells_oracle.iterate(arr.row, arr.col, end, case_query_rows(case), lo, hi)
        out.append(((lo, hi), cells, n_left))
    return out


def var_strings(arr, name):
    off, val = arr.var(name)
    b = val.astype(np.uint8).tobytes() if val.dtype != np.uint8 else val.tobytes()
    return lambda i: b[off[i]:off[i + 1]].decode()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_iterator_order_matches_reference_csv(case):
    """row, begin, end, REF, ALT of every line the reference's VariantCallPrintCSVOperator printed, in order"""
    arr, end = load(case)
    ref, alt = var_strings(arr, "REF"), var_strings(arr, "ALT")
    got = []
    for _, cells, _ in case_cells(case, arr, end):
        got += ["%d,%d,%d,%s,%s" % (r, b, e, ref(i), alt(i)) for i, r, b, e in cells]
    text = "\n" + open(os.path.join(CELLS, case["id"] + ".csv"), "rb").read().decode("latin-1")
    asa = case["id"] == "t0_1_2_all_asa_at_0"    # its 2-D fields are printed as raw bytes, newlines included
    if not asa:
        assert len(got) == text.count("\n") - 1
    pos = 0
    for g in got:                               # every line, in order
        nxt = text.find("\n" + g + ",", pos)
        assert nxt >= 0 and (asa or nxt == pos), (g, text[pos:pos + 80])
        pos = text.find("\n", nxt + 1)


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_histogram_matches_reference(case):
    arr, end = load(case)
    counts, n_bins = {}, 0
    for _, cells, _ in case_cells(case, arr, end):
        h, n_bins = cells_oracle.histogram(cells, 0, 4000000000, 100)
        for k, v in h.items():
            counts[k] = counts.get(k, 0) + v
    assert cells_oracle.histogram_text(counts, n_bins, 0, 100) == open(os.path.join(CELLS, case["id"] + ".hist")).read()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_allele_counts_match_reference(case):
    arr, end = load(case)
    ref, alt = var_strings(arr, "REF"), var_strings(arr, "ALT")
    goff, gval = arr.var("GT")
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    step = 2 if vid["GT"]["length"] == "PP" else 1
    text = ""
    for _, cells, _ in case_cells(case, arr, end):
        text += cells_oracle.allele_counts_text(cells_oracle.allele_counts(cells, ref, alt, lambda i: gval[goff[i]:goff[i + 1]], step))
    assert text == open(os.path.join(CELLS, case["id"] + ".ac")).read()


def test_left_cells_come_first_ordered_by_begin_then_row():
    # rows 0..2: calls [5,20] (row 0), [3,30] (row 1), begin cell at 10 (row 2); query [10, 40]
    row = np.array([1, 0, 2, 0, 1], dtype=np.int64)
    col = np.array([3, 5, 10, 20, 30], dtype=np.int64)
    end = np.array([30, 20, 10, 5, 3], dtype=np.int64)
    cells, n_left = cells_oracle.iterate(row, col, end, [0, 1, 2], 10, 40)
    assert n_left == 2
    assert cells == [(4, 1, 3, 30), (3, 0, 5, 20), (2, 2, 10, 10)]
    cells, n_left = cells_oracle.iterate(row, col, end, [0, 2], 10, 40)
    assert cells == [(3, 0, 5, 20), (2, 2, 10, 10)]
    cells, n_left = cells_oracle.iterate(row, col, end, [0, 1, 2])
    assert n_left == 0 and [c[0] for c in cells] == [0, 1, 2]


def test_restatement_matches_the_reference_tool_on_a_cohort_with_deletions(tmp_path):
    """the iterator's order, the histogram and the allele counts of a mid-array interval of a synthetic cohort whose variant calls
    are deletions to a third (tools/gdbsyn_deletions.py: calls that reach into the interval from the left, END duplicates,
    normalised deletion alleles) against the unmodified reference tool run here"""
    import subprocess
    from conftest import REPO
    from tools import gdbsyn, gdbsyn_deletions, tiledb_writer
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=120, n_sites=80, spacing=100, p_var=0.2, a_pool=5, with_fixed=True), share=0.35)
    info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d)
    lo, hi = d["col_range"]
    q_lo, q_hi = 1000 + 100 * 20 + 1, hi - 1500      # begins one base into the deletions of site 20
    qf = tiledb_writer.write_query_json(str(tmp_path / "q.json"), info, q_lo, q_hi, attributes=("REF", "ALT", "GT"))
    run = lambda flag: subprocess.run(["stdbuf", "-o0", ref_tool, "-j", qf, flag], capture_output=True, check=True).stdout.decode()
    arr = tiledb_fragment.Array(info["workspace"], "syn")
    end = arr.fixed("END").reshape(-1).astype(np.int64)
    cells, n_left = cells_oracle.iterate(arr.row, arr.col, end, list(range(120)), q_lo, q_hi)
    assert n_left > 100 and len(cells) > 1000
    ref, alt = var_strings(arr, "REF"), var_strings(arr, "ALT")
    got = ["%d,%d,%d,%s,%s" % (r, b, e, ref(i), alt(i)) for i, r, b, e in cells]
    want = [",".join(l.split(",")[:5]) for l in run("--print-csv").splitlines()]
    assert got == want
    h, n_bins = cells_oracle.histogram(cells, 0, 4000000000, 100)
    assert cells_oracle.histogram_text(h, n_bins, 0, 100) == run("--produce-histogram")
    goff, gval = arr.var("GT")
    assert cells_oracle.allele_counts_text(cells_oracle.allele_counts(cells, ref, alt, lambda i: gval[goff[i]:goff[i + 1]], 1)) == run("--print-AC")
