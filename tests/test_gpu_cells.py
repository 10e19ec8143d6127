"""GPU: the columnar iterator and its operators on the device (gdbgpu_scan_cells*, through the C ABI) against the
unmodified reference's outputs over the golden arrays and against the restatement on synthetic cohorts."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import cells_oracle, oracle, tiledb_fragment
from util import case_fields, case_query_rows

pytestmark = pytest.mark.gpu
CASES = golden_cases()
CELLS = os.path.join(GOLDEN, "cells")


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def open_case(ctx, case):
    from genomicsdb_b200 import capi
    fields, phased = case_fields(case)
    return capi.Array.open(ctx, os.path.join(GOLDEN, case["workspace"]), case["array"], fields, gt_phased=phased), phased


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_cells_histogram_allele_counts_match_reference(ctx, case):
    from genomicsdb_b200 import capi
    arr, phased = open_case(ctx, case)
    host = tiledb_fragment.Array(os.path.join(GOLDEN, case["workspace"]), case["array"])
    end = host.fixed("END").reshape(-1).astype(np.int64)
    roff, rval = host.var("REF")
    rbytes = rval.astype(np.uint8).tobytes()
    try:
        ac_text, counts, n_bins, csv_prefix = "", {}, 0, []
        for lo, hi in case["query_column_ranges"]:
            sc = capi.Scan(arr, lo, hi, query_rows=case_query_rows(case), flags=capi.SCAN_CELLS_ONLY)
            try:
                got = sc.cells()
                want, n_left = cells_oracle.iterate(host.row, host.col, end, case_query_rows(case), lo, hi)
                assert got["n_left"] == n_left
                assert list(zip(got["cell"].tolist(), got["row"].tolist(), got["begin"].tolist(), got["end"].tolist())) == want
                ac_text += cells_oracle.allele_counts_text(sc.cells_allele_counts())
                # a sparse view of the 40 M-bin histogram the reference's tool allocates
                h = sc.cells_histogram(0, 4000000000, 100)
                n_bins = len(h)
                for k in np.nonzero(h)[0]:
                    counts[int(k)] = counts.get(int(k), 0) + int(h[k])
                off, by = sc.cells_fetch(capi.CELL_FIELD_REF)
                assert [by[off[i]:off[i + 1]] for i in range(len(want))] == [rbytes[roff[c]:roff[c + 1]] for c, _, _, _ in want]
                with pytest.raises(capi.GdbGpuError):
                    sc.next_batch()
            finally:
                sc.close()
        assert ac_text == open(os.path.join(CELLS, case["id"] + ".ac")).read()
        assert cells_oracle.histogram_text(counts, n_bins, 0, 100) == open(os.path.join(CELLS, case["id"] + ".hist")).read()
    finally:
        arr.close()


def test_cells_on_a_synthetic_cohort(ctx):
    """1000 samples, mid-array interval, a row subset, the full scan: list, histogram, allele counts, column fetches"""
    from genomicsdb_b200 import capi
    from tools import gdbsyn
    d = gdbsyn.generate(n_samples=1000, n_sites=60, spacing=100, p_var=0.1, a_pool=8, seed=7)
    arr = gdbsyn.to_gpu_array(ctx, d)

    class View:     # per-cell accessors over the generated columns (byte offsets + values)
        row, col, end = d["coords"][:, 0], d["coords"][:, 1], d["end"]

        @staticmethod
        def cell_bytes(name, c):
            off, val = d[name]
            b = val.view(np.uint8)
            return b[int(off[c]):(int(off[c + 1]) if c + 1 < len(off) else b.size)].tobytes()
        ref_str = staticmethod(lambda c: View.cell_bytes("ref", c).decode())
        alt_str = staticmethod(lambda c: View.cell_bytes("alt", c).decode())
        gt_words = staticmethod(lambda c: np.frombuffer(View.cell_bytes("gt", c), dtype=np.int32))
        field_words = staticmethod(lambda f, c: np.frombuffer(View.cell_bytes(("ad", "pl")[f], c), dtype=np.int32))
    view = View
    lo0, hi0 = d["col_range"]
    rng = np.random.default_rng(5)
    subset = np.sort(rng.choice(1000, 137, replace=False))
    try:
        for lo, hi, rows in ((lo0 + 1234, lo0 + 3999, None), (lo0 + 2050, lo0 + 2050, subset), (None, None, None), (lo0, hi0, subset)):
            qrows = list(range(1000)) if rows is None else rows.tolist()
            sc = capi.Scan(arr, lo, hi, query_rows=rows, flags=capi.SCAN_CELLS_ONLY)
            try:
                got = sc.cells()
                want, n_left = cells_oracle.iterate(view.row, view.col, view.end, qrows, lo, hi)
                assert got["n_left"] == n_left and len(want) > 0
                assert list(zip(got["cell"].tolist(), got["row"].tolist(), got["begin"].tolist(), got["end"].tolist())) == want
                h = sc.cells_histogram(lo0, hi0, 250)
                hw, nb = cells_oracle.histogram(want, lo0, hi0, 250)
                assert nb == len(h) and {int(k): int(h[k]) for k in np.nonzero(h)[0]} == hw
                assert sc.cells_allele_counts() == cells_oracle.allele_counts(want, view.ref_str, view.alt_str, view.gt_words, 1)
                off, by = sc.cells_fetch(capi.CELL_FIELD_ALT)
                assert all(by[off[i]:off[i + 1]].decode() == view.alt_str(c) for i, (c, _, _, _) in enumerate(want[:500]))
                off, by = sc.cells_fetch(capi.CELL_FIELD_REMAPPED, 0)
                w = np.frombuffer(by, dtype=np.int32)
                for i, (c, _, _, _) in enumerate(want[:200]):
                    assert w[off[i] // 4:off[i + 1] // 4].tolist() == list(view.field_words(0, c))
                # paging: a window in the middle of the list (straddling the left part when there is one)
                first = max(0, n_left - 3) if n_left else len(want) // 3
                sc.cells_window(first, 1000)
                page = sc.cells()
                assert page["n_total"] == len(want) and page["first"] == first and page["n_left"] == min(max(n_left - first, 0), 1000)
                assert list(zip(page["cell"].tolist(), page["row"].tolist(), page["begin"].tolist(), page["end"].tolist())) == want[first:first + 1000]
                off, by = sc.cells_fetch(capi.CELL_FIELD_REF)
                assert [by[off[i]:off[i + 1]].decode() for i in range(len(page["cell"]))] == [view.ref_str(c) for c, _, _, _ in want[first:first + 1000]]
            finally:
                sc.close()
    finally:
        arr.close()


def test_equal_load_shards_from_the_device_histogram(ctx):
    """shard boundaries with ~equal cell counts, taken from the device histogram of a cohort whose call density
    varies along the genome; the shards' cell lists tile the whole interval's list"""
    from genomicsdb_b200 import capi
    from genomicsdb_b200.sharding import device_equal_load_cuts
    from tools import gdbsyn
    d = gdbsyn.generate(n_samples=2000, n_sites=200, spacing=100, p_var=0.15, a_pool=6, seed=11)
    arr = gdbsyn.to_gpu_array(ctx, d)
    lo, hi = d["col_range"]
    try:
        rows = np.arange(0, 2000, 3)          # uneven load: a third of the samples
        cuts = device_equal_load_cuts(arr, 4, 100, lo, hi, query_rows=rows)
        assert len(cuts) == 4 and cuts[0][0] == lo and cuts[-1][1] == hi
        whole = capi.Scan(arr, lo, hi, query_rows=rows, flags=capi.SCAN_CELLS_ONLY)
        all_cells = whole.cells()
        whole.close()
        loads, begun = [], []
        for a, b in cuts:
            sc = capi.Scan(arr, a, b, query_rows=rows, flags=capi.SCAN_CELLS_ONLY)
            c = sc.cells()
            sc.close()
            loads.append(len(c["cell"]) - c["n_left"])
            begun += c["cell"][c["n_left"]:].tolist()         # cells that begin inside the shard
        assert begun == all_cells["cell"][all_cells["n_left"]:].tolist()
        assert max(loads) - min(loads) <= 0.2 * (sum(loads) / 4), loads
    finally:
        arr.close()
