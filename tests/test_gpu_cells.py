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


def _random_array(rng, n_rows, span):
    """a small gVCF-shaped array: per row, non-overlapping calls [b, e] (some single-position, some overlapping the next
    call on purpose), each with its END copy at e when e > b; cells in column-major order; ALT pools with deletions"""
    cells = []   # (col, row, END, ref, alt, gt)
    for r in range(n_rows):
        pos = int(rng.integers(0, 6))
        while pos < span:
            length = int(rng.choice([1, 1, 2, 5, 17]))
            e = min(pos + length - 1, span + 5)
            kind = int(rng.integers(0, 4))
            if kind == 0:
                ref, alt, gt = "N", "&", [0, 0]
            elif kind == 1:
                ref, alt, gt = "A", "T|&", [0, 1]
            elif kind == 2:
                ref, alt, gt = "ATG", "A|ACG|&", [int(rng.integers(0, 3)), int(rng.integers(1, 4))]
            else:
                ref, alt, gt = "C", "CAA|G|&", [1, 2]
            cells.append((pos, r, e, ref, alt, gt))
            if e > pos:
                cells.append((e, r, pos, ref, alt, gt))      # END copy: coordinates hold the end, END the begin
            nxt = e + 1 + int(rng.choice([0, 0, 0, 3, 9]))
            if rng.random() < 0.1 and e > pos + 1:
                nxt = e - 1                                  # the next call starts inside this one
            pos = max(nxt, pos + 1)
    # one cell per coordinate (TileDB keeps the newest): drop collisions, column-major order
    seen, uniq = set(), []
    for c in sorted(cells, key=lambda c: (c[0], c[1])):
        if (c[0], c[1]) not in seen:
            seen.add((c[0], c[1]))
            uniq.append(c)
    coords = np.array([[c[1], c[0]] for c in uniq], dtype=np.int64)
    end = np.array([c[2] for c in uniq], dtype=np.int64)

    def var(strings, dtype=np.uint8):
        vals = [np.frombuffer(s.encode(), dtype=np.uint8) if dtype == np.uint8 else np.asarray(s, dtype=np.int32) for s in strings]
        off = np.cumsum([0] + [v.nbytes for v in vals])[:-1].astype(np.uint64)
        return off, (np.concatenate(vals) if vals else np.zeros(0, dtype=dtype))
    return coords, end, var([c[3] for c in uniq]), var([c[4] for c in uniq]), var([c[5] for c in uniq], np.int32), uniq


def test_cells_on_random_small_arrays(ctx):
    """300 random arrays x 4 queries: the device's list (left part from END copies, rows subset, empty intervals, calls that
    overlap the next one) and allele counts against the restatement"""
    from genomicsdb_b200 import capi
    rng = np.random.default_rng(20240611)
    for it in range(300):
        n_rows, span = int(rng.integers(1, 9)), int(rng.integers(8, 70))
        coords, end, ref, alt, gt, uniq = _random_array(rng, n_rows, span)
        arr = capi.Array.from_columns(ctx, coords, end, ref, alt, gt, [], 0, n_rows)
        try:
            for _ in range(4):
                lo = int(rng.integers(0, span + 8))
                hi = lo + int(rng.choice([0, 0, 3, 40]))
                rows = None if rng.random() < 0.5 else np.sort(rng.choice(n_rows, int(rng.integers(1, n_rows + 1)), replace=False))
                full = rng.random() < 0.15
                sc = capi.Scan(arr, None if full else lo, None if full else hi, query_rows=rows, flags=capi.SCAN_CELLS_ONLY)
                try:
                    got = sc.cells()
                    want, n_left = cells_oracle.iterate(coords[:, 0], coords[:, 1], end, range(n_rows) if rows is None else rows.tolist(),
                                                        None if full else lo, None if full else hi)
                    key = (it, lo, hi, None if rows is None else rows.tolist(), full)
                    assert got["n_left"] == n_left, key
                    assert list(zip(got["cell"].tolist(), got["row"].tolist(), got["begin"].tolist(), got["end"].tolist())) == want, key
                    assert sc.cells_allele_counts() == cells_oracle.allele_counts(want, lambda i: uniq[i][3], lambda i: uniq[i][4], lambda i: uniq[i][5], 1), key
                finally:
                    sc.close()
        finally:
            arr.close()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_site_summaries_match_the_reference_interesting_positions(ctx, case):
    """GDBGPU_SCAN_SITE_STATS | GDBGPU_SCAN_SUMMARY_ONLY: per site "begin, valid calls, reference-block calls, calls beginning
    here" -- what the unmodified reference's InterestingLocationsPrinter printed through scan_and_operate
    (gt_mpi_gather --produce-interesting-positions); no per-call data is handed out"""
    from genomicsdb_b200 import capi
    arr, _ = open_case(ctx, case)
    try:
        text = ""
        for lo, hi in case["query_column_ranges"]:
            sc = capi.Scan(arr, lo, hi, query_rows=case_query_rows(case), flags=capi.SCAN_SITE_STATS | capi.SCAN_SUMMARY_ONLY)
            try:
                while True:
                    b = sc.next_batch(max_sites=2)
                    if b is None:
                        break
                    assert b.live is None and b.site_stats is not None
                    for s in range(b.n_sites):
                        text += "%d %d %d %d\n" % (b.site_begin[s], b.n_live[s], b.site_stats[s, 0], b.site_stats[s, 1])
            finally:
                sc.close()
        assert text == open(os.path.join(CELLS, case["id"] + ".ipos")).read()
    finally:
        arr.close()
