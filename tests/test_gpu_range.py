"""GPU: tile-range staging (gdbgpu_array_open_range).  An array opened for a column range must answer every
query inside that range exactly as the whole array does, while staging only the tiles its book-keeping says can matter."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from tools import gdbsyn, tiledb_writer
from util import case_fields, cells_to_columns

pytestmark = pytest.mark.gpu
FIELDS = [("AD", "R", "int"), ("PL", "G", "int")]


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def dumps(arr, ranges, names, **kw):
    from tools import dump
    return dump.combine_dump(arr, ranges, names, **kw)


@pytest.mark.parametrize("compress", [False, True], ids=["plain", "gzip"])
def test_shards_of_an_on_disk_cohort(ctx, tmp_path, compress):
    """every shard opened on its own tile range equals the same interval queried on the whole array; the shards stage only
    their share of the fragment"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=400, n_sites=300, spacing=100, p_var=0.1, a_pool=8)
    ws = str(tmp_path / "ws")
    tiledb_writer.write_array(ws, "syn", d, capacity=1000, compress=compress)
    whole = capi.Array.open(ctx, ws, "syn", FIELDS)
    lo, hi = d["col_range"]
    cuts = [lo, lo + 4321, lo + 9000, lo + 17777, hi + 1]
    try:
        total_staged = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            b -= 1
            shard = capi.Array.open(ctx, ws, "syn", FIELDS, col_range=(a, b))
            try:
                si = shard.info()
                assert si.n_cells <= whole.info().n_cells
                if a == lo:
                    assert si.n_cells < whole.info().n_cells * 0.85   # a seventh of the columns + border tiles + the margin that holds the END copies
                total_staged += si.staged_bytes
                for q in ([a, b], [a + 100, a + 100], [a + 57, min(b, a + 2000)], [b - 10, b]):
                    want = dumps(whole, [q], ["GT", "AD", "PL"])
                    got = dumps(shard, [q], ["GT", "AD", "PL"])
                    assert got == want, (a, b, q)
            finally:
                shard.close()
        # (shards overlap by the tiles at their borders and by the stretch to the right that holds the END copies of the calls
        # spanning their right edge -- at 36 tiles in all, that is most of this array)
        assert total_staged > 0
    finally:
        whole.close()


def test_range_grows_to_the_right_for_a_call_spanning_the_whole_range(ctx, tmp_path):
    """sample 2 has one reference block over everything: inside a narrow range it is only found through its END copy, far to
    the right of the range -- the staged tile range has to grow until it shows up (gt_get_column's forward scan)"""
    from genomicsdb_b200 import capi
    cells = []
    for r in (0, 1):
        for i in range(3000):
            c = 1000 + 10 * i
            cells.append(dict(row=r, col=c, end=c, ref="A", alt="T|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0]))
    cells.append(dict(row=2, col=900, end=40000, ref="N", alt="&", gt=[0, 0], pl=[0, 5, 50], ad=[]))
    cells.append(dict(row=2, col=40000, end=900, ref="N", alt="&", gt=[0, 0], pl=[0, 5, 50], ad=[]))    # its END copy
    d = cells_to_columns(cells, 3)
    ws = str(tmp_path / "ws")
    tiledb_writer.write_array(ws, "syn", d, capacity=100)
    whole = capi.Array.open(ctx, ws, "syn", FIELDS)
    shard = capi.Array.open(ctx, ws, "syn", FIELDS, col_range=(5000, 5200))
    try:
        si, wi = shard.info(), whole.info()
        assert si.max_col == wi.max_col and si.n_cells < wi.n_cells   # grew to the end (the END copy is the array's last cell); the left part stays on disk
        q = [5000, 5200]
        assert dumps(shard, [q], ["GT", "AD", "PL"]) == dumps(whole, [q], ["GT", "AD", "PL"])
        n_live = [int(x) for x in capi.Scan(shard, 5000, 5200).next_batch().n_live]
        assert max(n_live) == 3
    finally:
        shard.close()
        whole.close()


CASES = [c for c in golden_cases() if c["id"] in ("t0_1_2_at_0", "t0_1_2_two_fragments_at_0", "t6_7_8_three_fragments_at_0", "t0_overlapping_at_0")]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_reference_arrays_by_range(ctx, case):
    """the reference's own arrays (GZIP tiles of 3 cells, several fragments): any sub-range staged on its own"""
    from genomicsdb_b200 import capi
    fields, phased = case_fields(case)
    names = ["GT"] + [f[0] for f in fields]
    ws = os.path.join(GOLDEN, case["workspace"])
    whole = capi.Array.open(ctx, ws, case["array"], fields, gt_phased=phased)
    try:
        coords, end = whole.cells()
        cols = np.unique(coords[:, 1])
        rng = np.random.RandomState(7)
        for _ in range(8):
            a, b = sorted(int(x) for x in rng.choice(cols, 2))
            a -= int(rng.randint(0, 3))
            shard = capi.Array.open(ctx, ws, case["array"], fields, gt_phased=phased, col_range=(a, b))
            try:
                assert dumps(shard, [[a, b]], names) == dumps(whole, [[a, b]], names), (a, b)
            finally:
                shard.close()
    finally:
        whole.close()
