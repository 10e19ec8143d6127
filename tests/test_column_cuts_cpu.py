"""CPU: gdbgpu_array_column_cuts (equal-load column partitions for N ranks from the fragments' book-keeping alone --
tile bounding coordinates and sizes, no attribute file is read, no device is needed)."""
import numpy as np
import pytest

from tools import gdbsyn, tiledb_writer


@pytest.mark.parametrize("writer_args", [{}, dict(compress=True), dict(n_fragments=4, dup_every=11)], ids=["plain", "gzip", "4-fragments"])
def test_cuts_tile_the_columns_with_about_equal_cell_counts(tmp_path, writer_args):
    from genomicsdb_b200 import sharding
    d = gdbsyn.generate(n_samples=300, n_sites=400, spacing=100, p_var=0.1, a_pool=4, with_fixed=True)
    ws = str(tmp_path / "ws")
    tiledb_writer.write_array(ws, "syn", d, capacity=500, **writer_args)
    col = d["coords"][:, 1]
    n = len(col)
    for world in (1, 2, 3, 8):
        cuts = sharding.tile_balanced_cuts(ws, "syn", world)
        assert len(cuts) == world
        assert cuts[0][0] <= int(col.min()) and cuts[-1][1] >= int(col.max())
        for (lo, hi), (lo2, hi2) in zip(cuts, cuts[1:]):
            assert lo <= hi and lo2 == hi + 1                     # contiguous, non-overlapping
        counts = [int(((col >= lo) & (col <= hi)).sum()) for lo, hi in cuts]
        assert sum(counts) == n
        # a part closes at a tile boundary: no part is off its share by more than a few tiles of 500 cells (per fragment)
        slack = 500 * 3 * writer_args.get("n_fragments", 1)
        assert max(counts) - n / world <= slack, (world, counts)


def test_cuts_of_a_missing_array_are_an_error(tmp_path):
    from genomicsdb_b200 import capi, sharding
    with pytest.raises(capi.GdbGpuError):
        sharding.tile_balanced_cuts(str(tmp_path), "nothing_here", 2)
