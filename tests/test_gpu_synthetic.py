"""GPU: CUDA path vs the oracle on GDBSYN-1 cohorts (same seeded inputs), sizes the oracle finishes in seconds."""
import numpy as np
import pytest

from oracle import oracle
from tools import gdbsyn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def both(ctx, d, ranges, rows=None, **kw):
    from genomicsdb_b200 import dump
    cols = gdbsyn.to_oracle_columns(d)
    want, n = oracle.combine_query(cols, ranges, query_rows=rows)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        got, n_got = dump.combine_dump(arr, ranges, ["GT", "AD", "PL"], query_rows=rows, **kw)
    finally:
        arr.close()
    assert n_got == n
    assert oracle.diff_dumps(want, got) == []
    assert got == want
    return n


@pytest.mark.parametrize("shape", [dict(n_samples=50, n_sites=200, spacing=100, p_var=0.1, a_pool=8),
                                   dict(n_samples=300, n_sites=100, spacing=100, p_var=0.1, a_pool=3),
                                   dict(n_samples=1000, n_sites=40, spacing=1000, p_var=0.15, a_pool=8),
                                   dict(n_samples=7, n_sites=500, spacing=3, p_var=0.4, a_pool=5)],
                         ids=["50x200", "300x100", "1000x40", "7x500dense"])
def test_full_scan_and_whole_interval(ctx, shape):
    d = gdbsyn.generate(**shape)
    n_full = both(ctx, d, [])
    n_int = both(ctx, d, [list(d["col_range"])])
    assert n_full == n_int


def test_mid_array_intervals_left_sweep(ctx):
    d = gdbsyn.generate(n_samples=200, n_sites=300, spacing=100, p_var=0.1, a_pool=8)
    lo, hi = d["col_range"]
    rng = np.random.RandomState(4321)
    ranges = []
    for _ in range(12):
        a = int(rng.randint(lo, hi))
        ranges.append([a, min(hi, a + int(rng.randint(0, 3000)))])
    ranges += [[lo + 1234, lo + 1234], [1000 + 100 * 7, 1000 + 100 * 7], [hi - 5, hi + 100], [lo - 50, lo + 10]]
    for r in ranges:
        both(ctx, d, [r])


def test_row_subset_and_small_batches(ctx):
    d = gdbsyn.generate(n_samples=120, n_sites=150, spacing=100, p_var=0.12, a_pool=8)
    rows = [3, 5, 8, 13, 21, 34, 55, 89, 119, 400]
    both(ctx, d, [], rows=rows)
    both(ctx, d, [[5000, 9000]], rows=rows, max_sites=3)
    both(ctx, d, [], max_bytes=200000)


def test_column_shard_equals_partition(ctx):
    """a shard generated on its own equals the slice of the full cohort with the same columns (multi-GPU sharding)"""
    full = gdbsyn.generate(n_samples=60, n_sites=120, spacing=100, p_var=0.1, a_pool=8)
    part = gdbsyn.generate(n_samples=60, n_sites=120, spacing=100, p_var=0.1, a_pool=8, site_lo=40, site_hi=80)
    lo, hi = part["col_range"]
    sel = (full["coords"][:, 1] >= lo) & (full["coords"][:, 1] <= hi)
    assert np.array_equal(full["coords"][sel], part["coords"])
    assert np.array_equal(full["end"][sel], part["end"])
    both(ctx, part, [[lo, hi]])
