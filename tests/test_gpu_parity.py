"""GPU: the CUDA path, called through the C ABI, against the reference's own dumps and against the oracle."""
import os

import pytest

from conftest import GOLDEN, golden_cases
from oracle import oracle
from util import case_fields, case_query_rows, oracle_dump

pytestmark = pytest.mark.gpu
CASES = golden_cases()


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def gpu_dump(ctx, case, **kw):
    from genomicsdb_b200 import capi
    from tools import dump
    fields, phased = case_fields(case)
    arr = capi.Array.open(ctx, os.path.join(GOLDEN, case["workspace"]), case["array"], fields, gt_phased=phased)
    try:
        flags = capi.SCAN_MIN_PL_GT if case["produce_GT_with_min_PL_value_for_spanning_deletions"] else 0
        names = ["GT"] + [f[0] for f in fields]
        return dump.combine_dump(arr, case["query_column_ranges"], names, query_rows=case_query_rows(case), flags=flags, **kw)
    finally:
        arr.close()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_gpu_matches_reference_dump(ctx, case):
    gold = open(os.path.join(GOLDEN, case["dump"]), "rb").read()
    got, n = gpu_dump(ctx, case)
    assert n == case["sites"]
    assert oracle.diff_dumps(gold, got) == []
    assert got == gold


@pytest.mark.parametrize("case", [c for c in CASES if c["sites"] > 2][:6], ids=lambda c: c["id"])
def test_gpu_batching_is_transparent(ctx, case):
    """one site per batch (the overflow/resume protocol, query_variants.cc:460-475) gives the same bytes"""
    gold = open(os.path.join(GOLDEN, case["dump"]), "rb").read()
    got, n = gpu_dump(ctx, case, max_sites=1)
    assert got == gold
