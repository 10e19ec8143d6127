"""CPU: pins the oracle (oracle/combine_oracle.c + oracle/tiledb_fragment.py) against the reference.

The golden dumps under tests/golden/dumps were written by the UNMODIFIED reference engine (oracle/_ref/bin/gdb_ref_dump,
see oracle/make_golden.py) on the reference's own CI scenarios (tests/run.py:185-870); the reference's VCF output on
the same arrays was md5-equal to its golden_outputs/ files when the fixtures were generated.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import oracle

CASES = golden_cases()


def _oracle_dump(case):
    cols = oracle.load_columns(os.path.join(GOLDEN, case["workspace"]), case["array"],
                               os.path.join(GOLDEN, case["vid_mapping_file"]), case["attributes"])
    rows = []
    for lo, hi in case["query_row_ranges"]:
        rows += list(range(lo, hi + 1))
    dump, n = oracle.combine_query(cols, case["query_column_ranges"], query_rows=rows,
                                   produce_GT_with_min_PL=case["produce_GT_with_min_PL_value_for_spanning_deletions"])
    return dump, n


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_dump(case):
    assert case["reference_vcf_md5_matches_golden"] in (True, None)
    gold = open(os.path.join(GOLDEN, case["dump"]), "rb").read()
    dump, n = _oracle_dump(case)
    assert n == case["sites"]
    assert oracle.diff_dumps(gold, dump) == []
    assert dump == gold
