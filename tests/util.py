"""Shared helpers for the parity tests."""
import os

from conftest import GOLDEN
from oracle import oracle


def case_query_rows(case):
    rows = []
    for lo, hi in case["query_row_ranges"]:
        rows += list(range(lo, hi + 1))
    return rows


def case_fields(case):
    """(name, kind, elem) of the allele-dependent 1-D fields in query order + GT layout, from the vid mapping."""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    fields = []
    for name in case["attributes"]:
        ent = vid.get(name)
        if ent is None or not isinstance(ent["length"], str) or ent["length"] not in oracle.KIND:
            continue
        fields.append((name, ent["length"], "float" if ent["type"] == "float" else "int"))
    return fields, vid["GT"]["length"] == "PP"


def oracle_dump(case):
    cols = oracle.load_columns(os.path.join(GOLDEN, case["workspace"]), case["array"],
                               os.path.join(GOLDEN, case["vid_mapping_file"]), case["attributes"])
    return oracle.combine_query(cols, case["query_column_ranges"], query_rows=case_query_rows(case),
                                produce_GT_with_min_PL=case["produce_GT_with_min_PL_value_for_spanning_deletions"])


def format_field_names(case, dump_names):
    """GT + the allele-dependent fields of the dump that are FORMAT fields of the VCF (vid vcf_field_class)."""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    return [n for n in dump_names if n == "GT" or "FORMAT" in (vid.get(n, {}).get("classes") or [])]
