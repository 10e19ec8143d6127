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
    # an INFO field whose combine operation is move_to_FORMAT is emitted as a FORMAT field (broad_combined_gvcf.cc:183-189)
    return [n for n in dump_names if n == "GT" or "FORMAT" in (vid.get(n, {}).get("classes") or [])
            or vid.get(n, {}).get("combine") == "move_to_FORMAT"]


def plain_format_fields(case):
    """(attribute, 'int'|'float'|'char', fixed length | 0) of the queried FORMAT attributes that are not allele-dependent
    and not GT (GQ, SB, PGT, PID, MIN_DP, DP_FORMAT ...), in query order."""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    out = []
    for name in case["attributes"]:
        ent = vid.get(name)
        if ent is None or name == "GT":
            continue
        base = name[:-7] if name.endswith("_FORMAT") else name
        classes = ent.get("classes") or []
        is_format = "FORMAT" in classes and (name.endswith("_FORMAT") or "INFO" not in classes)
        if not is_format or (isinstance(ent["length"], str) and ent["length"] in oracle.KIND):
            continue
        ln = ent["length"]
        fixed = 0 if (isinstance(ln, str) and not ln.isdigit()) else int(ln)
        elem = "char" if ent["type"] == "char" else ("float" if ent["type"] == "float" else "int")
        out.append((name, elem, fixed))
    return out


def vcf_format_name(attr):
    """DP_FORMAT is the FORMAT twin of the INFO field DP (vid_mapper.cc): it is written as FORMAT/DP"""
    return attr[:-7] if attr.endswith("_FORMAT") else attr


def format_order(case, remapped_names, plain):
    """VCF names in the order the reference emits FORMAT fields: GT first (vcf.c:3670-3676), then query order, DP last
    (broad_combined_gvcf.cc:168-240,688-720)"""
    names = []
    for a in case["attributes"]:
        if a in remapped_names and a != "GT":
            names.append(a)
        elif any(a == p[0] for p in plain) and a != "DP_FORMAT":
            names.append(a)
    return ["GT"] + names + (["DP"] if any(p[0] == "DP_FORMAT" for p in plain) else [])


KNOWN_INFO_OPS = {"BaseQRankSum": "median", "ClippingRankSum": "median", "MQRankSum": "median", "ReadPosRankSum": "median",
                  "MQ": "median", "MQ0": "median", "ExcessHet": "median", "RAW_MQ": "sum"}   # known_field_info.cc:284-308


def info_scalar_fields(case):
    """(attribute, 'int'|'float', op) of the queried scalar INFO attributes the Broad operator combines with median / sum /
    mean, in query order (m_INFO_fields_vec, broad_combined_gvcf.cc:168-215); DP and END are handled separately."""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    out = []
    for name in case["attributes"]:
        ent = vid.get(name)
        if ent is None or name in ("END", "DP") or "INFO" not in (ent.get("classes") or []):
            continue
        op = ent.get("combine") or KNOWN_INFO_OPS.get(name)
        if op not in ("median", "sum", "mean") or ent["length"] not in (1, "1", None) or ent["type"] not in ("int", "float"):
            continue
        out.append((name, ent["type"], op))
    return out


def info_vector_fields(case, dump_names, op="element_wise_sum"):
    """(remapped [(name, elem)], plain [(attribute, elem, fixed_len)]) of the queried INFO fields combined with an element-wise sum
    (or, op="concatenate", strung together)"""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    remapped, plain = [], []
    for name in case["attributes"]:
        ent = vid.get(name)
        if ent is None or ent.get("combine") != op or "INFO" not in (ent.get("classes") or []) or isinstance(ent["length"], list):
            continue   # (2-D fields such as the allele-specific annotations are not covered)
        elem = "float" if ent["type"] == "float" else "int"
        if name in dump_names:
            remapped.append((name, elem))
        else:
            ln = ent["length"]
            plain.append((name, elem, 0 if (isinstance(ln, str) and not ln.isdigit()) else int(ln)))
    return remapped, plain


def qual_field(case):
    """('QUAL', 'float', op) when the vid mapping gives QUAL a scalar combine operation and QUAL is queried"""
    vid = oracle.load_vid_fields(os.path.join(GOLDEN, case["vid_mapping_file"]))
    ent = vid.get("QUAL")
    if ent and "QUAL" in case["attributes"] and ent.get("combine") in ("median", "sum", "mean"):
        return ("QUAL", "float", ent["combine"])
    return None


def cells_to_columns(cells, num_rows):
    """Hand-made cohort in the generator's dict layout (tools/gdbsyn.generate) from a list of cells
    dict(row, col, end, ref, alt, gt=[..], pl=[..], ad=[..]); cells are put into TileDB's global order (column, then row)."""
    import numpy as np
    cells = sorted(cells, key=lambda c: (c["col"], c["row"]))
    n = len(cells)
    d = dict(coords=np.array([[c["row"], c["col"]] for c in cells], dtype=np.int64).reshape(n, 2),
             end=np.array([c["end"] for c in cells], dtype=np.int64), num_rows=num_rows)
    for key, dt, enc in (("ref", np.uint8, lambda c: np.frombuffer(c["ref"].encode(), dtype=np.uint8)),
                         ("alt", np.uint8, lambda c: np.frombuffer(c["alt"].encode(), dtype=np.uint8)),
                         ("gt", np.int32, lambda c: np.array(c.get("gt", []), dtype=np.int32)),
                         ("pl", np.int32, lambda c: np.array(c.get("pl", []), dtype=np.int32)),
                         ("ad", np.int32, lambda c: np.array(c.get("ad", []), dtype=np.int32))):
        vals = [enc(c) for c in cells]
        off = np.zeros(n, dtype=np.uint64)
        if n:
            off[1:] = np.cumsum([v.nbytes for v in vals])[:-1]
        d[key] = (off, np.concatenate(vals).astype(dt) if n and sum(v.size for v in vals) else np.zeros(0, dtype=dt))
    lo = min(c["col"] for c in cells)
    hi = max(max(c["col"], c["end"]) for c in cells)
    d["col_range"] = (lo, hi)
    return d
