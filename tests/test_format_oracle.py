"""CPU: pins the FORMAT collation + BCF typed-encoding restatement (oracle/format_blocks.py) against the reference.

tests/golden/bcf/<case>.bcf is the UNMODIFIED reference's own `gt_mpi_gather --produce-Broad-GVCF -O bu` output on the
golden arrays (oracle/make_golden_bcf.py).  For every record the blocks of GT and of every allele-dependent FORMAT field
(descriptor bytes and data bytes as htslib laid them into bcf1_t.indiv) must equal what the restatement derives from the
reference's hot-path dump of the same query.
"""
import os

import pytest

from conftest import GOLDEN, golden_cases
from oracle import bcf_reader, format_blocks, oracle
from oracle import tiledb_fragment as tf
from util import (case_fields, case_query_rows, format_field_names, format_order, info_scalar_fields, info_vector_fields,
                  plain_format_fields, qual_field, vcf_format_name)

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_format_blocks_match_reference_bcf(case):
    _, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    names, sites = oracle.parse_dump(open(os.path.join(GOLDEN, case["dump"]), "rb").read())
    assert len(recs) == len(sites) == case["sites"]
    fields, phased = case_fields(case)
    elem = {f[0]: f[2] for f in fields}
    fmt_names = format_field_names(case, names)
    rows = [r for r in case_query_rows(case) if r < len(samples)]
    assert len(rows) == len(samples)
    n_blocks = 0
    for rec, site in zip(recs, sites):
        want = {k: v for k, v in rec["fmt"].items() if k in fmt_names}
        got = format_blocks.site_blocks(site, fmt_names, rows, elem, phased, case["produce_GT_field"])
        assert sorted(got) == sorted(want), (site["begin"], sorted(got), sorted(want))
        for k, (bt, nps, data, hdr) in want.items():
            g = got[k]
            assert (g[0], g[1]) == (bt, nps), (site["begin"], k)
            assert g[2] == data, (site["begin"], k, g[2].hex(), data.hex())
            assert hdr.endswith(g[3]), (site["begin"], k)     # hdr = typed key + descriptor
            n_blocks += 1
    assert n_blocks > 0


@pytest.mark.parametrize("java", [False, True], ids=["htslib", "java"])
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_whole_indiv_matches_reference_bcf(case, java):
    """remapped + plain FORMAT fields, assembled in the reference's field order, equal bcf1_t.indiv byte for byte -- with
    htslib's conventions (gt_mpi_gather) and with the JNI path's (missing instead of vector_end, no-call GT)"""
    text, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf_java" if java else "bcf", case["id"] + ".bcf"))
    key_ids = bcf_reader.header_dict(text)
    names, sites = oracle.parse_dump(open(os.path.join(GOLDEN, case["dump"]), "rb").read())
    fields, phased = case_fields(case)
    elem = {f[0]: f[2] for f in fields}
    fmt_names = format_field_names(case, names)
    plain_fields = plain_format_fields(case)
    arr = tf.Array(os.path.join(GOLDEN, case["workspace"]), case["array"])
    plain = format_blocks.PlainColumns(arr, arr.fixed("END")[:, 0], plain_fields)
    rows = [r for r in case_query_rows(case) if r < len(samples)]
    order = format_order(case, fmt_names, plain_fields)
    for rec, site in zip(recs, sites):
        blocks = format_blocks.site_blocks(site, fmt_names, rows, elem, phased, case["produce_GT_field"], java=java)
        for attr, b in format_blocks.plain_site_blocks(site, rows, plain, java=java).items():
            blocks[vcf_format_name(attr)] = b
        assert sorted(blocks) == sorted(rec["fmt"]), (site["begin"], sorted(blocks), sorted(rec["fmt"]))
        assert [n for n in order if n in blocks] == rec["fmt_order"], (site["begin"], order, rec["fmt_order"])
        assert format_blocks.assemble_indiv(order, blocks, key_ids) == rec["indiv"], site["begin"]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_info_scalars_match_reference_bcf(case):
    """END, the median / sum combined INFO scalars and the GATK-style INFO/DP of every record"""
    _, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    names, sites = oracle.parse_dump(open(os.path.join(GOLDEN, case["dump"]), "rb").read())
    info_fields = info_scalar_fields(case)
    vec_remapped, vec_plain = info_vector_fields(case, names)
    qual = qual_field(case)
    arr = tf.Array(os.path.join(GOLDEN, case["workspace"]), case["array"])
    dp_names = tuple(n if n in arr.schema.attributes else None for n in ("DP", "MIN_DP", "DP_FORMAT"))
    extra = [(n, e, 1) for n, e, _ in info_fields] + [(n, "int", 1) for n in dp_names if n] + list(vec_plain) + ([("QUAL", "float", 1)] if qual else [])
    plain = format_blocks.PlainColumns(arr, arr.fixed("END")[:, 0], extra)
    n_checked = 0
    for rec, site in zip(recs, sites):
        vals, dp_sum = format_blocks.info_site_values(site, plain, info_fields + ([qual] if qual else []), dp_names)
        if qual:
            assert rec["qual_bits"] == vals.pop("QUAL", format_blocks.F32_MISSING), site["begin"]
        else:
            assert rec["qual_bits"] == format_blocks.F32_MISSING
        vecs = format_blocks.info_vector_sums(site, plain, vec_remapped, [(n, e) for n, e, _ in vec_plain])
        for (name, elem) in vec_remapped + [(n, e) for n, e, _ in vec_plain]:
            assert (name in rec["info"]) == (name in vecs), (site["begin"], name)
            if name in vecs:
                bt, n, raw = rec["info"][name]
                want = (format_blocks.enc_size(n, bt) + raw)
                got = format_blocks.enc_info_vector(vecs[name], elem)
                if n == 1:
                    got = got if elem == "float" else format_blocks.enc_int1(vecs[name][0] - (1 << 32) if vecs[name][0] >= (1 << 31) else vecs[name][0])
                assert got == want, (site["begin"], name, got.hex(), want.hex())
                n_checked += 1
        want = {k: v for k, v in rec["info"].items() if k in [f[0] for f in info_fields]}
        assert sorted(vals) == sorted(want), (site["begin"], sorted(vals), sorted(want))
        for name, elem, _ in info_fields:
            if name in want:
                bt, n, raw = want[name]
                assert bytes([n << 4 | bt]) + raw == format_blocks.enc_info_value(vals[name], elem), (site["begin"], name)
                n_checked += 1
        assert ("END" in rec["info"]) == (site["end"] > site["begin"])
        if "END" in rec["info"]:
            bt, n, raw = rec["info"]["END"]
            assert bytes([n << 4 | bt]) + raw == format_blocks.enc_int1(site["end"] + 1)
        assert ("DP" in rec["info"]) == (dp_sum > 0 and bool(site["remap"])), (site["begin"], dp_sum)
        if "DP" in rec["info"]:
            bt, n, raw = rec["info"]["DP"]
            assert bytes([n << 4 | bt]) + raw == format_blocks.enc_int1(dp_sum), (site["begin"], dp_sum)
    assert n_checked or not any(r["info"].keys() - {"END", "DP"} for r in recs)


def test_enc_int1_and_size_known_answers():
    # htslib vcf.h bcf_enc_int1 / bcf_enc_size thresholds
    assert format_blocks.enc_int1(127) == bytes([0x11, 0x7F])
    assert format_blocks.enc_int1(-120) == bytes([0x11, 0x88])
    assert format_blocks.enc_int1(-121) == bytes([0x12, 0x87, 0xFF])
    assert format_blocks.enc_int1(128) == bytes([0x12, 0x80, 0x00])
    assert format_blocks.enc_int1(40000) == bytes([0x13, 0x40, 0x9C, 0x00, 0x00])
    assert format_blocks.enc_size(3, 1) == bytes([0x31])
    assert format_blocks.enc_size(55, 2) == bytes([0xF2, 0x11, 55])
    bt, desc, data = format_blocks.enc_vint([0, 127, -126, format_blocks.I32_MISSING, format_blocks.I32_VEND, 5], 3)
    assert (bt, desc, data) == (1, bytes([0x31]), bytes([0, 127, 0x82, 0x80, 0x81, 5]))
    bt, desc, data = format_blocks.enc_vint([0, 128, 1, 2], 2)
    assert bt == 2 and data == bytes([0, 0, 128, 0, 1, 0, 2, 0])
    bt, desc, data = format_blocks.enc_vint([0, -127, 1, 2], 2)
    assert bt == 2
    bt, desc, data = format_blocks.enc_vint([0, 32768, 1, 2], 2)
    assert bt == 3


def test_info_concatenate_matches_reference_bcf():
    """the reference CI's info_ops1 scenario (vid_info_ops1.json: var_field combined with "concatenate")"""
    case = [c for c in CASES if c["id"] == "info_ops1_at_0"][0]
    _, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    names, sites = oracle.parse_dump(open(os.path.join(GOLDEN, case["dump"]), "rb").read())
    _, cat_plain = info_vector_fields(case, names, op="concatenate")
    assert [n for n, _, _ in cat_plain] == ["var_field"]
    arr = tf.Array(os.path.join(GOLDEN, case["workspace"]), case["array"])
    plain = format_blocks.PlainColumns(arr, arr.fixed("END")[:, 0], list(cat_plain))
    n_checked = 0
    for rec, site in zip(recs, sites):
        vecs = format_blocks.info_vector_concat(site, plain, [(n, e) for n, e, _ in cat_plain])
        for name, elem, _ in cat_plain:
            assert (name in rec["info"]) == (name in vecs)
            if name in vecs:
                bt, n, raw = rec["info"][name]
                assert format_blocks.enc_info_vector(vecs[name], elem) == format_blocks.enc_size(n, bt) + raw
                n_checked += 1
    assert n_checked == 1


@pytest.mark.parametrize("produce_gt", [False, True], ids=["no-GT", "GT"])
def test_format_blocks_match_reference_bcf_on_a_cohort_with_deletions(tmp_path, produce_gt):
    """the two restatements chained (combine_oracle.c -> format_blocks.py) against the unmodified reference's BCF on a synthetic
    cohort whose variant calls are deletions to a third (tools/gdbsyn_deletions.py): GT / AD / PL blocks of every record"""
    import json
    import subprocess
    from conftest import REPO
    from tools import gdbsyn, gdbsyn_deletions, tiledb_writer
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=70, n_sites=90, spacing=100, p_var=0.2, a_pool=5, with_fixed=True), share=0.35)
    info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(str(tmp_path / "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL"))
    if produce_gt:
        q = json.load(open(qf))
        q["produce_GT_field"] = True
        json.dump(q, open(qf, "w"))
    bcf = str(tmp_path / "ref.bcf")
    with open(bcf, "wb") as f:
        p = subprocess.run([ref_tool, "-j", qf, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-600:]
    _, samples, recs = bcf_reader.read_bcf(bcf)
    cols = oracle.load_columns(info["workspace"], "syn", info["vid"], ["END", "REF", "ALT", "GT", "AD", "PL"])
    dump, n = oracle.combine_query(cols, [[lo, hi]])
    names, sites = oracle.parse_dump(dump)
    assert len(recs) == len(sites) == n and n > 200
    rows = list(range(len(samples)))
    for rec, site in zip(recs, sites):
        want = {k: v for k, v in rec["fmt"].items() if k in ("GT", "AD", "PL")}
        got = format_blocks.site_blocks(site, ["GT", "AD", "PL"], rows, {"AD": "int", "PL": "int"}, False, produce_gt)
        assert sorted(got) == sorted(want), (site["begin"], sorted(got), sorted(want))
        for k, (bt, nps, data, hdr) in want.items():
            assert (got[k][0], got[k][1]) == (bt, nps) and got[k][2] == data, (site["begin"], k)
