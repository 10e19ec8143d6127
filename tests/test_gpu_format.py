"""GPU: FORMAT collation + BCF typed encoding on the device (GDBGPU_SCAN_BCF_FORMAT, kernels k_fmt_range / k_fmt_pack)
against (1) the UNMODIFIED reference's own BCF output on the golden arrays, (2) the oracle restatement on GDBSYN-1
cohorts, (3) the int32 output of the ordinary scan (decode round trip at a size the python oracle cannot reach)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import bcf_reader, format_blocks, oracle
from oracle import tiledb_fragment as tf
from tools import gdbsyn
from util import (case_fields, case_query_rows, format_field_names, format_order, info_scalar_fields, info_vector_fields,
                  plain_format_fields, qual_field, vcf_format_name)

pytestmark = pytest.mark.gpu
CASES = golden_cases()


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def gpu_blocks(arr, ranges, names, query_rows=None, flags=0, max_sites=0, max_bytes=0):
    """-> [ {field name: (type, nps, bytes)} per site ] through the C ABI"""
    from genomicsdb_b200 import capi
    out = []
    for r in (sorted(tuple(r) for r in ranges) if ranges else [None]):
        sc = capi.Scan(arr, None if r is None else r[0], None if r is None else r[1], query_rows=query_rows,
                       flags=flags | capi.SCAN_BCF_FORMAT)
        try:
            while True:
                bv = sc.next_batch(max_sites, max_bytes)
                if bv is None:
                    break
                assert bv.live is None and bv.fields[0]["data"] is None     # only the blocks travel
                for s in range(bv.n_sites):
                    d = {}
                    for o, nm in enumerate(names):
                        blk = bv.fmt_block(o, s)
                        if blk is not None:
                            d[nm] = blk
                    out.append(d)
        finally:
            sc.close()
    return out


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_blocks_match_reference_bcf(ctx, case):
    from genomicsdb_b200 import capi
    _, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    fields, phased = case_fields(case)
    names = ["GT"] + [f[0] for f in fields]
    fmt_names = format_field_names(case, names)
    flags = (capi.SCAN_MIN_PL_GT if case["produce_GT_with_min_PL_value_for_spanning_deletions"] else 0) | \
            (capi.SCAN_PRODUCE_GT if case["produce_GT_field"] else 0)
    arr = capi.Array.open(ctx, os.path.join(GOLDEN, case["workspace"]), case["array"], fields, gt_phased=phased)
    try:
        for kw in (dict(), dict(max_sites=1)):
            got = gpu_blocks(arr, case["query_column_ranges"], names, query_rows=case_query_rows(case), flags=flags, **kw)
            assert len(got) == len(recs) == case["sites"]
            for rec, g in zip(recs, got):
                want = {k: v[:3] for k, v in rec["fmt"].items() if k in fmt_names}
                g = {k: v for k, v in g.items() if k in fmt_names}
                assert sorted(g) == sorted(want), (rec["pos"], sorted(g), sorted(want))
                for k in want:
                    assert g[k] == want[k], (rec["pos"], k, g[k][:2], want[k][:2], g[k][2].hex(), want[k][2].hex())
    finally:
        arr.close()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_whole_indiv_matches_reference_bcf(ctx, case):
    """remapped + plain FORMAT fields (GQ, SB, PGT, PID, MIN_DP, DP_FORMAT -> DP) collated and encoded on the device; the
    blocks, laid out in the reference's field order behind their typed keys and size descriptors, equal the reference's
    bcf1_t.indiv byte for byte"""
    from genomicsdb_b200 import capi
    text, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    key_ids = bcf_reader.header_dict(text)
    fields, phased = case_fields(case)
    names = ["GT"] + [f[0] for f in fields]
    fmt_names = format_field_names(case, names)
    plain = plain_format_fields(case)
    order = format_order(case, fmt_names, plain)
    all_names = names + [p[0] for p in plain]
    flags = (capi.SCAN_MIN_PL_GT if case["produce_GT_with_min_PL_value_for_spanning_deletions"] else 0) | \
            (capi.SCAN_PRODUCE_GT if case["produce_GT_field"] else 0)
    arr = capi.Array.open(ctx, os.path.join(GOLDEN, case["workspace"]), case["array"], fields, gt_phased=phased, plain=plain)
    try:
        assert arr.info().n_plain_fields == len(plain)
        for kw in (dict(), dict(max_sites=2)):
            got = gpu_blocks(arr, case["query_column_ranges"], all_names, query_rows=case_query_rows(case), flags=flags, **kw)
            assert len(got) == len(recs)
            for rec, g in zip(recs, got):
                blocks = {vcf_format_name(k): v for k, v in g.items() if k in fmt_names or k in [p[0] for p in plain]}
                assert sorted(blocks) == sorted(rec["fmt"]), (rec["pos"], sorted(blocks), sorted(rec["fmt"]))
                assert format_blocks.assemble_indiv(order, blocks, key_ids) == rec["indiv"], rec["pos"]
    finally:
        arr.close()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_info_scalars_match_reference_bcf(ctx, case):
    """median / sum combined INFO scalars (k_info_collect + k_info_reduce) and the GATK-style INFO/DP sum of every record,
    as typed BCF values, against the reference's BCF output"""
    from genomicsdb_b200 import capi
    _, samples, recs = bcf_reader.read_bcf(os.path.join(GOLDEN, "bcf", case["id"] + ".bcf"))
    fields, phased = case_fields(case)
    info_fields = info_scalar_fields(case)
    names = ["GT"] + [f[0] for f in fields]
    vec_remapped, vec_plain = info_vector_fields(case, names)
    qual = qual_field(case)
    have = set(tf.Array(os.path.join(GOLDEN, case["workspace"]), case["array"]).schema.attributes)
    scalars = info_fields + ([qual] if qual else [])
    plain = [(n, e, 1, 1, op) for n, e, op in scalars]                           # INFO scalars (+ QUAL): staged, combined, no FORMAT block
    dp_names = [n for n in ("DP", "MIN_DP", "DP_FORMAT") if n in have]
    dp_at = {n: len(plain) + k + 1 for k, n in enumerate(dp_names)}
    plain += [(n, "int", 1, 1) for n in dp_names]                                  # the DP rule's inputs
    vec_at = {n: len(plain) + k for k, (n, e, fx) in enumerate(vec_plain)}
    plain += [(n, e, fx, 1, "element_wise_sum") for n, e, fx in vec_plain]
    n_remapped = len(fields)
    mask = sum(1 << [f[0] for f in fields].index(n) for n, _ in vec_remapped)
    n_slots = n_remapped + len(plain)
    arr = capi.Array.open(ctx, os.path.join(GOLDEN, case["workspace"]), case["array"], fields, gt_phased=phased, plain=plain)
    try:
        for kw in (dict(), dict(max_sites=1)):
            got = []
            for r in sorted(tuple(r) for r in case["query_column_ranges"]):
                sc = capi.Scan(arr, r[0], r[1], query_rows=case_query_rows(case), flags=capi.SCAN_BCF_FORMAT,
                               dp_fields=(dp_at.get("DP", 0), dp_at.get("MIN_DP", 0), dp_at.get("DP_FORMAT", 0)), esum_remapped_mask=mask)
                while True:
                    bv = sc.next_batch(**kw)
                    if bv is None:
                        break
                    for s in range(bv.n_sites):
                        vecs = {n: bv.info_vector([f[0] for f in fields].index(n), s, n_slots) for n, _ in vec_remapped}
                        vecs.update({n: bv.info_vector(n_remapped + vec_at[n], s, n_slots) for n, _, _ in vec_plain})
                        got.append(dict(begin=int(bv.site_begin[s]), end=int(bv.site_end[s]), remap=bool(bv.site_flags[s] & capi.SITE_REMAPPED),
                                        dp=int(bv.info_dp_sum[s]), vecs={k: v for k, v in vecs.items() if len(v)},
                                        vals={scalars[k][0]: int(bv.info_value[k, s]) for k in range(len(scalars)) if bv.info_valid[k, s]}))
                sc.close()
            assert len(got) == len(recs)
            for rec, g in zip(recs, got):
                assert rec["qual_bits"] == g["vals"].pop("QUAL", format_blocks.F32_MISSING), g["begin"]
                want = {k: v for k, v in rec["info"].items() if k in [f[0] for f in info_fields]}
                assert sorted(g["vals"]) == sorted(want), (g["begin"], sorted(g["vals"]), sorted(want))
                for name, elem, _ in info_fields:
                    if name in want:
                        bt, n, raw = want[name]
                        assert bytes([n << 4 | bt]) + raw == format_blocks.enc_info_value(g["vals"][name], elem), (g["begin"], name)
                for name, elem in vec_remapped + [(n, e) for n, e, _ in vec_plain]:
                    assert (name in rec["info"]) == (name in g["vecs"]), (g["begin"], name)
                    if name in g["vecs"]:
                        bt, n, raw = rec["info"][name]
                        words = [int(x) for x in g["vecs"][name]]
                        enc = format_blocks.enc_info_vector(words, elem) if (n > 1 or elem == "float") else \
                            format_blocks.enc_int1(words[0] - (1 << 32) if words[0] >= (1 << 31) else words[0])
                        assert enc == format_blocks.enc_size(n, bt) + raw, (g["begin"], name)
                assert ("END" in rec["info"]) == (g["end"] > g["begin"])
                assert ("DP" in rec["info"]) == (g["dp"] > 0 and g["remap"]), (g["begin"], g["dp"])
                if "DP" in rec["info"]:
                    bt, n, raw = rec["info"]["DP"]
                    assert bytes([n << 4 | bt]) + raw == format_blocks.enc_int1(g["dp"])
    finally:
        arr.close()


def oracle_blocks(d, ranges, rows, produce_gt):
    cols = gdbsyn.to_oracle_columns(d)
    dump, n = oracle.combine_query(cols, ranges, query_rows=rows)
    names, sites = oracle.parse_dump(dump)
    q = sorted(set(r for r in (rows if rows is not None else range(d["num_rows"])) if r < d["num_rows"]))
    return [format_blocks.site_blocks(s, names, q, {"AD": "int", "PL": "int"}, False, produce_gt) for s in sites]


@pytest.mark.parametrize("shape,produce_gt", [(dict(n_samples=50, n_sites=120, spacing=100, p_var=0.1, a_pool=8), True),
                                              (dict(n_samples=300, n_sites=60, spacing=100, p_var=0.1, a_pool=3), False),
                                              (dict(n_samples=7, n_sites=300, spacing=3, p_var=0.4, a_pool=5), True),
                                              (dict(n_samples=1, n_sites=50, spacing=10, p_var=0.3, a_pool=4), True)],
                         ids=["50x120", "300x60", "7x300dense", "1sample"])
def test_blocks_match_oracle_on_synthetic(ctx, shape, produce_gt):
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(**shape)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        flags = capi.SCAN_PRODUCE_GT if produce_gt else 0
        for ranges, rows, kw in (([], None, {}), ([list(d["col_range"])], None, dict(max_bytes=100000)),
                                 ([[1500, 4000]], [0, 2, 3, 5, 40, 41, 290], dict(max_sites=2))):
            want = oracle_blocks(d, ranges, rows, produce_gt)
            got = gpu_blocks(arr, ranges, ["GT", "AD", "PL"], query_rows=rows, flags=flags, **kw)
            assert len(got) == len(want)
            for i, (g, w) in enumerate(zip(got, want)):
                assert sorted(g) == sorted(w), (i, sorted(g), sorted(w))
                for k in w:
                    assert g[k] == w[k][:3], (i, k, g[k][:2], w[k][:2])
    finally:
        arr.close()


def test_plain_fields_and_dp_rule_match_oracle_on_synthetic(ctx):
    """GQ / MIN_DP blocks and the INFO/DP sum on a GDBSYN-1 cohort (MIN_DP exists on reference blocks only, DP everywhere)"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=200, n_sites=80, spacing=100, p_var=0.15, a_pool=8, with_fixed=True)
    arr = gdbsyn.to_gpu_array(ctx, d, plain=True)
    try:
        view = gdbsyn.OracleArrayView(d)
        plain = format_blocks.PlainColumns(view, d["end"], [("DP", "int", 1), ("GQ", "int", 1), ("MIN_DP", "int", 1)])
        fmt_plain = format_blocks.PlainColumns(view, d["end"], [("GQ", "int", 1), ("MIN_DP", "int", 1)])
        for ranges, rows, kw in (([], None, {}), ([[2500, 6000]], list(range(0, 200, 3)), dict(max_sites=3))):
            cols = gdbsyn.to_oracle_columns(d)
            dump, n = oracle.combine_query(cols, ranges, query_rows=rows)
            names, sites = oracle.parse_dump(dump)
            q = sorted(rows) if rows is not None else list(range(200))
            got_blocks, got_dp = [], []
            for r in (ranges or [None]):
                sc = capi.Scan(arr, None if r is None else r[0], None if r is None else r[1], query_rows=rows, flags=capi.SCAN_BCF_FORMAT,
                               dp_fields=gdbsyn.PLAIN_DP_RULE)
                while True:
                    bv = sc.next_batch(**kw)
                    if bv is None:
                        break
                    for s in range(bv.n_sites):
                        got_blocks.append({nm: bv.fmt_block(3 + k, s) for k, nm in enumerate(["DP", "GQ", "MIN_DP"])})
                        got_dp.append(int(bv.info_dp_sum[s]))
                sc.close()
            assert len(got_blocks) == len(sites)
            for i, site in enumerate(sites):
                want = format_blocks.plain_site_blocks(site, q, fmt_plain)
                assert got_blocks[i]["DP"] is None                      # no_format: staged for the DP rule only
                for nm in ("GQ", "MIN_DP"):
                    assert (got_blocks[i][nm] is None) == (nm not in want), (i, nm)
                    if nm in want:
                        assert got_blocks[i][nm] == want[nm][:3], (i, nm)
                _, dp_sum = format_blocks.info_site_values(site, plain, [], ("DP", "MIN_DP", None))
                assert got_dp[i] == dp_sum, (i, got_dp[i], dp_sum)
    finally:
        arr.close()


def test_blocks_decode_to_the_int32_output(ctx):
    """10 k samples: every block, widened again, equals the collated int32 vectors of the ordinary scan"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=10000, n_sites=24, spacing=1000, p_var=0.1, a_pool=8)
    arr = gdbsyn.to_gpu_array(ctx, d)
    widths = set()
    try:
        a, b = capi.Scan(arr), capi.Scan(arr, flags=capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT)
        n = 0
        while True:
            x = a.next_batch(5, 0)
            y = b.next_batch(5, 0)
            if x is None:
                assert y is None
                break
            assert x.n_sites == y.n_sites
            nq = x.nq
            for s in range(x.n_sites):
                for o in range(3):
                    f = x.fields[o]
                    stride, ln = int(f["stride"][s]), f["len"][s].astype(np.int64)
                    blk = y.fmt_block(o, s)
                    if ln.max() == 0:
                        assert blk is None
                        continue
                    t, nps, raw = blk
                    widths.add((o, t))
                    base = int(f["base"][s])
                    v = f["data"][base:base + nq * stride].view(np.int32).reshape(nq, stride)[:, :nps].astype(np.int64).copy()
                    if o == 0:
                        valid = (v != format_blocks.I32_MISSING) & (v != format_blocks.I32_VEND)
                        v = np.where(valid, (v + 1) << 1, v)
                    else:
                        v[ln == 0, 0] = format_blocks.I32_MISSING
                    dt = {1: "<i1", 2: "<i2", 3: "<i4"}[t]
                    w = np.frombuffer(raw, dtype=dt).reshape(nq, nps).astype(np.int64)
                    lo = {1: -128, 2: -32768, 3: format_blocks.I32_MISSING}[t]
                    w = np.where(w == lo, format_blocks.I32_MISSING, np.where(w == lo + 1, format_blocks.I32_VEND, w))
                    assert np.array_equal(v, w), (n + s, o)
                    valid = v[(v != format_blocks.I32_MISSING) & (v != format_blocks.I32_VEND)]
                    if t == 1:
                        assert valid.max() <= 127 and valid.min() > -127
                    elif t == 2:
                        assert valid.max() <= 32767 and not (valid.max() <= 127 and valid.min() > -127)
            n += x.n_sites
        a.close()
        b.close()
    finally:
        arr.close()
    assert (0, 1) in widths and (2, 2) in widths     # GT int8, PL int16
