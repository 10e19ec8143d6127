"""GPU: the paths round 1 left untested -- genotype-length fields dropped above 50 ALT alleles (with the reference's warning),
wide sites through the generic remap kernel, a five-field query, the illegal-overlap error and every engine limit.
Ground truth is the unmodified reference (oracle/_ref) wherever it can run the case, else the plain-C restatement."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO
from oracle import oracle
from test_gpu_scale import OURS, REF, compare_with_reference
from tools import gdbsyn, tiledb_writer
from util import cells_to_columns

pytestmark = pytest.mark.gpu

WARNING = "has too many alleles in the combined VCF record"


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def needs_ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")


def test_pl_dropped_above_50_alt_alleles_with_the_reference_warning(tmp_path):
    """variant_operations.cc:613-629: sites with more than 50 merged ALT alleles lose PL; one stderr line per site and field"""
    needs_ref()
    d = gdbsyn.generate(n_samples=1500, n_sites=10, spacing=100, p_var=0.5, a_pool=64)
    compare_with_reference(tmp_path, d, n_intervals=2)
    q = str(tmp_path / "ws" / "q0.json")
    ref = subprocess.run([REF, "-j", q, "--mode", "dump", "-o", "/dev/null"], capture_output=True, text=True)
    ours = subprocess.run([OURS, "-j", q, "-o", "/dev/null"], capture_output=True, text=True)
    want = sorted(l for l in ref.stderr.splitlines() if WARNING in l)
    got = sorted(l for l in ours.stderr.splitlines() if WARNING in l)
    assert len(want) >= 3, "the case must contain sites above the limit"
    assert got == want


def test_wide_sites_go_through_the_generic_kernel(tmp_path):
    """45 to 50 merged alleles: above the tile kernels' 44, PL kept (1275 genotypes per sample)"""
    needs_ref()
    d = gdbsyn.generate(n_samples=500, n_sites=8, spacing=100, p_var=0.5, a_pool=47)
    compare_with_reference(tmp_path, d, n_intervals=2)
    cols = gdbsyn.to_oracle_columns(d)
    dump, n = oracle.combine_query(cols, [])
    widths = [len(s["alleles"]) for s in oracle.parse_dump(dump)[1]]
    assert max(widths) > 44 and max(widths) <= 51, widths


def five_fields(d):
    ad, pl = d["ad"], d["pl"]
    return [dict(name="AD", kind="R", elem="int", col=ad), dict(name="PL", kind="G", elem="int", col=pl),
            dict(name="AD2", kind="R", elem="float", col=ad), dict(name="PL2", kind="G", elem="int", col=pl),
            dict(name="AD3", kind="R", elem="int", col=ad)]


def test_five_remapped_fields(ctx):
    """more remapped fields than the tile kernels take (4): every site goes through k_remap_generic; against the restatement"""
    from genomicsdb_b200 import capi
    from tools import dump
    d = gdbsyn.generate(n_samples=150, n_sites=40, spacing=100, p_var=0.2, a_pool=8)
    fields = five_fields(d)

    def el(col):
        return np.concatenate([col[0].astype(np.int64), [col[1].nbytes]]) // 4
    ofields = [dict(name=f["name"], kind=f["kind"], elem=f["elem"], off=el(f["col"]), val=f["col"][1].view(np.uint32)) for f in fields]
    cols = oracle.Columns(d["coords"][:, 0], d["coords"][:, 1], d["end"], np.concatenate([d["ref"][0].astype(np.int64), [d["ref"][1].nbytes]]), d["ref"][1],
                          np.concatenate([d["alt"][0].astype(np.int64), [d["alt"][1].nbytes]]), d["alt"][1], el(d["gt"]), d["gt"][1], False, ofields, d["num_rows"], 0)
    want, n = oracle.combine_query(cols, [])
    arr = capi.Array.from_columns(ctx, d["coords"], d["end"], d["ref"], d["alt"], d["gt"], fields, 0, d["num_rows"])
    try:
        got, n_got = dump.combine_dump(arr, [], ["GT"] + [f["name"] for f in fields])
    finally:
        arr.close()
    assert n_got == n and got == want


def overlap_cohort():
    """sample 1: a SNP call that claims [100, 110] is overlapped by a call of the same sample at 105 -- neither a deletion nor a
    reference block: query_variants.cc:533-535 throws"""
    cells = [dict(row=0, col=100, end=100, ref="A", alt="T|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0]),
             dict(row=1, col=100, end=110, ref="A", alt="T|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0]),
             dict(row=1, col=105, end=105, ref="C", alt="G|&", gt=[1, 1], pl=[90, 20, 0, 30, 40, 50], ad=[0, 7, 0]),
             dict(row=2, col=120, end=120, ref="G", alt="&", gt=[0, 0], pl=[0, 5, 50], ad=[])]
    return cells_to_columns(cells, 3)


def test_illegal_overlap_is_an_error_on_both_sides(ctx, tmp_path):
    from genomicsdb_b200 import capi
    d = overlap_cohort()
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        with pytest.raises(capi.GdbGpuError) as ei:
            sc = capi.Scan(arr)
            while sc.next_batch() is not None:
                pass
        assert ei.value.code == capi.ERR_OVERLAP and "overlapping variants" in str(ei.value)
        # an interval that ends before the offending call is fine
        sc = capi.Scan(arr, 90, 104)
        assert sc.next_batch().n_sites >= 1
        sc.close()
    finally:
        arr.close()
    if os.path.exists(REF):   # the mirror raises where the reference does
        info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d, capacity=1000)
        q = tiledb_writer.write_query_json(str(tmp_path / "ws" / "q.json"), info)
        ref = subprocess.run([REF, "-j", q, "--mode", "dump", "-o", "/dev/null"], capture_output=True, text=True)
        ours = subprocess.run([OURS, "-j", q, "-o", "/dev/null"], capture_output=True, text=True)
        assert ref.returncode != 0 and "Unhandled overlapping variants" in ref.stderr
        assert ours.returncode != 0 and "overlapping variants" in ours.stderr


def limit_error(ctx, d, fields=None, scan=True, **scan_kw):
    from genomicsdb_b200 import capi
    with pytest.raises(capi.GdbGpuError) as ei:
        if fields is None:
            arr = gdbsyn.to_gpu_array(ctx, d)
        else:
            arr = capi.Array.from_columns(ctx, d["coords"], d["end"], d["ref"], d["alt"], d["gt"], fields, 0, d["num_rows"])
        try:
            if scan:
                sc = capi.Scan(arr, **scan_kw)
                try:
                    while sc.next_batch() is not None:
                        pass
                finally:
                    sc.close()
        finally:
            arr.close()
    assert ei.value.code == capi.ERR_LIMIT, str(ei.value)
    return str(ei.value)


def test_limit_more_than_15_remapped_fields(ctx):
    d = gdbsyn.generate(n_samples=5, n_sites=3, spacing=100, p_var=0.5, a_pool=3)
    fields = [dict(name="F%d" % i, kind="R", elem="int", col=d["ad"]) for i in range(16)]
    assert "15" in limit_error(ctx, d, fields=fields, scan=False)


def test_limit_more_than_512_distinct_alleles_at_a_site(ctx):
    bases = "ACGT"
    cells = []
    for r in range(600):
        alt = "A" + "".join(bases[(r >> (2 * k)) & 3] for k in range(5))
        cells.append(dict(row=r, col=100, end=100, ref="A", alt=alt + "|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0]))
    assert "512" in limit_error(ctx, cells_to_columns(cells, 600))


def test_limit_more_than_1023_alt_alleles_in_a_call(ctx):
    bases = "ACGT"
    alts = ["A" + "".join(bases[(i >> (2 * k)) & 3] for k in range(6)) for i in range(1100)]
    cells = [dict(row=0, col=100, end=100, ref="A", alt="|".join(alts) + "|&", gt=[0, 1], pl=[1, 2, 3], ad=[1, 2]),
             dict(row=1, col=100, end=100, ref="A", alt="T|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0])]
    assert "1023" in limit_error(ctx, cells_to_columns(cells, 2))


def test_limit_ploidy_above_8(ctx):
    cells = [dict(row=0, col=100, end=100, ref="A", alt="T|&", gt=[0, 1, 0, 1, 0, 1, 0, 1, 0], pl=list(range(55)), ad=[3, 4, 0]),
             dict(row=1, col=100, end=100, ref="A", alt="G|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0])]
    assert "ploidy" in limit_error(ctx, cells_to_columns(cells, 2))


def test_limit_column_span(ctx):
    cells = [dict(row=0, col=100, end=100, ref="A", alt="T|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0]),
             dict(row=1, col=(1 << 38), end=(1 << 38), ref="A", alt="G|&", gt=[0, 1], pl=[10, 0, 20, 30, 40, 50], ad=[3, 4, 0])]
    assert "span" in limit_error(ctx, cells_to_columns(cells, 2))


def test_query_far_left_of_the_data_costs_nothing(ctx):
    """the breakpoint bitmap is based at the data, not at the query's begin: [0, 2^40] over a small cohort works and equals the full scan"""
    from tools import dump
    d = gdbsyn.generate(n_samples=30, n_sites=20, spacing=100, p_var=0.2, a_pool=4, first_col=5_000_000_000)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        full, n = dump.combine_dump(arr, [], ["GT", "AD", "PL"])
        wide, n2 = dump.combine_dump(arr, [[0, 1 << 40]], ["GT", "AD", "PL"])
    finally:
        arr.close()
    cols = gdbsyn.to_oracle_columns(d)
    want, n3 = oracle.combine_query(cols, [[0, 1 << 40]])
    assert n2 == n3 and wide == want
    assert n == n2
