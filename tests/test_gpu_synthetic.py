"""GPU: CUDA path vs the oracle on GDBSYN-1 cohorts (same seeded inputs), sizes the oracle finishes in seconds."""
import os

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
    from tools import dump
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


def test_empty_inputs(ctx):
    """empty array, interval past the data, interval before the data, no queried row inside the array: zero sites, no error --
    in the plain and in the BCF-typed mode"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=20, n_sites=30, spacing=100, p_var=0.2, a_pool=4)
    arr = gdbsyn.to_gpu_array(ctx, d)
    lo, hi = d["col_range"]
    try:
        for flags in (0, capi.SCAN_BCF_FORMAT, capi.SCAN_DEVICE_OUTPUT):
            for a, b, rows in ((hi + 1000, hi + 2000, None), (0, 3, None), (lo, hi, [100, 200, 4000])):
                sc = capi.Scan(arr, a, b, query_rows=rows, flags=flags)
                assert sc.info().n_sites == 0
                assert sc.next_batch() is None
                assert sc.next_batch() is None          # asking again after the end is fine
                sc.close()
    finally:
        arr.close()
    e = dict(coords=np.zeros((0, 2), dtype=np.int64), end=np.zeros(0, dtype=np.int64), num_rows=5)
    for k, dt in (("ref", np.uint8), ("alt", np.uint8), ("gt", np.int32), ("pl", np.int32), ("ad", np.int32)):
        e[k] = (np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=dt))
    arr = gdbsyn.to_gpu_array(ctx, e)
    try:
        assert arr.info().n_cells == 0
        for flags in (0, capi.SCAN_BCF_FORMAT):
            sc = capi.Scan(arr, flags=flags)
            assert sc.next_batch() is None
            sc.close()
    finally:
        arr.close()


def test_seek_and_rescan_give_the_same_batches(ctx):
    """gdbgpu_scan_seek (resume after a consumer-side overflow) with the prefetch in flight: the re-read sites are identical"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=64, n_sites=60, spacing=100, p_var=0.15, a_pool=8)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        sc = capi.Scan(arr)
        first = sc.next_batch(7, 0)
        snap = [first.fields[2]["data"].copy(), first.site_begin.copy()]
        second = sc.next_batch(7, 0)
        s2 = [second.fields[2]["data"].copy(), second.site_begin.copy()]
        sc.seek(0)
        again = sc.next_batch(7, 0)
        assert np.array_equal(again.site_begin, snap[1]) and np.array_equal(again.fields[2]["data"], snap[0])
        sc.seek(3)                                   # into the middle of a batch
        mid = sc.next_batch(4, 0)
        assert np.array_equal(mid.site_begin, snap[1][3:7])
        nxt = sc.next_batch(7, 0)
        assert np.array_equal(nxt.site_begin, s2[1]) and np.array_equal(nxt.fields[2]["data"], s2[0])
        sc.close()
    finally:
        arr.close()


def test_full_width_100k_samples(ctx):
    """BASELINE config #4's width (100 k samples, up to 8 ALT alleles): (1) a 300-row query over the 100 k-row array equals the
    oracle byte for byte (sites are cut by all samples, the union runs over the queried ones); (2) over all rows the remapped
    output does not depend on how the scan is batched, every row's PL vector has the length its ploidy and the merged allele
    count dictate, and every remapped value is one of the call's input values or missing."""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=100000, n_sites=10000, spacing=1000, p_var=0.1, a_pool=8, site_lo=100, site_hi=124)
    lo, hi = d["col_range"]
    rng = np.random.RandomState(7)
    rows = sorted(set(int(r) for r in rng.randint(0, 100000, size=300)))
    both(ctx, d, [[lo, hi]], rows=rows)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        def collect(max_bytes):
            sc = capi.Scan(arr, lo, hi)
            out = []
            while True:
                bv = sc.next_batch(0, max_bytes)
                if bv is None:
                    break
                for s in range(bv.n_sites):
                    rec = [int(bv.site_begin[s]), int(bv.site_end[s]), int(bv.n_alleles[s]), int(bv.site_flags[s])]
                    for f in bv.fields:
                        st, base = int(f["stride"][s]), int(f["base"][s])
                        need = int(f["len"][s].max())
                        rec.append(f["data"][base:base + bv.nq * st].reshape(bv.nq, st)[:, :need].copy())
                        rec.append(f["len"][s].copy())
                    out.append(rec)
            sc.close()
            return out
        a, b = collect(1 << 30), collect(64 << 20)
        assert len(a) == len(b) == 49
        n_remapped = 0
        for x, y in zip(a, b):
            assert x[:4] == y[:4]
            for u, v in zip(x[4:], y[4:]):
                assert np.array_equal(u, v)
            n_alleles, flags = x[2], x[3]
            pl, pl_len = x[8], x[9]
            if flags & capi.SITE_REMAPPED:
                n_remapped += 1
                assert set(np.unique(pl_len)) <= {0, n_alleles * (n_alleles + 1) // 2}
                assert (x[7] == n_alleles).all() or set(np.unique(x[7])) <= {0, n_alleles}        # AD: one value per merged allele
                vals = pl.view(np.int32)
                assert ((vals >= 0) & (vals <= 3000) | (vals == -2147483648) | (vals == -2147483647)).all()
        assert n_remapped == 24
    finally:
        arr.close()


@pytest.mark.parametrize("capacity", [1000, 37])
def test_compressed_fragment_is_inflated_on_the_device(ctx, tmp_path, capacity):
    """a GZIP-compressed fragment (one zlib stream per tile, dynamic Huffman blocks at 1000 cells per tile) written in the
    reference's on-disk format: gdbgpu_array_open inflates the tiles on the device (k_inflate); the combine over it equals
    the oracle over the generated columns, and equals the open of the uncompressed twin"""
    from genomicsdb_b200 import capi
    from tools import dump
    from tools import tiledb_writer
    d = gdbsyn.generate(n_samples=300, n_sites=200, spacing=100, p_var=0.1, a_pool=8, with_fixed=True)
    want, n = oracle.combine_query(gdbsyn.to_oracle_columns(d), [])
    got = {}
    for comp in (True, False):
        ws = str(tmp_path / ("ws%d" % comp))
        tiledb_writer.write_array(ws, "syn", d, capacity=capacity, compress=comp)
        arr = capi.Array.open(ctx, ws, "syn", [("AD", "R", "int"), ("PL", "G", "int")], plain=[("GQ", "int", 1), ("MIN_DP", "int", 1)])
        try:
            info = arr.info()
            assert info.n_cells == len(d["end"])
            got[comp] = (dump.combine_dump(arr, [], ["GT", "AD", "PL"]), info.staged_bytes)
        finally:
            arr.close()
    assert got[True][0][1] == n
    assert got[True][0][0] == want
    assert got[False][0][0] == want
    assert got[True][1] < got[False][1]          # fewer bytes crossed PCIe


def test_corrupt_compressed_tile_is_reported(ctx, tmp_path):
    from genomicsdb_b200 import capi
    from tools import tiledb_writer
    d = gdbsyn.generate(n_samples=50, n_sites=100, spacing=100, p_var=0.1, a_pool=4, with_fixed=True)
    ws = str(tmp_path / "ws")
    tiledb_writer.write_array(ws, "syn", d, capacity=500, compress=True)
    frag = [p for p in os.listdir(os.path.join(ws, "syn")) if p.startswith("__") and os.path.isdir(os.path.join(ws, "syn", p))][0]
    path = os.path.join(ws, "syn", frag, "PL_var.tdb")
    b = bytearray(open(path, "rb").read())
    b[len(b) // 2] ^= 0x5A
    open(path, "wb").write(bytes(b))
    with pytest.raises(capi.GdbGpuError) as e:
        capi.Array.open(ctx, ws, "syn", [("AD", "R", "int"), ("PL", "G", "int")])
    assert e.value.code == capi.ERR_IO and "inflate" in str(e.value)


def test_cached_device_memory_is_reused_and_can_be_released(ctx):
    """array open / close cycles reuse the context's cached device blocks; gdbgpu_ctx_release_cached_memory gives them back"""
    import torch
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=2000, n_sites=400, spacing=100, p_var=0.1, a_pool=8)
    want = None
    for it in range(4):
        arr = gdbsyn.to_gpu_array(ctx, d)
        sc = capi.Scan(arr)
        chunks = []
        while True:
            bv = sc.next_batch()
            if bv is None:
                break
            chunks.append(bv.fields[2]["data"].copy())
        sc.close()
        arr.close()
        got = np.concatenate(chunks)
        if want is None:
            want = got
        assert np.array_equal(got, want)
        if it == 1:
            free0 = torch.cuda.mem_get_info()[0]
    assert torch.cuda.mem_get_info()[0] == free0          # steady state: nothing new from the driver
    ctx.release_cached_memory()
    assert torch.cuda.mem_get_info()[0] > free0


def _scan_arrays(arr, flags, max_sites=0, max_bytes=0):
    """one record per SITE (whatever the batches were): tables, alleles, live row, every field's stride / lengths / vectors"""
    from genomicsdb_b200 import capi
    sc = capi.Scan(arr, flags=flags)
    out = []
    try:
        while True:
            b = sc.next_batch(max_sites, max_bytes)
            if b is None:
                break
            for s in range(b.n_sites):
                rec = [int(b.site_begin[s]), int(b.site_end[s]), int(b.n_alleles[s]), np.array(b.live[s]), b"|".join(b.alleles(s))]
                for f in b.fields:
                    base, stride = int(f["base"][s]), int(f["stride"][s])
                    rec += [stride, np.array(f["len"][s]), np.array(f["data"][base:base + b.nq * stride])]
                out.append(rec)
    finally:
        sc.close()
    return out


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert len(x) == len(y)
        for u, v in zip(x, y):
            assert np.array_equal(u, v) if isinstance(u, np.ndarray) else u == v


def test_signature_rows_equal_per_call_luts(ctx):
    """scans of >= 4096 samples remap the non-trivial calls through one permutation row per call signature (k_verify_alleles,
    k_sig_perm, nt_apply); a scan that also asks for the per-call LUTs rebuilds every call's LUT instead (nt_entry) -- the
    remapped vectors must not differ, and the LUT scan is the one the oracle comparison runs on"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=6000, n_sites=48, spacing=1000, p_var=0.15, a_pool=8)
    both(ctx, d, [])                                   # per-call LUT path against the oracle
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        _same(_scan_arrays(arr, 0, max_sites=7), _scan_arrays(arr, capi.SCAN_WANT_LUT, max_sites=7))
    finally:
        arr.close()


def test_many_front_windows(ctx, monkeypatch):
    """K3/K4 run once per window of sites (6 GiB of front state by default: one window in every other test); with a budget of a
    few sites per window the batches stop at window borders, with and without prefetch, in tiny and in large batches -- the sites
    and every vector are the same"""
    from genomicsdb_b200 import capi
    d = gdbsyn.generate(n_samples=5000, n_sites=60, spacing=1000, p_var=0.15, a_pool=8)
    arr = gdbsyn.to_gpu_array(ctx, d)
    try:
        want = _scan_arrays(arr, 0)
        monkeypatch.setenv("GDBGPU_FRONT_BYTES", str(24 * 5000 * 7))          # 7 sites per window
        _same(_scan_arrays(arr, 0, max_sites=5), want)
        _same(_scan_arrays(arr, capi.SCAN_NO_PREFETCH, max_sites=5), want)
        _same(_scan_arrays(arr, 0), want)
        _same(_scan_arrays(arr, 0, max_bytes=3 << 20), want)
    finally:
        monkeypatch.delenv("GDBGPU_FRONT_BYTES", raising=False)
        arr.close()
    monkeypatch.setenv("GDBGPU_FRONT_BYTES", str(24 * 300 * 3))               # and against the oracle: 3 sites per window
    d2 = gdbsyn.generate(n_samples=300, n_sites=100, spacing=100, p_var=0.1, a_pool=3)
    both(ctx, d2, [])
    both(ctx, d2, [], max_sites=2)
