"""Python face of the GDBSYN-1 generator (tools/gdbsyn.c): synthetic cohorts as TileDB-layout columns.

Tooling for bench.py / tests: builds libgdbsyn.so with gcc on first use, returns numpy columns that feed
  * gdbgpu_array_from_columns (the product path),
  * oracle.Columns (the checker),
  * tools/tiledb_writer.py (fragments on disk for the unmodified reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libgdbsyn.so")

# BASELINE.json configs.  n_sites = positions / spacing.
CONFIGS = {
    "config2": dict(n_samples=1000, n_sites=10000, spacing=100, p_var=0.1, a_pool=3),
    "config3": dict(n_samples=10000, n_sites=10000, spacing=100, p_var=0.1, a_pool=3),
    "config4": dict(n_samples=100000, n_sites=10000, spacing=1000, p_var=0.1, a_pool=8),
}


class Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_samples", C.c_int64), ("n_sites", C.c_int64), ("spacing", C.c_int64),
                ("first_col", C.c_int64), ("p_var", C.c_double), ("a_pool", C.c_int32), ("site_lo", C.c_int64),
                ("site_hi", C.c_int64)]


class Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("coords", "end", "ref_off", "ref", "alt_off", "alt", "gt_off", "gt", "pl_off",
                                          "pl", "ad_off", "ad", "dp", "gq", "min_dp")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "gdbsyn.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["gcc", "-O3", "-std=c99", "-fopenmp", "-fPIC", "-shared", "-o", LIB, src])
        _lib = C.CDLL(LIB)
        _lib.gdbsyn_num_groups.restype = C.c_int64
    return _lib


def generate(n_samples, n_sites, spacing, p_var, a_pool, site_lo=0, site_hi=None, seed=1234, first_col=1000,
             with_fixed=False):
    """-> dict(coords[n,2], end[n], ref/alt/gt/pl/ad = (uint64 byte offsets[n], values), num_rows, params)"""
    L = lib()
    site_hi = n_sites if site_hi is None else site_hi
    p = Params(seed, n_samples, n_sites, spacing, first_col, p_var, a_pool, site_lo, site_hi)
    ng = L.gdbsyn_num_groups(C.byref(p))
    counts = np.zeros((ng, 6), dtype=np.uint64)
    L.gdbsyn_plan(C.byref(p), counts.ctypes.data_as(C.c_void_p))
    starts = np.zeros_like(counts)
    starts[1:] = np.cumsum(counts, axis=0)[:-1]
    tot = counts.sum(axis=0)
    n = int(tot[0])
    d = dict(coords=np.empty((n, 2), dtype=np.int64), end=np.empty(n, dtype=np.int64))
    arrays = {}
    for k, (name, dt) in enumerate([("ref", np.uint8), ("alt", np.uint8), ("gt", np.int32), ("pl", np.int32), ("ad", np.int32)], start=1):
        arrays[name] = (np.empty(n, dtype=np.uint64), np.empty(int(tot[k]), dtype=dt))
    fixed = {nm: (np.empty(n, dtype=np.int32) if with_fixed else None) for nm in ("dp", "gq", "min_dp")}
    o = Out()
    o.coords, o.end = d["coords"].ctypes.data, d["end"].ctypes.data
    for name in ("ref", "alt", "gt", "pl", "ad"):
        setattr(o, name + "_off", arrays[name][0].ctypes.data)
        setattr(o, name, arrays[name][1].ctypes.data if arrays[name][1].size else None)
    for nm in fixed:
        if fixed[nm] is not None:
            setattr(o, nm, fixed[nm].ctypes.data)
    L.gdbsyn_fill(C.byref(p), starts.ctypes.data_as(C.c_void_p), C.byref(o))
    d.update(arrays)
    d.update(fixed)
    d["num_rows"] = n_samples
    d["params"] = dict(n_samples=n_samples, n_sites=n_sites, spacing=spacing, p_var=p_var, a_pool=a_pool,
                       site_lo=site_lo, site_hi=site_hi, seed=seed, first_col=first_col)
    # column interval a query over exactly this shard uses (what the rank's column partition covers)
    # shards tile the genome: boundaries sit half a spacing before the first site of the next shard
    lo = first_col - 2 if site_lo == 0 else first_col + site_lo * spacing - spacing // 2
    hi = first_col + (n_sites - 1) * spacing + 2 if site_hi == n_sites else first_col + site_hi * spacing - spacing // 2 - 1
    d["col_range"] = (lo, hi)
    return d


def algorithmic_bytes(d):
    """B_in = hot-path attribute bytes of every cell, read once (begin cells: coords, END, REF,
    ALT, GT, PL, AD incl. their 8-byte offsets; END copies: coords + END only)."""
    col, end = d["coords"][:, 1], d["end"]
    n = len(end)
    is_copy = col > end
    total_var = sum(d[k][1].nbytes for k in ("ref", "alt", "gt", "pl", "ad"))
    # END copies duplicate their begin cell's payload; subtract it (they cost 24 bytes each)
    copy_var = 0
    for k in ("ref", "alt", "gt", "pl", "ad"):
        off = d[k][0].astype(np.int64)
        ln = np.diff(np.concatenate([off, [d[k][1].nbytes]]))
        copy_var += int(ln[is_copy].sum())
    n_copy = int(is_copy.sum())
    b_in = (n - n_copy) * (16 + 8 + 5 * 8) + (total_var - copy_var) + n_copy * 24
    return dict(n_cells=n, n_begin_cells=n - n_copy, n_end_copies=n_copy, b_in=int(b_in))


PLAIN_FIELDS = [("DP", "int", 1, 1), ("GQ", "int", 1, 0), ("MIN_DP", "int", 1, 0)]   # (attribute, elem, fixed_len, no_format) of generate(with_fixed=True)
PLAIN_DP_RULE = (1, 3, 0)    # 1-based plain indices of (INFO DP, MIN_DP, DP_FORMAT) for capi.Scan(dp_fields=...)


def to_gpu_array(ctx, d, plain=False):
    """plain=True (needs generate(with_fixed=True)): also stage DP (INFO, input of the DP rule), GQ and MIN_DP as plain fields"""
    from genomicsdb_b200 import capi
    fields = [dict(name="AD", kind="R", elem="int", col=d["ad"]), dict(name="PL", kind="G", elem="int", col=d["pl"])]
    pl = []
    if plain:
        pl = [dict(name=n, elem=e, fixed_len=k, no_format=nf, col=(None, d[n.lower()])) for n, e, k, nf in PLAIN_FIELDS]
    return capi.Array.from_columns(ctx, d["coords"], d["end"], d["ref"], d["alt"], d["gt"], fields, 0, d["num_rows"], plain=pl)


class OracleArrayView:
    """the accessors oracle.format_blocks.PlainColumns takes from a TileDB array, over a generated cohort"""

    def __init__(self, d):
        self.d, self.row, self.col = d, d["coords"][:, 0], d["coords"][:, 1]

    def fixed(self, name):
        return self.d[name.lower()].reshape(-1, 1)

    def var(self, name):
        off, val = self.d[name.lower()]
        item = val.dtype.itemsize
        return np.concatenate([off.astype(np.int64), [val.nbytes]]) // item, val


def to_oracle_columns(d):
    from oracle import oracle

    def el(col, itemsize):
        off = np.concatenate([col[0].astype(np.int64), [col[1].nbytes]]) // itemsize
        return off

    fields = [dict(name="AD", kind="R", elem="int", off=el(d["ad"], 4), val=d["ad"][1].view(np.uint32)),
              dict(name="PL", kind="G", elem="int", off=el(d["pl"], 4), val=d["pl"][1].view(np.uint32))]
    return oracle.Columns(d["coords"][:, 0], d["coords"][:, 1], d["end"], el(d["ref"], 1), d["ref"][1], el(d["alt"], 1), d["alt"][1],
                          el(d["gt"], 4), d["gt"][1], False, fields, d["num_rows"], 0)
