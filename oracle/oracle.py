"""TEST INFRASTRUCTURE ONLY: python face of the CPU oracle (oracle/combine_oracle.c + oracle/tiledb_fragment.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; genomicsdb_b200/ never does.
It produces the canonical hot-path dump (layout: oracle/ref_harness/gdb_ref_dump.cc) for a query over a TileDB array
or over in-memory columns, and parses dumps for field-level diffs.
"""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np

from . import tiledb_fragment as tf

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
MAGIC = 0x31444247
KIND = {"R": 0, "A": 1, "G": 2}


def build(force=False):
    src = os.path.join(HERE, "combine_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-Wall", "-o", LIB, src])
    return LIB


class _Field(C.Structure):
    _fields_ = [("off", C.c_void_p), ("val", C.c_void_p), ("kind", C.c_int32), ("elem", C.c_int32), ("is_PL", C.c_int32)]


class _Input(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("row", C.c_void_p), ("col", C.c_void_p), ("end", C.c_void_p),
                ("ref_off", C.c_void_p), ("ref", C.c_void_p), ("alt_off", C.c_void_p), ("alt", C.c_void_p),
                ("gt_off", C.c_void_p), ("gt", C.c_void_p), ("gt_phased", C.c_int32),
                ("n_fields", C.c_int32), ("fields", C.c_void_p),
                ("scan_full", C.c_int32), ("col_lo", C.c_int64), ("col_hi", C.c_int64),
                ("n_query_rows", C.c_int64), ("query_rows", C.c_void_p),
                ("max_alt", C.c_int32), ("handle_spanning_deletions", C.c_int32), ("produce_GT_with_min_PL", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_combine.restype = C.c_int
        _lib.oracle_nCr.restype = C.c_uint64
        _lib.oracle_nCr.argtypes = [C.c_int64, C.c_int64]
        _lib.oracle_genotype_index.restype = C.c_uint64
        _lib.oracle_num_genotypes.restype = C.c_uint32
        _lib.oracle_enumerate_genotypes.restype = C.c_uint64
    return _lib


# ------------------------------------------------------------------ vid mapping (field descriptors)
def load_vid_fields(path):
    """src/main/cpp/src/utils/vid_mapper.cc (file-based mapper): name -> dict(type, length, classes)."""
    d = json.load(open(path))
    fl = d["fields"]
    items = fl.items() if isinstance(fl, dict) else [(x["name"], x) for x in fl]
    out = {}
    for name, v in items:
        classes = v.get("vcf_field_class", [])
        ent = dict(type=v.get("type", "int"), length=v.get("length", 1), classes=classes,
                   combine=v.get("VCF_field_combine_operation"))
        out[name] = ent
        if "FORMAT" in classes and "INFO" in classes:   # vid_mapper.cc: FORMAT twin gets the _FORMAT suffix
            out[name + "_FORMAT"] = ent
    return out


class Columns:
    """Decoded hot-path columns of an array in global cell order (what both the oracle and gdbgpu_array_from_columns eat)."""

    def __init__(self, row, col, end, ref_off, ref, alt_off, alt, gt_off, gt, gt_phased, fields, num_rows, lb_row=0):
        self.row = np.ascontiguousarray(row, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.end = np.ascontiguousarray(end, dtype=np.int64)
        self.ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        self.ref = np.ascontiguousarray(ref, dtype=np.uint8)
        self.alt_off = np.ascontiguousarray(alt_off, dtype=np.int64)
        self.alt = np.ascontiguousarray(alt, dtype=np.uint8)
        self.gt_off = None if gt_off is None else np.ascontiguousarray(gt_off, dtype=np.int64)
        self.gt = None if gt is None else np.ascontiguousarray(gt, dtype=np.int32)
        self.gt_phased = bool(gt_phased)
        # fields: list of dict(name, kind 'R'|'A'|'G', elem 'int'|'float', off, val(uint32 view))
        self.fields = fields
        self.num_rows = int(num_rows)
        self.lb_row = int(lb_row)

    @property
    def n_cells(self):
        return len(self.row)


def load_columns(workspace, array_name, vid_path, attributes):
    """Array + vid + queried attribute list -> Columns (GT first, then allele-dependent 1-D fields in query order)."""
    arr = tf.Array(workspace, array_name)
    vid = load_vid_fields(vid_path)
    end = arr.fixed("END")[:, 0]
    ref_off, ref = arr.var("REF")
    alt_off, alt = arr.var("ALT")
    gt_off = gt = None
    gt_phased = False
    if "GT" in arr.schema.attributes:          # do_query_bookkeeping(require_alleles) adds GT (query_variants.cc:578-685)
        gt_off, gt = arr.var("GT")
        gt_phased = (vid["GT"]["length"] == "PP")
    fields = []
    for name in attributes:
        ent = vid.get(name)
        if ent is None or not isinstance(ent["length"], str) or ent["length"] not in KIND:
            continue
        off, val = arr.var(name)
        elem = "float" if ent["type"] == "float" else "int"
        fields.append(dict(name=name, kind=ent["length"], elem=elem, off=off.astype(np.int64),
                           val=np.ascontiguousarray(val).view(np.uint32)))
    lb = arr.lb_row_idx or 0
    num_rows = (arr.max_valid_row_idx - lb + 1) if arr.max_valid_row_idx is not None else int(arr.row.max() + 1)
    return Columns(arr.row, arr.col, end, ref_off, ref, alt_off, alt, gt_off, gt, gt_phased, fields, num_rows, lb)


def dump_header(cols):
    names = (["GT"] if cols.gt_off is not None else []) + [f["name"] for f in cols.fields]
    b = struct.pack("<II", MAGIC, len(names))
    for n in names:
        b += struct.pack("<I", len(n)) + n.encode()
    return b


def combine(cols, col_lo=None, col_hi=None, query_rows=None, max_alt=50, handle_spanning_deletions=True,
            produce_GT_with_min_PL=False):
    """One column interval ([col_lo, col_hi] or full scan when col_lo is None) -> (dump bytes without header, n_sites)."""
    L = lib()
    if query_rows is None:
        query_rows = np.arange(cols.lb_row, cols.lb_row + cols.num_rows, dtype=np.int64)
    query_rows = np.ascontiguousarray(sorted(set(int(r) for r in query_rows
                                                  if cols.lb_row <= r < cols.lb_row + cols.num_rows)), dtype=np.int64)
    farr = (_Field * max(1, len(cols.fields)))()
    keep = []
    for i, f in enumerate(cols.fields):
        off = np.ascontiguousarray(f["off"], dtype=np.int64)
        val = np.ascontiguousarray(f["val"], dtype=np.uint32)
        keep += [off, val]
        farr[i] = _Field(off.ctypes.data, val.ctypes.data, KIND[f["kind"]], 1 if f["elem"] == "float" else 0,
                         1 if f["name"] == "PL" else 0)
    inp = _Input()
    inp.n_cells = cols.n_cells
    inp.row, inp.col, inp.end = cols.row.ctypes.data, cols.col.ctypes.data, cols.end.ctypes.data
    inp.ref_off, inp.ref = cols.ref_off.ctypes.data, cols.ref.ctypes.data
    inp.alt_off, inp.alt = cols.alt_off.ctypes.data, cols.alt.ctypes.data
    if cols.gt_off is not None:
        inp.gt_off, inp.gt = cols.gt_off.ctypes.data, cols.gt.ctypes.data
    inp.gt_phased = int(cols.gt_phased)
    inp.n_fields = len(cols.fields)
    inp.fields = C.cast(farr, C.c_void_p)
    inp.scan_full = 1 if col_lo is None else 0
    inp.col_lo = 0 if col_lo is None else int(col_lo)
    inp.col_hi = 0 if col_hi is None else int(col_hi)
    inp.n_query_rows = len(query_rows)
    inp.query_rows = query_rows.ctypes.data
    inp.max_alt = max_alt
    inp.handle_spanning_deletions = int(handle_spanning_deletions)
    inp.produce_GT_with_min_PL = int(produce_GT_with_min_PL)
    out = C.POINTER(C.c_uint8)()
    out_len = C.c_uint64()
    n_sites = C.c_int64()
    rc = L.oracle_combine(C.byref(inp), C.byref(out), C.byref(out_len), C.byref(n_sites))
    data = C.string_at(out, out_len.value)
    L.oracle_free(out)
    if rc != 0:
        raise RuntimeError("oracle_combine error %d (1: unhandled overlapping variants)" % rc)
    return data, n_sites.value


def combine_query(cols, column_ranges, **kw):
    """Full canonical dump (with header) for a list of [lo,hi] column intervals, in the reference's (sorted) order."""
    out = dump_header(cols)
    total = 0
    if not column_ranges:
        d, n = combine(cols, **kw)
        return out + d, n
    for lo, hi in sorted(tuple(r) for r in column_ranges):
        d, n = combine(cols, lo, hi, **kw)
        out += d
        total += n
    return out, total


# ------------------------------------------------------------------ dump parsing (for readable diffs)
def parse_dump(b):
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, b, off)
        off += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def rd_str():
        nonlocal off
        n = rd("I")
        s = b[off:off + n]
        off += n
        return s.decode()

    assert rd("I") == MAGIC
    names = [rd_str() for _ in range(rd("I"))]
    sites = []
    while off < len(b):
        s = dict(begin=rd("q"), end=rd("q"), nonref=rd("B"), remap=rd("B"))
        n_alt = rd("H")
        s["alleles"] = [rd_str() for _ in range(1 + n_alt)]
        calls = []
        for _ in range(rd("I")):
            c = dict(row=rd("q"), begin=rd("q"), end=rd("q"))
            a_in = rd("H")
            c["in2m"] = list(struct.unpack_from("<%di" % a_in, b, off))
            off += 4 * a_in
            c["fields"] = {}
            for nm in names:
                n = rd("I")
                c["fields"][nm] = np.frombuffer(b, dtype=np.uint32, count=n, offset=off).copy()
                off += 4 * n
            calls.append(c)
        s["calls"] = calls
        sites.append(s)
    return names, sites


def diff_dumps(a, b, max_report=10):
    """Human-readable differences between two dumps ([] when byte-identical)."""
    if a == b:
        return []
    na, sa = parse_dump(a)
    nb, sb = parse_dump(b)
    msgs = []
    if na != nb:
        msgs.append("field lists differ: %s vs %s" % (na, nb))
    if len(sa) != len(sb):
        msgs.append("site counts differ: %d vs %d" % (len(sa), len(sb)))
        msgs.append("a sites: %s" % [(s["begin"], s["end"]) for s in sa][:40])
        msgs.append("b sites: %s" % [(s["begin"], s["end"]) for s in sb][:40])
    for x, y in zip(sa, sb):
        if len(msgs) >= max_report:
            break
        hx = {k: x[k] for k in ("begin", "end", "nonref", "remap", "alleles")}
        hy = {k: y[k] for k in ("begin", "end", "nonref", "remap", "alleles")}
        if hx != hy:
            msgs.append("site header: %s vs %s" % (hx, hy))
            continue
        if len(x["calls"]) != len(y["calls"]):
            msgs.append("site %d: #calls %d vs %d" % (x["begin"], len(x["calls"]), len(y["calls"])))
            continue
        for cx, cy in zip(x["calls"], y["calls"]):
            for k in ("row", "begin", "end", "in2m"):
                if cx[k] != cy[k]:
                    msgs.append("site %d row %d: %s %s vs %s" % (x["begin"], cx["row"], k, cx[k], cy[k]))
            for nm in na:
                if nm in cy["fields"] and not np.array_equal(cx["fields"][nm], cy["fields"][nm]):
                    msgs.append("site %d row %d field %s: %s vs %s" % (x["begin"], cx["row"], nm,
                                                                         cx["fields"][nm].view(np.int32), cy["fields"][nm].view(np.int32)))
    return msgs or ["dumps differ in bytes but parse equal"]
