"""ctypes binding of include/gdbgpu.h (the C-ABI drop-in boundary).  Plumbing only: no computation happens here.

The library is built in-tree by __graft_entry__.build() (genomicsdb_b200/lib/libgdbgpu.so).  Loading fails loudly when
it is missing; every computing entry point fails with GDBGPU_ERR_CUDA when there is no CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GDBGPU_LIB") or os.path.join(HERE, "lib", "libgdbgpu.so")   # GDBGPU_LIB: a differently tuned build of the same library

OK, ERR_ARG, ERR_IO, ERR_CUDA, ERR_OVERLAP, ERR_LIMIT = 0, -1, -2, -3, -4, -5
LEN_R, LEN_A, LEN_G = 0, 1, 2
ELEM_INT32, ELEM_FLOAT32, ELEM_CHAR = 0, 1, 2
SCAN_FULL, SCAN_NO_DELETIONS, SCAN_MIN_PL_GT, SCAN_WANT_LUT, SCAN_DEVICE_OUTPUT = 1, 2, 4, 8, 16
SCAN_BCF_FORMAT, SCAN_PRODUCE_GT, SCAN_NO_PREFETCH, SCAN_HTSJDK_CONVENTIONS = 32, 64, 128, 256
SCAN_CELLS_ONLY, SCAN_SITE_STATS, SCAN_SUMMARY_ONLY = 512, 1024, 2048
CELL_FIELD_REF, CELL_FIELD_ALT, CELL_FIELD_GT, CELL_FIELD_REMAPPED, CELL_FIELD_PLAIN = 0, 1, 2, 3, 4
BT_SIZE = {1: 1, 2: 2, 3: 4, 5: 4, 7: 1}
SITE_NONREF, SITE_REMAPPED, SITE_G_SUPPRESSED = 1, 2, 4
CALL_REDUCED = 0x40000000
KIND = {"R": LEN_R, "A": LEN_A, "G": LEN_G}

EXPORTS = ["gdbgpu_version", "gdbgpu_ctx_create", "gdbgpu_ctx_destroy", "gdbgpu_last_error", "gdbgpu_ctx_stream",
           "gdbgpu_ctx_launch_count", "gdbgpu_ctx_synchronize", "gdbgpu_ctx_release_cached_memory", "gdbgpu_ctx_set_profiling",
           "gdbgpu_ctx_reset_kernel_stats", "gdbgpu_ctx_kernel_stats", "gdbgpu_array_open", "gdbgpu_array_open_ex", "gdbgpu_array_open_range", "gdbgpu_array_column_cuts", "gdbgpu_inflate_streams", "gdbgpu_array_from_columns", "gdbgpu_array_get_info",
           "gdbgpu_array_copy_cells", "gdbgpu_array_close", "gdbgpu_scan_begin", "gdbgpu_scan_get_info",
           "gdbgpu_scan_next_batch", "gdbgpu_scan_seek", "gdbgpu_scan_end", "gdbgpu_scan_cells", "gdbgpu_scan_cells_histogram",
           "gdbgpu_scan_cells_allele_counts", "gdbgpu_scan_cells_fetch", "gdbgpu_scan_cells_window",
           "gdbgpu_stream_open", "gdbgpu_stream_close", "gdbgpu_stream_num_bytes_available", "gdbgpu_stream_read_next_byte",
           "gdbgpu_stream_read", "gdbgpu_stream_skip", "gdbgpu_stream_last_error"]


class FieldDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("length_kind", C.c_int32), ("elem_type", C.c_int32)]


class PlainFieldDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("elem_type", C.c_int32), ("fixed_len", C.c_int32), ("no_format", C.c_int32),
                ("info_op", C.c_int32)]


class VarColumn(C.Structure):
    _fields_ = [("offsets", C.c_void_p), ("values", C.c_void_p), ("values_bytes", C.c_uint64)]


class Columns(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("coords", C.c_void_p), ("end", C.c_void_p),
                ("ref", VarColumn), ("alt", VarColumn), ("gt", VarColumn), ("gt_phased", C.c_int32),
                ("n_fields", C.c_int32), ("field_descs", C.POINTER(FieldDesc)), ("fields", C.POINTER(VarColumn)),
                ("lb_row_idx", C.c_int64), ("num_rows", C.c_int64),
                ("n_plain_fields", C.c_int32), ("plain_descs", C.POINTER(PlainFieldDesc)), ("plain_fields", C.POINTER(VarColumn))]


class ArrayInfo(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_begin_cells", C.c_uint64), ("n_fragments", C.c_uint32),
                ("lb_row_idx", C.c_int64), ("num_rows", C.c_int64), ("min_col", C.c_int64), ("max_col", C.c_int64),
                ("max_end", C.c_int64), ("staged_bytes", C.c_uint64), ("hbm_bytes", C.c_uint64),
                ("n_fields", C.c_int32), ("has_gt", C.c_int32), ("gt_phased", C.c_int32), ("n_plain_fields", C.c_int32)]


class ScanParams(C.Structure):
    _fields_ = [("col_lo", C.c_int64), ("col_hi", C.c_int64), ("query_rows", C.c_void_p), ("n_query_rows", C.c_uint64),
                ("flags", C.c_uint32), ("max_diploid_alt_alleles", C.c_uint32),
                ("dp_info_field", C.c_uint32), ("dp_min_dp_field", C.c_uint32), ("dp_format_field", C.c_uint32),
                ("esum_remapped_mask", C.c_uint32)]


class ScanInfo(C.Structure):
    _fields_ = [("n_sites", C.c_uint64), ("n_breakpoints", C.c_uint64), ("n_query_rows", C.c_uint64),
                ("n_forward_cells", C.c_uint64)]


class FieldOut(C.Structure):
    _fields_ = [("data", C.c_void_p), ("base", C.c_void_p), ("stride", C.c_void_p), ("len", C.c_void_p)]


class CellList(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_left", C.c_uint64), ("n_total", C.c_uint64), ("first", C.c_uint64), ("cell", C.c_void_p), ("row", C.c_void_p), ("begin", C.c_void_p), ("end", C.c_void_p)]


class AlleleCounts(C.Structure):
    _fields_ = [("n", C.c_uint64), ("column", C.c_void_p), ("count", C.c_void_p), ("ref_pos", C.c_void_p), ("alt_pos", C.c_void_p), ("bytes", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("first_site", C.c_uint64), ("n_sites", C.c_uint32), ("n_query_rows", C.c_uint64),
                ("site_begin", C.c_void_p), ("site_end", C.c_void_p), ("site_flags", C.c_void_p),
                ("n_alleles", C.c_void_p), ("n_live", C.c_void_p),
                ("allele_first", C.c_void_p), ("allele_pos", C.c_void_p), ("allele_bytes", C.c_void_p),
                ("live", C.c_void_p), ("n_fields", C.c_int32), ("fields", FieldOut * 16),
                ("lut_first", C.c_void_p), ("lut_len", C.c_void_p), ("lut_in2m", C.c_void_p), ("out_bytes", C.c_uint64),
                ("n_fmt_fields", C.c_int32), ("fmt_type", C.c_void_p), ("fmt_nps", C.c_void_p), ("fmt_offset", C.c_void_p), ("fmt_bytes", C.c_void_p),
                ("fmt_total_bytes", C.c_uint64), ("info_valid", C.c_void_p), ("info_value", C.c_void_p),
                ("info_dp_sum", C.c_void_p), ("vec_len", C.c_void_p), ("vec_offset", C.c_void_p), ("vec_values", C.c_void_p),
                ("site_stats", C.c_void_p)]


_lib = None


def lib():
    """Loads libgdbgpu.so; raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libgdbgpu.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.gdbgpu_version.restype = C.c_char_p
        L.gdbgpu_last_error.restype = C.c_char_p
        L.gdbgpu_last_error.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.gdbgpu_ctx_destroy.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_stream.restype = C.c_void_p
        L.gdbgpu_ctx_stream.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_launch_count.restype = C.c_uint64
        L.gdbgpu_ctx_launch_count.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_synchronize.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_release_cached_memory.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.gdbgpu_ctx_reset_kernel_stats.argtypes = [C.c_void_p]
        L.gdbgpu_ctx_kernel_stats.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.gdbgpu_array_open.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(FieldDesc), C.c_int32, C.c_int32,
                                        C.POINTER(C.c_void_p)]
        L.gdbgpu_array_open_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(FieldDesc), C.c_int32,
                                           C.POINTER(PlainFieldDesc), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.gdbgpu_array_open_range.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(FieldDesc), C.c_int32,
                                              C.POINTER(PlainFieldDesc), C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        L.gdbgpu_array_column_cuts.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_int64)]
        L.gdbgpu_inflate_streams.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_double)]
        L.gdbgpu_array_from_columns.argtypes = [C.c_void_p, C.POINTER(Columns), C.POINTER(C.c_void_p)]
        L.gdbgpu_array_get_info.argtypes = [C.c_void_p, C.POINTER(ArrayInfo)]
        L.gdbgpu_array_copy_cells.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.gdbgpu_array_close.argtypes = [C.c_void_p]
        L.gdbgpu_scan_begin.argtypes = [C.c_void_p, C.POINTER(ScanParams), C.POINTER(C.c_void_p)]
        L.gdbgpu_scan_get_info.argtypes = [C.c_void_p, C.POINTER(ScanInfo)]
        L.gdbgpu_scan_next_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(Batch)]
        L.gdbgpu_scan_seek.argtypes = [C.c_void_p, C.c_uint64]
        L.gdbgpu_scan_cells.argtypes = [C.c_void_p, C.POINTER(CellList)]
        L.gdbgpu_scan_cells_window.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.gdbgpu_scan_cells_histogram.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.gdbgpu_scan_cells_allele_counts.argtypes = [C.c_void_p, C.POINTER(AlleleCounts)]
        L.gdbgpu_scan_cells_fetch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.gdbgpu_scan_end.argtypes = [C.c_void_p]
        L.gdbgpu_stream_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.gdbgpu_stream_close.argtypes = [C.c_void_p]
        L.gdbgpu_stream_num_bytes_available.restype = C.c_uint64
        L.gdbgpu_stream_num_bytes_available.argtypes = [C.c_void_p]
        L.gdbgpu_stream_read_next_byte.argtypes = [C.c_void_p]
        L.gdbgpu_stream_read.restype = C.c_int64
        L.gdbgpu_stream_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.gdbgpu_stream_skip.restype = C.c_int64
        L.gdbgpu_stream_skip.argtypes = [C.c_void_p, C.c_uint64]
        L.gdbgpu_stream_last_error.restype = C.c_char_p
        _lib = L
    return _lib


class GdbGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gdbgpu error %d: %s" % (code, msg))
        self.code = code


def _view(ptr, dtype, count):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (int(count) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(count))


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().gdbgpu_ctx_create(device, C.byref(self.h))
        if rc != OK:
            raise GdbGpuError(rc, "cannot create a context on CUDA device %d (no GPU? there is no CPU fallback)" % device)

    def check(self, rc):
        if rc != OK:
            raise GdbGpuError(rc, lib().gdbgpu_last_error(self.h).decode())

    @property
    def stream(self):
        return lib().gdbgpu_ctx_stream(self.h)

    @property
    def launches(self):
        return lib().gdbgpu_ctx_launch_count(self.h)

    def synchronize(self):
        self.check(lib().gdbgpu_ctx_synchronize(self.h))

    def release_cached_memory(self):
        self.check(lib().gdbgpu_ctx_release_cached_memory(self.h))

    def set_profiling(self, on):
        lib().gdbgpu_ctx_set_profiling(self.h, int(on))

    def reset_kernel_stats(self):
        lib().gdbgpu_ctx_reset_kernel_stats(self.h)

    def kernel_stats(self, name):
        ms, n = C.c_double(), C.c_uint64()
        self.check(lib().gdbgpu_ctx_kernel_stats(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def inflate_streams(self, streams, out_sizes):
        """K0 on its own: `streams` (list of bytes, one zlib stream each) -> (list of inflated bytes, kernel ms)"""
        import numpy as np
        n = len(streams)
        in_off = np.zeros(n + 1, dtype=np.uint64)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        in_off[1:] = np.cumsum([len(x) for x in streams], dtype=np.uint64)
        out_off[1:] = np.cumsum(list(out_sizes), dtype=np.uint64)
        blob = b"".join(streams)
        out = np.zeros(max(int(out_off[-1]), 1), dtype=np.uint8)
        ms = C.c_double(0.0)
        self.check(lib().gdbgpu_inflate_streams(self.h, blob, in_off.ctypes.data_as(C.POINTER(C.c_uint64)), out.ctypes.data_as(C.c_void_p),
                                                out_off.ctypes.data_as(C.POINTER(C.c_uint64)), n, C.byref(ms)))
        return [out[int(out_off[i]):int(out_off[i + 1])].tobytes() for i in range(n)], ms.value

    def close(self):
        if self.h:
            lib().gdbgpu_ctx_destroy(self.h)
            self.h = C.c_void_p()


def _field_descs(fields):
    arr = (FieldDesc * max(1, len(fields)))()
    for i, (name, kind, elem) in enumerate(fields):
        arr[i] = FieldDesc(name.encode(), KIND[kind] if isinstance(kind, str) else kind,
                           ELEM_FLOAT32 if elem in ("float", ELEM_FLOAT32) else ELEM_INT32)
    return arr


def _plain_descs(plain):
    """plain: list of (attribute name, 'int'|'float'|'char', fixed length or 0 for variable[, no_format[, info_op]])
    info_op: None | 'median' | 'sum' | 'mean'"""
    arr = (PlainFieldDesc * max(1, len(plain)))()
    code = {"int": ELEM_INT32, "float": ELEM_FLOAT32, "char": ELEM_CHAR}
    ops = {None: 0, "median": 1, "sum": 2, "mean": 3, "element_wise_sum": 4}
    for i, p in enumerate(plain):
        name, elem, fixed = p[:3]
        arr[i] = PlainFieldDesc(name.encode(), code.get(elem, elem), int(fixed), int(bool(p[3])) if len(p) > 3 else 0,
                                ops.get(p[4], p[4]) if len(p) > 4 else 0)
    return arr


class Array:
    """fields: list of (attribute name, 'R'|'A'|'G', 'int'|'float')"""

    def __init__(self, ctx, handle, keep=None):
        self.ctx, self.h, self._keep = ctx, handle, keep

    @classmethod
    def open(cls, ctx, workspace, array_name, fields, gt_phased=False, plain=(), col_range=None):
        """col_range=(lo, hi): stage only the tiles that can hold cells of those columns (gdbgpu_array_open_range)"""
        h = C.c_void_p()
        fd = _field_descs(fields)
        pd = _plain_descs(list(plain))
        if col_range is None:
            ctx.check(lib().gdbgpu_array_open_ex(ctx.h, workspace.encode(), array_name.encode(), fd, len(fields), pd, len(plain),
                                                 int(gt_phased), C.byref(h)))
        else:
            ctx.check(lib().gdbgpu_array_open_range(ctx.h, workspace.encode(), array_name.encode(), fd, len(fields), pd, len(plain),
                                                    int(gt_phased), int(col_range[0]), int(col_range[1]), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_columns(cls, ctx, coords, end, ref, alt, gt, fields, lb_row, num_rows, gt_phased=False, plain=()):
        """ref/alt/gt/fields[i]['col'] = (byte_offsets uint64[n], values ndarray); fields = list of dict(name, kind, elem, col);
        plain = list of dict(name, elem, fixed_len, col) with col = (None, values) for a fixed-length field"""
        keep = []

        def vc(col):
            if col is None:
                return VarColumn(None, None, 0)
            off = np.ascontiguousarray(col[0], dtype=np.uint64)
            val = np.ascontiguousarray(col[1])
            keep.extend([off, val])
            return VarColumn(off.ctypes.data, val.ctypes.data if val.size else None, val.nbytes)

        coords = np.ascontiguousarray(coords, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        keep += [coords, end]
        c = Columns()
        c.n_cells = len(end)
        c.coords, c.end = coords.ctypes.data, end.ctypes.data
        c.ref, c.alt, c.gt = vc(ref), vc(alt), vc(gt)
        c.gt_phased = int(gt_phased)
        c.n_fields = len(fields)
        fd = _field_descs([(f["name"], f["kind"], f["elem"]) for f in fields])
        fv = (VarColumn * max(1, len(fields)))()
        for i, f in enumerate(fields):
            fv[i] = vc(f["col"])
        keep += [fd, fv]
        c.field_descs, c.fields = fd, fv
        c.lb_row_idx, c.num_rows = lb_row, num_rows
        plain = list(plain)
        pd = _plain_descs([(f["name"], f["elem"], f["fixed_len"], f.get("no_format", 0), f.get("info_op")) for f in plain])
        pv = (VarColumn * max(1, len(plain)))()
        for i, f in enumerate(plain):
            if f["fixed_len"]:
                val = np.ascontiguousarray(f["col"][1])
                keep.append(val)
                pv[i] = VarColumn(None, val.ctypes.data if val.size else None, val.nbytes)
            else:
                pv[i] = vc(f["col"])
        keep += [pd, pv]
        c.n_plain_fields, c.plain_descs, c.plain_fields = len(plain), pd, pv
        h = C.c_void_p()
        ctx.check(lib().gdbgpu_array_from_columns(ctx.h, C.byref(c), C.byref(h)))
        return cls(ctx, h, keep=None)

    def info(self):
        i = ArrayInfo()
        self.ctx.check(lib().gdbgpu_array_get_info(self.h, C.byref(i)))
        return i

    def cells(self):
        n = self.info().n_cells
        coords = np.zeros((n, 2), dtype=np.int64)
        end = np.zeros(n, dtype=np.int64)
        if n:
            self.ctx.check(lib().gdbgpu_array_copy_cells(self.h, 0, n, coords.ctypes.data, end.ctypes.data))
        return coords, end

    def close(self):
        if self.h:
            lib().gdbgpu_array_close(self.h)
            self.h = C.c_void_p()


class BatchView:
    """numpy views over one gdbgpu_batch (host output mode); valid until the next next_batch() call."""

    def __init__(self, b, n_out):
        ns, nq = b.n_sites, b.n_query_rows
        self.first_site, self.n_sites, self.nq, self.out_bytes = b.first_site, ns, nq, b.out_bytes
        self.site_begin = _view(b.site_begin, np.int64, ns)
        self.site_end = _view(b.site_end, np.int64, ns)
        self.site_flags = _view(b.site_flags, np.uint8, ns)
        self.n_alleles = _view(b.n_alleles, np.uint32, ns)
        self.n_live = _view(b.n_live, np.uint32, ns)
        self.site_stats = _view(b.site_stats, np.uint32, 2 * ns).reshape(-1, 2) if b.site_stats else None
        self.allele_first = _view(b.allele_first, np.uint32, ns)
        tot = int((self.allele_first + self.n_alleles).max()) if ns else 0
        self.allele_pos = _view(b.allele_pos, np.uint32, 2 * tot).reshape(-1, 2)
        nbytes = int((self.allele_pos[:, 0] + self.allele_pos[:, 1]).max()) if tot else 0
        self.allele_bytes = _view(b.allele_bytes, np.uint8, nbytes)
        self.live = _view(b.live, np.int32, ns * nq).reshape(ns, nq) if b.live else None
        self.fields = []
        for o in range(n_out):
            f = b.fields[o]
            base = _view(f.base, np.uint64, ns + 1)
            stride = _view(f.stride, np.uint32, ns)
            ln = _view(f.len, np.uint16, ns * nq).reshape(ns, nq) if f.len else None
            data = _view(f.data, np.uint32, int(base[ns]) if ns else 0) if f.data else None
            self.fields.append(dict(base=base, stride=stride, len=ln, data=data))
        self.fmt_total_bytes = b.fmt_total_bytes
        self.n_fmt_fields = b.n_fmt_fields
        if b.fmt_type:
            nf = b.n_fmt_fields
            self.fmt_type = _view(b.fmt_type, np.uint8, nf * ns).reshape(nf, ns)
            self.fmt_nps = _view(b.fmt_nps, np.uint32, nf * ns).reshape(nf, ns)
            self.fmt_offset = _view(b.fmt_offset, np.uint64, nf * ns).reshape(nf, ns)
            self.fmt_bytes = _view(b.fmt_bytes, np.uint8, b.fmt_total_bytes)
            npl = nf - n_out
            self.info_valid = _view(b.info_valid, np.uint8, npl * ns).reshape(npl, ns)
            self.info_value = _view(b.info_value, np.uint32, npl * ns).reshape(npl, ns)
            self.info_dp_sum = _view(b.info_dp_sum, np.int32, ns)
            self._vec = (b.vec_len, b.vec_offset, b.vec_values)
        else:
            self.fmt_type = None
        self.lut_first = _view(b.lut_first, np.uint32, ns * nq).reshape(ns, nq) if b.lut_first else None
        self.lut_len = _view(b.lut_len, np.uint16, ns * nq).reshape(ns, nq) if b.lut_len else None
        self._lut_ptr = b.lut_in2m

    def alleles(self, s):
        out = []
        for m in range(int(self.n_alleles[s])):
            off, ln = self.allele_pos[int(self.allele_first[s]) + m]
            out.append(bytes(self.allele_bytes[off:off + ln]))
        return out

    def fmt_block(self, o, s):
        """(BCF type, values per sample, block bytes) of output field o at site s, or None when it is not emitted"""
        t = int(self.fmt_type[o, s])
        if t == 0:
            return None
        nps, off = int(self.fmt_nps[o, s]), int(self.fmt_offset[o, s])
        return t, nps, bytes(self.fmt_bytes[off:off + self.nq * nps * BT_SIZE[t]])

    def info_vector(self, slot, s, n_slots):
        """element-wise sum of vector slot `slot` (remapped field f, or n_remapped + plain index) at site s as uint32 words"""
        ln = _view(self._vec[0], np.uint32, n_slots * self.n_sites).reshape(n_slots, self.n_sites)
        off = _view(self._vec[1], np.uint64, n_slots * self.n_sites).reshape(n_slots, self.n_sites)
        n = int(ln[slot, s])
        return _view(self._vec[2] + 4 * int(off[slot, s]), np.uint32, n).copy() if n else np.zeros(0, dtype=np.uint32)

    def lut(self, s, q):
        first = int(self.lut_first[s, q])
        return _view(self._lut_ptr + 2 * first, np.int16, int(self.lut_len[s, q]))


class Scan:
    def __init__(self, arr, col_lo=None, col_hi=None, query_rows=None, flags=0, max_alt=0, dp_fields=(0, 0, 0), esum_remapped_mask=0):
        """dp_fields: 1-based plain-field indices of (INFO DP, MIN_DP, DP_FORMAT) for the DP rule, 0 = absent;
        esum_remapped_mask: bit f = element-wise sum of remapped field f over the valid calls"""
        self.arr, self.ctx = arr, arr.ctx
        p = ScanParams()
        if col_lo is None:
            flags |= SCAN_FULL
        else:
            p.col_lo, p.col_hi = int(col_lo), int(col_hi)
        self._rows = None
        if query_rows is not None:
            self._rows = np.ascontiguousarray(query_rows, dtype=np.int64)
            p.query_rows, p.n_query_rows = self._rows.ctypes.data, len(self._rows)
        p.flags, p.max_diploid_alt_alleles = flags, max_alt
        p.dp_info_field, p.dp_min_dp_field, p.dp_format_field = dp_fields
        p.esum_remapped_mask = esum_remapped_mask
        self.h = C.c_void_p()
        self.ctx.check(lib().gdbgpu_scan_begin(arr.h, C.byref(p), C.byref(self.h)))
        i = arr.info()
        self.n_out = (1 if i.has_gt else 0) + i.n_fields

    def info(self):
        i = ScanInfo()
        self.ctx.check(lib().gdbgpu_scan_get_info(self.h, C.byref(i)))
        return i

    def next_batch(self, max_sites=0, max_bytes=0, raw=False):
        b = Batch()
        self.ctx.check(lib().gdbgpu_scan_next_batch(self.h, max_sites, max_bytes, C.byref(b)))
        if raw:
            return b
        return BatchView(b, self.n_out) if b.n_sites else None

    def seek(self, site):
        self.ctx.check(lib().gdbgpu_scan_seek(self.h, site))

    # ---- the columnar iterator (gdbgpu_scan_cells*) ----
    def cells(self):
        """-> dict(n_left, cell, row, begin, end): the cells the reference's columnar iterator visits, in its order"""
        cl = CellList()
        self.ctx.check(lib().gdbgpu_scan_cells(self.h, C.byref(cl)))
        n = cl.n_cells
        return dict(n_left=int(cl.n_left), n_total=int(cl.n_total), first=int(cl.first), cell=_view(cl.cell, np.uint64, n).copy(), row=_view(cl.row, np.int64, n).copy(),
                    begin=_view(cl.begin, np.int64, n).copy(), end=_view(cl.end, np.int64, n).copy())

    def cells_window(self, first, count=0):
        """page through the list: cells() / cells_fetch() hand out entries [first, first + count) from now on"""
        self.ctx.check(lib().gdbgpu_scan_cells_window(self.h, first, count))

    def cells_histogram(self, begin, end, bin_size):
        n_bins = (end - begin) // bin_size + 1
        counts = np.zeros(n_bins, dtype=np.uint64)
        self.ctx.check(lib().gdbgpu_scan_cells_histogram(self.h, begin, end, bin_size, counts.ctypes.data, n_bins))
        return counts

    def cells_allele_counts(self):
        """-> [(column, REF, ALT, count)] in the order AlleleCountOperator::print_allele_counts prints"""
        ac = AlleleCounts()
        self.ctx.check(lib().gdbgpu_scan_cells_allele_counts(self.h, C.byref(ac)))
        n = ac.n
        col, cnt = _view(ac.column, np.int64, n), _view(ac.count, np.uint64, n)
        rp, ap = _view(ac.ref_pos, np.uint32, 2 * n), _view(ac.alt_pos, np.uint32, 2 * n)
        nb = int(ap[2 * n - 2] + ap[2 * n - 1]) if n else 0
        by = bytes(_view(ac.bytes, np.uint8, nb)) if nb else b""
        return [(int(col[k]), by[rp[2 * k]:rp[2 * k] + rp[2 * k + 1]].decode(), by[ap[2 * k]:ap[2 * k] + ap[2 * k + 1]].decode(), int(cnt[k])) for k in range(n)]

    def cells_fetch(self, kind, index=0):
        """-> (offsets[n_cells + 1], bytes) of one column over the listed cells"""
        n = self.cells()["row"].size
        off, by = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().gdbgpu_scan_cells_fetch(self.h, kind, index, C.byref(off), C.byref(by)))
        o = _view(off.value, np.uint64, n + 1).copy()
        return o, (bytes(_view(by.value, np.uint8, int(o[-1]))) if n and o[-1] else b"")

    def close(self):
        if self.h:
            lib().gdbgpu_scan_end(self.h)
            self.h = C.c_void_p()


class QueryStream:
    """The Java reader's query stream through the C ABI (gdbgpu_stream_*: one call per native method of
    GenomicsDBQueryStream, jni/src/genomicsdb_GenomicsDBQueryStream.cc:29-111): an uncompressed BCF2 stream."""

    def __init__(self, loader_json, query_json, chr="", start=0, end=0, rank=0, buffer_capacity=1 << 20, segment_size=1 << 20,
                 is_bcf=True, produce_header_only=False, use_missing_values_only_not_vector_end=False, keep_idx_fields_in_bcf_header=True):
        L = lib()
        h = C.c_void_p()
        rc = L.gdbgpu_stream_open((loader_json or "").encode(), query_json.encode(), (chr or "").encode(), start, end, rank, buffer_capacity,
                                  segment_size, int(is_bcf), int(produce_header_only), int(use_missing_values_only_not_vector_end),
                                  int(keep_idx_fields_in_bcf_header), C.byref(h))
        if rc != 0:
            raise GdbGpuError(rc, L.gdbgpu_stream_last_error().decode())
        self._h = h

    def num_bytes_available(self):
        return lib().gdbgpu_stream_num_bytes_available(self._h)

    def read(self, n):
        buf = (C.c_uint8 * n)()
        k = lib().gdbgpu_stream_read(self._h, buf, 0, n)
        if k < 0:
            raise GdbGpuError(ERR_IO, lib().gdbgpu_stream_last_error().decode())
        return bytes(buf[:k])

    def read_next_byte(self):
        return lib().gdbgpu_stream_read_next_byte(self._h)

    def skip(self, n):
        return lib().gdbgpu_stream_skip(self._h, n)

    def close(self):
        if self._h:
            lib().gdbgpu_stream_close(self._h)
            self._h = None
