"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the TileDB 0.5.x sparse-array *read* path that GenomicsDB
sits on.  Never imported by the product (genomicsdb_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline leg may use it.

Restates, for sparse column-major arrays opened in read mode:
  * ArraySchema::deserialize            dependencies/TileDB/core/src/array/array_schema.cc:948-1115 (layout :569-594)
  * BookKeeping::load                   dependencies/TileDB/core/src/fragment/book_keeping.cc:217-243,368-422
  * ReadState::prepare_tile_for_reading{,_var}{_cmp,_cmp_none}   .../fragment/read_state.cc:2442-2889
    (+ gunzip = one zlib stream per tile, .../misc/utils.cc:466-512; var offsets are ABSOLUTE offsets into the
     uncompressed <ATTR>_var.tdb stream, read_state.cc:2991-3020)
  * StorageManager fragment discovery / sort_fragment_names      .../storage_manager/storage_manager.cc:2169-2207
  * ArrayReadState fragment merge: global (col,row) order, newest fragment wins ties
    .../array/array_read_state.cc:1836-2132, comparator :2369-2402
  * GenomicsDB array metadata (lb_row_idx / max_valid_row_idx_in_array)
    src/main/cpp/src/genomicsdb/variant_storage_manager.cc:313-377
"""
import glob
import json
import os
import struct
import zlib

import numpy as np

TILEDB_INT32, TILEDB_INT64, TILEDB_FLOAT32, TILEDB_FLOAT64, TILEDB_CHAR = 0, 1, 2, 3, 4
TILEDB_VAR_NUM = 0x7FFFFFFF
TILEDB_NO_COMPRESSION, TILEDB_GZIP = 0, 1
TILEDB_COL_MAJOR = 1
_NP = {TILEDB_INT32: np.int32, TILEDB_INT64: np.int64, TILEDB_FLOAT32: np.float32, TILEDB_FLOAT64: np.float64,
       TILEDB_CHAR: np.uint8}


class ArraySchema:
    """array_schema.cc:948-1115"""

    def __init__(self, path):
        b = open(path, "rb").read()
        off = 0

        def rd(fmt):
            nonlocal off
            v = struct.unpack_from("<" + fmt, b, off)
            off += struct.calcsize("<" + fmt)
            return v[0] if len(v) == 1 else v

        def rd_str():
            nonlocal off
            n = rd("i")
            s = b[off:off + n].decode()
            off += n
            return s

        tag = struct.unpack_from("<I", b, 0)[0]
        if tag >= 0xFFFFFFF0:  # version tag present -> workspace string follows (array_schema.cc:964-977)
            off = 4
            self.workspace = rd_str()
        self.array_name = rd_str()
        self.dense = bool(rd("B"))
        self.tile_order = rd("b")
        self.cell_order = rd("b")
        self.capacity = rd("q")
        nattr = rd("i")
        self.attributes = [rd_str() for _ in range(nattr)]
        ndim = rd("i")
        self.dimensions = [rd_str() for _ in range(ndim)]
        dom_size = rd("i")
        off += dom_size
        ext_size = rd("i")
        off += ext_size
        self.types = [rd("b") for _ in range(nattr + 1)]
        self.cell_val_num = [rd("i") for _ in range(nattr)]
        self.compression = [rd("b") for _ in range(nattr + 1)]
        assert off == len(b), "schema size mismatch"
        assert not self.dense and ndim == 2 and self.types[nattr] == TILEDB_INT64
        assert self.cell_order == TILEDB_COL_MAJOR
        self.attribute_num = nattr

    def attr_id(self, name):
        return self.attributes.index(name)

    def is_var(self, i):
        return self.cell_val_num[i] == TILEDB_VAR_NUM

    def cell_size(self, i):
        if i == self.attribute_num:
            return 16
        if self.is_var(i):
            return 8
        return self.cell_val_num[i] * np.dtype(_NP[self.types[i]]).itemsize


class BookKeeping:
    """book_keeping.cc:217-243 (format), :368-422 (load)"""

    def __init__(self, path, schema):
        b = zlib.decompress(open(path, "rb").read(), 16 + zlib.MAX_WBITS)
        off = 0
        n = schema.attribute_num

        def arr(dtype, count):
            nonlocal off
            a = np.frombuffer(b, dtype=dtype, count=count, offset=off)
            off += a.nbytes
            return a

        dom_size = int(arr(np.uint64, 1)[0])
        self.non_empty_domain = arr(np.int64, dom_size // 8)
        mbr_num = int(arr(np.int64, 1)[0])
        self.mbrs = arr(np.int64, mbr_num * 4).reshape(-1, 4)
        bc_num = int(arr(np.int64, 1)[0])
        self.bounding_coords = arr(np.int64, bc_num * 4).reshape(-1, 4)
        self.tile_offsets = []
        for _ in range(n + 1):
            k = int(arr(np.int64, 1)[0])
            self.tile_offsets.append(arr(np.int64, k))
        self.tile_var_offsets = []
        for _ in range(n):
            k = int(arr(np.int64, 1)[0])
            self.tile_var_offsets.append(arr(np.int64, k))
        self.tile_var_sizes = []
        for _ in range(n):
            k = int(arr(np.int64, 1)[0])
            self.tile_var_sizes.append(arr(np.uint64, k))
        self.last_tile_cell_num = int(arr(np.int64, 1)[0])
        assert off == len(b), "book-keeping size mismatch"
        self.tile_num = mbr_num
        self.cell_num = (mbr_num - 1) * schema.capacity + self.last_tile_cell_num if mbr_num else 0


def _inflate_tiles(raw, offsets, sizes=None):
    """one zlib stream per tile, located by book-keeping tile offsets (read_state.cc:2468-2560); var tiles whose
    uncompressed size is 0 are not inflated at all (read_state.cc:2719-2723)"""
    out = []
    for i in range(len(offsets)):
        lo = int(offsets[i])
        hi = int(offsets[i + 1]) if i + 1 < len(offsets) else len(raw)
        if sizes is not None and int(sizes[i]) == 0:
            continue
        out.append(zlib.decompress(raw[lo:hi]))
        assert sizes is None or len(out[-1]) == int(sizes[i])
    return b"".join(out)


class Fragment:
    def __init__(self, path, schema):
        self.path = path
        self.schema = schema
        self.bk = BookKeeping(os.path.join(path, "__book_keeping.tdb.gz"), schema)
        self.cell_num = self.bk.cell_num

    def _fixed_bytes(self, attr_id):
        s = self.schema
        name = "__coords" if attr_id == s.attribute_num else s.attributes[attr_id]
        raw = open(os.path.join(self.path, name + ".tdb"), "rb").read()
        if s.compression[attr_id] == TILEDB_NO_COMPRESSION:
            return raw
        assert s.compression[attr_id] == TILEDB_GZIP, "only GZIP / none are produced by GenomicsDB"
        return _inflate_tiles(raw, self.bk.tile_offsets[attr_id])

    def coords(self):
        """-> (row[n], col[n]) int64; cell = (samples, position) pairs (vid_mapper.cc:354-442)"""
        a = np.frombuffer(self._fixed_bytes(self.schema.attribute_num), dtype=np.int64).reshape(-1, 2)
        assert a.shape[0] == self.cell_num
        return a[:, 0].copy(), a[:, 1].copy()

    def fixed(self, name):
        i = self.schema.attr_id(name)
        a = np.frombuffer(self._fixed_bytes(i), dtype=_NP[self.schema.types[i]])
        return a.reshape(self.cell_num, self.schema.cell_val_num[i]).copy()

    def var(self, name):
        """-> (offsets[n+1] in ELEMENTS, values[]) ; on-disk offsets are absolute BYTE offsets (read_state.cc:2991-3001)"""
        s = self.schema
        i = s.attr_id(name)
        offs = np.frombuffer(self._fixed_bytes(i), dtype=np.uint64).astype(np.int64)
        raw = open(os.path.join(self.path, s.attributes[i] + "_var.tdb"), "rb").read()
        if s.compression[i] != TILEDB_NO_COMPRESSION:
            raw = _inflate_tiles(raw, self.bk.tile_var_offsets[i], self.bk.tile_var_sizes[i])
        dt = np.dtype(_NP[s.types[i]])
        vals = np.frombuffer(raw, dtype=dt)
        offs = np.concatenate([offs, [len(raw)]]) // dt.itemsize
        assert len(offs) == self.cell_num + 1
        return offs, vals.copy()


def fragment_dirs(array_dir):
    """storage_manager.cc:2169-2207: visible '__*' dirs holding __tiledb_fragment.tdb, ordered by trailing timestamp"""
    ds = [d for d in glob.glob(os.path.join(array_dir, "__*")) if os.path.isdir(d)
          and os.path.exists(os.path.join(d, "__tiledb_fragment.tdb"))]

    def ts(d):
        name = os.path.basename(d)
        return int(name[name.index("_", 2) + 1:])
    return [d for _, _, d in sorted((ts(d), k, d) for k, d in enumerate(sorted(ds)))]


def read_array_meta(array_dir):
    """variant_storage_manager.cc:313-377: lb_row_idx = min over files, max_valid_row_idx_in_array = max over files"""
    lb, ub = None, None
    for f in glob.glob(os.path.join(array_dir, "genomicsdb_meta_dir", "genomicsdb_meta_*.json")):
        d = json.load(open(f))
        if "lb_row_idx" in d:
            lb = d["lb_row_idx"] if lb is None else min(lb, d["lb_row_idx"])
        if "max_valid_row_idx_in_array" in d:
            ub = d["max_valid_row_idx_in_array"] if ub is None else max(ub, d["max_valid_row_idx_in_array"])
    return lb, ub


class Array:
    """All fragments of one array merged into global cell order.

    The merge restates ArrayReadState::sort_fragment_cell_ranges (array_read_state.cc:1836-2132): cells ordered by
    (col,row) [column-major: array_schema.cc:1660-1704]; when two fragments hold the same coordinates the one with the
    larger fragment id (newer) wins and the older cell is dropped (comparator array_read_state.cc:2369-2402)."""

    def __init__(self, workspace, array_name):
        self.dir = os.path.join(workspace, array_name)
        self.schema = ArraySchema(os.path.join(self.dir, "__array_schema.tdb"))
        self.fragments = [Fragment(d, self.schema) for d in fragment_dirs(self.dir)]
        self.lb_row_idx, self.max_valid_row_idx = read_array_meta(self.dir)
        rows, cols, frag = [], [], []
        for k, f in enumerate(self.fragments):
            r, c = f.coords()
            rows.append(r)
            cols.append(c)
            frag.append(np.full(len(r), k, dtype=np.int64))
        if rows:
            rows, cols, frag = np.concatenate(rows), np.concatenate(cols), np.concatenate(frag)
        else:
            rows = cols = frag = np.zeros(0, dtype=np.int64)
        idx_in_frag = np.concatenate([np.arange(f.cell_num) for f in self.fragments]) if self.fragments else rows
        order = np.lexsort((-frag, rows, cols))          # (col, row, newest fragment first)
        rows, cols, frag, idx_in_frag = rows[order], cols[order], frag[order], idx_in_frag[order]
        if len(rows):
            keep = np.ones(len(rows), dtype=bool)
            keep[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
            rows, cols, frag, idx_in_frag = rows[keep], cols[keep], frag[keep], idx_in_frag[keep]
        self.row, self.col, self._frag, self._idx = rows, cols, frag, idx_in_frag
        self.cell_num = len(rows)

    def fixed(self, name):
        parts = [f.fixed(name) for f in self.fragments]
        base = np.cumsum([0] + [f.cell_num for f in self.fragments])
        allv = np.concatenate(parts) if parts else np.zeros((0, 1))
        return allv[base[self._frag] + self._idx]

    def var(self, name):
        """-> (offsets[n+1], values) in merged global cell order"""
        offs_l, vals_l = zip(*[f.var(name) for f in self.fragments]) if self.fragments else ((), ())
        lens = [np.diff(o) for o in offs_l]
        base = np.cumsum([0] + [f.cell_num for f in self.fragments])
        vbase = np.cumsum([0] + [len(v) for v in vals_l])
        all_len = np.concatenate(lens) if lens else np.zeros(0, dtype=np.int64)
        all_start = np.concatenate([o[:-1] + vbase[k] for k, o in enumerate(offs_l)]) if lens else all_len
        allv = np.concatenate(vals_l) if vals_l else np.zeros(0, dtype=np.int32)
        sel = base[self._frag] + self._idx
        ln, st = all_len[sel], all_start[sel]
        out_off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
        if len(ln) == 0:
            return out_off, allv[:0]
        gather = np.repeat(st - out_off[:-1], ln) + np.arange(out_off[-1])
        return out_off, allv[gather]
