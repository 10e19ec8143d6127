"""TEST INFRASTRUCTURE ONLY: CPU restatement of the FORMAT collation + BCF typed encoding of the Broad combine operator.

Input: the canonical hot-path dump (remapped GT / allele-dependent fields per valid call); output: for every site and
field the block htslib would put into bcf1_t.indiv.  Restates, for the fields of the hot path:
  VariantFieldHandler<T>::collect_and_extend_fields   src/main/cpp/src/genomicsdb/variant_field_handler.cc:804-871
      nps = max length over valid calls with a valid field (no such call: the field is not emitted); every query row
      gets nps values: its vector, or [missing] ([vector_end] for GT) when the row has no valid field, then vector_end
  encode_GT_vector<phase?, produce_GT?>                src/main/cpp/src/query_operations/broad_combined_gvcf.cc:53-138,646-658
      allele a -> (a + 1) << 1 | phased; phase slots dropped; !produce_GT_field: every valid allele -> no-call
  G-length fields dropped above max_diploid_alt_alleles  broad_combined_gvcf.cc:620-623
  bcf_enc_vint / bcf_enc_int1 / bcf_enc_size           dependencies/htslib/vcf.c:1885-1972
      int8 when max <= 127 and min > -127, int16 when max <= 32767 and min > -32767 (missing / vector_end excluded), else int32
Pure-python loops: small cases only.  Pinned against the reference's own BCF output (tests/golden/bcf, tests/test_format_oracle.py).
"""
import struct

import numpy as np

I32_MISSING, I32_VEND = -2147483648, -2147483647
F32_MISSING, F32_VEND = 0x7F800001, 0x7F800002
BT_INT8, BT_INT16, BT_INT32, BT_FLOAT, BT_CHAR = 1, 2, 3, 5, 7
TDB_NULL_INT, TDB_NULL_FLT, TDB_NULL_CHAR = 0x7FFFFFFF, 0x7F7FFFFF, 127
STR_MISSING, STR_VEND = 0x07, 0x00


def enc_int1(x):
    """bcf_enc_int1 (vcf.h): one typed integer"""
    if x == I32_VEND:
        return bytes([1 << 4 | BT_INT8, 0x81])
    if x == I32_MISSING:
        return bytes([1 << 4 | BT_INT8, 0x80])
    if -120 <= x <= 127:
        return bytes([1 << 4 | BT_INT8]) + struct.pack("<b", x)
    if -32760 <= x <= 32767:
        return bytes([1 << 4 | BT_INT16]) + struct.pack("<h", x)
    return bytes([1 << 4 | BT_INT32]) + struct.pack("<i", x)


def enc_size(n, bt):
    """bcf_enc_size (vcf.h)"""
    if n >= 15:
        return bytes([15 << 4 | bt]) + enc_int1(n)
    return bytes([n << 4 | bt])


def enc_vint(a, nps):
    """bcf_enc_vint (vcf.c:1926-1972) over int32 values `a`, nps per sample -> (type, descriptor bytes, data bytes)"""
    a = np.asarray(a, dtype=np.int64)
    n = len(a)
    if n == 1:
        b = enc_int1(int(a[0]))
        return b[0] & 0xF, b[:1], b[1:]
    valid = a[(a != I32_MISSING) & (a != I32_VEND)]
    mx = int(valid.max()) if len(valid) else I32_VEND
    mn = int(valid.min()) if len(valid) else 2147483647
    if mx <= 127 and mn > -127:
        out = np.where(a == I32_VEND, -127, np.where(a == I32_MISSING, -128, a)).astype("<i1")
        return BT_INT8, enc_size(nps, BT_INT8), out.tobytes()
    if mx <= 32767 and mn > -32767:
        out = np.where(a == I32_VEND, -32767, np.where(a == I32_MISSING, -32768, a)).astype("<i2")
        return BT_INT16, enc_size(nps, BT_INT16), out.tobytes()
    return BT_INT32, enc_size(nps, BT_INT32), a.astype("<i4").tobytes()


def encode_gt(vec, phased_layout, produce_gt):
    """encode_GT_vector over one sample's padded vector (int32 values) -> BCF genotype ints"""
    def valid(v):
        return v != I32_MISSING and v != I32_VEND

    def enc(v, ph):
        if not valid(v):
            return v
        if not produce_gt:
            return 1 if ph else 0
        return ((((v + 1) << 1) | (1 if ph else 0)) + 2 ** 31) % 2 ** 32 - 2 ** 31
    if not len(vec):
        return []
    out = [enc(vec[0], False)]
    if phased_layout:
        for k in range(2, len(vec), 2):
            out.append(enc(vec[k], vec[k - 1] > 0))
    else:
        for k in range(1, len(vec)):
            out.append(enc(vec[k], False))
    return out


def site_blocks(site, names, query_rows, field_elem, gt_phased, produce_gt, g_fields=(), max_alt=50, java=False):
    """-> {field name: (type, nps, data bytes, descriptor bytes)} for one parsed dump site (oracle.parse_dump).
    java: the conventions of the JNI path (use_missing_values_only_not_vector_end, variant_field_handler.cc:845-866): rows are
    padded with missing instead of vector_end, a row without GT gets the no-call allele"""
    by_row = {c["row"]: c for c in site["calls"]}
    n_alt = len(site["alleles"]) - 1
    out = {}
    for nm in names:
        if nm in g_fields and site["remap"] and n_alt > max_alt:
            continue
        is_gt = nm == "GT"
        is_float = field_elem.get(nm) == "float"
        lens = [len(c["fields"][nm]) for c in site["calls"]]
        nps = max(lens) if lens else 0
        if nps == 0:
            continue
        miss, vend = (F32_MISSING, F32_VEND) if is_float else (I32_MISSING, I32_VEND)
        rows = []
        for r in query_rows:
            c = by_row.get(r)
            v = [] if c is None else list(c["fields"][nm].view(np.uint32 if is_float else np.int32))
            v = [int(x) for x in v]
            if not v:
                v = [(-1 if java else vend) if is_gt else miss]
            v += [miss if java else vend] * (nps - len(v))
            rows.append(v)
        if is_gt:
            rows = [encode_gt(v, gt_phased, produce_gt) for v in rows]
            nps = len(rows[0])
        flat = [x for v in rows for x in v]
        if is_float:
            out[nm] = (BT_FLOAT, nps, np.asarray(flat, dtype="<u4").tobytes(), enc_size(nps, BT_FLOAT))
        else:
            bt, desc, data = enc_vint(flat, nps)
            out[nm] = (bt, nps, data, desc)
    return out


# ---------------------------------------------------------------------------------------------- plain FORMAT fields
class PlainColumns:
    """Values of the FORMAT attributes that are not allele-dependent, per cell of an array in global cell order
    (oracle.tiledb_fragment.Array), and the (row, begin column) -> begin cell map the dump's calls are looked up with."""

    def __init__(self, arr, end, fields):
        """fields: list of (attribute name, 'int'|'float'|'char', fixed length or 0)"""
        self.fields = fields
        self.cols = {}
        for name, elem, fixed in fields:
            if fixed:
                v = np.ascontiguousarray(arr.fixed(name))
                self.cols[name] = (None, v.view(np.uint32) if elem != "char" else v.view(np.uint8))
            else:
                off, val = arr.var(name)
                val = np.ascontiguousarray(val)
                self.cols[name] = (off, val.view(np.uint8) if elem == "char" else val.view(np.uint32))
        begin = arr.col <= end
        self.cell_of = {(int(r), int(c)): int(i) for i, (r, c, b) in enumerate(zip(arr.row, arr.col, begin)) if b}
        self._ref = arr.var("REF")
        self._alt = arr.var("ALT")

    def contains_deletion(self, cell):
        """VariantUtils::contains_deletion (src/main/cpp/src/utils/known_field_info.cc:310-319) over the cell's REF / ALT"""
        ro, rv = self._ref
        ao, av = self._alt
        ref_len = int(ro[cell + 1] - ro[cell])
        if ref_len <= 1:
            return False
        alt = bytes(np.asarray(av[ao[cell]:ao[cell + 1]], dtype=np.uint8)).split(b"\0")[0]
        for t in alt.split(b"|"):
            if not t:
                continue
            symbolic = t[:1] == b"&" or t == b"*" or (t[:1] == b"<" and t[-1:] == b">") or b"[" in t or b"]" in t
            if not symbolic and len(t) < ref_len:
                return True
        return False

    def values(self, name, elem, cell):
        """-> list of the field's elements in `cell`, [] when the field is invalid (every element TileDB-null)"""
        off, val = self.cols[name]
        v = val[cell] if off is None else val[off[cell]:off[cell + 1]]
        v = [int(x) for x in np.atleast_1d(v)]
        null = {"int": TDB_NULL_INT, "float": TDB_NULL_FLT, "char": TDB_NULL_CHAR}[elem]
        return v if any(x != null for x in v) else []


def plain_site_blocks(site, query_rows, plain, java=False):
    """collect_and_extend_fields + bcf_update_format for the plain FORMAT fields of one parsed dump site
    -> {attribute name: (type, nps, data bytes, descriptor bytes)}"""
    by_row = {c["row"]: c for c in site["calls"]}
    out = {}
    for name, elem, _ in plain.fields:
        vecs = {}
        for r, c in by_row.items():
            vecs[r] = plain.values(name, elem, plain.cell_of[(r, c["begin"])])
        nps = max([len(v) for v in vecs.values()] or [0])
        if nps == 0:
            continue
        if elem == "char":
            miss, vend = STR_MISSING, STR_VEND
        elif elem == "float":
            miss, vend = F32_MISSING, F32_VEND
        else:
            miss, vend = I32_MISSING, I32_VEND
        flat = []
        for r in query_rows:
            v = list(vecs.get(r, []))
            if elem == "int":
                v = [x - (1 << 32) if x >= (1 << 31) else x for x in v]
            if not v:
                v = [vend if (java and elem == "char") else miss]          # strings: use_vector_end_only in the JNI path
            flat += v + [miss if (java and elem != "char") else vend] * (nps - len(v))
        if elem == "char":
            out[name] = (BT_CHAR, nps, bytes(flat), enc_size(nps, BT_CHAR))
        elif elem == "float":
            out[name] = (BT_FLOAT, nps, np.asarray(flat, dtype="<u4").tobytes(), enc_size(nps, BT_FLOAT))
        else:
            bt, desc, data = enc_vint(flat, nps)
            out[name] = (bt, nps, data, desc)
    return out


def assemble_indiv(order, blocks, key_ids):
    """bcf1_t.indiv as htslib serialises it: per emitted field typed key, size descriptor, data; `order` = VCF names in
    m_FORMAT_fields_vec order with GT first and DP last (broad_combined_gvcf.cc:168-240, vcf.c:3670-3676)"""
    b = b""
    for vcf_name in order:
        if vcf_name in blocks:
            bt, nps, data = blocks[vcf_name][:3]
            b += enc_int1(key_ids[vcf_name]) + enc_size(nps, bt) + data
    return b


# ---------------------------------------------------------------------------------------------- INFO scalars
def info_site_values(site, plain, info_fields, dp_fields=(None, None, None)):
    """BroadCombinedGVCFOperator::handle_INFO_fields for scalar fields (broad_combined_gvcf.cc:524-545) + the DP rule
    (:688-727) at one parsed dump site.  info_fields: [(attribute, 'int'|'float', 'median'|'sum'|'mean')];
    dp_fields: attribute names of (INFO DP, MIN_DP, DP_FORMAT) or None.
    -> ({attribute: 4-byte result as uint32 bits}, dp_sum)"""
    calls = sorted(site["calls"], key=lambda c: c["row"])
    cells = [plain.cell_of[(c["row"], c["begin"])] for c in calls]
    # a call seen past its begin that contains a deletion had its INFO fields invalidated (broad_combined_gvcf.cc:1067-1075)
    reduced = [site["begin"] > c["begin"] and plain.contains_deletion(cell) for c, cell in zip(calls, cells)]
    out = {}
    for name, elem, op in info_fields:
        bad = (F32_MISSING, F32_VEND) if elem == "float" else (I32_MISSING & 0xFFFFFFFF, I32_VEND & 0xFFFFFFFF)
        vals = []
        for cell, red in zip(cells, reduced):
            v = plain.values(name, elem, cell)
            if red or not v or v[0] in bad:
                continue
            vals.append(v[0])
        if not vals:
            continue
        a = np.asarray(vals, dtype=np.uint32).view(np.float32 if elem == "float" else np.int32)
        if op == "median":                      # std::nth_element at n/2 (variant_field_handler.cc:529-560)
            r = np.sort(a)[len(a) // 2]
        else:                                   # left-to-right sum in the field's own type (:562-607)
            r = a.dtype.type(0)
            with np.errstate(over="ignore"):
                for x in a:
                    r = a.dtype.type(r + x)
                if op == "mean":
                    r = a.dtype.type(r / a.dtype.type(len(a))) if elem == "float" else np.int32(np.uint32(r) // np.uint32(len(a)))
        out[name] = int(np.asarray(r).view(np.uint32))
    dp_sum = 0
    dp, min_dp, dp_format = dp_fields

    def scalar(name, cell):
        if name is None:
            return None
        v = plain.values(name, "int", cell)
        if not v:
            return None
        x = v[0] - (1 << 32) if v[0] >= (1 << 31) else v[0]
        return None if x in (I32_MISSING, I32_VEND) else x
    if dp is not None or dp_format is not None:
        for cell in cells:
            a = scalar(dp, cell)
            if a is None:
                a = scalar(min_dp, cell)
                if a is None:
                    a = scalar(dp_format, cell)
            dp_sum += a or 0
    dp_sum = (dp_sum + 2 ** 31) % 2 ** 32 - 2 ** 31
    return out, dp_sum


def info_vector_sums(site, plain, remapped_fields=(), plain_fields=()):
    """compute_valid_element_wise_sum (variant_field_handler.cc:618-660) at one parsed dump site: remapped_fields = [(dump
    field name, 'int'|'float')] take each valid call's remapped vector, plain_fields = [(attribute, elem)] the cell's own.
    -> {name: list of uint32 words} (fields without any valid value are absent)"""
    calls = sorted(site["calls"], key=lambda c: c["row"])
    cells = [plain.cell_of[(c["row"], c["begin"])] for c in calls]
    reduced = [site["begin"] > c["begin"] and plain.contains_deletion(cell) for c, cell in zip(calls, cells)]
    out = {}
    for name, elem, is_remapped in [(n, e, True) for n, e in remapped_fields] + [(n, e, False) for n, e in plain_fields]:
        bad = (F32_MISSING, F32_VEND) if elem == "float" else (I32_MISSING & 0xFFFFFFFF, I32_VEND & 0xFFFFFFFF)
        dt = np.float32 if elem == "float" else np.int32
        res, n_valid = [], 0
        for c, cell, red in zip(calls, cells, reduced):
            if red:
                continue
            vec = [int(x) for x in c["fields"][name]] if is_remapped else plain.values(name, elem, cell)
            for i, x in enumerate(vec):
                if x in bad:
                    continue
                while len(res) <= i:
                    res.append(None)
                v = np.asarray([x], dtype=np.uint32).view(dt)[0]
                with np.errstate(over="ignore"):
                    res[i] = v if res[i] is None else dt(res[i] + v)
                n_valid = max(n_valid, i + 1)
        if n_valid:
            miss = F32_MISSING if elem == "float" else I32_MISSING & 0xFFFFFFFF
            out[name] = [miss if r is None else int(np.asarray([r], dtype=dt).view(np.uint32)[0]) for r in res[:n_valid]]
    return out


def info_vector_concat(site, plain, plain_fields=()):
    """VariantFieldHandler::concatenate_field (variant_field_handler.cc:743-800) at one parsed dump site: the vectors of the
    valid calls appended in row order, nothing dropped (a missing value inside a vector stays).  plain_fields = [(attribute,
    elem)].  -> {name: list of uint32 words} (fields no valid call has are absent)"""
    calls = sorted(site["calls"], key=lambda c: c["row"])
    out = {}
    for name, elem in plain_fields:
        res = []
        for c in calls:
            res += plain.values(name, elem, plain.cell_of[(c["row"], c["begin"])])
        if res:
            out[name] = res
    return out


def enc_info_vector(words, elem):
    """typed vector of an INFO field as bcf_update_info writes it (bcf_enc_vint with wsize = -1 / bcf_enc_vfloat)"""
    if elem == "float":
        return enc_size(len(words), BT_FLOAT) + np.asarray(words, dtype="<u4").tobytes()
    vals = [w - (1 << 32) if w >= (1 << 31) else w for w in words]
    bt, desc, data = enc_vint(vals, len(vals))
    return desc + data


def enc_info_value(bits, elem):
    """typed value of a one-element INFO field as bcf_update_info writes it: ints through bcf_enc_int1, floats as float32"""
    if elem == "float":
        return bytes([1 << 4 | BT_FLOAT]) + struct.pack("<I", bits)
    return enc_int1(bits - (1 << 32) if bits >= (1 << 31) else bits)
