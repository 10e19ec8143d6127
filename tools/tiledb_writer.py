"""Writes GDBSYN-1 columns as a GenomicsDB/TileDB 0.5.x array on disk (one uncompressed fragment) so that the
UNMODIFIED reference (oracle/_ref) can read exactly the bytes the GPU path is given.

Tooling for bench.py's reference arm and the tests.  Formats (restated from the reference):
  __array_schema.tdb        dependencies/TileDB/core/src/array/array_schema.cc:569-594
  __book_keeping.tdb.gz     dependencies/TileDB/core/src/fragment/book_keeping.cc:217-243
  <ATTR>.tdb/_var.tdb       packed values / absolute uint64 byte offsets + value stream
  genomicsdb_meta_dir/*.json  src/main/cpp/src/genomicsdb/variant_storage_manager.cc:382-432
plus the vid / callset mapping JSON the query needs (src/main/cpp/src/utils/vid_mapper.cc).
"""
import gzip
import json
import os
import struct
import zlib

import numpy as np

T_INT32, T_INT64, T_FLOAT32, T_CHAR = 0, 1, 2, 4
VAR = 0x7FFFFFFF
I64_MAX = (1 << 63) - 1

# (name, tiledb type, cell_val_num) in the order GenomicsDB builds them (vid_mapper.cc:354-442)
ATTRS = [("END", T_INT64, 1), ("REF", T_CHAR, VAR), ("ALT", T_CHAR, VAR), ("QUAL", T_FLOAT32, 1), ("FILTER", T_INT32, VAR),
         ("DP", T_INT32, 1), ("GQ", T_INT32, 1), ("AD", T_INT32, VAR), ("PL", T_INT32, VAR), ("MIN_DP", T_INT32, 1),
         ("GT", T_INT32, VAR)]

VID = {
    "fields": {
        "END": {"vcf_field_class": ["INFO"], "type": "int"},
        "DP": {"vcf_field_class": ["FORMAT"], "type": "int"},
        "GQ": {"vcf_field_class": ["FORMAT"], "type": "int"},
        "AD": {"vcf_field_class": ["FORMAT"], "type": "int", "length": "R"},
        "PL": {"vcf_field_class": ["FORMAT"], "type": "int", "length": "G"},
        "MIN_DP": {"vcf_field_class": ["FORMAT"], "type": "int"},
        "GT": {"vcf_field_class": ["FORMAT"], "type": "int", "length": "P"},
    },
    "contigs": {"1": {"length": 249250621, "tiledb_column_offset": 0}},
}

VCF_HEADER = """##fileformat=VCFv4.2
##ALT=<ID=NON_REF,Description="Represents any possible alternative allele at this location">
##FILTER=<ID=LowQual,Description="Low quality">
##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic depths for the ref and alt alleles in the order listed">
##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Approximate read depth">
##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="Minimum DP observed within the GVCF block">
##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Normalized, Phred-scaled likelihoods for genotypes">
##INFO=<ID=END,Number=1,Type=Integer,Description="Stop position of the interval">
##contig=<ID=1,length=249250621>
#CHROM	POS	ID	REF	ALT	QUAL	FILTER	INFO
"""


def _schema_bytes(array_path, capacity, compress=False):
    b = struct.pack("<I", 0xFFFFFFFF) + struct.pack("<i", 0)            # version tag, empty workspace
    name = array_path.encode()
    b += struct.pack("<i", len(name)) + name
    b += struct.pack("<Bbbq", 0, 0, 1, capacity)                          # sparse, tile order row, cell order col-major
    b += struct.pack("<i", len(ATTRS))
    for n, _, _ in ATTRS:
        b += struct.pack("<i", len(n)) + n.encode()
    b += struct.pack("<i", 2)
    for n in ("samples", "position"):
        b += struct.pack("<i", len(n)) + n.encode()
    b += struct.pack("<i", 32) + struct.pack("<qqqq", 0, I64_MAX - 1, 0, I64_MAX - 1)
    b += struct.pack("<i", 0)                                             # no tile extents
    b += bytes([t for _, t, _ in ATTRS] + [T_INT64])
    for _, _, c in ATTRS:
        b += struct.pack("<i", c)
    b += bytes([1 if compress else 0] * (len(ATTRS) + 1))                 # compression per attribute (+ coords): 1 = GZIP
    return b


def _book_keeping_bytes(coords, capacity, tables=None):
    """tables (compressed fragments): dict(tile_offsets=[per attr + coords], var_offsets=[per attr], var_sizes=[per attr])"""
    n = len(coords)
    tiles = (n + capacity - 1) // capacity
    b = struct.pack("<Q", 32) + struct.pack("<qqqq", 0, I64_MAX - 1, 0, I64_MAX - 1)
    pad = tiles * capacity - n
    rows = np.concatenate([coords[:, 0], np.full(pad, coords[-1, 0])]).reshape(tiles, capacity)
    cols = np.concatenate([coords[:, 1], np.full(pad, coords[-1, 1])]).reshape(tiles, capacity)
    mbr = np.stack([rows.min(1), rows.max(1), cols.min(1), cols.max(1)], axis=1).astype("<i8")
    first = coords[np.arange(tiles) * capacity]
    last = coords[np.minimum(np.arange(1, tiles + 1) * capacity, n) - 1]
    bc = np.concatenate([first, last], axis=1).astype("<i8")
    b += struct.pack("<q", tiles) + mbr.tobytes()
    b += struct.pack("<q", tiles) + bc.tobytes()
    if tables is None:
        b += struct.pack("<q", 0) * (len(ATTRS) + 1)      # tile offsets: empty when uncompressed
        b += struct.pack("<q", 0) * len(ATTRS)            # var tile offsets
        b += struct.pack("<q", 0) * len(ATTRS)            # var tile sizes
    else:                                                 # book_keeping.cc:217-243
        for t in tables["tile_offsets"]:
            b += struct.pack("<q", len(t)) + np.asarray(t, dtype="<i8").tobytes()
        for t in tables["var_offsets"]:
            b += struct.pack("<q", len(t)) + np.asarray(t, dtype="<i8").tobytes()
        for t in tables["var_sizes"]:
            b += struct.pack("<q", len(t)) + np.asarray(t, dtype="<u8").tobytes()
    last_n = n - (tiles - 1) * capacity
    b += struct.pack("<q", last_n)
    return b


def _write_tiles(path, data, tile_bytes, level=6):
    """one zlib stream per tile (ReadState::decompress_tile_gzip); returns the tile offsets in the file"""
    offs, pos = [], 0
    with open(path, "wb") as f:
        for lo in range(0, len(data), tile_bytes) if len(data) else []:
            z = zlib.compress(data[lo:lo + tile_bytes], level)
            offs.append(pos)
            f.write(z)
            pos += len(z)
    return offs


def subset_cells(d, idx, bump_pl=None):
    """the cells `idx` (ascending) of a GDBSYN-1 column set as a column set of its own; bump_pl: positions in idx whose first PL value
    is raised by one (a re-imported call that differs from the one it shadows)"""
    idx = np.asarray(idx, dtype=np.int64)
    out = dict(d)
    out["coords"] = d["coords"][idx]
    out["end"] = d["end"][idx]
    for key in ("dp", "gq", "min_dp"):
        if d.get(key) is not None:
            out[key] = d[key][idx]
    for key in ("ref", "alt", "gt", "ad", "pl"):
        off, val = d[key]
        vb = val.view(np.uint8)
        ends = np.concatenate([off[1:], [vb.nbytes]]).astype(np.int64)
        lens = ends[idx] - off[idx].astype(np.int64)
        noff = np.zeros(len(idx), dtype=np.uint64)
        if len(idx):
            noff[1:] = np.cumsum(lens)[:-1]
        total = int(lens.sum())
        # gather the byte ranges: index of every output byte
        src = np.repeat(off[idx].astype(np.int64) - noff.astype(np.int64), lens) + np.arange(total, dtype=np.int64)
        nval = vb[src].copy()
        if key == "pl" and bump_pl is not None and len(bump_pl):
            w = nval.view("<i4")
            at = (noff[bump_pl] // 4).astype(np.int64)
            at = at[lens[bump_pl] >= 4]
            w[at] += 1
        out[key] = (noff, nval.view(val.dtype) if total % val.dtype.itemsize == 0 else nval)
    return out


def _write_fragment(fdir, d, capacity, compress):
    os.makedirs(fdir, exist_ok=True)
    open(os.path.join(fdir, "__tiledb_fragment.tdb"), "wb").close()
    n = len(d["end"])
    qual = np.full(n, np.float32(3.4028234663852886e38), dtype="<f4")    # TileDB NULL
    fixed = {"END": d["end"].astype("<i8"), "QUAL": qual}
    for name, key in (("DP", "dp"), ("GQ", "gq"), ("MIN_DP", "min_dp")):
        fixed[name] = (d[key] if d.get(key) is not None else np.full(n, 0x7FFFFFFF, dtype=np.int32)).astype("<i4")
    var = {name: (d[key][0].astype("<u8"), d[key][1]) for name, key in (("REF", "ref"), ("ALT", "alt"), ("GT", "gt"), ("AD", "ad"), ("PL", "pl"))}
    var["FILTER"] = (np.zeros(n, dtype="<u8"), np.zeros(0, dtype=np.uint8))
    if not compress:
        d["coords"].astype("<i8").tofile(os.path.join(fdir, "__coords.tdb"))
        for name, v in fixed.items():
            v.tofile(os.path.join(fdir, name + ".tdb"))
        for name, (off, val) in var.items():
            off.tofile(os.path.join(fdir, name + ".tdb"))
            val.tofile(os.path.join(fdir, name + "_var.tdb"))
        tables = None
    else:
        tables = dict(tile_offsets=[], var_offsets=[], var_sizes=[])
        for name, _, cvn in ATTRS:
            if cvn != VAR:
                b = fixed[name].tobytes()
                tables["tile_offsets"].append(_write_tiles(os.path.join(fdir, name + ".tdb"), b, capacity * fixed[name].itemsize))
                tables["var_offsets"].append([])
                tables["var_sizes"].append([])
                continue
            off, val = var[name]
            vb = val.tobytes()
            tables["tile_offsets"].append(_write_tiles(os.path.join(fdir, name + ".tdb"), off.tobytes(), capacity * 8))
            # the value bytes of the cells of every tile form one stream; sizes are the decompressed ones
            bounds = [int(off[i]) for i in range(0, n, capacity)] + [len(vb)]
            voffs, vsizes, pos = [], [], 0
            with open(os.path.join(fdir, name + "_var.tdb"), "wb") as f:
                for lo, hi in zip(bounds[:-1], bounds[1:]):
                    voffs.append(pos)
                    vsizes.append(hi - lo)
                    if hi > lo:
                        z = zlib.compress(vb[lo:hi], 6)
                        f.write(z)
                        pos += len(z)
            tables["var_offsets"].append(voffs)
            tables["var_sizes"].append(vsizes)
        tables["tile_offsets"].append(_write_tiles(os.path.join(fdir, "__coords.tdb"), d["coords"].astype("<i8").tobytes(), capacity * 16))
    with gzip.open(os.path.join(fdir, "__book_keeping.tdb.gz"), "wb", compresslevel=1) as f:
        f.write(_book_keeping_bytes(d["coords"], capacity, tables))
    return tables


def write_array(workspace, array_name, d, capacity=1000, compress=False, n_fragments=1, dup_every=0):
    """d: tools.gdbsyn.generate(..., with_fixed=True).  Returns dict(workspace, array, vid, callsets, vcf_header, fasta).
    n_fragments > 1: the cells go to that many fragments as batches of samples would (fragment f holds the rows r with r % n == f, each
    spans every column), named so that a higher f is newer; dup_every > 0: fragment f also holds every dup_every-th cell of fragment
    f - 1 again with a changed PL (a re-import), so that the older copy is shadowed (array_read_state.cc:2369-2402)."""
    os.makedirs(workspace, exist_ok=True)
    open(os.path.join(workspace, "__tiledb_workspace.tdb"), "wb").close()
    adir = os.path.join(workspace, array_name)
    os.makedirs(os.path.join(adir, "genomicsdb_meta_dir"), exist_ok=True)
    open(os.path.join(adir, ".__consolidation_lock"), "wb").close()
    open(os.path.join(adir, "__array_schema.tdb"), "wb").write(_schema_bytes(os.path.abspath(adir), capacity, compress))
    json.dump({"lb_row_idx": 0, "max_valid_row_idx_in_array": int(d["num_rows"]) - 1},
              open(os.path.join(adir, "genomicsdb_meta_dir", "genomicsdb_meta_00000000-0000-4000-8000-000000000000.json"), "w"))
    if n_fragments <= 1:
        fdir = os.path.join(adir, "__00000000-0000-4000-8000-0000000000001_1700000000000")
        tables = _write_fragment(fdir, d, capacity, compress)
    else:
        rows = d["coords"][:, 0]
        begin = d["coords"][:, 1] <= d["end"]          # not an END copy
        for f in range(n_fragments):
            own = np.nonzero(rows % n_fragments == f)[0]
            extra = np.zeros(0, dtype=np.int64)
            if dup_every and f > 0:
                extra = np.nonzero((rows % n_fragments == f - 1) & begin)[0][::dup_every]
            idx = np.sort(np.concatenate([own, extra]))
            bump = np.nonzero(np.isin(idx, extra))[0]
            fdir = os.path.join(adir, "__00000000-0000-4000-8000-%013d_%d" % (f + 1, 1700000000000 + f))
            tables = _write_fragment(fdir, subset_cells(d, idx, bump), capacity, compress)
    # mapping files + template header + a reference FASTA long enough for the columns used
    vid_path = os.path.join(workspace, "vid.json")
    json.dump(VID, open(vid_path, "w"))
    cs_path = os.path.join(workspace, "callsets.json")
    with open(cs_path, "w") as f:
        f.write('{"callsets":{')
        f.write(",".join('"S%d":{"row_idx":%d,"idx_in_file":0,"filename":"s%d.vcf.gz"}' % (r, r, r) for r in range(int(d["num_rows"]))))
        f.write("}}")
    hdr_path = os.path.join(workspace, "template_vcf_header.vcf")
    open(hdr_path, "w").write(VCF_HEADER)
    fasta = os.path.join(workspace, "ref.fasta")
    length = int(max(d["coords"][:, 1].max(), d["end"].max())) + 1000
    if not os.path.exists(fasta) or os.path.getsize(fasta) < length:
        line = 60
        nlines = (length + line - 1) // line
        with open(fasta, "wb") as f:
            f.write(b">1\n")
            f.write((b"ACGT" * 15 + b"\n") * nlines)
        open(fasta + ".fai", "w").write("1\t%d\t3\t%d\t%d\n" % (nlines * line, line, line + 1))
    return dict(workspace=workspace, array=array_name, vid=vid_path, callsets=cs_path, vcf_header=hdr_path, fasta=fasta,
                fragment_dir=fdir, tile_tables=tables)


def write_query_json(path, info, col_lo=None, col_hi=None, attributes=("REF", "ALT", "GT", "AD", "PL"), segment_size=10485760):
    q = {"workspace": info["workspace"], "array_name": info["array"], "vid_mapping_file": info["vid"],
         "callset_mapping_file": info["callsets"], "vcf_header_filename": [info["vcf_header"]],
         "reference_genome": info["fasta"], "attributes": list(attributes), "segment_size": segment_size}
    if col_lo is None:
        q["scan_full"] = True
    else:
        q["query_column_ranges"] = [{"range_list": [{"low": int(col_lo), "high": int(col_hi)}]}]
    json.dump(q, open(path, "w"))
    return path
