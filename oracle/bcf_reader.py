"""TEST INFRASTRUCTURE ONLY: minimal reader of uncompressed BCF2 (what the reference writes with `-O bu`).

Restates the BCF2 record layout htslib serialises (dependencies/htslib/vcf.c: bcf_write1 / bcf_serialize side:
vcf.c:1805-1972 typed values, vcf.c:3607-3700 bcf_update_format) just far enough to cut a record into its FORMAT blocks:
    file   = "BCF\\2\\2", u32 l_text, header text, records
    record = u32 l_shared, u32 l_indiv, shared[l_shared], indiv[l_indiv]
    shared = i32 chrom, i32 pos, i32 rlen, f32 qual, u32 n_allele<<16|n_info, u32 n_fmt<<24|n_sample, ...
    indiv  = n_fmt x ( typed-int key, typed descriptor (nps, type), n_sample * nps * size bytes )
"""
import re
import struct

BT_SIZE = {1: 1, 2: 2, 3: 4, 5: 4, 7: 1}


def _typed_int(b, off):
    """one typed scalar integer (bcf_enc_int1's output) -> (value, new offset)"""
    t = b[off]
    assert (t >> 4) == 1, "expected a single typed int"
    bt = t & 0xF
    off += 1
    if bt == 1:
        return struct.unpack_from("<b", b, off)[0], off + 1
    if bt == 2:
        return struct.unpack_from("<h", b, off)[0], off + 2
    assert bt == 3
    return struct.unpack_from("<i", b, off)[0], off + 4


def _descriptor(b, off):
    """typed descriptor byte(s) (bcf_enc_size's output) -> (n, type, new offset)"""
    t = b[off]
    n, bt = t >> 4, t & 0xF
    off += 1
    if n == 15:
        n, off = _typed_int(b, off)
    return n, bt, off


def header_dict(text):
    """BCF dictionary of strings: ID -> index for FILTER/INFO/FORMAT lines (IDX= when present, else order of appearance)."""
    ids = {}
    nxt = 0
    for line in text.split("\n"):
        m = re.match(r"##(FILTER|INFO|FORMAT)=<(.*)>", line)
        if not m:
            continue
        body = m.group(2)
        name = re.search(r"ID=([^,>]+)", body).group(1)
        idx = re.search(r"IDX=(\d+)", body)
        if name in ids:
            continue
        if idx:
            ids[name] = int(idx.group(1))
            nxt = max(nxt, ids[name] + 1)
        else:
            ids[name] = nxt
            nxt += 1
    return ids


def read_bcf(path):
    """-> (header text, sample names, [record dict(pos, n_allele, n_sample, alleles, indiv=raw bytes, fmt_order,
    fmt={name: (type, nps, bytes, key_and_descriptor_bytes)})])"""
    b = open(path, "rb").read()
    assert b[:5] == b"BCF\x02\x02", "not an uncompressed BCF2 file"
    l_text = struct.unpack_from("<I", b, 5)[0]
    text = b[9:9 + l_text].rstrip(b"\0").decode()
    ids = header_dict(text)
    names = {v: k for k, v in ids.items()}
    samples = []
    for line in text.split("\n"):
        if line.startswith("#CHROM"):
            samples = line.split("\t")[9:]
    off = 9 + l_text
    recs = []
    while off < len(b):
        l_shared, l_indiv = struct.unpack_from("<II", b, off)
        off += 8
        sh = b[off:off + l_shared]
        ind = b[off + l_shared:off + l_shared + l_indiv]
        off += l_shared + l_indiv
        chrom, pos, rlen, qual, na_ni, nf_ns = struct.unpack_from("<iiifII", sh, 0)
        n_allele, n_info = na_ni >> 16, na_ni & 0xFFFF
        n_fmt, n_sample = nf_ns >> 24, nf_ns & 0xFFFFFF
        o = 24
        n, bt, o = _descriptor(sh, o)      # ID
        rec_id = sh[o:o + n].decode()
        o += n
        alleles = []
        for _ in range(n_allele):
            n, bt, o = _descriptor(sh, o)
            alleles.append(sh[o:o + n].decode())
            o += n
        n, bt, o = _descriptor(sh, o)      # FILTER: typed vector of dictionary indices
        flt = [int.from_bytes(sh[o + k * BT_SIZE[bt]:o + (k + 1) * BT_SIZE[bt]], "little", signed=True) for k in range(n)] if n else []
        o += n * (BT_SIZE[bt] if n else 0)
        info = {}
        for _ in range(n_info):
            key, o = _typed_int(sh, o)
            n, bt, o = _descriptor(sh, o)
            nb = n * (BT_SIZE[bt] if n else 0)
            info[names[key]] = (bt, n, sh[o:o + nb])
            o += nb
        assert o == len(sh)
        fmt = {}
        o = 0
        for _ in range(n_fmt):
            o0 = o
            key, o = _typed_int(ind, o)
            nps, bt, o = _descriptor(ind, o)
            nbytes = n_sample * nps * BT_SIZE[bt]
            fmt[names[key]] = (bt, nps, ind[o:o + nbytes], ind[o0:o])
            o += nbytes
        assert o == len(ind)
        recs.append(dict(chrom=chrom, pos=pos, rlen=rlen, qual_bits=struct.unpack_from("<I", sh, 12)[0], n_allele=n_allele, n_sample=n_sample,
                         id=rec_id, alleles=alleles, filter=flt, info=info, info_order=list(info), fmt=fmt, fmt_order=list(fmt),
                         shared=bytes(sh), indiv=bytes(ind)))
    return text, samples, recs
