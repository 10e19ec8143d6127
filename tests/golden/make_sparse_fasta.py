#!/usr/bin/env python3
"""Writes tests/golden/inputs/ref_sparse.fasta.gz (+ .fai, .gzi): a bgzip-compressed, indexed FASTA of contig "1" that holds the
reference bases the golden scenarios look up (tests/golden/inputs/ref_bases.tsv: 0-based position, base -- read from the
reference's own tests/inputs/chr1_10MB.fasta.gz when the goldens were made) and N everywhere else.  It stands in for that
2.8 MB file in the tests of the generated-header / reference-genome path: same format (BGZF blocks, .gzi block index, .fai),
a few kilobytes.  Usage: python tests/golden/make_sparse_fasta.py"""
import os
import struct
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
LINE = 80


def bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    bsize = 18 + len(body) + 8
    hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def main():
    bases = {}
    for ln in open(os.path.join(HERE, "inputs", "ref_bases.tsv")):
        contig, pos, base = ln.split()
        assert contig == "1"
        bases[int(pos)] = base
    length = (max(bases) // LINE + 2) * LINE
    seq = bytearray(b"N" * length)
    for p, b in bases.items():
        seq[p] = ord(b)
    text = bytearray(b">1\n")
    for i in range(0, length, LINE):
        text += seq[i:i + LINE] + b"\n"
    out = os.path.join(HERE, "inputs", "ref_sparse.fasta.gz")
    index = []
    with open(out, "wb") as f:
        u = 0
        for i in range(0, len(text), 65280):
            if i:
                index.append((f.tell(), u))
            chunk = bytes(text[i:i + 65280])
            f.write(bgzf_block(chunk))
            u += len(chunk)
        f.write(bgzf_block(b""))   # the empty block that ends a BGZF file
    with open(out + ".gzi", "wb") as f:
        f.write(struct.pack("<Q", len(index)))
        for c, uo in index:
            f.write(struct.pack("<QQ", c, uo))
    with open(out + ".fai", "w") as f:
        f.write("1\t%d\t3\t%d\t%d\n" % (length, LINE, LINE + 1))
    print(out, os.path.getsize(out), "bytes,", len(index) + 1, "blocks")


if __name__ == "__main__":
    main()
