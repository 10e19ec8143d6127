"""Canonical hot-path dump (layout: oracle/ref_harness/gdb_ref_dump.cc) written from gdbgpu batches.

Host-side formatting only: every number in the dump was produced by the CUDA kernels and read through the C ABI.
"""
import struct

import numpy as np

from genomicsdb_b200 import capi

MAGIC = 0x31444247


def header(field_names):
    b = struct.pack("<II", MAGIC, len(field_names))
    for n in field_names:
        b += struct.pack("<I", len(n)) + n.encode()
    return b


def batch_records(bv, coords, end):
    """Serialises one BatchView; coords/end are the array's merged cell coordinates (Array.cells())."""
    out = []
    for s in range(bv.n_sites):
        alleles = bv.alleles(s)
        fl = int(bv.site_flags[s])
        rec = [struct.pack("<qqBBH", int(bv.site_begin[s]), int(bv.site_end[s]), fl & 1, (fl >> 1) & 1, len(alleles) - 1)]
        for a in alleles:
            rec.append(struct.pack("<I", len(a)) + a)
        live = bv.live[s]
        rows = np.nonzero(live >= 0)[0]
        rec.append(struct.pack("<I", len(rows)))
        for q in rows:
            e = int(live[q])
            cell = e & (capi.CALL_REDUCED - 1)
            row, col = int(coords[cell, 0]), int(coords[cell, 1])
            END = int(end[cell])
            begin, stop = (END, col) if col > END else (col, END)
            lut = bv.lut(s, q)
            rec.append(struct.pack("<qqqH", row, begin, stop, len(lut)))
            rec.append(lut.astype("<i4").tobytes())
            for f in bv.fields:
                n = int(f["len"][s, q])
                rec.append(struct.pack("<I", n))
                if n:
                    o = int(f["base"][s]) + int(q) * int(f["stride"][s])
                    rec.append(f["data"][o:o + n].tobytes())
        out.append(b"".join(rec))
    return b"".join(out)


def combine_dump(arr, column_ranges, field_names, query_rows=None, flags=0, max_sites=0, max_bytes=0):
    """Runs the scan(s) on the device and returns (canonical dump incl. header, number of sites)."""
    coords, end = arr.cells()
    out = [header(field_names)]
    total = 0
    ranges = sorted(tuple(r) for r in column_ranges) if column_ranges else [None]
    for r in ranges:
        sc = capi.Scan(arr, None if r is None else r[0], None if r is None else r[1], query_rows=query_rows,
                       flags=flags | capi.SCAN_WANT_LUT)
        try:
            while True:
                bv = sc.next_batch(max_sites, max_bytes)
                if bv is None:
                    break
                out.append(batch_records(bv, coords, end))
                total += bv.n_sites
        finally:
            sc.close()
    return b"".join(out), total
