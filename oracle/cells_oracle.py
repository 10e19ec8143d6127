"""TEST INFRASTRUCTURE ONLY: sequential restatement of the reference's columnar iterator and the operators on it.

Follows (GenomicsDB v0.10.2, paths under src/main/cpp/):
    SingleCellTileDBIterator                src/genomicsdb/genomicsdb_iterators.cc:181-300 (intervals), 425-472 (markers),
                                            474-510 (operator++), 535-636 (useful cells)
    GenomicsDBLiveCellMarker PQ order       include/genomicsdb/genomicsdb_iterators.h:133-139
    VariantQueryProcessor::iterate_over_cells   src/genomicsdb/query_variants.cc:557-576
    ColumnHistogramOperator                 src/query_operations/variant_operations.cc:732-790
    AlleleCountOperator                     src/query_operations/variant_operations.cc:900-1089
    VariantCallPrintCSVOperator / GenomicsDBColumnarCell::print_csv    src/genomicsdb/variant_cell.cc:167-185,
                                            src/genomicsdb/genomicsdb_columnar_field.cc:102-196
Pinned (tests/test_cells_oracle.py) on the unmodified reference's own --print-csv / --print-AC / --produce-histogram output
over the golden arrays (tests/golden/cells/, oracle/make_golden_cells.py), whose --print-calls twins equal the
reference CI's *_calls_* goldens.  Plain Python loops: small cases only.
"""
import heapq


def iterate(row, col, end, query_rows, col_lo=None, col_hi=None):
    """Cells in merged (column-major) order -> [(cell index, row, begin, end)] as the iterator hands them out for ONE
    query interval (col_lo None = full scan: simple traversal only), and the number of left (PQ) cells."""
    n = len(row)
    rows = set(int(r) for r in query_rows)
    out = []
    n_left = 0
    first = 0
    if col_lo is not None:
        while first < n and col[first] < col_lo:
            first += 1
        # find-intersecting-intervals mode: per queried row, its first cell at or after col_lo decides
        initialised = set()
        pq = []
        i = first
        while i < n and len(initialised) < len(rows):
            r = int(row[i])
            if r in rows and r not in initialised:
                initialised.add(r)
                c, e = int(col[i]), int(end[i])
                if c > e and e < col_lo:          # END copy of a call that began before the interval: swap back
                    heapq.heappush(pq, (e, r, i, c))
            i += 1
        while pq:
            b, r, i, e = heapq.heappop(pq)
            out.append((i, r, b, e))
        n_left = len(out)
    for i in range(first, n):
        c = int(col[i])
        if col_hi is not None and c > col_hi:
            break
        r = int(row[i])
        if r not in rows or c > int(end[i]):      # rows outside the query, END copies
            continue
        out.append((i, r, c, int(end[i])))
    return out, n_left


def histogram(cells, begin, end, bin_size):
    """-> {bin index: count} (sparse: the reference's tool allocates 40 M bins), number of bins"""
    n_bins = (end - begin) // bin_size + 1
    counts = {}
    for _, _, b, _ in cells:
        k = 0 if b <= begin else (n_bins - 1 if b >= end else (b - begin) // bin_size)
        counts[k] = counts.get(k, 0) + 1
    return counts, n_bins


def histogram_text(counts, n_bins, begin, bin_size, equi_bins=(128, 64, 32, 16, 8, 4, 2)):
    """produce_column_histogram + equi_partition_and_print_bins (gt_mpi_gather.cc:404-411, variant_operations.cc:767-790):
    bins are merged left to right until each group holds at least total / #bins cells"""
    out = []
    total = sum(counts.values())
    nz = sorted(k for k, v in counts.items() if v)
    for nb in equi_bins:
        if nb >= n_bins:
            continue
        per = total / nb
        out.append("Total %d #bins %d count/bins %.1f\n" % (total, nb, per))
        i = p = 0
        while i < n_bins:
            cur, j = 0, n_bins
            while p < len(nz):
                cur += counts[nz[p]]
                p += 1
                if cur >= per:
                    j = nz[p - 1] + 1
                    break
            out.append("%d,%d,%d\n" % (begin + i * bin_size, begin + j * bin_size - 1, cur))
            i = j
        out.append("\n")
    return "".join(out)


def _symbolic(a):
    return a in ("&", "*") or (len(a) > 0 and ((a[0] == "<" and a[-1] == ">") or "[" in a or "]" in a))


def normalize(ref, alt):
    """AlleleCountOperator::normalize_REF_ALT_pair"""
    if len(ref) > 1 and alt:
        if _symbolic(alt):
            return ref[:1], alt
        if len(alt) >= len(ref):
            suffix = len(ref) - 1
        elif len(alt) > 1:
            suffix = len(alt) - 1
        else:
            suffix = 0
        return ref[:len(ref) - suffix], alt[:len(alt) - suffix]
    return ref, alt


INT32_MISSING, INT32_VECTOR_END = -2147483648, -2147483647


def allele_counts(cells, ref, alt, gt, gt_step):
    """ref / alt / gt: per-cell accessors (cell index -> str / str / list of int) -> sorted [(column, REF, ALT, count)]"""
    m = {}
    for i, _, b, _ in cells:
        r, a, g = ref(i), alt(i), gt(i)
        if not r or not a or not len(g):
            continue
        toks = a.split("|")
        for k in range(0, len(g), gt_step):
            v = int(g[k])
            if v in (INT32_MISSING, INT32_VECTOR_END) or v <= 0:
                continue
            key = (b,) + normalize(r, toks[v - 1])
            m[key] = m.get(key, 0) + 1
    return [k + (m[k],) for k in sorted(m)]


def allele_counts_text(recs):
    return "".join("%d %s %s %d\n" % r for r in recs)
