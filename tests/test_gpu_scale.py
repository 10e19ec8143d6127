"""GPU: the CUDA path against the UNMODIFIED reference at BASELINE sizes.

Both sides are native programs reading the same on-disk TileDB fragment (written by tools/tiledb_writer.py from the
GDBSYN-1 generator) and writing the canonical hot-path dump (layout: oracle/ref_harness/gdb_ref_dump.cc):
    oracle/_ref/bin/gdb_ref_dump --mode dump     the reference's scan_and_operate + handle_deletions + GA4GHOperator::operate
    genomicsdb_b200/bin/gt_gpu_gather            the host mirror's scan_and_operate over libgdbgpu (kernels), same operator seam
The dumps run to gigabytes, so each program writes to a pipe and only the md5 of the stream is kept; the column range is
cut into intervals so that the reference's single-threaded processes use the box's cores side by side.  Every queried
row and every site of every interval is compared -- nothing is sampled.
"""
import hashlib
import json
import os
import subprocess
import threading

import pytest

from conftest import REPO
from tools import gdbsyn, tiledb_writer

pytestmark = pytest.mark.gpu

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")
OURS = os.path.join(REPO, "genomicsdb_b200", "bin", "gt_gpu_gather")


def _md5_of(cmd, out):
    """runs cmd + ["-o", <a pipe>] (both tools also print a JSON summary on stdout, so the dump gets a descriptor of its own),
    leaves (md5, bytes, returncode, stderr tail) in out"""
    r, w = os.pipe()
    p = subprocess.Popen(cmd + ["-o", "/dev/fd/%d" % w], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, pass_fds=(w,))
    os.close(w)
    err = []
    t = threading.Thread(target=lambda: err.append(p.stderr.read()))
    t.start()
    h, n = hashlib.md5(), 0
    with os.fdopen(r, "rb") as f:
        while True:
            b = f.read(1 << 22)
            if not b:
                break
            h.update(b)
            n += len(b)
    t.join()
    out.extend([h.hexdigest(), n, p.wait(), err[0].decode(errors="replace")[-600:]])


def _intervals(lo, hi, k):
    step = (hi - lo + k) // k
    return [(a, min(hi, a + step - 1)) for a in range(lo, hi + 1, step)]


def compare_with_reference(tmp_path, d, n_intervals, max_parallel=None, **writer_args):
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000, **writer_args)
    lo, hi = d["col_range"]
    queries = []
    for i, (a, b) in enumerate(_intervals(lo, hi, n_intervals)):
        queries.append((a, b, tiledb_writer.write_query_json(os.path.join(ws, "q%d.json" % i), info, a, b)))
    # the reference: one process per interval, as many at a time as there are cores
    ref = [[] for _ in queries]
    par = max_parallel or max(1, min(len(queries), os.cpu_count() or 1))
    for first in range(0, len(queries), par):
        ts = [threading.Thread(target=_md5_of, args=([REF, "-j", q, "--mode", "dump"], ref[i]))
              for i, (_, _, q) in enumerate(queries) if first <= i < first + par]
        [t.start() for t in ts]
        [t.join() for t in ts]
    # ours: one process per interval, one after the other (one GPU)
    total = 0
    for i, (a, b, q) in enumerate(queries):
        got = []
        _md5_of([OURS, "-j", q], got)
        assert ref[i][2] == 0, "reference failed on [%d, %d]: %s" % (a, b, ref[i][3])
        assert got[2] == 0, "gt_gpu_gather failed on [%d, %d]: %s" % (a, b, got[3])
        assert got[1] == ref[i][1] and got[0] == ref[i][0], "dump of columns [%d, %d] differs from the reference's (%d vs %d bytes)" % (a, b, got[1], ref[i][1])
        total += got[1]
    return total


@pytest.fixture(scope="module", autouse=True)
def tools_present():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built (it is built where /root/reference exists and travels with the repo)")
    assert os.path.exists(OURS), "gt_gpu_gather is not built: run __graft_entry__.build()"


def test_config2_all_rows_all_sites(tmp_path):
    """BASELINE config #2 whole: 1 k samples x 1 M positions (10 k candidate sites, ~20 k records), every row, every site"""
    d = gdbsyn.generate(**gdbsyn.CONFIGS["config2"])
    n = compare_with_reference(tmp_path, d, n_intervals=16)
    assert n > 1 << 30   # gigabytes of dump went through the comparison


def test_config3_column_slice(tmp_path):
    """BASELINE config #3, a 6 % column slice: 10 k samples x 600 candidate sites, every row"""
    d = gdbsyn.generate(**gdbsyn.CONFIGS["config3"], site_lo=4000, site_hi=4600)
    compare_with_reference(tmp_path, d, n_intervals=16)


def test_config4_width_all_rows(tmp_path):
    """BASELINE config #4's width: 100 k samples x 12 candidate sites (max 8 ALT alleles, 10 merged alleles), all 100 k rows"""
    d = gdbsyn.generate(**gdbsyn.CONFIGS["config4"], site_lo=700, site_hi=712)
    compare_with_reference(tmp_path, d, n_intervals=6)


@pytest.mark.parametrize("n_fragments,compress", [(8, False), (64, False), (5, True)], ids=["8-fragments", "64-fragments", "5-gzip-fragments"])
def test_many_fragments_newest_wins(tmp_path, n_fragments, compress):
    """an array imported in many batches (the production default): every fragment spans all columns and holds its batch's samples,
    and re-imports every 7th call of the batch before it with a changed PL -- the k-way merge (K1: merge path over the
    fragments' sorted runs) must interleave the runs in (column, row) order and let the newest copy win, as the reference's
    ArrayReadState does (array_read_state.cc:1836-2132, :2369-2402)"""
    d = gdbsyn.generate(n_samples=2000, n_sites=10000, spacing=100, p_var=0.1, a_pool=4, site_lo=100, site_hi=700, with_fixed=True)
    compare_with_reference(tmp_path, d, n_intervals=4, n_fragments=n_fragments, dup_every=7, compress=compress)


def test_spanning_deletions_at_cohort_scale(tmp_path):
    """2 000 samples, 300 candidate sites, a third of the variant calls turned into 2-4 base deletions (tools/gdbsyn_deletions.py):
    handle_deletions, the `*` allele, 1-bp stepping while a deletion is live and the reduced PL / AD / GT of every spanning call,
    every row and every site against the unmodified reference"""
    from tools import gdbsyn_deletions
    d = gdbsyn.generate(n_samples=2000, n_sites=10000, spacing=100, p_var=0.1, a_pool=6, site_lo=200, site_hi=500, with_fixed=True)
    dd = gdbsyn_deletions.with_deletions(d, share=0.3)
    assert dd["n_deletion_calls"] > 10000
    compare_with_reference(tmp_path, dd, n_intervals=8)


@pytest.mark.parametrize("deletions,writer_args", [(False, {}), (True, {}), (True, dict(compress=True, n_fragments=5, dup_every=9))],
                         ids=["plain", "deletions", "deletions-5-gzip-fragments"])
def test_broad_gvcf_file_of_a_synthetic_cohort(tmp_path, deletions, writer_args):
    """tool against tool on the same on-disk array: the unmodified `gt_mpi_gather --produce-Broad-GVCF -O bu` and
    `gt_gpu_gather --produce-Broad-GVCF` (header generated in-library, reference bases from the indexed FASTA) write the same
    BCF file, byte for byte -- 600 samples, 150 candidate sites, with GQ / MIN_DP / DP; once with a third of the variant calls
    turned into deletions (spanning `*` alleles, min-PL GT off), once more from five GZIP fragments with re-imported calls"""
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn.generate(n_samples=600, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=50, site_hi=200, with_fixed=True)
    if deletions:
        d = gdbsyn_deletions.with_deletions(d, share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000, **writer_args)
    lo, hi = d["col_range"]
    q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP", "DP"))
    ref_out, our_out = str(tmp_path / "ref.bcf"), str(tmp_path / "ours.bcf")
    with open(ref_out, "wb") as f:
        p = subprocess.run([ref_tool, "-j", q, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    p2 = subprocess.run([OURS, "-j", q, "--produce-Broad-GVCF", "-o", our_out], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = open(ref_out, "rb").read(), open(our_out, "rb").read()
    assert len(want) > 1 << 20
    assert len(got) == len(want) and got == want


QUERY_VARIANTS = {
    "produce_GT": {"produce_GT_field": True},
    "min_PL_GT": {"produce_GT_field": True, "produce_GT_with_min_PL_value_for_spanning_deletions": True},
    "sites_only": {"sites_only_query": True},
    "row_subset": {"query_row_ranges": [{"range_list": [{"low": 7, "high": 211}, {"low": 400, "high": 590}]}]},
    "two_intervals": "two_intervals",
    "max_alt_2": {"max_diploid_alt_alleles_that_can_be_genotyped": 2, "produce_GT_field": True},
    "inside_deletions": "inside_deletions",      # intervals that begin one and three bases into the deletions of a site
    "paged_small_windows": "paged",              # a 150 KB output buffer (overflow / resume all the way) over 3-site front windows
    "scan_full": "scan_full",                    # no query_column_ranges: the whole array
    "left_and_right_of_the_data": "outside",     # intervals that hold no call: header-only file
    "list_syntax_unsorted_overlapping": "lists",  # [[ [lo, hi], pos, ... ]] per-rank list syntax, unsorted, overlapping intervals
}


@pytest.mark.parametrize("variant", sorted(QUERY_VARIANTS))
def test_broad_gvcf_file_query_options_on_a_cohort_with_deletions(tmp_path, variant):
    """the options of the query JSON that change the records (GT with alleles, the min-PL genotype of spanning deletions, sites-only
    records, a row subset, several column intervals, PL dropped above 2 ALT alleles), each tool against tool on the 600-sample cohort
    with deletions: same BCF file"""
    import json
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=600, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=50, site_hi=170, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP"))
    q = json.load(open(qf))
    opt = QUERY_VARIANTS[variant]
    extra, env = [], None
    if opt == "two_intervals":
        mid = (lo + hi) // 2
        q["query_column_ranges"] = [{"range_list": [{"low": lo + 950, "high": mid - 333}, {"low": mid + 1, "high": hi - 1200}]}]
    elif opt == "inside_deletions":
        c = 1000 + 100 * 90           # the column of candidate site 90 (first_col + site * spacing)
        q["query_column_ranges"] = [{"range_list": [{"low": c + 1, "high": c + 2}, {"low": c + 3, "high": c + 450}, {"low": c + 501, "high": c + 501}]}]
        q["produce_GT_field"] = True
    elif opt == "scan_full":
        del q["query_column_ranges"]
        q["scan_full"] = True
    elif opt == "lists":
        c = 1000 + 100 * 100
        q["query_column_ranges"] = [[[c + 700, c + 1500], c - 350, [c - 900, c + 900], [c + 100, c + 100]]]
        q["query_row_ranges"] = [[[100, 300], 7, [450, 599]]]
    elif opt == "outside":
        q["query_column_ranges"] = [{"range_list": [{"low": 10, "high": 500}, {"low": hi + 5000, "high": hi + 9000}]}]
    elif opt == "paged":
        extra, env = ["-p", "150000"], dict(os.environ, GDBGPU_FRONT_BYTES=str(24 * 600 * 3))
    else:
        q.update(opt)
    json.dump(q, open(qf, "w"))
    ref_out, our_out = str(tmp_path / "ref.bcf"), str(tmp_path / "ours.bcf")
    with open(ref_out, "wb") as f:
        p = subprocess.run([ref_tool, "-j", qf, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    p2 = subprocess.run([OURS, "-j", qf, "--produce-Broad-GVCF", "-o", our_out] + extra, capture_output=True, text=True, env=env)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = open(ref_out, "rb").read(), open(our_out, "rb").read()
    assert len(want) > (1000 if opt == "outside" else 10000)
    assert len(got) == len(want) and got == want


@pytest.mark.parametrize("produce_gt", [False, True], ids=["no-GT", "GT"])
def test_jni_conventions_on_a_cohort_with_deletions(tmp_path, produce_gt):
    """the records the JNI path hands to the Java reader (BroadCombinedGVCFOperator constructed with
    use_missing_values_only_not_vector_end: missing instead of vector_end padding, no-call GT for rows without a call): the
    unmodified reference classes driven by oracle/ref_harness (gdb_ref_dump --mode bcf-java) against
    gt_gpu_gather --produce-Broad-GVCF --htsjdk-conventions on the 600-sample cohort with deletions"""
    import json
    from tools import gdbsyn_deletions
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=600, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=50, site_hi=170, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP"))
    if produce_gt:
        q = json.load(open(qf))
        q["produce_GT_field"] = True
        json.dump(q, open(qf, "w"))
    p = subprocess.run([REF, "-j", qf, "--mode", "bcf-java"], capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")[-800:]
    our_out = str(tmp_path / "ours.bcf")
    p2 = subprocess.run([OURS, "-j", qf, "--produce-Broad-GVCF", "--htsjdk-conventions", "-o", our_out], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = p.stdout, open(our_out, "rb").read()
    assert len(want) > 100000
    assert len(got) == len(want) and got == want


@pytest.mark.parametrize("flag", ["--range-query", "--print-calls", "--print-csv", "--print-AC", "--produce-histogram", "--produce-interesting-positions"])
def test_tool_commands_on_a_cohort_with_deletions(tmp_path, flag):
    """the reference tool's other commands on a 250-sample cohort with deletions, stdout against stdout: the GA4GH range query (its
    default command; here hundreds of variants hold several calls whose ALT lists name the same alleles in different orders, so
    the re-indexing of merged variants is not the identity), the four columnar-iterator commands, the interesting positions"""
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=250, n_sites=10000, spacing=100, p_var=0.15, a_pool=4, site_lo=20, site_hi=60, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo + 700, hi - 900, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP", "DP"))
    ref_cmd = ["stdbuf", "-o0", ref_tool, "-j", qf] + ([] if flag == "--range-query" else [flag])   # (this build of the tool leaves without flushing a block-buffered stdout)
    p = subprocess.run(ref_cmd, capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")[-800:]
    out = str(tmp_path / "ours.txt")
    p2 = subprocess.run([OURS, "-j", qf, "-o", out, flag], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = p.stdout, open(out, "rb").read()
    assert len(want) > 1000
    assert got == want


@pytest.mark.parametrize("shape", [dict(n_samples=300, a_pool=64, sites=(100, 160)), dict(n_samples=5000, a_pool=8, sites=(300, 318))],
                         ids=["64-allele-pool", "5000-samples"])
def test_broad_gvcf_file_wide_sites_and_wide_cohorts(tmp_path, shape):
    """whole BCF files tool against tool where (i) sites carry 45-60 merged alleles (PL dropped above 50 ALT alleles, the generic
    remap kernel) and (ii) the cohort is wide enough (>= 4096 samples) for the signature-row remap path, both with deletions"""
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn.generate(n_samples=shape["n_samples"], n_sites=10000, spacing=100, p_var=0.15, a_pool=shape["a_pool"],
                        site_lo=shape["sites"][0], site_hi=shape["sites"][1], with_fixed=True)
    d = gdbsyn_deletions.with_deletions(d, share=0.25)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP"))
    ref_out, our_out = str(tmp_path / "ref.bcf"), str(tmp_path / "ours.bcf")
    with open(ref_out, "wb") as f:
        p = subprocess.run([ref_tool, "-j", q, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    p2 = subprocess.run([OURS, "-j", q, "--produce-Broad-GVCF", "-o", our_out], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = open(ref_out, "rb").read(), open(our_out, "rb").read()
    assert len(want) > 100000
    assert len(got) == len(want) and got == want


ATTRIBUTE_SETS = {
    "PL_only": ("REF", "ALT", "PL"),
    "AD_GQ": ("REF", "ALT", "AD", "GQ"),
    "GT_only": ("GT",),
    "no_alleles_listed": ("GT", "AD", "PL", "MIN_DP"),
    "MIN_DP_GQ_only": ("GQ", "MIN_DP"),
    "PL_before_GT": ("PL", "AD", "GT", "REF", "ALT"),
}


@pytest.mark.parametrize("which", sorted(ATTRIBUTE_SETS))
def test_broad_gvcf_file_attribute_sets(tmp_path, which):
    """what the query lists decides which kernels are instantiated (with / without GT, one / two remapped fields) and which
    attributes the bookkeeping adds on its own (END, REF / ALT for allele-dependent fields, GT for genotype-dependent ones) and in
    which order the record's fields come: a few attribute lists, tool against tool on the cohort with deletions"""
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=400, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=50, site_hi=130, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=ATTRIBUTE_SETS[which])
    ref_out, our_out = str(tmp_path / "ref.bcf"), str(tmp_path / "ours.bcf")
    with open(ref_out, "wb") as f:
        p = subprocess.run([ref_tool, "-j", q, "--produce-Broad-GVCF", "-O", "bu"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    p2 = subprocess.run([OURS, "-j", q, "--produce-Broad-GVCF", "-o", our_out], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    want, got = open(ref_out, "rb").read(), open(our_out, "rb").read()
    assert len(want) > 10000
    assert len(got) == len(want) and got == want


@pytest.mark.parametrize("flag", ["--range-query", "--print-calls", "--print-AC", "--produce-interesting-positions"])
def test_tool_commands_row_subset_and_intervals(tmp_path, flag):
    """the same commands with a row subset and three query intervals (one of them a single position inside a deletion)"""
    import json
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=250, n_sites=10000, spacing=100, p_var=0.15, a_pool=4, site_lo=20, site_hi=60, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ"))
    q = json.load(open(qf))
    c = 1000 + 100 * 40
    q["query_column_ranges"] = [{"range_list": [{"low": lo + 333, "high": c - 77}, {"low": c + 1, "high": c + 1}, {"low": c + 2, "high": hi - 500}]}]
    q["query_row_ranges"] = [{"range_list": [{"low": 3, "high": 101}, {"low": 200, "high": 249}]}]
    json.dump(q, open(qf, "w"))
    p = subprocess.run(["stdbuf", "-o0", ref_tool, "-j", qf] + ([] if flag == "--range-query" else [flag]), capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")[-800:]
    out = str(tmp_path / "ours.txt")
    p2 = subprocess.run([OURS, "-j", qf, "-o", out, flag], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    assert len(p.stdout) > 1000 and open(out, "rb").read() == p.stdout


def test_jni_stream_of_a_cohort_with_deletions(tmp_path):
    """GenomicsDBBCFGenerator of the mirror pulled in 64 KB reads with a 1 MB buffer (what the JNI methods sit on) over the
    600-sample cohort with deletions: the stream is the reference's BCF under the JNI conventions, header first"""
    from tools import gdbsyn_deletions
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=600, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=50, site_hi=170, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, lo, hi, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ", "MIN_DP"))
    p = subprocess.run([REF, "-j", qf, "--mode", "bcf-java"], capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")[-800:]
    out = str(tmp_path / "ours.bcf")
    p2 = subprocess.run([OURS, "-j", qf, "--bcf-generator", "65536", "-p", "1000000", "--htsjdk-conventions", "-o", out], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    got = open(out, "rb").read()
    assert len(got) == len(p.stdout) and got == p.stdout


@pytest.mark.parametrize("scan_full", [False, True], ids=["interval", "scan_full"])
def test_hot_path_dump_row_subset_on_a_cohort_with_deletions(tmp_path, scan_full):
    """the hot-path dump (scan_and_operate + handle_deletions + GA4GHOperator::operate on both sides) for a subset of the rows:
    calls of rows outside the query still cut intervals (query_variants.cc:489-505), their deletions do not step.  With
    "scan_full" in the JSON the reference reads neither column nor row ranges (json_config.cc:253-260): whole array, every row"""
    import json
    from tools import gdbsyn_deletions
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=500, n_sites=10000, spacing=100, p_var=0.12, a_pool=6, site_lo=300, site_hi=420, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    lo, hi = d["col_range"]
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, None if scan_full else lo + 1234, None if scan_full else hi - 777)
    q = json.load(open(qf))
    q["query_row_ranges"] = [{"range_list": [{"low": 11, "high": 60}, {"low": 123, "high": 123}, {"low": 300, "high": 480}]}]
    json.dump(q, open(qf, "w"))
    ref, got = [], []
    _md5_of([REF, "-j", qf, "--mode", "dump"], ref)
    _md5_of([OURS, "-j", qf], got)
    assert ref[2] == 0 and got[2] == 0, (ref[3], got[3])
    assert ref[1] > 100000 and (got[0], got[1]) == (ref[0], ref[1])


@pytest.mark.parametrize("flag", ["--print-calls", "--print-AC", "--range-query"])
def test_tool_commands_scan_full(tmp_path, flag):
    """the columnar commands over the whole array ("scan_full" in the query JSON)"""
    from tools import gdbsyn_deletions
    ref_tool = os.path.join(REPO, "oracle", "_ref", "bin", "gt_mpi_gather")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref is not built")
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=120, n_sites=10000, spacing=100, p_var=0.15, a_pool=4, site_lo=20, site_hi=50, with_fixed=True), share=0.3)
    ws = str(tmp_path / "ws")
    info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
    qf = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, None, None, attributes=("REF", "ALT", "GT", "AD", "PL", "GQ"))
    p = subprocess.run(["stdbuf", "-o0", ref_tool, "-j", qf] + ([] if flag == "--range-query" else [flag]), capture_output=True)
    assert p.returncode == 0, p.stderr.decode(errors="replace")[-800:]
    out = str(tmp_path / "ours.txt")
    p2 = subprocess.run([OURS, "-j", qf, "-o", out, flag], capture_output=True, text=True)
    assert p2.returncode == 0, p2.stderr[-800:]
    # the range query walks the query's column intervals, and "scan_full" leaves none: an empty variant list
    assert len(p.stdout) > (0 if flag == "--range-query" else 1000) and open(out, "rb").read() == p.stdout
