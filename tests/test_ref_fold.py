"""merge_reference_allele's length register (variant_operations.cc:73-122): when an N-leading REF of several bases precedes the
first real REF of a site, the reference replaces the merged string but keeps the old length, ignores later REFs that are not longer
than that stale length and appends the tails of longer ones from the stale offset.  REFs like that do not occur in gVCFs; the
restatement (oracle/combine_oracle.c) and the kernels (k_union) reproduce the fold all the same, and both are held against the
UNMODIFIED reference reading the same on-disk array here."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO
from tools import gdbsyn, tiledb_writer

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")
OURS = os.path.join(REPO, "genomicsdb_b200", "bin", "gt_gpu_gather")

# (REF, ALT) of the eight samples' calls at one column, in row order; every ALT allele is as long as its REF (no deletions)
STALE = [(b"NNN", None), (b"CT", None), (b"CTG", None), (b"CTGA", None), (b"C", None), (b"CTGAC", None), (b"NNNNNN", None), (b"CT", None)]
CONTROL = [(b"N", None), (b"CT", None), (b"CTG", None), (b"CTGA", None), (b"C", None), (b"CTGAC", None), (b"NNNNNN", None), (b"CT", None)]


def _strings(col):
    off, val = col
    vb = val.view(np.uint8)
    ends = list(off[1:]) + [vb.nbytes]
    return [bytes(vb[int(off[i]):int(ends[i])]) for i in range(len(off))]


def _pack(strs):
    off = np.zeros(len(strs), dtype=np.uint64)
    pos = 0
    for i, s in enumerate(strs):
        off[i] = pos
        pos += len(s)
    return off, np.frombuffer(b"".join(strs), dtype=np.uint8).copy()


def crafted(plan, column=1100):
    d = gdbsyn.generate(n_samples=8, n_sites=8, spacing=100, p_var=0.9, a_pool=4, with_fixed=True)
    refs, alts = _strings(d["ref"]), _strings(d["alt"])
    rows = {int(d["coords"][i, 0]): i for i in range(len(refs)) if d["coords"][i, 1] == column and d["coords"][i, 1] <= d["end"][i]}
    assert sorted(rows) == list(range(8)), "the generator's cohort changed: every sample must have a call at the column"
    bases = b"GTAC"
    for r, (ref, _) in enumerate(plan):
        i = rows[r]
        n_alt = alts[i].count(b"|")          # tokens before the trailing <NON_REF>
        refs[i] = ref
        alts[i] = b"|".join([bytes([bases[(r + k) % 4]]) + b"A" * (len(ref) - 1) for k in range(n_alt)] + [b"&"])
    d["ref"], d["alt"] = _pack(refs), _pack(alts)
    return d


def _dump(cmd, out):
    p = subprocess.run(cmd + ["-o", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-600:]
    return open(out, "rb").read()


@pytest.fixture(scope="module")
def ref_present():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")


@pytest.mark.parametrize("plan", [STALE, CONTROL], ids=["stale-length", "control"])
def test_restatement_equals_the_reference(ref_present, tmp_path, plan):
    from oracle import oracle
    d = crafted(plan)
    info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d)
    q = tiledb_writer.write_query_json(str(tmp_path / "q.json"), info)
    want = _dump([REF, "-j", q, "--mode", "dump"], str(tmp_path / "ref.bin"))
    cols = oracle.load_columns(info["workspace"], "syn", info["vid"], ["END", "REF", "ALT", "GT", "AD", "PL"])
    got, n = oracle.combine_query(cols, [])
    assert got == want
    if plan is STALE:   # the fold really took the odd path: CT, then A from CTGA, C from CTGAC, N from NNNNNN
        assert b"CTACN" in want and b"CTGAC" not in want.replace(b"CTGACN", b"")


@pytest.mark.gpu
@pytest.mark.parametrize("plan", [STALE, CONTROL], ids=["stale-length", "control"])
def test_kernels_equal_the_reference(ref_present, tmp_path, plan):
    d = crafted(plan)
    info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d)
    q = tiledb_writer.write_query_json(str(tmp_path / "q.json"), info)
    want = _dump([REF, "-j", q, "--mode", "dump"], str(tmp_path / "ref.bin"))
    got = _dump([OURS, "-j", q], str(tmp_path / "ours.bin"))
    assert hashlib.md5(got).hexdigest() == hashlib.md5(want).hexdigest() and got == want


def test_restatement_equals_the_reference_on_a_cohort_with_deletions(ref_present, tmp_path):
    """the C restatement against the unmodified reference on a synthetic cohort whose variant calls are deletions to a third
    (tools/gdbsyn_deletions.py): spanning-deletion form, 1-bp stepping, reduced vectors"""
    from oracle import oracle
    from tools import gdbsyn_deletions
    d = gdbsyn_deletions.with_deletions(gdbsyn.generate(n_samples=60, n_sites=80, spacing=100, p_var=0.2, a_pool=5, with_fixed=True))
    info = tiledb_writer.write_array(str(tmp_path / "ws"), "syn", d)
    q = tiledb_writer.write_query_json(str(tmp_path / "q.json"), info)
    want = _dump([REF, "-j", q, "--mode", "dump"], str(tmp_path / "ref.bin"))
    got, n = oracle.combine_query(oracle.load_columns(info["workspace"], "syn", info["vid"], ["END", "REF", "ALT", "GT", "AD", "PL"]), [])
    assert n > 200 and got == want
