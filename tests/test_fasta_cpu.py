"""CPU: IndexedFasta -- the library's reader of the reference genome (the N -> reference base substitution of the Broad operator)
-- against the reference's ReferenceGenomeInfo over its own htslib faidx (oracle/_ref/bin/gdb_ref_dump --mode fasta;
vcf_adapter.cc:31-58): contigs with a description after the name, one-base and exact-multiple-of-the-line contigs, upper and
lower case, line boundaries, the reader's 32 KB window crossed forwards and backwards."""
import os
import random
import subprocess

import pytest

from conftest import REPO

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path_factory.mktemp("probe") / "fasta_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "fasta_probe.cc"), "-o", exe,
                           "-I" + os.path.join(REPO, "include"), "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


CONTIGS = [("chr1 a description after the name", 1000, 60), ("2", 75, 60), ("one_base", 1, 60), ("wide", 100000, 70), ("two_full_lines", 120, 60)]


def test_bases_as_the_reference_reads_them(probe, tmp_path):
    rng = random.Random(5)
    fa = tmp_path / "ref.fa"
    seqs = {}
    with open(fa, "w") as f:
        for name, n, width in CONTIGS:
            s = "".join(rng.choice("ACGTacgtNn") for _ in range(n))
            seqs[name.split()[0]] = s
            f.write(">" + name + "\n")
            for i in range(0, n, width):
                f.write(s[i:i + width] + "\n")
    queries = []
    for name, n, width in CONTIGS:
        key = name.split()[0]
        wanted = [0, 1, width - 1, width, width + 1, n - 1, n // 2, n - 2] + [rng.randrange(n) for _ in range(12)]
        queries += [(key, p) for p in wanted if 0 <= p < n]
    rng.shuffle(queries)
    queries += [("wide", p) for p in (0, 32759, 32760, 32761, 32767, 32768, 40000, 39999, 99999, 5)]   # around the reader's window
    pf = tmp_path / "positions.txt"
    pf.write_text("".join("%s %d\n" % q for q in queries))
    ref = subprocess.run([REF, "--mode", "fasta", "-j", str(fa), "-l", str(pf)], capture_output=True, text=True)   # builds ref.fa.fai
    assert ref.returncode == 0, ref.stderr[-400:]
    assert os.path.exists(str(fa) + ".fai")
    ours = subprocess.run([probe, str(fa), str(pf)], capture_output=True, text=True)
    assert ours.returncode == 0, ours.stderr[-400:]
    assert ours.stdout == ref.stdout
    got = [l.split() for l in ours.stdout.splitlines()]
    assert len(got) == len(queries)
    for (key, p), (c, q, base) in zip(queries, got):
        assert (c, int(q)) == (key, p) and base == seqs[key][p]   # bases as the file holds them, case kept
