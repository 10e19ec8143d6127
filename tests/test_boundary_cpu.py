"""CPU: the parts of the boundary that need no device -- header lines (htslib's bookkeeping restated), the indexed FASTA, the
loader configuration -- and a compile of the reference's own driver loop against the mirror's header."""
import os
import re
import subprocess
import sys

import pytest

from conftest import GOLDEN, REPO, golden_cases

CSRC = os.path.join(REPO, "genomicsdb_b200", "csrc")
PROBE = r'''
#include <fstream>
#include <iostream>
#include <sstream>
#include "gdb_host_api.h"
#include "gdb_vcf_header.h"
using namespace genomicsdb_b200;
int main(int argc, char** argv) {
  const std::string mode = argv[1];
  if (mode == "header") {   // template, vid, callsets -> the header a query of all rows with the template's own fields gets
    std::ifstream f(argv[2]); std::stringstream ss; ss << f.rdbuf();
    VCFHeaderLines h; h.parse(ss.str(), false);
    VidMapper vid(argv[3]); vid.read_callsets_file(argv[4]);
    for (unsigned i = 0; i < vid.get_num_contigs(); ++i) { const ContigInfo& c = vid.get_contig_info(i); if (h.name2id(c.name) < 0) h.append("##contig=<ID=" + c.name + ",length=" + std::to_string(c.length) + ">"); }
    std::string name;
    for (int64_t r = 0; r <= vid.get_max_callset_row_idx(); ++r) { vid.get_callset_name(r, name); h.add_sample(name); }
    VCFHeaderLines h2; h2.parse(h.format(false), false);
    std::cout << h2.format(true);
  } else if (mode == "fasta") {
    IndexedFasta fa; fa.open(argv[2]);
    std::ifstream t(argv[3]); std::string c, b; long long p; int bad = 0, n = 0;
    while (t >> c >> p >> b) { ++n; if (fa.base_at(c, p) != b[0]) ++bad; }
    std::cout << n << " " << bad << " " << fa.base_at("1", 5) << fa.base_at("1", 1000000000) << "\n";
  } else if (mode == "loader") {
    GenomicsDBImportConfig ld; ld.read_from_file(argv[2], 0);
    for (int r = 0; r < 3; ++r) { auto p = ld.get_column_partition(r); std::cout << p.first << " " << p.second << " " << ld.get_workspace(r) << " " << ld.get_array_name(r) << "\n"; }
    VariantQueryConfig q; q.update_from_loader(ld, 1);
    q.add_column_interval_to_query(5, 50); q.add_column_interval_to_query(150, 160); q.add_column_interval_to_query(99, 100); q.add_column_interval_to_query(1000, 2000);
    q.subset_query_column_ranges_based_on_partition(ld, 1);
    std::cout << q.get_workspace(1) << " " << q.get_array_name(1);
    for (unsigned i = 0; i < q.get_num_column_intervals(); ++i) std::cout << " [" << q.get_column_begin(i) << "," << q.get_column_end(i) << "]";
    std::cout << "\n";
  }
  return 0;
}
'''


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    d = tmp_path_factory.mktemp("probe")
    src = d / "probe.cc"
    src.write_text(PROBE)
    exe = d / "probe"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + CSRC, str(src), "-o", str(exe), "-L" + os.path.dirname(lib), "-lgdbgpu",
                           "-Wl,-rpath," + os.path.dirname(lib)])
    return str(exe)


def golden_header(cid):
    b = open(os.path.join(GOLDEN, "bcf", cid + ".bcf"), "rb").read()
    n = int.from_bytes(b[5:9], "little")
    return b[9:9 + n - 1].decode()


@pytest.mark.parametrize("cid,callsets", [("t0_1_2_at_0", "t0_1_2.json"), ("haploid_triploid_deletion", "t0_haploid_triploid_1_2_3_triploid_deletion.json"),
                                          ("min_PL_spanning_deletion_min_PL", "min_PL_spanning_deletion.json"), ("t0_overlapping_at_0", "t0_overlapping.json")])
def test_header_text_equals_the_reference(probe, cid, callsets):
    """template header + contigs of the vid mapping + samples, written without IDX, read back, written for BCF: the text the
    unmodified reference (htslib) put at the head of its BCF file"""
    out = subprocess.run([probe, "header", os.path.join(GOLDEN, "inputs", "template_vcf_header.vcf"), os.path.join(GOLDEN, "inputs", "vid.json"),
                          os.path.join(GOLDEN, "inputs", "callsets", callsets)], capture_output=True, text=True, check=True).stdout
    assert out == golden_header(cid)


def test_indexed_bgzip_fasta(probe):
    out = subprocess.run([probe, "fasta", os.path.join(GOLDEN, "inputs", "ref_sparse.fasta.gz"), os.path.join(GOLDEN, "inputs", "ref_bases.tsv")],
                         capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "25" and out[1] == "0" and out[2] == "NN"


def test_loader_configuration(probe, tmp_path):
    import json
    ld = {"column_partitions": [{"begin": 100, "workspace": "/w1", "array": "a1"}, {"begin": 0, "workspace": "/w0", "array": "a0"},
                                {"begin": 1000, "end": 5000, "workspace": "/w2", "array": "a2"}], "size_per_column_partition": 1}
    f = tmp_path / "loader.json"
    f.write_text(json.dumps(ld))
    out = subprocess.run([probe, "loader", str(f)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "100 999 /w1 a1"        # a partition ends where the next one (in column order) begins
    assert out[1] == "0 99 /w0 a0"
    assert out[2] == "1000 5000 /w2 a2"
    # rank 1 = partition [0, 99].  The subset edits the base configuration's list of ranges (what the query FILE named), not the
    # intervals already handed to the query: as in the reference, the scan keeps all of them (tests/test_query_config_cpu.py)
    assert out[3] == "/w0 a0 [5,50] [99,100] [150,160] [1000,2000]"


@pytest.mark.skipif(not os.path.exists("/root/reference/tools/src/gt_mpi_gather.cc"), reason="the reference's sources are not on this machine")
def test_reference_driver_loop_compiles_against_the_mirror(tmp_path):
    """the body of scan_and_produce_Broad_GVCF (tools/src/gt_mpi_gather.cc) is taken from the reference's file as it lies there and
    compiled, unchanged, against genomicsdb_b200/csrc/gdb_host_api.h: class names, constructor and method signatures,
    the overflow / scan-state protocol and the adapter types line up"""
    src = open("/root/reference/tools/src/gt_mpi_gather.cc").read()
    m = re.search(r"void scan_and_produce_Broad_GVCF\(.*?\n}\n", src, re.S)
    assert m
    tu = '#include <algorithm>\n#include <iostream>\n#include "gdb_host_api.h"\nusing namespace genomicsdb_b200;\n' + m.group(0)
    f = tmp_path / "driver.cc"
    f.write_text(tu)
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + CSRC, str(f)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
