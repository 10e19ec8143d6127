// CPU probe: IndexedFasta (the library's reader of the reference genome) at the "contig pos" lines of argv[2], printed the way
// `oracle/_ref/bin/gdb_ref_dump --mode fasta` prints the reference's ReferenceGenomeInfo.
#include <fstream>
#include <iostream>
#include "../../genomicsdb_b200/csrc/gdb_vcf_header.h"
using namespace genomicsdb_b200;
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    IndexedFasta fa;
    fa.open(argv[1]);
    std::ifstream l(argv[2]); std::string contig; long long pos;
    while (l >> contig >> pos) std::cout << contig << " " << pos << " " << fa.base_at(contig, pos) << "\n";
  } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
  return 0;
}
