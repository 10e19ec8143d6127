// CPU probe: how the mirror's VidMapper reads multi-dimensional and tuple fields (no CUDA call is made).
#include <cstdio>
#include <iostream>
#include "../../genomicsdb_b200/csrc/gdb_host_api.h"
using namespace genomicsdb_b200;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  VidMapper vid(argv[1]);
  std::cout << "{";
  bool first = true;
  for (int a = 2; a < argc; ++a) {
    const FieldInfo* f = vid.get_field_info(argv[a]);
    if (!first) std::cout << ",";
    first = false;
    std::cout << "\"" << argv[a] << "\":";
    if (!f) { std::cout << "null"; continue; }
    std::cout << "{\"multid\":" << (f->is_multid() ? 1 : 0) << ",\"composite\":" << (f->is_composite() ? 1 : 0) << ",\"type\":\"" << f->type << "\",\"parent\":\"" << f->parent
              << "\",\"tuple_index\":" << f->tuple_index << ",\"vcf_name\":\"" << f->vcf_name() << "\",\"op\":\"" << f->combine_op << "\",\"dim0_allele\":" << (f->dim0_allele_dependent() ? 1 : 0)
              << ",\"dims\":[";
    for (size_t i = 0; i < f->dim_lengths.size(); ++i) std::cout << (i ? "," : "") << "\"" << f->dim_lengths[i] << "\"";
    std::cout << "],\"delims\":\"";
    for (char c : f->vcf_delimiters) std::cout << c;
    std::cout << "\",\"allele_dependent\":" << (f->is_allele_dependent() ? 1 : 0) << "}";
  }
  std::cout << "}\n";
  return 0;
}
