// CPU probe: the vid mapping (argv[1]; argv[2], optional, a callset mapping file) as the mirror's VidMapper holds it, printed
// the way `oracle/_ref/bin/gdb_ref_dump --mode vid` prints the reference's FileBasedVidMapper.  A refused file exits 1.
#include <iostream>
#include "../../genomicsdb_b200/csrc/gdb_host_api.h"
using namespace genomicsdb_b200;
static const char* len_name(const std::string& l) { return l == "VAR" ? "var" : l.c_str(); }
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    VidMapper vid(argv[1]);
    if (argc > 2) vid.read_callsets_file(argv[2]);
    std::cout << "{\"fields\": [";
    bool first = true;
    for (auto& f : vid.fields()) {
      std::string vcf_name = f.vcf_name();
      if (f.is_format_twin && vcf_name.size() > 7 && vcf_name.compare(vcf_name.size() - 7, 7, "_FORMAT") == 0) vcf_name.resize(vcf_name.size() - 7);
      const bool both = f.is_info && f.is_format;
      std::cout << (first ? "" : ", ") << "{\"name\": \"" << f.name << "\", \"idx\": " << f.global_idx << ", \"vcf_name\": \"" << vcf_name << "\", \"class\": \""
                << ((f.is_info && !(both && f.is_format_twin)) ? "I" : "") << ((f.is_format && !(both && !f.is_format_twin)) ? "F" : "") << (f.is_filter ? "T" : "") << "\", \"type\": [";
      first = false;
      if (f.is_composite()) for (size_t j = 0; j < f.tuple_types.size(); ++j) std::cout << (j ? ", " : "") << "\"" << f.tuple_types[j] << "\"";
      else std::cout << "\"" << f.type << "\"";
      std::cout << "], \"length\": [";
      if (!f.dim_lengths.empty()) for (size_t d = 0; d < f.dim_lengths.size(); ++d) std::cout << (d ? ", " : "") << "\"" << len_name(f.dim_lengths[d]) << "\"";
      else std::cout << "\"" << len_name(f.length) << "\"";
      static const char* ops[] = {"sum", "mean", "median", "DP", "move_to_FORMAT", "element_wise_sum", "concatenate", "histogram_sum"};
      int op = 8;
      for (int i = 0; i < 8; ++i) if (f.combine_op == ops[i]) op = i;
      if (both && f.is_format_twin) op = 8;   // the reference keeps the operation on the INFO side only
      std::cout << "], \"op\": " << op << ", \"flattened\": " << (f.parent.empty() ? 0 : 1);
      if (f.dim_lengths.size() > 1) { std::cout << ", \"delims\": \""; for (size_t d = 0; d < f.dim_lengths.size() && d < f.vcf_delimiters.size(); ++d) std::cout << f.vcf_delimiters[d]; std::cout << "\""; }
      std::cout << "}";
    }
    std::cout << "], \"num_fields\": " << vid.get_num_fields() << ", \"contigs\": [";
    for (unsigned i = 0; i < vid.get_num_contigs(); ++i) { auto& c = vid.get_contig_info(i); std::cout << (i ? ", " : "") << "[\"" << c.name << "\", " << c.tiledb_column_offset << ", " << c.length << "]"; }
    std::cout << "], \"callsets\": [";
    if (vid.is_callset_mapping_initialized())
      for (int64_t r = 0; r <= vid.get_max_callset_row_idx(); ++r) { std::string n; vid.get_callset_name(r, n); std::cout << (r ? ", " : "") << "\"" << n << "\""; }
    std::cout << "], \"locate\": [";
    first = true;
    for (unsigned i = 0; i < vid.get_num_contigs(); ++i) {
      auto& c = vid.get_contig_info(i);
      for (int64_t col : {c.tiledb_column_offset - 1, c.tiledb_column_offset, c.tiledb_column_offset + c.length - 1, c.tiledb_column_offset + c.length}) {
        const ContigInfo* at = col >= 0 ? vid.get_contig_location(col) : nullptr;
        std::cout << (first ? "" : ", ") << "[" << col << ", \"" << (at ? at->name : "") << "\", " << (at ? col - at->tiledb_column_offset : -1) << "]";
        first = false;
      }
    }
    std::cout << "]}" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
