// CPU probe: VCFHeaderLines (the library's restatement of htslib's header bookkeeping) on a header text (argv[1]; empty file:
// what bcf_hdr_init("w") starts with) plus lines to append (argv[2], optional; "SAMPLE <name>" adds a sample), then the clean-up
// the Broad operator does (format without IDX, parse again, format as BCF).  Prints what
// `oracle/_ref/bin/gdb_ref_dump --mode header` prints for the reference's own htslib.
#include <fstream>
#include <iostream>
#include <sstream>
#include "../../genomicsdb_b200/csrc/gdb_host_api.h"
#include "../../genomicsdb_b200/csrc/gdb_vcf_header.h"
using namespace genomicsdb_b200;
namespace genomicsdb_b200 { bool add_field_to_hdr_if_missing(VCFHeaderLines& hdr, const VidMapper& id_mapper, const std::string& field_name, const std::string& vcf_name, int field_type_idx); }
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
  std::ifstream f(argv[1]); std::stringstream ss; ss << f.rdbuf();
  const std::string text = ss.str();
  VCFHeaderLines h(text.empty());
  if (!text.empty() && !h.parse(text, false)) { std::cerr << "parse failed\n"; return 1; }
  if (argc > 2) {
    std::ifstream l(argv[2]); std::string line;
    while (std::getline(l, line)) {
      if (line.compare(0, 6, "FIELD ") == 0 && line.size() > 8) {   // add_field_to_hdr_if_missing with the vid mapping argv[3]
        static VidMapper* vid = nullptr;
        if (!vid) vid = new VidMapper(argv[3]);
        const int type = line[6] == 'I' ? HL_INFO : line[6] == 'F' ? HL_FMT : HL_FLT;
        const std::string name = line.substr(8);
        try { std::cout << "field " << (add_field_to_hdr_if_missing(h, *vid, name, name, type) ? 1 : 0) << "\n"; }
        catch (const std::exception&) { std::cout << "field refused\n"; }
      } else if (line.compare(0, 7, "SAMPLE ") == 0) std::cout << "add_sample " << (h.add_sample(line.substr(7)) ? 0 : -1) << "\n";
      else std::cout << "append " << (h.append(line) ? 0 : -1) << "\n";
    }
  }
  VCFHeaderLines h2(false);
  if (!h2.parse(h.format(false), false)) { std::cerr << "second parse failed\n"; return 1; }
  std::cout << h2.format(true);
  for (auto& kv : h2.ids()) std::cout << "ID " << kv.first << " " << kv.second << "\n";
  for (auto& kv : h2.contigs()) std::cout << "CTG " << kv.first << " " << kv.second << "\n";
  std::cout << "nsamples " << h2.nsamples() << "\n";
  } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
  return 0;
}
