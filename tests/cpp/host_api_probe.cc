// CPU-only probe of the host mirror's parsing logic (no CUDA call is made): prints what VCFAdapter, VidMapper and
// VariantQueryConfig derive from a BCF header / vid mapping / query JSON as one JSON line, so that a python test can check it
// against the same facts read with the test-side readers.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../genomicsdb_b200/csrc/gdb_host_api.h"

using namespace genomicsdb_b200;

int main(int argc, char** argv) {
  if (argc < 4) { std::cerr << "usage: host_api_probe file.bcf vid.json query.json\n"; return 2; }
  try {
    std::ifstream f(argv[1], std::ios::binary);
    std::stringstream ss; ss << f.rdbuf();
    const std::string all = ss.str();
    uint32_t l_text; memcpy(&l_text, all.data() + 5, 4);
    VCFAdapter a;
    a.set_vcf_header_text(std::string(all.data() + 9, strnlen(all.data() + 9, l_text)));
    VidMapper vid(argv[2]);
    VariantQueryConfig qc;
    qc.read_from_file(argv[3]);
    printf("{\"ids\": {");
    const char* names[] = {"PASS", "LowQual", "END", "DP", "GT", "GQ", "AD", "PL", "MIN_DP", "SB", "PGT", "PID", "BaseQRankSum", "MQ", "RAW_MQ", "no_such_field"};
    bool first = true;
    for (const char* n : names) { printf("%s\"%s\": %d", first ? "" : ", ", n, a.get_dictionary_id(n)); first = false; }
    printf("}, \"contig_1\": %d, \"contig_2\": %d", a.get_contig_id("1"), a.get_contig_id("2"));
    const ContigInfo* c = vid.get_contig_location(249250621 + 5);
    printf(", \"contig_of_col\": \"%s\", \"contig_offset\": %lld", c ? c->name.c_str() : "", c ? (long long)c->tiledb_column_offset : -1LL);
    const FieldInfo* pl = vid.get_field_info("PL"); const FieldInfo* mq = vid.get_field_info("MQ"); const FieldInfo* dpf = vid.get_field_info("DP_FORMAT");
    printf(", \"PL_length\": \"%s\", \"MQ_op\": \"%s\", \"DP_FORMAT_is_format\": %d", pl ? pl->length.c_str() : "", mq ? mq->combine_op.c_str() : "", dpf && dpf->is_format ? 1 : 0);
    printf(", \"intervals\": [");
    for (unsigned i = 0; i < qc.get_num_column_intervals(); ++i) printf("%s[%lld, %lld]", i ? ", " : "", (long long)qc.get_column_begin(i), (long long)qc.get_column_end(i));
    printf("], \"n_rows\": %zu, \"produce_GT\": %d, \"n_attributes\": %zu}\n", qc.get_rows_to_query().size(), qc.produce_GT_field() ? 1 : 0, qc.get_query_attributes().size());
  } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
  return 0;
}
