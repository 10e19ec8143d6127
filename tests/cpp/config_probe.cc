// CPU probe: the query configuration as the mirror holds it after the calls the reference's tool makes (gt_mpi_gather.cc:547-557:
// the loader's file first, then the query's, then the partition subset).  Prints what `oracle/_ref/bin/gdb_ref_dump --mode config`
// prints for the reference's own classes; a refused configuration exits 1 with the reason on stderr.  No CUDA call is made.
#include <cstdio>
#include <cstring>
#include <iostream>
#include "../../genomicsdb_b200/csrc/gdb_host_api.h"
using namespace genomicsdb_b200;
int main(int argc, char** argv) {
  std::string loader, query;
  int rank = 0;
  for (int a = 1; a + 1 < argc; a += 2) {
    if (!strcmp(argv[a], "-l")) loader = argv[a + 1];
    else if (!strcmp(argv[a], "-j")) query = argv[a + 1];
    else if (!strcmp(argv[a], "-r")) rank = atoi(argv[a + 1]);
  }
  if (query.empty()) return 2;
  try {
    VariantQueryConfig qc;
    GenomicsDBImportConfig lc;
    if (!loader.empty()) { lc.read_from_file(loader, rank); qc.update_from_loader(lc, rank); }
    qc.read_from_file(query, rank);
    if (!loader.empty()) qc.subset_query_column_ranges_based_on_partition(lc, rank);
    std::cout << "{\"workspace\": \"" << qc.get_workspace(rank) << "\", \"array\": \"" << qc.get_array_name(rank) << "\", \"intervals\": [";
    for (unsigned i = 0; i < qc.get_num_column_intervals(); ++i) std::cout << (i ? ", " : "") << "[" << (long long)qc.get_column_begin(i) << ", " << (long long)qc.get_column_end(i) << "]";
    std::cout << "], \"base_ranges\": [";
    auto& base = qc.get_query_column_ranges(rank);
    for (size_t i = 0; i < base.size(); ++i) std::cout << (i ? ", " : "") << "[" << (long long)base[i].first << ", " << (long long)base[i].second << "]";
    std::cout << "], \"all_rows\": " << (qc.query_all_rows() ? "true" : "false") << ", \"rows\": [";
    auto& rows = qc.get_rows_to_query();
    for (size_t i = 0; i < rows.size(); ++i) std::cout << (i ? ", " : "") << (long long)rows[i];
    std::cout << "], \"attributes\": [";
    auto& at = qc.get_query_attributes();
    for (size_t i = 0; i < at.size(); ++i) std::cout << (i ? ", " : "") << "\"" << at[i] << "\"";
    std::cout << "], \"segment_size\": " << qc.get_segment_size()
              << ", \"max_diploid_alt_alleles_that_can_be_genotyped\": " << qc.get_max_diploid_alt_alleles_that_can_be_genotyped()
              << ", \"produce_GT_field\": " << qc.produce_GT_field() << ", \"produce_FILTER_field\": " << qc.produce_FILTER_field()
              << ", \"sites_only_query\": " << qc.sites_only_query()
              << ", \"produce_GT_with_min_PL_value_for_spanning_deletions\": " << qc.produce_GT_with_min_PL_value_for_spanning_deletions()
              << ", \"vcf_header_filename\": \"" << qc.get_vcf_header_filename() << "\", \"reference_genome\": \"" << qc.get_reference_genome()
              << "\", \"loader_partition\": [" << (loader.empty() ? 0ll : (long long)lc.get_column_partition(rank).first) << ", " << (loader.empty() ? 0ll : (long long)lc.get_column_partition(rank).second) << "]"
              << ", \"vcf_output_filename\": \"" << qc.get_vcf_output_filename() << "\"}" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
