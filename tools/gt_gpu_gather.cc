// gt_gpu_gather -- the reference's query driver for the combine path (tools/src/gt_mpi_gather.cc:322-366,413-614),
// restated over the host mirror (genomicsdb_b200/csrc/gdb_host_api.h).  Same JSON query file, same rank convention
// (-r, no MPI needed: the combine path has no collectives), same scan_state / overflow loop.  Output: the canonical
// hot-path dump (layout in oracle/ref_harness/gdb_ref_dump.cc) written by an operator plugged into
// SingleVariantOperatorBase::operate -- the same seam the reference harness uses, so the two programs read alike.
// With --produce-Broad-GVCF the operator is BroadCombinedGVCFOperator (as in scan_and_produce_Broad_GVCF) and the output is
// an uncompressed BCF2 file: the header is taken verbatim from --bcf-header (a BCF written by the reference for the same
// query: producing the header is htslib's job on the unchanged host side), the records come from the kernels' FORMAT
// blocks / INFO scalars through the host mirror.
#include <fcntl.h>
#include <getopt.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../genomicsdb_b200/csrc/gdb_host_api.h"

using namespace genomicsdb_b200;

class HotPathDumpOperator : public GA4GHOperator {
 public:
  HotPathDumpOperator(const VariantQueryConfig& qc, FILE* out, size_t overflow_every) : m_out(out), m_overflow_every(overflow_every) {
    put<uint32_t>(0x31444247u);
    put<uint32_t>(qc.get_num_remapped_fields());
    for (unsigned i = 0; i < qc.get_num_remapped_fields(); ++i) put_string(qc.get_remapped_field_name(i));
    m_num_fields = qc.get_num_remapped_fields();
  }
  void operate(Variant& variant, const VariantQueryConfig& query_config) override {
    GA4GHOperator::operate(variant, query_config);
    ++m_num_sites; ++m_since_drain;
    put<int64_t>(variant.get_column_begin());
    put<int64_t>(variant.get_column_end());
    put<uint8_t>(m_NON_REF_exists ? 1 : 0);
    put<uint8_t>(m_remapping_needed ? 1 : 0);
    put<uint16_t>(m_merged_alt_alleles.size());
    put_string(m_merged_reference_allele);
    for (const auto& alt : m_merged_alt_alleles) put_string(alt);
    const Variant& src = get_remapped_variant();
    uint32_t L = 0;
    for (auto it = src.begin(); it != src.end(); ++it) ++L;
    put<uint32_t>(L);
    for (auto it = src.begin(); it != src.end(); ++it) {
      const VariantCall& call = *it;
      put<int64_t>(call.get_row_idx());
      put<int64_t>(call.get_column_begin());
      put<int64_t>(call.get_column_end());
      put<uint16_t>(call.get_num_input_alleles());
      for (unsigned i = 0; i < call.get_num_input_alleles(); ++i) put<int32_t>(call.get_alleles_LUT()[i]);
      for (unsigned f = 0; f < m_num_fields; ++f) {
        put<uint32_t>(call.get_field_length(f));
        if (call.get_field_length(f)) fwrite(call.get_field_words(f), 4u, call.get_field_length(f), m_out);
      }
    }
  }
  // emulates a consumer with a bounded buffer (VCFSerializedBufferAdapter::overflow, vcf_adapter.h:190-195)
  bool overflow() const override { return m_overflow_every && m_since_drain >= m_overflow_every; }
  void drain() { m_since_drain = 0; }
  uint64_t num_sites() const { return m_num_sites; }
 private:
  template <class T> void put(T v) { fwrite(&v, sizeof(T), 1u, m_out); }
  void put_string(const std::string& s) { put<uint32_t>(s.length()); if (s.length()) fwrite(s.data(), 1u, s.length(), m_out); }
  FILE* m_out; unsigned m_num_fields = 0; uint64_t m_num_sites = 0; size_t m_overflow_every, m_since_drain = 0;
};

// header text of a BCF2 file ("BCF\2\2", u32 l_text, text) or of a plain VCF header file
static std::string read_header_text(const std::string& path, std::string& raw_prefix) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  char magic[9] = {0};
  f.read(magic, 9);
  if (f.gcount() == 9 && memcmp(magic, "BCF\2\2", 5) == 0) {   // only the header is read: the file may hold gigabytes of records
    uint32_t l_text; memcpy(&l_text, magic + 5, 4);
    raw_prefix.assign(magic, 9);
    raw_prefix.resize(9 + (size_t)l_text);
    f.read(&raw_prefix[9], l_text);
    if ((size_t)f.gcount() != l_text) throw std::runtime_error("truncated BCF header in " + path);
    return std::string(raw_prefix.data() + 9, strnlen(raw_prefix.data() + 9, l_text));
  }
  f.clear(); f.seekg(0);
  std::stringstream ss; ss << f.rdbuf();
  const std::string all = ss.str();
  uint32_t l_text = (uint32_t)all.size() + 1;
  raw_prefix = std::string("BCF\2\2", 5) + std::string((const char*)&l_text, 4) + all + std::string(1, '\0');
  return all;
}

static int produce_broad_gvcf(const std::string& query_json, const std::string& loader_json, int rank, int device, const std::string& header_file, const std::string& ref_bases_file,
                              const std::string& out_file, size_t buffer_capacity, bool htsjdk_conventions) {
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const auto t0 = now();
  VariantQueryConfig query_config;
  GenomicsDBImportConfig loader_config;   // tools/src/gt_mpi_gather.cc:560-575
  if (!loader_json.empty()) { loader_config.read_from_file(loader_json, rank); query_config.update_from_loader(loader_config, rank); }
  query_config.read_from_file(query_json, rank);
  if (!loader_json.empty()) query_config.subset_query_column_ranges_based_on_partition(loader_config, rank);
  VariantStorageManager sm(query_config.get_workspace(rank), 10u * 1024u * 1024u, device);
  const auto t1 = now();   // CUDA context is up
  VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
  qp.do_query_bookkeeping(query_config, true, /*stage_broad_gvcf_fields=*/true);
  const auto t2 = now();   // fragment files are staged in HBM
  VCFAdapter vcf_adapter;
  std::string raw_prefix;
  // the header: generated from the configuration (template header + queried fields + contigs + samples), as the reference does
  // (VCFAdapter::initialize); or, with --bcf-header, taken verbatim from a BCF / VCF file
  if (header_file.empty()) vcf_adapter.initialize(query_config);
  else vcf_adapter.set_vcf_header_text(read_header_text(header_file, raw_prefix));
  vcf_adapter.set_buffer_capacity(buffer_capacity);
  if (!ref_bases_file.empty()) {
    std::ifstream rb(ref_bases_file);
    std::string contig, base; long long pos;
    while (rb >> contig >> pos >> base) vcf_adapter.set_reference_base(contig, pos, base[0]);
  }
  // records leave through writev(): header bytes from small owned chunks, FORMAT blocks straight from the batch's pinned memory
  const int fd = out_file == "-" ? 1 : open(out_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { std::cerr << "cannot open " << out_file << "\n"; return -1; }
  vcf_adapter.set_output_fd(fd);
  vcf_adapter.append(raw_prefix.data(), raw_prefix.size(), false);
  BroadCombinedGVCFOperator op(vcf_adapter, query_config.get_vid_mapper(), query_config, htsjdk_conventions);
  auto drain = [&] { vcf_adapter.flush(); };
  // tools/src/gt_mpi_gather.cc:350-362
  VariantQueryProcessorScanState scan_state;
  for (auto i = 0u; i < std::max(1u, query_config.get_num_column_intervals()); ++i) {
    while (!scan_state.end()) {
      qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state);
      drain();
    }
    scan_state.reset();
  }
  drain();
  if (fd != 1) close(fd);
  const auto t3 = now();
  fprintf(fd == 1 ? stderr : stdout, "{\"records\": %llu, \"bytes\": %llu, \"seconds\": {\"context\": %.3f, \"open_array\": %.3f, \"scan_and_write\": %.3f}}\n",
          (unsigned long long)op.get_num_records(), (unsigned long long)vcf_adapter.bytes_written(), secs(t0, t1), secs(t1, t2), secs(t2, t3));
  return 0;
}

// What the Java reader does through JNI (GenomicsDBQueryStream: jniGenomicsDBInit / Read / ReadNextByte / Skip / Close over a
// GenomicsDBBCFGenerator): pull the serialised BCF stream in `chunk`-byte reads, every 5th read a single byte.
static int bcf_generator_stream(const std::string& query_json, int rank, int device, const std::string& header_file, const std::string& ref_bases_file,
                                const std::string& out_file, size_t buffer_capacity, size_t chunk, bool htsjdk_conventions,
                                const std::string& chr, int start, int end, bool header_only, const std::string& loader_json) {
  std::string raw_prefix;
  if (!header_file.empty()) read_header_text(header_file, raw_prefix);
  std::unique_ptr<GenomicsDBBCFGenerator> gen_ptr;
  if (header_file.empty())   // the reference's own constructor: everything from the loader / query configuration
    gen_ptr.reset(new GenomicsDBBCFGenerator(loader_json, query_json, chr.c_str(), start, end, rank, buffer_capacity ? buffer_capacity : 1048576u, 1048576u, "bu", header_only,
                                             htsjdk_conventions, true));
  else
    gen_ptr.reset(new GenomicsDBBCFGenerator(loader_json, query_json, raw_prefix, chr.c_str(), start, end, rank, buffer_capacity ? buffer_capacity : 1048576u, 1048576u, "bu", header_only,
                                             htsjdk_conventions, true, device, ref_bases_file));
  GenomicsDBBCFGenerator& gen = *gen_ptr;
  std::ofstream out(out_file, std::ios::binary);
  if (!out) { std::cerr << "cannot open " << out_file << "\n"; return -1; }
  std::vector<uint8_t> buf(chunk);
  uint64_t reads = 0, bytes = 0;
  while (!gen.end()) {
    if (++reads % 5 == 0) { if (gen.get_num_bytes_available() == 0) { gen.read_and_advance(nullptr, 0, SIZE_MAX); continue; } uint8_t b = gen.read_next_byte(); out.put((char)b); ++bytes; continue; }
    const size_t n = gen.read_and_advance(buf.data(), 0, chunk);
    out.write((const char*)buf.data(), n); bytes += n;
  }
  printf("{\"bytes\": %llu, \"reads\": %llu}\n", (unsigned long long)bytes, (unsigned long long)reads);
  return 0;
}

// gt_mpi_gather.cc:322-366 with PRODUCE_BROAD_GVCF_PRODUCE_INTERESTING_POSITIONS: the Broad scan loop with an operator that
// needs per-site counts only
static int interesting_positions(const std::string& query_json, int rank, int device, const std::string& out_file) {
  VariantQueryConfig query_config;
  query_config.read_from_file(query_json, rank);
  VariantStorageManager sm(query_config.get_workspace(rank), 10u * 1024u * 1024u, device);
  VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
  qp.do_query_bookkeeping(query_config, true);
  std::ofstream file;
  if (out_file != "-") { file.open(out_file, std::ios::binary); if (!file) { std::cerr << "cannot open " << out_file << "\n"; return -1; } }
  std::ostream& out = out_file == "-" ? std::cout : file;
  InterestingLocationsPrinter op(out);
  VariantQueryProcessorScanState scan_state;
  for (auto i = 0u; i < std::max(1u, query_config.get_num_column_intervals()); ++i) {
    while (!scan_state.end()) qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state);
    scan_state.reset();
  }
  out.flush();
  return 0;
}

// MaxAllelesCountOperator (variant_operations.h:410-475) plugged into the same loop: the top_count sites by number of merged ALT alleles
static int max_alleles(const std::string& query_json, int rank, int device, const std::string& out_file, unsigned top_count) {
  VariantQueryConfig query_config;
  query_config.read_from_file(query_json, rank);
  VariantStorageManager sm(query_config.get_workspace(rank), 10u * 1024u * 1024u, device);
  VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
  qp.do_query_bookkeeping(query_config, true);
  MaxAllelesCountOperator op(top_count);
  VariantQueryProcessorScanState scan_state;
  for (auto i = 0u; i < std::max(1u, query_config.get_num_column_intervals()); ++i) {
    while (!scan_state.end()) qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state);
    scan_state.reset();
  }
  std::ofstream out(out_file, std::ios::binary);
  if (!out) { std::cerr << "cannot open " << out_file << "\n"; return -1; }
  op.print(out);
  return 0;
}

// gt_mpi_gather.cc:369-411: the four commands that sit on the columnar iterator
static int columnar_command(const std::string& query_json, int rank, int device, const std::string& out_file, int command) {
  VariantQueryConfig query_config;
  query_config.read_from_file(query_json, rank);
  VariantStorageManager sm(query_config.get_workspace(rank), 10u * 1024u * 1024u, device);
  VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
  qp.do_query_bookkeeping(query_config, /*alleles_required=*/command == 'Q', false, /*stage_columnar_fields=*/true);   // gt_mpi_gather.cc:583-585
  std::ofstream file;
  if (out_file != "-") { file.open(out_file, std::ios::binary); if (!file) { std::cerr << "cannot open " << out_file << "\n"; return -1; } }
  std::ostream& out = out_file == "-" ? std::cout : file;
  switch (command) {
    case 'C': {
      const std::string indent_prefix = "    ";
      out << "{\n" << indent_prefix << "\"variant_calls\": [\n";
      VariantCallPrintOperator printer(out, indent_prefix + indent_prefix, &query_config.get_vid_mapper());
      qp.iterate_over_cells(qp.get_array_descriptor(), query_config, printer, true);
      out << "\n" << indent_prefix << "]\n" << "}\n";
      break;
    }
    case 'V': {
      VariantCallPrintCSVOperator printer(out);
      qp.iterate_over_cells(qp.get_array_descriptor(), query_config, printer, true);
      break;
    }
    case 'A': {
      AlleleCountOperator AC_counter(query_config.get_vid_mapper(), query_config);
      qp.iterate_over_cells(qp.get_array_descriptor(), query_config, AC_counter, true);
      AC_counter.print_allele_counts(out);
      break;
    }
    case 'Q': {   // run_range_query (gt_mpi_gather.cc:102-319), the reference tool's default command: the GA4GH range query
      std::vector<GA4GHVariant> variants;
      for (auto i = 0u; i < query_config.get_num_column_intervals(); ++i) qp.gt_get_column_interval(qp.get_array_descriptor(), query_config, i, variants);
      print_variants(variants, "default", query_config, out, &query_config.get_vid_mapper());
      break;
    }
    case 'G': {
      ColumnHistogramOperator histogram_op(0, 4000000000ull, 100);
      qp.iterate_over_cells(qp.get_array_descriptor(), query_config, histogram_op, true);
      for (uint64_t val : {128, 64, 32, 16, 8, 4, 2}) histogram_op.equi_partition_and_print_bins(val, out);
      break;
    }
  }
  out.flush();
  return 0;
}

int main(int argc, char** argv) {
  std::string query_json, loader_json, out_file, header_file, ref_bases_file; int rank = 0, device = 0; size_t page = 0; bool broad = false, htsjdk = false; int columnar = 0; size_t stream_chunk = 0; bool header_only = false; unsigned max_alleles_top = 0; std::string chr; int chr_start = 0, chr_end = 0;
  static struct option long_options[] = {{"query", 1, 0, 'j'}, {"rank", 1, 0, 'r'}, {"output", 1, 0, 'o'}, {"device", 1, 0, 'd'}, {"page-size", 1, 0, 'p'},
                                         {"produce-Broad-GVCF", 0, 0, 'B'}, {"loader-json-config", 1, 0, 'l'}, {"bcf-header", 1, 0, 'H'}, {"reference-bases", 1, 0, 'R'}, {"htsjdk-conventions", 0, 0, 'J'},
                                         {"produce-interesting-positions", 0, 0, 'I'}, {"bcf-generator", 1, 0, 'S'}, {"header-only", 0, 0, 'h'}, {"max-alleles", 1, 0, 'M'}, {"chromosome", 1, 0, 'c'}, {"begin", 1, 0, 'b'}, {"end", 1, 0, 'e'}, {"print-calls", 0, 0, 'C'}, {"print-csv", 0, 0, 'V'}, {"print-AC", 0, 0, 'a'}, {"produce-histogram", 0, 0, 'G'}, {"range-query", 0, 0, 'Q'},
                                         // the reference tool's own spellings (tools/src/gt_mpi_gather.cc:438-455)
                                         {"json-config", 1, 0, 'j'}, {"output-format", 1, 0, 'O'}, {"workspace", 1, 0, 'w'}, {"array", 1, 0, 'A'}, {"segment-size", 1, 0, 's'},
                                         {"skip-query-on-root", 0, 0, 'k'}, {"version", 0, 0, 'v'}, {0, 0, 0, 0}};
  int c;
  std::string output_format;
  while ((c = getopt_long(argc, argv, "j:l:r:o:d:p:BH:R:JO:w:A:s:", long_options, NULL)) >= 0) {
    switch (c) {
      case 'O': output_format = optarg; break;   // gt_mpi_gather -O: only uncompressed BCF ("bu") is produced here
      case 'w': case 'A': break;                 // read and, as in the reference (gt_mpi_gather.cc:564-565), overridden by the configuration's workspace / array
      case 's': break;                           // TileDB segment size: fragment files are staged whole, there is no segment
      case 'k': break;                           // MPI gather detail of the reference tool: one process per GPU here, every rank queries
      case 'v': std::cout << gdbgpu_version() << "\n"; return 0;
      case 'a': columnar = 'A'; break;
      case 'j': query_json = optarg; break;
      case 'l': loader_json = optarg; break;   // gt_mpi_gather -l: workspace / array / mappings / column partition of the rank from the loader's file
      case 'r': rank = atoi(optarg); break;
      case 'o': out_file = optarg; break;
      case 'd': device = atoi(optarg); break;
      case 'p': page = strtoull(optarg, 0, 10); break;   // operate() calls (dump) / buffered bytes (Broad GVCF) between two "drains" (exercises overflow/resume)
      case 'B': broad = true; break;
      case 'H': header_file = optarg; break;
      case 'R': ref_bases_file = optarg; break;
      case 'J': htsjdk = true; break;
      case 'S': stream_chunk = strtoull(optarg, 0, 10); break;   // pull the stream through GenomicsDBBCFGenerator in reads of this many bytes
      case 'h': header_only = true; break;
      case 'M': max_alleles_top = (unsigned)atoi(optarg); break;   // MaxAllelesCountOperator(top_count) over the query, its report to the output file
      case 'c': chr = optarg; break;
      case 'b': chr_start = atoi(optarg); break;
      case 'e': chr_end = atoi(optarg); break;
      case 'C': case 'V': case 'G': case 'I': case 'Q': columnar = c; break;   // records as GenomicsDBBCFGenerator (JNI) makes them for the Java reader
      default: return -1;
    }
  }
  if (query_json.empty()) { std::cerr << "Query JSON file (-j) is a mandatory argument - unspecified\n"; return -1; }
  if (!output_format.empty() && output_format != "bu") { std::cerr << "gt_gpu_gather: output format '" << output_format << "' is not produced on this path (uncompressed BCF, \"bu\", only)\n"; return -1; }
  try {
    if (out_file.empty()) {
      // the reference tool writes to stdout, the Broad GVCF to the configuration's "vcf_output_filename" (default "-": stdout)
      if (broad && !stream_chunk) { VariantQueryConfig qc; if (!loader_json.empty()) { GenomicsDBImportConfig lc; lc.read_from_file(loader_json, rank); qc.update_from_loader(lc, rank); } qc.read_from_file(query_json, rank); out_file = qc.get_vcf_output_filename(); }
      else if (columnar) out_file = "-";
      else { std::cerr << "usage: gt_gpu_gather -j query.json -o dump.bin [-r rank] [-d cuda_device] [-p sites_per_page]\n"; return -1; }
    }
    if (stream_chunk) {
      return bcf_generator_stream(query_json, rank, device, header_file, ref_bases_file, out_file, page, stream_chunk, htsjdk, chr, chr_start, chr_end, header_only, loader_json);
    }
    if (max_alleles_top) return max_alleles(query_json, rank, device, out_file, max_alleles_top);
    if (columnar == 'I') return interesting_positions(query_json, rank, device, out_file);
    if (columnar) return columnar_command(query_json, rank, device, out_file, columnar);
    if (broad) {
      return produce_broad_gvcf(query_json, loader_json, rank, device, header_file, ref_bases_file, out_file, page, htsjdk);
    }
    VariantQueryConfig query_config;
    query_config.read_from_file(query_json, rank);
    VariantStorageManager sm(query_config.get_workspace(rank), 10u * 1024u * 1024u, device);
    VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
    qp.do_query_bookkeeping(query_config, true);
    FILE* out = fopen(out_file.c_str(), "wb");
    if (!out) { std::cerr << "cannot open " << out_file << "\n"; return -1; }
    HotPathDumpOperator op(query_config, out, page);
    VariantQueryProcessorScanState scan_state;
    for (auto i = 0u; i < std::max(1u, query_config.get_num_column_intervals()); ++i) {
      while (!scan_state.end()) {
        qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state, true);
        op.drain();
      }
      scan_state.reset();
    }
    fclose(out);
    printf("{\"sites\": %llu}\n", (unsigned long long)op.num_sites());
  } catch (const std::exception& e) {
    std::cerr << "gt_gpu_gather: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
