/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref): drives the UNMODIFIED reference engine.
 *
 * gdb_ref_dump links the reference's own objects (VariantQueryProcessor::scan_and_operate,
 * BroadCombinedGVCFOperator::handle_deletions, GA4GHOperator::operate) and plugs an operator subclass into the
 * reference's plugin seam (SingleVariantOperatorBase::operate, variant_operations.h:368).  Modes:
 *   --mode dump  : run handle_deletions + GA4GHOperator::operate per site, then write the canonical hot-path dump
 *                  (layout below; the same layout gdb_b200 and oracle/combine_oracle.c emit) to -o FILE
 *   --mode hot   : same work without the dump (CPU timing of the hot path alone)
 *   --mode null  : the full BroadCombinedGVCFOperator::operate with a VCFAdapter whose hand-off is a no-op
 *                  (INFO/FORMAT collation + bcf1_t build, no serialisation)
 *   --mode fasta  : bases of an indexed FASTA (-j) at the "contig pos" lines of -l, through ReferenceGenomeInfo
 *   --mode header : htslib's header bookkeeping on a header text (-j) plus lines to append (-l)
 *   --mode vid    : print the vid mapping (-j vid.json [-l callsets.json]) as FileBasedVidMapper holds it
 *   --mode config : print the query configuration as the reference's classes hold it after parsing (no storage opened)
 *   --mode bcf-java : the reference's own BroadCombinedGVCFOperator + VCFAdapter writing uncompressed BCF to stdout, constructed
 *                  as the JNI path constructs it (genomicsdb_bcf_generator.cc:60-83: use_missing_values_only_not_vector_end =
 *                  true, the conventions htsjdk's BCF2 codec needs); gt_mpi_gather has no switch for it
 * All modes print one JSON line on stdout: {"mode":..,"sites":N,"calls":M,"seconds":T}.
 *
 * Driver structure mirrors tools/src/gt_mpi_gather.cc:322-366,535-606 (scan_and_produce_Broad_GVCF / main).
 *
 * Canonical dump, little-endian:
 *   header : u32 magic 0x31444247 ("GBD1"), u32 nfields, nfields x { u32 len, bytes name }     (field 0 is GT if queried)
 *   site   : i64 col_begin, i64 col_end, u8 nonref_exists, u8 remap_needed, u16 n_alt,
 *            (1+n_alt) x { u32 len, bytes }   REF then ALTs, NON_REF spelled <NON_REF>
 *            u32 L, then L valid calls in query-row order:
 *              i64 row, i64 call_begin, i64 call_end, u16 A_in, i32 in2merged[A_in],
 *              nfields x { u32 n, u32 word[n] }      n==0 : field invalid / suppressed
 */
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <cstdio>
#include <getopt.h>
#include <mpi.h>
#include "json_config.h"
#include "timer.h"
#include "query_variants.h"
#include "broad_combined_gvcf.h"
#include "vid_mapper_pb.h"

class NullVCFAdapter : public VCFAdapter
{
  public:
    NullVCFAdapter() : VCFAdapter(false) { }
    void handoff_output_bcf_line(bcf1_t*& line, const size_t bcf_record_size) override { }
    void print_header() override { }
};

enum HarnessMode { MODE_DUMP=0, MODE_HOT, MODE_NULL };

class HotPathDumpOperator : public BroadCombinedGVCFOperator
{
  public:
    HotPathDumpOperator(VCFAdapter& adapter, const VidMapper& vid, const VariantQueryConfig& qc, HarnessMode mode, FILE* out)
      : BroadCombinedGVCFOperator(adapter, vid, qc, qc.get_max_diploid_alt_alleles_that_can_be_genotyped()),
      m_mode(mode), m_out(out), m_num_sites(0), m_num_calls(0)
    {
      if(m_GT_query_idx != UNDEFINED_ATTRIBUTE_IDX_VALUE)
        m_dump_fields.push_back(m_GT_query_idx);
      for(auto idx : m_remapped_fields_query_idxs)
        if(qc.get_length_descriptor_for_query_attribute_idx(idx).get_num_dimensions() == 1u)
          m_dump_fields.push_back(idx);
      if(m_mode == MODE_DUMP)
      {
        put<uint32_t>(0x31444247u);
        put<uint32_t>(m_dump_fields.size());
        for(auto idx : m_dump_fields)
          put_string(qc.get_query_attribute_name(idx));
      }
    }
    void operate(Variant& variant, const VariantQueryConfig& query_config) override
    {
      ++m_num_sites;
      if(m_mode == MODE_NULL)
      {
        BroadCombinedGVCFOperator::operate(variant, query_config);
        return;
      }
      handle_deletions(variant, query_config);
      GA4GHOperator::operate(variant, query_config);
      if(m_mode == MODE_HOT)
      {
        for(auto iter=variant.begin();iter!=variant.end();++iter)
          ++m_num_calls;
        return;
      }
      put<int64_t>(variant.get_column_begin());
      put<int64_t>(variant.get_column_end());
      put<uint8_t>(m_NON_REF_exists ? 1 : 0);
      put<uint8_t>(m_remapping_needed ? 1 : 0);
      put<uint16_t>(m_merged_alt_alleles.size());
      put_string(m_merged_reference_allele);
      for(const auto& alt : m_merged_alt_alleles)
        put_string(IS_NON_REF_ALLELE(alt) ? g_vcf_NON_REF : alt);
      uint32_t L = 0;
      for(auto iter=variant.begin();iter!=variant.end();++iter)
        ++L;
      put<uint32_t>(L);
      m_num_calls += L;
      const Variant& src = m_remapping_needed ? m_remapped_variant : variant;
      const unsigned num_merged_alleles = m_merged_alt_alleles.size()+1u;
      for(auto iter=variant.begin();iter!=variant.end();++iter)
      {
        const auto& call = *iter;
        auto call_idx = iter.get_call_idx_in_variant();
        put<int64_t>(call.get_row_idx());
        put<int64_t>(call.get_column_begin());
        put<int64_t>(call.get_column_end());
        const auto& alt_vec = get_known_field<VariantFieldALTData, true>(call, query_config, GVCF_ALT_IDX)->get();
        uint16_t A_in = alt_vec.size()+1u;
        put<uint16_t>(A_in);
        for(auto i=0u;i<A_in;++i)
        {
          auto v = m_alleles_LUT.get_merged_idx_for_input(call_idx, i);
          put<int32_t>(CombineAllelesLUT::is_missing_value(v) ? -1 : static_cast<int32_t>(v));
        }
        for(auto idx : m_dump_fields)
        {
          auto& ld = query_config.get_length_descriptor_for_query_attribute_idx(idx);
          const auto& fptr = src.get_call(call_idx).get_field(idx);
          bool suppressed = (m_remapping_needed && ld.is_length_genotype_dependent()
              && too_many_alt_alleles_for_genotype_length_fields(num_merged_alleles-1u));
          if(!fptr.get() || !fptr->is_valid() || suppressed)
          {
            put<uint32_t>(0u);
            continue;
          }
          uint32_t n = fptr->length();
          put<uint32_t>(n);
          if(n)
            fwrite(fptr->get_raw_pointer(), 4u, n, m_out);
        }
      }
    }
    uint64_t num_sites() const { return m_num_sites; }
    uint64_t num_calls() const { return m_num_calls; }
  private:
    template<class T> void put(T v) { fwrite(&v, sizeof(T), 1u, m_out); }
    void put_string(const std::string& s) { put<uint32_t>(s.length()); if(s.length()) fwrite(s.data(), 1u, s.length(), m_out); }
    HarnessMode m_mode;
    FILE* m_out;
    uint64_t m_num_sites;
    uint64_t m_num_calls;
    std::vector<unsigned> m_dump_fields;
};

int main(int argc, char** argv)
{
  std::string loader_json, query_json, out_file, mode_str = "dump";
  int rank = 0;
  uint64_t segment_size = 0;
  static struct option long_options[] = {
    {"loader", 1, 0, 'l'}, {"query", 1, 0, 'j'}, {"rank", 1, 0, 'r'}, {"output", 1, 0, 'o'},
    {"mode", 1, 0, 'm'}, {"segment-size", 1, 0, 's'}, {0,0,0,0} };
  int c;
  while((c = getopt_long(argc, argv, "l:j:r:o:m:s:", long_options, NULL)) >= 0)
  {
    switch(c)
    {
      case 'l': loader_json = optarg; break;
      case 'j': query_json = optarg; break;
      case 'r': rank = strtol(optarg, 0, 10); break;
      case 'o': out_file = optarg; break;
      case 'm': mode_str = optarg; break;
      case 's': segment_size = strtoull(optarg, 0, 10); break;
      default: std::cerr << "unknown option\n"; return -1;
    }
  }
  if(mode_str == "bcf-java")
  {
    VariantQueryConfig query_config;
    VCFAdapter vcf_adapter;
    query_config.read_from_file(query_json, rank);
    query_config.set_vcf_output_format("bu");
    vcf_adapter.initialize(query_config);
    VariantStorageManager sm(query_config.get_workspace(rank), segment_size ? segment_size : query_config.get_segment_size());
    VariantQueryProcessor qp(&sm, query_config.get_array_name(rank), query_config.get_vid_mapper());
    qp.do_query_bookkeeping(qp.get_array_schema(), query_config, query_config.get_vid_mapper(), true);
    {
      BroadCombinedGVCFOperator op(vcf_adapter, query_config.get_vid_mapper(), query_config,
          query_config.get_max_diploid_alt_alleles_that_can_be_genotyped(), true);
      VariantQueryProcessorScanState scan_state;
      for(auto i=0u;i<std::max(1u, query_config.get_num_column_intervals());++i)
      {
        while(!scan_state.end())
          qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state);
        scan_state.reset();
      }
    }
    sm.close_array(qp.get_array_descriptor());
    GenomicsDBProtoBufInitAndCleanup::shutdown_protobuf_library();
    return 0;
  }
  if(mode_str == "fasta")
  {
    // ReferenceGenomeInfo (vcf_adapter.cc:31-58) on an indexed FASTA (-j; the .fai is built when missing): one line "contig pos"
    // of -l per query, in the file's order (the reader keeps a 32 KB window), the base printed per line
    ReferenceGenomeInfo info;
    info.initialize(query_json);
    std::ifstream l(loader_json); std::string contig; int pos;
    while(l >> contig >> pos)
      std::cout << contig << " " << pos << " " << info.get_reference_base_at_position(contig.c_str(), pos) << "\n";
    return 0;
  }
  if(mode_str == "header")
  {
    // htslib's header bookkeeping (the reference's own copy): -j names a header text ("" lines: bcf_hdr_init("w")), -l, optional,
    // a file of lines to append one by one (a line "SAMPLE <name>" adds a sample); then what the Broad operator does before it
    // writes the header (broad_combined_gvcf.cc:314-325): format without IDX, parse again, format as BCF.  Printed: the final
    // text, then "ID <name> <idx>" per dictionary entry.
    std::ifstream f(query_json); std::stringstream ss; ss << f.rdbuf();
    auto text = ss.str();
    size_t hdr_len = 0;
    bcf_hdr_t* h = bcf_hdr_init(text.empty() ? "w" : "r");
    if(!text.empty() && bcf_hdr_parse_required_sample_line(h, &(text[0]), &hdr_len, 0) != 0) { std::cerr << "parse failed\n"; return 1; }
    if(!loader_json.empty())
    {
      std::ifstream l(loader_json); std::string line;
      while(std::getline(l, line))
      {
        if(line.compare(0, 6, "FIELD ") == 0 && line.size() > 8)   // "FIELD I|F|T <name>": VCFAdapter::add_field_to_hdr_if_missing with the vid mapping named by -o
        {
          static FileBasedVidMapper* vid = 0;
          if(!vid) vid = new FileBasedVidMapper(out_file);
          const int type = line[6] == 'I' ? BCF_HL_INFO : line[6] == 'F' ? BCF_HL_FMT : BCF_HL_FLT;
          try { std::cout << "field " << VCFAdapter::add_field_to_hdr_if_missing(h, vid, line.substr(8), type) << "\n"; }
          catch(const std::exception& e) { std::cout << "field refused\n"; }
        }
        else if(line.compare(0, 7, "SAMPLE ") == 0) std::cout << "add_sample " << bcf_hdr_add_sample(h, line.c_str()+7) << "\n";
        else std::cout << "append " << bcf_hdr_append(h, line.c_str()) << "\n";
      }
      bcf_hdr_add_sample(h, 0);
      bcf_hdr_sync(h);
    }
    kstring_t a = {0, 0, 0};
    bcf_hdr_format(h, 0, &a);
    bcf_hdr_t* h2 = bcf_hdr_init("r");
    if(bcf_hdr_parse_required_sample_line(h2, a.s, &hdr_len, 0) != 0) { std::cerr << "second parse failed\n"; return 1; }
    kstring_t b = {0, 0, 0};
    bcf_hdr_format(h2, 1, &b);
    std::cout << b.s;
    for(int t : { BCF_DT_ID, BCF_DT_CTG })
      for(int i=0;i<h2->n[t];++i)
        if(h2->id[t][i].key) std::cout << (t == BCF_DT_ID ? "ID " : "CTG ") << h2->id[t][i].key << " " << i << "\n";
    std::cout << "nsamples " << bcf_hdr_nsamples(h2) << "\n";
    return 0;
  }
  if(mode_str == "vid")
  {
    // the vid mapping (-j names the file; -l, optional, a callset mapping file) as the reference's FileBasedVidMapper holds it:
    // every field in index order (FORMAT twins, flattened tuple elements and the mandatory fields included), contigs, callsets
    FileBasedVidMapper vid(query_json);
    if(!loader_json.empty())
      vid.parse_callsets_json(loader_json, true);
    auto ht_name = [](const int ht) { return ht == BCF_HT_FLAG ? "flag" : ht == BCF_HT_INT ? "int" : ht == BCF_HT_REAL ? "float" : (ht == BCF_HT_STR || ht == BCF_HT_CHAR) ? "char" : "?"; };
    std::cout << "{\"fields\": [";
    for(auto i=0u;i<vid.get_num_fields();++i)
    {
      auto& f = vid.get_field_info(i);
      std::cout << (i ? ", " : "") << "{\"name\": \"" << f.m_name << "\", \"idx\": " << f.m_field_idx << ", \"vcf_name\": \"" << f.m_vcf_name
        << "\", \"class\": \"" << (f.m_is_vcf_INFO_field ? "I" : "") << (f.m_is_vcf_FORMAT_field ? "F" : "") << (f.m_is_vcf_FILTER_field ? "T" : "")
        << "\", \"type\": [";
      auto& t = f.get_genomicsdb_type();
      for(auto j=0u;j<t.get_num_elements_in_tuple();++j)
      {
        auto ht = t.get_tuple_element_bcf_ht_type(j);
        std::cout << (j ? ", " : "") << "\"" << ht_name(ht) << "\"";
      }
      std::cout << "], \"length\": [";
      auto& l = f.m_length_descriptor;
      for(auto d=0u;d<l.get_num_dimensions();++d)
      {
        std::cout << (d ? ", " : "") << "\"";
        if(l.is_fixed_size_dimension(d))
          std::cout << l.get_num_elements_in_dimension(d);
        else
          std::cout << g_length_descriptor_int_to_string[l.get_length_descriptor(d)];
        std::cout << "\"";
      }
      std::cout << "], \"op\": " << f.m_VCF_field_combine_operation << ", \"flattened\": " << f.is_flattened_field();
      if(l.get_num_dimensions() > 1u)
      {
        std::cout << ", \"delims\": \"";
        for(auto d=0u;d<l.get_num_dimensions();++d)
          std::cout << l.get_vcf_delimiter(d);
        std::cout << "\"";
      }
      std::cout << "}";
    }
    std::cout << "], \"contigs\": [";
    for(auto i=0u;i<vid.get_num_contigs();++i)
    {
      auto& c = vid.get_contig_info(i);
      std::cout << (i ? ", " : "") << "[\"" << c.m_name << "\", " << c.m_tiledb_column_offset << ", " << c.m_length << "]";
    }
    std::cout << "], \"callsets\": [";
    if(vid.is_callset_mapping_initialized())
      for(auto r=0ll;r<=vid.get_max_callset_row_idx();++r)
      {
        std::string name;
        vid.get_callset_name(r, name);
        std::cout << (r ? ", " : "") << "\"" << name << "\"";
      }
    std::cout << "], \"locate\": [";
    // a few columns -> (contig, 0-based position), -1: none
    auto first = true;
    for(auto i=0u;i<vid.get_num_contigs();++i)
    {
      auto& c = vid.get_contig_info(i);
      for(auto col : { c.m_tiledb_column_offset - 1, c.m_tiledb_column_offset, c.m_tiledb_column_offset + c.m_length - 1, c.m_tiledb_column_offset + c.m_length })
      {
        std::string name;
        int64_t pos = -1;
        auto found = (col >= 0) && vid.get_contig_location(col, name, pos);
        std::cout << (first ? "" : ", ") << "[" << col << ", \"" << (found ? name : "") << "\", " << (found ? pos : -1) << "]";
        first = false;
      }
    }
    std::cout << "]}" << std::endl;
    return 0;
  }
  if(mode_str == "config")
  {
    // the query configuration as the reference's own classes hold it after the calls gt_mpi_gather.cc:547-557 makes
    // (loader first, the query file, then the partition subset): one JSON line, no storage is opened
    VariantQueryConfig query_config;
    GenomicsDBImportConfig loader_config;
    if(!loader_json.empty())
    {
      loader_config.read_from_file(loader_json, rank);
      query_config.update_from_loader(loader_config, rank);
    }
    query_config.read_from_file(query_json, rank);
    if(!loader_json.empty())
      query_config.subset_query_column_ranges_based_on_partition(loader_config, rank);
    std::cout << "{\"workspace\": \"" << query_config.get_workspace(rank) << "\", \"array\": \"" << query_config.get_array_name(rank) << "\", \"intervals\": [";
    for(auto i=0u;i<query_config.get_num_column_intervals();++i)
      std::cout << (i ? ", " : "") << "[" << (long long)query_config.get_column_begin(i) << ", " << (long long)query_config.get_column_end(i) << "]";
    std::cout << "], \"base_ranges\": [";
    auto& base = query_config.get_query_column_ranges(rank);
    for(auto i=0u;i<base.size();++i)
      std::cout << (i ? ", " : "") << "[" << (long long)base[i].first << ", " << (long long)base[i].second << "]";
    std::cout << "], \"all_rows\": " << (query_config.query_all_rows() ? "true" : "false") << ", \"rows\": [";
    auto& rows = query_config.get_rows_to_query();
    for(auto i=0u;i<rows.size();++i)
      std::cout << (i ? ", " : "") << (long long)rows[i];
    std::cout << "], \"attributes\": [";
    for(auto i=0u;i<query_config.get_num_queried_attributes();++i)
      std::cout << (i ? ", " : "") << "\"" << query_config.get_query_attribute_name(i) << "\"";
    std::cout << "], \"segment_size\": " << query_config.get_segment_size()
      << ", \"max_diploid_alt_alleles_that_can_be_genotyped\": " << query_config.get_max_diploid_alt_alleles_that_can_be_genotyped()
      << ", \"produce_GT_field\": " << query_config.produce_GT_field() << ", \"produce_FILTER_field\": " << query_config.produce_FILTER_field()
      << ", \"sites_only_query\": " << query_config.sites_only_query()
      << ", \"produce_GT_with_min_PL_value_for_spanning_deletions\": " << query_config.produce_GT_with_min_PL_value_for_spanning_deletions()
      << ", \"vcf_header_filename\": \"" << query_config.get_vcf_header_filename() << "\", \"reference_genome\": \"" << query_config.get_reference_genome()
      << "\", \"loader_partition\": [" << (loader_json.empty() ? 0ll : (long long)loader_config.get_column_partition(rank).first) << ", "
      << (loader_json.empty() ? 0ll : (long long)loader_config.get_column_partition(rank).second) << "]"
      << ", \"vcf_output_filename\": \"" << query_config.get_vcf_output_filename()
      << "\", \"vcf_output_format\": \"" << query_config.get_vcf_output_format() << "\"}" << std::endl;
    GenomicsDBProtoBufInitAndCleanup::shutdown_protobuf_library();
    return 0;
  }
  HarnessMode mode = (mode_str == "hot") ? MODE_HOT : (mode_str == "null") ? MODE_NULL : MODE_DUMP;
  if(query_json.empty() || (mode == MODE_DUMP && out_file.empty()))
  {
    std::cerr << "usage: gdb_ref_dump [-l loader.json] -j query.json [-r rank] [-s segment_size] --mode dump|hot|null [-o out.bin]\n";
    return -1;
  }
  FILE* out = 0;
  if(mode == MODE_DUMP)
  {
    out = fopen(out_file.c_str(), "wb");
    if(!out) { std::cerr << "cannot open "<<out_file<<"\n"; return -1; }
  }
  {
    VariantQueryConfig query_config;
    NullVCFAdapter vcf_adapter;
    GenomicsDBImportConfig loader_config;
    if(!loader_json.empty())
    {
      loader_config.read_from_file(loader_json, rank);
      query_config.update_from_loader(loader_config, rank);
    }
    query_config.read_from_file(query_json, rank);
    if(!loader_json.empty())
      query_config.subset_query_column_ranges_based_on_partition(loader_config, rank);
    vcf_adapter.initialize(query_config);
    auto workspace = query_config.get_workspace(rank);
    auto array_name = query_config.get_array_name(rank);
    if(segment_size == 0)
      segment_size = query_config.get_segment_size();
    VariantStorageManager sm(workspace, segment_size);
    VariantQueryProcessor qp(&sm, array_name, query_config.get_vid_mapper());
    qp.do_query_bookkeeping(qp.get_array_schema(), query_config, query_config.get_vid_mapper(), true);
    {
      HotPathDumpOperator op(vcf_adapter, query_config.get_vid_mapper(), query_config, mode, out);
      Timer timer;
      timer.start();
      VariantQueryProcessorScanState scan_state;
      for(auto i=0u;i<std::max(1u, query_config.get_num_column_intervals());++i)
      {
        while(!scan_state.end())
          qp.scan_and_operate(qp.get_array_descriptor(), query_config, op, i, true, &scan_state);
        scan_state.reset();
      }
      timer.stop();
      printf("{\"mode\": \"%s\", \"sites\": %llu, \"calls\": %llu, \"seconds\": %.6f}\n", mode_str.c_str(),
          (unsigned long long)op.num_sites(), (unsigned long long)op.num_calls(), timer.get_last_interval_wall_clock_time()*1e-6);
    }
    sm.close_array(qp.get_array_descriptor());
  }
  if(out)
    fclose(out);
  GenomicsDBProtoBufInitAndCleanup::shutdown_protobuf_library();
  return 0;
}
