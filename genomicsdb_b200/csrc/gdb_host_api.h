// Host-side C++ mirror of the reference's plugin surface for the variant-combine path, written on top of the C ABI
// (include/gdbgpu.h) only.  Same class names, method names, argument meaning and error behaviour (C++ exceptions) as
// GenomicsDB v0.10.2, so that tools/src/gt_mpi_gather.cc:322-366 (scan_and_produce_Broad_GVCF) and
// src/main/cpp/src/vcf/genomicsdb_bcf_generator.cc:83,135 read the same against either implementation:
//
//   reference                                                         here
//   VariantStorageManager(workspace, segment_size)                    VariantStorageManager(workspace, segment_size, device)
//     variant_storage_manager.h:160-235
//   VariantQueryConfig::read_from_file(json)                          same keys (json_config.cc:207-656), see .cc
//   VariantQueryProcessor(&sm, array_name, vid_mapper)                same
//   qp.do_query_bookkeeping(schema, config, vid_mapper, alleles)      same role: opens + stages the array on the GPU
//     query_variants.cc:578-685
//   qp.scan_and_operate(ad, config, op, interval_idx,                 same signature (query_variants.h:241-243); batches
//       handle_spanning_deletions, &scan_state)                       come from gdbgpu_scan_next_batch, overflow()/resume
//                                                                     protocol identical (query_variants.cc:305,415,460-475)
//   SingleVariantOperatorBase::operate(Variant&, config)              same virtual; the merged REF/ALT, NON_REF flag and
//     variant_operations.h:349-388                                    remapping flag are already computed by the kernels
//   GA4GHOperator::operate / get_remapped_variant()                   same; the remapped GT / allele-dependent fields are
//     variant_operations.h:648-686                                    views into the batch (no per-element virtual calls)
//
// Difference that matters to an operator author: the Variant handed to operate() carries the REMAPPED fields of every
// valid call (what GA4GHOperator::m_remapped_variant holds in the reference); the un-remapped inputs stay in HBM.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <iostream>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gdbgpu.h"

namespace gdb { struct JsonValue; }
namespace genomicsdb_b200 {

class GenomicsDBGPUException : public std::runtime_error {
 public:
  GenomicsDBGPUException(int code, const std::string& m) : std::runtime_error(m), code_(code) {}
  int code() const { return code_; }
 private:
  int code_;
};
class VariantOperationException : public GenomicsDBGPUException { public: explicit VariantOperationException(const std::string& m) : GenomicsDBGPUException(GDBGPU_ERR_ARG, m) {} };
// query_variants.cc:533-535
class VariantQueryProcessorException : public GenomicsDBGPUException { using GenomicsDBGPUException::GenomicsDBGPUException; };

// ---- field / contig / callset mapping (src/main/cpp/src/utils/vid_mapper.cc, file-based mapper)
struct FieldInfo {
  std::string name;
  std::string type;        // "int", "float", "char", "flag"
  std::string vcf_type;    // the type the VCF header declares ("vcf_type" of the vid mapping; the field's own type when absent)
  std::string length;      // "1", "R", "A", "G", "P", "PP", "VAR", "<n>"; 2-D descriptors are kept as "2D"
  bool is_format = false, is_info = false, is_filter = false;
  int global_idx = -1;     // the reference's field index (what FILTER cells store): file order, FORMAT twins and flattened tuple elements counted
  bool is_format_twin = false;   // <name>_FORMAT of a field that is both INFO and FORMAT (or a tuple element of that twin)
  std::string combine_op;  // "VCF_field_combine_operation" of the vid mapping, else the known field's default (known_field_info.cc:284-308)
  bool is_allele_dependent() const { return length == "R" || length == "A" || length == "G"; }
  // multi-dimensional fields and fields whose elements are tuples (vid_mapper.cc, genomicsdb_multid_vector_field.h:66-83): length
  // stays "2D", the per-dimension descriptors, VCF delimiters and tuple element types are kept here.  A composite field (tuple
  // of n > 1 elements) is stored as n flattened attributes <name>_tuple_element_<j> (VidMapper::get_flattened_field_info); each
  // gets its own FieldInfo naming its parent.
  std::vector<std::string> dim_lengths;    // {"R", "var"}
  std::vector<char> vcf_delimiters;        // {'|', ','}
  std::vector<std::string> tuple_types;    // {"float", "int"}
  std::string parent;                      // flattened element: the composite field's name (== its VCF name)
  int tuple_index = 0;
  bool is_multid() const { return length == "2D"; }
  bool is_composite() const { return tuple_types.size() > 1 && parent.empty(); }
  bool dim0_allele_dependent() const { return !dim_lengths.empty() && (dim_lengths[0] == "R" || dim_lengths[0] == "A"); }
  const std::string& vcf_name() const { return parent.empty() ? name : parent; }
};
struct ContigInfo { std::string name; int64_t tiledb_column_offset = 0, length = 0; };
class VidMapper {
 public:
  VidMapper() = default;
  explicit VidMapper(const std::string& vid_json_file);
  bool is_initialized() const { return initialized_; }
  unsigned get_num_fields() const { return (unsigned)num_fields_; }   // with the mandatory fields the file does not list
  const FieldInfo* get_field_info(const std::string& name) const;
  const std::vector<FieldInfo>& fields() const { return fields_; }
  // vid_mapper.cc get_contig_location: the contig a TileDB column falls into (nullptr: unknown position)
  const ContigInfo* get_contig_location(int64_t column) const;
  const ContigInfo* get_contig_info(const std::string& name) const;   // nullptr: no such contig
  bool get_contig_info(const std::string& name, ContigInfo& info) const { const ContigInfo* c = get_contig_info(name); if (c) info = *c; return c != nullptr; }
  // contigs in the order of the vid mapping file (VidMapper::get_contig_info(idx): the order contig lines are added to a header in)
  unsigned get_num_contigs() const { return (unsigned)contigs_file_order_.size(); }
  const ContigInfo& get_contig_info(unsigned idx) const { return contigs_file_order_[idx]; }
  // callset mapping (vid_mapper.cc:686-711,319-325): sample name of every TileDB row
  void read_callsets_file(const std::string& callset_mapping_file);
  void read_callsets(const gdb::JsonValue& callsets);   // a dictionary name -> info or a list of info dictionaries
  bool is_callset_mapping_initialized() const { return callsets_initialized_; }
  bool get_callset_name(int64_t row_idx, std::string& callset_name) const;
  int64_t get_max_callset_row_idx() const { return (int64_t)callset_names_.size() - 1; }
 private:
  std::vector<FieldInfo> fields_;
  std::vector<ContigInfo> contigs_;   // sorted by tiledb_column_offset
  std::vector<ContigInfo> contigs_file_order_;
  std::vector<std::string> callset_names_;   // by row
  std::vector<bool> callset_rows_named_;
  int num_fields_ = 0;
  bool initialized_ = false, callsets_initialized_ = false;
};

// ---- loader configuration (include/config/genomicsdb_config_base.h GenomicsDBImportConfig, src/config/json_config.cc:340-440,
// 718-790): what the query side takes from it -- workspace / array / mapping files of the rank's column partition and the
// partition's column range
class GenomicsDBImportConfig {
 public:
  void read_from_file(const std::string& filename, const int rank = 0);
  bool is_partitioned_by_column() const { return !row_based_; }
  std::pair<int64_t, int64_t> get_column_partition(const int rank, const unsigned idx = 0u) const;
  const std::string& get_workspace(const int rank) const;
  const std::string& get_array_name(const int rank) const;
  const VidMapper& get_vid_mapper() const { return vid_; }
  const std::string& get_callset_mapping_file() const { return callset_mapping_file_; }
  const std::string& get_vcf_header_filename() const { return vcf_header_filename_; }
  const std::string& get_reference_genome() const { return reference_genome_; }
  size_t get_segment_size() const { return segment_size_; }
 private:
  bool row_based_ = false, partitions_specified_ = false, single_workspace_ = false, single_array_ = false;
  std::vector<std::pair<int64_t, int64_t>> partitions_;   // in file order, ends fixed against the sorted order
  std::vector<std::string> workspaces_, array_names_;
  VidMapper vid_;
  std::string callset_mapping_file_, vcf_header_filename_, reference_genome_;
  size_t segment_size_ = 10u * 1024u * 1024u;
};

// ---- query configuration (include/config/variant_query_config.h)
class VariantQueryConfig {
 public:
  void read_from_file(const std::string& filename, int rank = 0);
  // programmatic setters, same names as the reference
  void set_workspace(const std::string& w) { workspace_ = w; }
  void set_array_name(const std::string& a) { array_name_ = a; }
  void set_attributes_to_query(const std::vector<std::string>& a) { attributes_ = a; }
  void add_column_interval_to_query(int64_t b, int64_t e);
  void set_column_interval_to_query(int64_t b, int64_t e) { intervals_.assign(1, {b, e}); }
  void set_rows_to_query(const std::vector<int64_t>& rows);
  void set_vid_mapper(const VidMapper& v) { vid_ = v; }
  void set_max_diploid_alt_alleles_that_can_be_genotyped(unsigned v) { max_alt_ = v; }
  void set_produce_GT_with_min_PL_value_for_spanning_deletions(bool v) { min_pl_gt_ = v; }
  void set_produce_GT_field(bool v) { produce_gt_ = v; }
  bool produce_GT_field() const { return produce_gt_; }
  void set_produce_FILTER_field(bool v) { produce_filter_ = v; }
  bool produce_FILTER_field() const { return produce_filter_; }
  void set_sites_only_query(bool v) { sites_only_ = v; }
  bool sites_only_query() const { return sites_only_; }
  // genomicsdb_config_base.cc:179-223: what a query takes from the loader's configuration (gt_mpi_gather -l, the JNI path)
  void update_from_loader(const GenomicsDBImportConfig& loader_config, const int rank = 0);
  void subset_query_column_ranges_based_on_partition(const GenomicsDBImportConfig& loader_config, const int rank = 0);
  void set_vcf_output_format(const std::string& f) { vcf_output_format_ = f; }
  const std::string& get_vcf_output_format() const { return vcf_output_format_; }
  const std::string& get_vcf_header_filename() const { return vcf_header_filename_; }
  // where the Broad GVCF goes ("-": stdout, the default; json_config.cc:577-617: one name, one per rank, or the rank's column partition's)
  const std::string& get_vcf_output_filename() const { return vcf_output_filename_; }
  void set_vcf_header_filename(const std::string& f) { vcf_header_filename_ = f; }
  const std::string& get_reference_genome() const { return reference_genome_; }
  void set_reference_genome(const std::string& f) { reference_genome_ = f; }
  size_t get_segment_size() const { return segment_size_; }
  // rows of the query in query order (all rows of the callset mapping when none are given)
  uint64_t get_num_rows_to_query() const { return rows_.empty() ? (uint64_t)(vid_.get_max_callset_row_idx() + 1) : rows_.size(); }
  int64_t get_array_row_idx_for_query_row_idx(uint64_t i) const { return rows_.empty() ? (int64_t)i : rows_[i]; }

  const std::string& get_workspace(int = 0) const { return workspace_; }
  const std::string& get_array_name(int = 0) const { return array_name_; }
  // the rank's ranges as the file lists them (GenomicsDBConfigBase::get_query_column_ranges): what the partition subset edits
  const std::vector<std::pair<int64_t, int64_t>>& get_query_column_ranges(int = 0) const { return base_column_ranges_; }
  unsigned get_num_column_intervals() const { return (unsigned)intervals_.size(); }
  int64_t get_column_begin(unsigned i) const { return intervals_[i].first; }
  int64_t get_column_end(unsigned i) const { return intervals_[i].second; }
  bool query_all_rows() const { return rows_.empty(); }
  const std::vector<int64_t>& get_rows_to_query() const { return rows_; }
  const std::vector<std::string>& get_query_attributes() const { return attributes_; }
  // after do_query_bookkeeping: END / ALT / REF / GT added as the reference adds them, END, REF, ALT swapped to the front
  // (query_variants.cc:578-637, variant_query_config.cc:161-185) -- the order operators see the fields in
  const std::vector<std::string>& get_ordered_query_attributes() const { return ordered_attributes_; }
  const VidMapper& get_vid_mapper() const { return vid_; }
  unsigned get_max_diploid_alt_alleles_that_can_be_genotyped() const { return max_alt_; }
  bool produce_GT_with_min_PL_value_for_spanning_deletions() const { return min_pl_gt_; }
  bool is_bookkeeping_done() const { return bookkeeping_done_; }
  // after do_query_bookkeeping: the fields the operator sees (GT first when present), in query order
  unsigned get_num_remapped_fields() const { return (unsigned)remapped_names_.size(); }
  const std::string& get_remapped_field_name(unsigned i) const { return remapped_names_[i]; }

 private:
  friend class VariantQueryProcessor;
  std::string workspace_, array_name_;
  std::vector<std::pair<int64_t, int64_t>> intervals_, base_column_ranges_;
  std::vector<int64_t> rows_;
  std::vector<std::string> attributes_, ordered_attributes_;
  VidMapper vid_;
  unsigned max_alt_ = 50;
  bool min_pl_gt_ = false, produce_gt_ = false, sites_only_ = false, produce_filter_ = false;
  bool bookkeeping_done_ = false;
  std::string vcf_header_filename_, reference_genome_, vcf_output_format_ = "bu", callset_mapping_file_, vcf_output_filename_ = "-";
  size_t segment_size_ = 10u * 1024u * 1024u;
  std::vector<std::string> remapped_names_;
  // plain fields staged for the Broad operator (do_query_bookkeeping): FORMAT fields that are not allele-dependent, scalar
  // INFO fields with a median / sum / mean combine operation, the inputs of the DP rule
  struct PlainField { std::string attribute, vcf_name; int elem = 0, fixed_len = 0; bool format = false; int info_op = 0; bool is_qual = false, is_id = false, is_filter = false, is_concat = false, is_raw = false; };   // is_raw: multi-dimensional field (or the REF / ALT its remap needs), combined by the operator
  std::vector<PlainField> plain_;
  unsigned dp_fields_[3] = {0, 0, 0};   // 1-based plain indices of INFO DP, MIN_DP, DP_FORMAT
  unsigned esum_remapped_mask_ = 0;     // remapped fields (GT not counted) combined with an element-wise sum
 public:
  const std::vector<PlainField>& get_plain_fields() const { return plain_; }
  const unsigned* get_dp_rule_fields() const { return dp_fields_; }
  unsigned get_element_wise_sum_mask() const { return esum_remapped_mask_; }
  // the queried attributes in the order the reference's bookkeeping leaves them (END, REF, ALT first:
  // variant_query_config.cc:161-185); filled by do_query_bookkeeping(..., stage_columnar_fields = true)
  struct QueryAttribute { std::string name; int kind = -1; int index = 0; char elem = 'i'; int fixed_len = 0; bool multid = false; };   // kind: -1 END, -2 multi-dimensional (not staged), else GDBGPU_CELL_FIELD_*
  unsigned get_num_queried_attributes() const { return (unsigned)query_attrs_.size(); }
  const std::string& get_query_attribute_name(unsigned i) const { return query_attrs_[i].name; }
  const QueryAttribute& get_query_attribute(unsigned i) const { return query_attrs_[i]; }
  bool get_query_idx_for_name(const std::string& name, unsigned& idx) const {
    for (unsigned i = 0; i < query_attrs_.size(); ++i) if (query_attrs_[i].name == name) { idx = i; return true; }
    return false;
  }
 private:
  std::vector<QueryAttribute> query_attrs_;
};

// ---- data model (include/genomicsdb/variant.h).  Views into the current batch: valid until the next operate() call.
class Variant;
class VariantCall {
 public:
  bool is_valid() const { return valid_; }
  uint64_t get_row_idx() const { return row_; }
  uint64_t get_column_begin() const { return (uint64_t)begin_; }
  uint64_t get_column_end() const { return (uint64_t)end_; }
  bool is_spanning_deletion_form() const { return reduced_; }   // handle_deletions already applied (REF N, ALT *[,NON_REF])
  // remapped field i (0 = GT when the array has GT): words and length; length 0 == field invalid for this call
  const uint32_t* get_field_words(unsigned i) const { return fields_[i].first; }
  unsigned get_field_length(unsigned i) const { return fields_[i].second; }
  template <class T> const T* get_field(unsigned i) const { return reinterpret_cast<const T*>(fields_[i].first); }
  // input->merged allele index LUT of this call (only when the scan asked for it)
  const int16_t* get_alleles_LUT() const { return lut_; }
  unsigned get_num_input_alleles() const { return lut_len_; }
 private:
  friend class VariantQueryProcessor;
  bool valid_ = false, reduced_ = false;
  uint64_t row_ = 0; int64_t begin_ = 0, end_ = 0;
  std::vector<std::pair<const uint32_t*, unsigned>> fields_;
  const int16_t* lut_ = nullptr; unsigned lut_len_ = 0;
};

class Variant {
 public:
  uint64_t get_column_begin() const { return (uint64_t)begin_; }
  uint64_t get_column_end() const { return (uint64_t)end_; }
  uint64_t get_num_calls() const { return calls_.size(); }
  VariantCall& get_call(uint64_t i) { return calls_[i]; }
  const VariantCall& get_call(uint64_t i) const { return calls_[i]; }
  // iteration over valid calls in query-row order (variant.h:290-320)
  class valid_calls_iterator {
   public:
    valid_calls_iterator(const std::vector<VariantCall>* c, uint64_t i) : c_(c), i_(i) { skip(); }
    bool operator!=(const valid_calls_iterator& o) const { return i_ != o.i_; }
    const VariantCall& operator*() const { return (*c_)[i_]; }
    valid_calls_iterator& operator++() { ++i_; skip(); return *this; }
    uint64_t get_call_idx_in_variant() const { return i_; }
   private:
    void skip() { while (i_ < c_->size() && !(*c_)[i_].is_valid()) ++i_; }
    const std::vector<VariantCall>* c_; uint64_t i_;
  };
  valid_calls_iterator begin() const { return valid_calls_iterator(&calls_, 0); }
  valid_calls_iterator end() const { return valid_calls_iterator(&calls_, calls_.size()); }
 private:
  friend class VariantQueryProcessor;
  friend class SingleVariantOperatorBase;
  friend class GA4GHOperator;
  int64_t begin_ = 0, end_ = 0;
  std::vector<VariantCall> calls_;
  // merged-site results of the kernels, picked up by SingleVariantOperatorBase::operate
  std::string merged_ref_;
  std::vector<std::string> merged_alts_;
  bool nonref_ = false, remapped_ = false, g_suppressed_ = false;
};

// ---- operators (include/query_operations/variant_operations.h:349-388, 648-686)
class SingleVariantOperatorBase {
 public:
  virtual ~SingleVariantOperatorBase() = default;
  virtual void operate(Variant& variant, const VariantQueryConfig& query_config);
  virtual bool overflow() const { return false; }
  // Opt-in fast path: an operator that consumes the kernels' BCF-typed FORMAT blocks and combined INFO scalars
  // per site instead of a materialised Variant; scan_and_operate then scans with GDBGPU_SCAN_BCF_FORMAT.
  virtual bool wants_bcf_blocks() const { return false; }
  virtual bool wants_htsjdk_conventions() const { return false; }   // FORMAT blocks as htsjdk's BCF2 codec needs them (JNI path)
  virtual void operate_on_bcf_site(const gdbgpu_batch&, uint32_t /*site*/, const VariantQueryConfig&) {}
  // called before the memory of the current batch is reused: anything that still points into it must be consumed now
  virtual void release_batch() {}
  // Opt-in for operators that need per-site summaries only (merged alleles, counts over the live calls: n_live and
  // GDBGPU_SCAN_SITE_STATS): no per-call data crosses PCIe, operate_on_site_summary replaces operate
  virtual bool wants_site_summaries() const { return false; }
  virtual void operate_on_site_summary(const gdbgpu_batch&, uint32_t /*site*/, const VariantQueryConfig&) {}
 protected:
  bool m_NON_REF_exists = false, m_remapping_needed = false, m_is_reference_block_only = false;
  std::string m_merged_reference_allele;
  std::vector<std::string> m_merged_alt_alleles;
};
// variant_operations.h:391-408, variant_operations.cc:380-400: "position valid-calls reference-block-calls calls-beginning-here"
class InterestingLocationsPrinter : public SingleVariantOperatorBase {
 public:
  explicit InterestingLocationsPrinter(std::ostream& fptr) : m_fptr(&fptr) {}
  void operate(Variant& variant, const VariantQueryConfig& query_config) override;
  bool wants_site_summaries() const override { return true; }
  void operate_on_site_summary(const gdbgpu_batch& b, uint32_t s, const VariantQueryConfig&) override;
 protected:
  std::ostream* m_fptr;
};
// variant_operations.h:410-475: the top_count sites with the most merged ALT alleles
class MaxAllelesCountOperator : public SingleVariantOperatorBase {
 public:
  explicit MaxAllelesCountOperator(unsigned top_count = 25u) : m_top_count(top_count) {}
  bool wants_site_summaries() const override { return true; }
  void operate_on_site_summary(const gdbgpu_batch& b, uint32_t s, const VariantQueryConfig&) override;
  void print(std::ostream& fptr) const;   // what the reference's destructor writes to std::cerr
  uint64_t get_total_lines() const { return m_total_lines; }
 private:
  struct AlleleTracker { std::string ref; std::vector<std::string> alt; int64_t column; };
  unsigned m_top_count;
  std::vector<AlleleTracker> m_heap;   // min-heap on the ALT count (std::priority_queue with AlleleTrackerCompare)
  uint64_t m_total_lines = 0;
};
class GA4GHOperator : public SingleVariantOperatorBase {
 public:
  void operate(Variant& variant, const VariantQueryConfig& query_config) override;
  const Variant& get_remapped_variant() const { return *m_remapped_variant; }
  bool too_many_alt_alleles_for_genotype_length_fields(unsigned num_alt_alleles) const { return num_alt_alleles > m_max_diploid_alt_alleles_that_can_be_genotyped; }
 protected:
  void warn_too_many_alleles(const VariantQueryConfig& query_config, int64_t column, unsigned num_alt) const;
  const Variant* m_remapped_variant = nullptr;
  unsigned m_max_diploid_alt_alleles_that_can_be_genotyped = 50;
};

// ---- what BroadCombinedGVCFOperator takes from the reference's VCFAdapter (include/vcf/vcf_adapter.h:84-190): the BCF header's
// dictionary (bcf_hdr_id2int), the reference genome (get_reference_base_at_position) and the output buffer whose fill level
// drives overflow() (VCFSerializedBufferAdapter, vcf_adapter.h:190-260).  The header itself is produced by the unchanged
// host code (htslib); here it is read back from its serialised form.
class VCFHeaderLines;   // gdb_vcf_header.h
class IndexedFasta;
class VCFAdapter {
 public:
  explicit VCFAdapter(bool keep_idx_fields_in_bcf_header = true);
  virtual ~VCFAdapter();
  VCFAdapter(const VCFAdapter&) = delete;
  VCFAdapter& operator=(const VCFAdapter&) = delete;
  // vcf_adapter.cc:309-380: the template VCF header (vcf_header_filename of the configuration; none: the default header) and
  // the reference genome (an indexed FASTA: <file>.fai).  BroadCombinedGVCFOperator's constructor then adds the queried
  // fields, the contigs of the vid mapping and the samples of the query, and print_header() puts the serialised BCF header
  // at the head of the output -- header generation needs nothing from outside this library.
  void initialize(const VariantQueryConfig& config);
  bool is_initialized_from_config() const { return lines_ != nullptr; }
  VCFHeaderLines* header_lines() { return lines_.get(); }
  // takes the header as the lines stand, sets the dictionary used by the records
  void finalize_header(bool with_samples);
  virtual void print_header();   // "BCF\2\2", l_text, text
  // `header` = the BCF header text (bcf_hdr_t as htslib writes it: ##FILTER / ##INFO / ##FORMAT / ##contig lines, #CHROM line)
  void set_vcf_header_text(const std::string& header);
  const std::string& get_vcf_header_text() const { return header_; }
  int get_dictionary_id(const std::string& name) const;   // -1: not in the header
  bool has_info_line(const std::string& name) const { return info_ids_.count(name) != 0; }   // the header declares ##INFO=<ID=name,...>
  int get_contig_id(const std::string& name) const;
  void set_reference_base(const std::string& contig, int64_t pos, char base) { ref_bases_[contig + ":" + std::to_string(pos)] = base; }
  char get_reference_base_at_position(const std::string& contig, int64_t pos) const;   // 'N' when unknown
  void set_buffer_capacity(size_t bytes) { capacity_ = bytes; }
  std::vector<uint8_t>& get_buffer() { return buffer_; }
  bool overflow() const { return capacity_ && buffer_.size() + pending_bytes_ >= capacity_; }
  // Zero-copy output: with a file descriptor set, records are kept as a list of byte ranges -- the few header bytes of
  // every record and field in small owned chunks, the FORMAT blocks as ranges of the batch's own (pinned) memory -- and
  // written with writev(); nothing of the 12 MB-per-site blocks is copied on the host.  Borrowed ranges live as long as the
  // batch: scan_and_operate flushes them before it asks for the next one.  Without a descriptor everything is appended to
  // get_buffer() (a copy), as VCFSerializedBufferAdapter does.
  void set_output_fd(int fd) { fd_ = fd; }
  bool zero_copy() const { return fd_ >= 0; }
  void append(const void* p, size_t n, bool borrowed);
  void flush();            // write the pending ranges (zero-copy mode); no-op otherwise
  uint64_t bytes_written() const { return bytes_written_; }
 private:
  struct Range { const uint8_t* p; size_t n; };
  std::vector<Range> pending_;
  std::vector<std::unique_ptr<uint8_t[]>> chunks_; size_t chunk_used_ = 0, chunk_cap_ = 0;
  size_t pending_bytes_ = 0; uint64_t bytes_written_ = 0; int fd_ = -1;
  std::string header_;
  std::vector<std::pair<std::string, int>> ids_, contigs_;
  std::set<std::string> info_ids_;   // IDs with an ##INFO line
  void collect_info_ids();
  std::map<std::string, char> ref_bases_;
  std::unique_ptr<VCFHeaderLines> lines_;
  bool keep_idx_ = true;
  mutable std::unique_ptr<IndexedFasta> fasta_;   // the reference genome (ReferenceGenomeInfo, vcf_adapter.cc:31-58)
  bool header_printed_ = false;
  std::vector<uint8_t> buffer_;
  size_t capacity_ = 0;
};

// vcf_adapter.h:190-260: the adapter the command line tool uses -- records accumulate in an RWBuffer the caller owns and
// do_output() writes them to the output file
struct RWBuffer {
  std::vector<uint8_t> m_buffer;
  size_t m_num_valid_bytes = 0, m_next_read_idx = 0;
};
class VCFSerializedBufferAdapter : public VCFAdapter {
 public:
  explicit VCFSerializedBufferAdapter(bool keep_idx_fields_in_bcf_header = true, bool do_output = true) : VCFAdapter(keep_idx_fields_in_bcf_header), m_do_output(do_output) {}
  void set_buffer(RWBuffer& buffer) { m_rw_buffer = &buffer; }
  void set_output_file(FILE* f) { m_output_fptr = f; }
  void do_output();    // writes what scan_and_operate left behind (and forgets it); the caller resets m_num_valid_bytes
 private:
  RWBuffer* m_rw_buffer = nullptr; FILE* m_output_fptr = stdout; bool m_do_output;
};
// include/utils/timer.h: wall + cpu seconds, printed as the reference prints them (GENOMICSDB_TIMER lines)
class Timer {
 public:
  Timer() { start(); }
  void start();
  void stop();
  void print(const std::string& prefix = "", std::ostream& fptr = std::cerr) const;
 private:
  double wall0_ = 0, cpu0_ = 0, wall_ = 0, cpu_ = 0, cum_wall_ = 0, cum_cpu_ = 0;
};
enum ProduceBroadGVCFSubOperation { PRODUCE_BROAD_GVCF_UNKNOWN = 0, PRODUCE_BROAD_GVCF_PRODUCE_GVCF, PRODUCE_BROAD_GVCF_PRODUCE_INTERESTING_POSITIONS };

// broad_combined_gvcf.h: one BCF2 record per operate() call, appended to the adapter's buffer in serialised form
// (u32 l_shared, u32 l_indiv, shared, indiv -- what bcf_write emits for an uncompressed BCF).  FORMAT blocks, INFO scalars
// and the DP sum come from the kernels (GDBGPU_SCAN_BCF_FORMAT); this class adds the typed keys / size descriptors and the
// fixed part of the record.  Not supported: 2-D fields and the histogram / concatenate INFO operations on them (throws).
class BroadCombinedGVCFOperator : public GA4GHOperator {
 public:
  // use_missing_values_only_not_vector_end: what GenomicsDBBCFGenerator passes for the Java reader (genomicsdb_bcf_generator.cc:60-83)
  BroadCombinedGVCFOperator(VCFAdapter& vcf_adapter, const VidMapper& id_mapper, const VariantQueryConfig& query_config,
                            bool use_missing_values_only_not_vector_end = false);
  // the reference's own signature (broad_combined_gvcf.h:80-83)
  BroadCombinedGVCFOperator(VCFAdapter& vcf_adapter, const VidMapper& id_mapper, const VariantQueryConfig& query_config,
                            const unsigned max_diploid_alt_alleles_that_can_be_genotyped, const bool use_missing_values_only_not_vector_end = false)
      : BroadCombinedGVCFOperator(vcf_adapter, id_mapper, query_config, use_missing_values_only_not_vector_end) {
    m_max_diploid_alt_alleles_that_can_be_genotyped = max_diploid_alt_alleles_that_can_be_genotyped;
  }
  bool wants_bcf_blocks() const override { return true; }
  bool wants_htsjdk_conventions() const override { return m_use_missing_values_not_vector_end; }
  void operate_on_bcf_site(const gdbgpu_batch& batch, uint32_t site, const VariantQueryConfig& query_config) override;
  bool overflow() const override { return m_vcf_adapter->overflow(); }
  void release_batch() override { m_vcf_adapter->flush(); }
  uint64_t get_num_records() const { return m_num_records; }
  // broad_combined_gvcf.cc:168-330: queried fields, contigs and samples into the adapter's template header; header printed
  static void build_header(VCFAdapter& vcf_adapter, const VidMapper& id_mapper, const VariantQueryConfig& query_config);
 private:
  VCFAdapter* m_vcf_adapter;
  const VidMapper* m_vid_mapper;
  struct FormatField { std::string vcf_name; int key = -1; int fmt_index = -1; };   // fmt_index: index into the batch's fmt_* tables
  std::vector<FormatField> m_FORMAT_fields_vec;   // emission order: GT first, query order, DP last (broad_combined_gvcf.cc:168-240)
  struct InfoField { std::string vcf_name; int key = -1; int plain_index = -1; bool is_float = false; int vector_slot = -1; int concat_fmt_index = -1;
                     int raw_fmt_index = -1; const FieldInfo* fi = nullptr; };   // vector_slot >= 0: element-wise sum; concat_fmt_index >= 0: concatenate; raw_fmt_index >= 0: element-wise sum of a 2-D field
  std::vector<InfoField> m_INFO_fields_vec;
  // histogram_sum fields (broad_combined_gvcf.cc:194-217,544-585): the two flattened elements of a tuple field, in the order of the
  // parents' vid mapping indices (the reference walks a std::map keyed by them) -- after every other INFO field
  struct HistogramField { int parent_idx = 0; int key = -1; int raw_fmt_index[2] = {-1, -1}; const FieldInfo* fi[2] = {nullptr, nullptr}; };
  std::vector<HistogramField> m_INFO_histogram_fields;
  int m_raw_REF_fmt_index = -1, m_raw_ALT_fmt_index = -1;   // the live calls' own REF / ALT strings (rows of a block), for the allele remap of 2-D fields
  // one call's 2-D value cut into its entries (remapped onto the merged alleles where the reference remaps it); false: the call has no value
  bool multid_entries(const gdbgpu_batch& b, uint32_t s, uint64_t q, int raw_fmt_index, const FieldInfo& fi, std::vector<std::pair<const uint8_t*, uint64_t>>& entries) const;
  int m_END_key = -1, m_DP_key = -1, m_QUAL_plain = -1;
  int m_ID_fmt_index = -1;        // the ID strings of the live calls, collated like a FORMAT field; merged per site (merge_ID_field)
  std::string m_ID_value;
  int m_FILTER_fmt_index = -1;    // the FILTER vectors of the live calls, collated like an int FORMAT field; united per site
  std::vector<int> m_global_field_idx_to_hdr_idx, m_FILTER_idx_vec;
  unsigned m_num_vector_slots = 0;
  bool m_dp_rule = false, m_use_missing_values_not_vector_end = false;
  uint64_t m_num_records = 0;
};

// ---- storage + engine
// ---- the columnar iterator surface (include/genomicsdb/variant_cell.h GenomicsDBColumnarCell, genomicsdb_iterators.h,
// include/query_operations/variant_operations.h:675-768).  The cells come from gdbgpu_scan_cells* (the device builds the
// list in the reference iterator's order and gathers the queried columns); histogram and allele counts run as kernels.
struct VariantArraySchema {};   // the operators' third argument; nothing on this path reads it
class GenomicsDBColumnarCell {
 public:
  const int64_t* get_coordinates() const { return coords_; }   // row, begin column
  int64_t get_end_position() const { return end_; }   // the cell's END
  bool at_new_query_column_interval() const { return at_new_interval_; }
  uint64_t get_current_query_column_interval_idx() const { return interval_idx_; }
  bool is_valid(unsigned query_idx) const;
  size_t get_field_length(unsigned query_idx) const;            // elements
  const uint8_t* get_field_ptr_for_query_idx(unsigned query_idx) const;
  template <class T> const T* get_field_ptr_for_query_idx(unsigned query_idx) const { return reinterpret_cast<const T*>(get_field_ptr_for_query_idx(query_idx)); }
  void print(std::ostream& fptr, const VariantQueryConfig* query_config, const std::string& indent_prefix, const VidMapper* vid_mapper = nullptr) const;
  void print_csv(std::ostream& fptr, const VariantQueryConfig* query_config) const;
 private:
  friend class VariantQueryProcessor;
  struct Column { const uint64_t* off = nullptr; const uint8_t* bytes = nullptr; };
  const VariantQueryConfig* qc_ = nullptr;
  const std::vector<Column>* columns_ = nullptr;   // one per query attribute
  uint64_t idx_ = 0;                               // position in the cell list
  int64_t coords_[2] = {0, 0}, end_ = 0;
  bool at_new_interval_ = false; uint64_t interval_idx_ = 0;
};

class SingleCellOperatorBase {
 public:
  virtual ~SingleCellOperatorBase() = default;
  virtual void operate_on_columnar_cell(const GenomicsDBColumnarCell&, const VariantQueryConfig&, const VariantArraySchema&) {
    throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Sub-classes should override operate_on_columnar_cell()");
  }
  virtual void finalize() {}
  // operators whose whole work is a kernel over the device's cell list take the interval's scan instead of its cells
  virtual bool runs_on_device() const { return false; }
  virtual void operate_on_device_interval(gdbgpu_scan*, unsigned /*column_interval_idx*/, const VariantQueryConfig&) {}
};

class ColumnHistogramOperator : public SingleCellOperatorBase {
 public:
  ColumnHistogramOperator(uint64_t begin, uint64_t end, uint64_t bin_size);
  void operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig&, const VariantArraySchema&) override;
  bool runs_on_device() const override { return true; }
  void operate_on_device_interval(gdbgpu_scan* scan, unsigned, const VariantQueryConfig&) override;
  bool equi_partition_and_print_bins(uint64_t num_bins, std::ostream& fptr = std::cout) const;
  const std::vector<uint64_t>& get_bin_counts() const { return m_bin_counts_vector; }
 private:
  std::vector<uint64_t> m_bin_counts_vector, m_scratch;
  uint64_t m_begin_column, m_end_column, m_bin_size;
};

class VariantCallPrintOperator : public SingleCellOperatorBase {
 public:
  VariantCallPrintOperator(std::ostream& fptr = std::cout, const std::string& indent_prefix = "", const VidMapper* vid_mapper = nullptr);
  void operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig& query_config, const VariantArraySchema&) override;
  void finalize() override;
 private:
  uint64_t m_num_calls_printed = 0, m_num_query_intervals_printed = 0;
  std::string m_indent_prefix, m_indent_prefix_plus_one, m_indent_prefix_plus_two;
  std::ostream* m_fptr;
  const VidMapper* m_vid_mapper;
};

class VariantCallPrintCSVOperator : public SingleCellOperatorBase {
 public:
  explicit VariantCallPrintCSVOperator(std::ostream& fptr = std::cout) : m_fptr(&fptr) {}
  void operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig& query_config, const VariantArraySchema&) override { cell.print_csv(*m_fptr, &query_config); }
 private:
  std::ostream* m_fptr;
};

class AlleleCountOperator : public SingleCellOperatorBase {
 public:
  AlleleCountOperator(const VidMapper& vid_mapper, const VariantQueryConfig& query_config);
  bool runs_on_device() const override { return true; }
  void operate_on_device_interval(gdbgpu_scan* scan, unsigned, const VariantQueryConfig&) override;
  void print_allele_counts(std::ostream& fptr = std::cout) const;
  struct Record { int64_t column; std::string REF, ALT; uint64_t count; };
  const std::vector<Record>& get_records() const { return m_records; }   // per query interval, then (column, REF, ALT)
 private:
  std::vector<Record> m_records;
};

class VariantStorageManager {
 public:
  VariantStorageManager(const std::string& workspace, unsigned segment_size = 10u * 1024u * 1024u, int cuda_device = 0);
  ~VariantStorageManager();
  VariantStorageManager(const VariantStorageManager&) = delete;
  VariantStorageManager& operator=(const VariantStorageManager&) = delete;
  const std::string& get_workspace() const { return workspace_; }
  gdbgpu_ctx* ctx() const { return ctx_; }
 private:
  std::string workspace_;
  gdbgpu_ctx* ctx_ = nullptr;
};

// query_variants.h:126-191
// ---- GA4GH range query results (include/genomicsdb/variant.h Variant / VariantCall as gt_get_column_interval fills them).  Owning
// objects: the calls come from the columnar iterator's cell list, a field is kept as the text Variant::print gives it.
struct GA4GHVariantCall {
  uint64_t row = 0; int64_t begin = 0, end = 0;
  std::string REF; std::vector<std::string> ALT;     // as stored (<NON_REF> is "&")
  std::vector<std::string> fields;                   // by query attribute index; empty: the field is not valid for this call
  void print(std::ostream& fptr, const VariantQueryConfig* query_config, const std::string& indent_prefix, const VidMapper* vid_mapper = nullptr) const;
};
struct GA4GHVariant {
  int64_t begin = 0, end = 0;
  std::vector<GA4GHVariantCall> calls;
  bool has_common_fields = false; std::string common_REF; std::vector<std::string> common_ALT;   // set by GA4GHOperator for variants of several calls
  uint64_t get_num_calls() const { return calls.size(); }
  uint64_t get_column_begin() const { return (uint64_t)begin; }
  uint64_t get_column_end() const { return (uint64_t)end; }
  void print(std::ostream& fptr, const VariantQueryConfig* query_config, const std::string& indent_prefix, const VidMapper* vid_mapper = nullptr) const;   // variant.cc:427-478
};
// variant.cc:942-1020, the default output format: { "variants": [ ... ] }
void print_variants(const std::vector<GA4GHVariant>& variants, const std::string& output_format, const VariantQueryConfig& query_config, std::ostream& fptr,
                    const VidMapper* id_mapper = nullptr);

class VariantQueryProcessorScanState {
 public:
  ~VariantQueryProcessorScanState() { reset(); }
  bool end() const { return m_done; }
  void reset();
  uint64_t get_num_sites_done() const { return sites_done_; }
 private:
  friend class VariantQueryProcessor;
  bool m_done = false;
  gdbgpu_scan* scan_ = nullptr;
  gdbgpu_batch batch_{};
  bool have_batch_ = false;
  uint32_t pos_in_batch_ = 0;
  uint64_t sites_done_ = 0;
};

class VariantQueryProcessor {
 public:
  VariantQueryProcessor(VariantStorageManager* storage_manager, const std::string& array_name, const VidMapper& vid_mapper);
  ~VariantQueryProcessor();
  // Adds END/REF/ALT/GT, resolves the allele-dependent fields among the queried attributes, stages the array in HBM.
  // stage_broad_gvcf_fields: also stage what BroadCombinedGVCFOperator combines besides the remapped fields (plain FORMAT
  // fields, scalar INFO fields, the DP rule's inputs)
  // stage_columnar_fields: stage EVERY queried attribute (what the columnar iterator's operators may print) and fix the
  // query attribute order as the reference does (END, REF, ALT first)
  void do_query_bookkeeping(VariantQueryConfig& query_config, bool alleles_required = true, bool stage_broad_gvcf_fields = false,
                            bool stage_columnar_fields = false);
  // the reference's own signature (query_variants.h:216-217): everything an operator of the combine path may want is staged
  void do_query_bookkeeping(const VariantArraySchema&, VariantQueryConfig& query_config, const VidMapper&, const bool alleles_required) {
    do_query_bookkeeping(query_config, alleles_required, /*stage_broad_gvcf_fields=*/true);
  }
  // query_variants.cc:557-576: every cell of every query interval (or of the array) goes to the operator, then finalize()
  void iterate_over_cells(const int ad, const VariantQueryConfig& query_config, SingleCellOperatorBase& variant_operator,
                          const bool use_common_array_object = true) const;
  const VariantArraySchema& get_array_schema() const { return schema_; }
  int get_array_descriptor() const { return 0; }
  void scan_and_operate(const int ad, const VariantQueryConfig& query_config, SingleVariantOperatorBase& variant_operator,
                        unsigned column_interval_idx = 0, bool handle_spanning_deletions = false,
                        VariantQueryProcessorScanState* scan_state = nullptr, bool want_alleles_LUT = false) const;
  gdbgpu_array* array() const { return arr_; }
  // query_variants.cc:687-843 (the GA4GH range query): the calls of one query interval -- those reaching into it
  // from the left, then those beginning inside it -- grouped into variants of calls that share begin, end, REF and the SET of
  // ALT alleles (GA4GHCallInfoToVariantIdx, variant.cc:26-70); a variant of several calls goes through GA4GHOperator (merged
  // REF / ALT as its common fields, GT and the allele-dependent fields re-indexed onto the merged alleles).  Appends to `variants`.
  // Needs do_query_bookkeeping(config, true, false, /*stage_columnar_fields=*/true).
  void gt_get_column_interval(const int ad, const VariantQueryConfig& query_config, unsigned column_interval_idx, std::vector<GA4GHVariant>& variants) const;
 private:
  void check(int rc) const;
  void iterate_over_cells_of_interval(const VariantQueryConfig& query_config, SingleCellOperatorBase& variant_operator, unsigned column_interval_idx) const;
  void remap_ga4gh_variant(const VariantQueryConfig& query_config, GA4GHVariant& variant) const;
  void fill_variant(const gdbgpu_batch& b, uint32_t s, Variant& v) const;
  VariantStorageManager* sm_;
  std::string array_name_;
  VidMapper vid_;
  gdbgpu_array* arr_ = nullptr;
  std::vector<int64_t> coords_, end_;   // merged cell coordinates (row, col) and END, host copy
  int n_out_ = 0;
  VariantArraySchema schema_;
};

// ---- the JNI path's pump (include/vcf/genomicsdb_bcf_generator.h, src/vcf/genomicsdb_bcf_generator.cc:26-185): what
// GenomicsDBQueryStream's six native methods sit on.  It owns storage manager, query processor, adapter and the Broad
// operator, runs scan_and_operate until the adapter's buffer reports overflow, and hands the serialised BCF stream
// (header first) out through read_and_advance / read_next_byte.  The BCF header is htslib's product on the unchanged host
// side; it comes in serialised ("BCF\2\2", l_text, text) -- the one argument the reference's constructor does not have.
class GenomicsDBBCFGenerator {
 public:
  // the reference's own signature (genomicsdb_bcf_generator.h:58-64): the header is generated here (template header of the
  // configuration + queried fields + contigs + samples), the reference bases come from the configuration's indexed FASTA
  GenomicsDBBCFGenerator(const std::string& loader_config_file, const std::string& query_config_file,
                         const char* chr, const int start, const int end, int my_rank = 0, size_t buffer_capacity = 1048576u,
                         size_t tiledb_segment_size = 1048576u, const char* output_format = "bu", const bool produce_header_only = false,
                         const bool use_missing_values_only_not_vector_end = false, const bool keep_idx_fields_in_bcf_header = true);
  GenomicsDBBCFGenerator(const std::string& loader_config_file, const std::string& query_config_file, const std::string& serialized_bcf_header,
                         const char* chr, const int start, const int end, int my_rank = 0, size_t buffer_capacity = 1048576u,
                         size_t tiledb_segment_size = 1048576u, const char* output_format = "bu", const bool produce_header_only = false,
                         const bool use_missing_values_only_not_vector_end = false, const bool keep_idx_fields_in_bcf_header = true,
                         int cuda_device = 0, const std::string& reference_bases_tsv = "");   // contig, 0-based position, base per line: the bases
                                                                                               // VCFAdapter::get_reference_base_at_position serves
  ~GenomicsDBBCFGenerator();
  GenomicsDBBCFGenerator(const GenomicsDBBCFGenerator&) = delete;
  GenomicsDBBCFGenerator& operator=(const GenomicsDBBCFGenerator&) = delete;
  size_t get_buffer_capacity() const { return m_buffer_capacity; }
  // n == SIZE_MAX: produce the next batch without advancing; dst may be NULL (skip)
  size_t read_and_advance(uint8_t* dst, size_t offset, size_t n);
  uint8_t read_next_byte();
  size_t get_num_bytes_available() const { return m_read_buffer.size() - m_next_read_idx; }   // in the batch being read
  bool end() const { return m_done; }
 private:
  void produce_next_batch();
  void construct(const std::string& loader_config_file, const std::string& query_config_file, const std::string* serialized_bcf_header, const char* chr, int start, int end,
                 int my_rank, const char* output_format, bool use_missing_values_only_not_vector_end, int cuda_device, const std::string& reference_bases_tsv);
  bool m_done = false, m_produce_header_only;
  size_t m_buffer_capacity;
  std::unique_ptr<VariantStorageManager> m_storage_manager;
  std::unique_ptr<VariantQueryProcessor> m_query_processor;
  VariantQueryConfig m_query_config;
  VCFAdapter m_vcf_adapter;
  unsigned m_query_column_interval_idx = 0;
  VariantQueryProcessorScanState m_scan_state;
  std::unique_ptr<BroadCombinedGVCFOperator> m_combined_bcf_operator;
  std::vector<uint8_t> m_read_buffer; size_t m_next_read_idx = 0;   // the batch being read (the adapter's buffer is the one being written)
};

}  // namespace genomicsdb_b200
