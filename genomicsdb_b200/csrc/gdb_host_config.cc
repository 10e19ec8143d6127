// Host side of the boundary that needs no kernel: the loader configuration and what a query takes from it, the callset
// mapping, the VCF / BCF header (template + queried fields + contigs + samples, gdb_vcf_header.*), the reference genome, the
// adapters of the command-line tool, and the C ABI of the Java reader's query stream.
//   loader JSON            src/main/cpp/src/config/json_config.cc:207-470,553-700,718-800
//   update_from_loader     src/main/cpp/src/config/genomicsdb_config_base.cc:179-223
//   callsets               src/main/cpp/src/utils/vid_mapper.cc:319-325,686-711
//   header                 src/main/cpp/src/vcf/vcf_adapter.cc:59-248,309-400, src/query_operations/broad_combined_gvcf.cc:158-330
//   query stream           src/main/jni/src/genomicsdb_GenomicsDBQueryStream.cc:29-111, src/vcf/genomicsdb_bcf_generator.cc
#include <sys/resource.h>
#include <sys/time.h>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>

#include "gdb_host_api.h"
#include "gdb_json.h"
#include "gdb_vcf_header.h"

namespace genomicsdb_b200 {

using gdb::JsonValue;

// ------------------------------------------------------------------------------------------------ callsets
void VidMapper::read_callsets_file(const std::string& file) {
  JsonValue d = gdb::json_parse_file(file);
  const JsonValue* cs = d.find("callsets");
  if (!cs) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "no \"callsets\" in " + file);
  callset_names_.clear();
  auto add = [&](const std::string& name, const JsonValue& v, int64_t dflt) {
    const JsonValue* r = v.find("row_idx");
    const int64_t row = r ? r->as_int() : dflt;
    if (row < 0) return;
    if ((size_t)row >= callset_names_.size()) callset_names_.resize(row + 1);
    callset_names_[row] = name;
  };
  int64_t i = 0;
  if (cs->type == JsonValue::OBJ) for (auto& kv : cs->obj) add(kv.first, kv.second, i++);
  else for (auto& v : cs->arr) add(v.has("sample_name") ? v.at("sample_name").str : v.at("name").str, v, i++);
}
bool VidMapper::get_callset_name(int64_t row_idx, std::string& callset_name) const {
  if (row_idx < 0 || (size_t)row_idx >= callset_names_.size()) return false;
  callset_name = callset_names_[row_idx];
  return true;
}

// ------------------------------------------------------------------------------------------------ loader configuration
namespace {
std::string per_rank_string(const JsonValue& v, int rank) {
  if (v.type == JsonValue::ARR) return v.arr.empty() ? std::string() : v.arr[std::min<size_t>((size_t)rank, v.arr.size() - 1)].str;
  return v.str;
}
// { "chr" : pos } / { "chr" : [ b, e ] } (json_config.cc extract_contig_interval_from_object)
std::pair<int64_t, int64_t> contig_interval(const JsonValue& o, const VidMapper& vid) {
  if (o.type != JsonValue::OBJ || o.obj.size() != 1) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "a contig position must be { \"<contig>\" : <position> }");
  const ContigInfo* c = vid.get_contig_info(o.obj[0].first);
  if (!c) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "unknown contig " + o.obj[0].first);
  const JsonValue& v = o.obj[0].second;
  if (v.type == JsonValue::ARR && v.arr.size() == 2) return {c->tiledb_column_offset + v.arr[0].as_int() - 1, c->tiledb_column_offset + v.arr[1].as_int() - 1};
  if (v.type == JsonValue::NUM) return {c->tiledb_column_offset + v.as_int() - 1, c->tiledb_column_offset + v.as_int() - 1};
  return {c->tiledb_column_offset, c->tiledb_column_offset + c->length - 1};
}
}  // namespace

void GenomicsDBImportConfig::read_from_file(const std::string& filename, const int rank) {
  JsonValue d = gdb::json_parse_file(filename);
  row_based_ = d.has("row_based_partitioning") && d.at("row_based_partitioning").as_bool();
  if (const JsonValue* v = d.find("vid_mapping_file")) vid_ = VidMapper(per_rank_string(*v, rank));
  if (const JsonValue* v = d.find("callset_mapping_file")) { callset_mapping_file_ = per_rank_string(*v, rank); if (vid_.is_initialized() && !callset_mapping_file_.empty()) vid_.read_callsets_file(callset_mapping_file_); }
  if (const JsonValue* v = d.find("vcf_header_filename")) vcf_header_filename_ = per_rank_string(*v, rank);
  if (const JsonValue* v = d.find("reference_genome")) reference_genome_ = per_rank_string(*v, rank);
  if (const JsonValue* v = d.find("segment_size")) segment_size_ = (size_t)v->as_int();
  workspaces_.clear(); array_names_.clear();
  if (const JsonValue* w = d.find("workspace")) { if (w->type == JsonValue::ARR) for (auto& x : w->arr) workspaces_.push_back(x.str); else workspaces_.push_back(w->str); }
  const JsonValue* an = d.find("array"); if (!an) an = d.find("array_name");
  if (an) { if (an->type == JsonValue::ARR) for (auto& x : an->arr) array_names_.push_back(x.str); else array_names_.push_back(an->str); }
  partitions_.clear(); partitions_specified_ = false;
  if (row_based_) { if (!d.has("row_partitions")) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "row based partitioning without \"row_partitions\""); return; }
  const JsonValue* cp = d.find("column_partitions");
  if (!cp || cp->type != JsonValue::ARR) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "the loader configuration has no \"column_partitions\"");
  partitions_specified_ = true;
  const std::string ws0 = workspaces_.size() == 1 ? workspaces_[0] : "", an0 = array_names_.size() == 1 ? array_names_[0] : "";
  const size_t n = cp->arr.size();
  partitions_.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const JsonValue& p = cp->arr[i];
    if (!p.has("begin")) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "a column partition needs \"begin\"");
    const JsonValue& b = p.at("begin");
    partitions_[i].first = b.type == JsonValue::NUM ? b.as_int() : contig_interval(b, vid_).first;
    partitions_[i].second = INT64_MAX - 1;
    if (const JsonValue* e = p.find("end")) partitions_[i].second = e->type == JsonValue::NUM ? e->as_int() : contig_interval(*e, vid_).first;
    if (partitions_[i].first > partitions_[i].second) std::swap(partitions_[i].first, partitions_[i].second);
    if (const JsonValue* w = p.find("workspace")) { if (workspaces_.size() < n) workspaces_.resize(n, ws0); workspaces_[i] = w->str; }
    const JsonValue* a = p.find("array"); if (!a) a = p.find("array_name");
    if (a) { if (array_names_.size() < n) array_names_.resize(n, an0); array_names_[i] = a->str; }
  }
  // a partition ends where the next one (in column order) begins
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return partitions_[x].first < partitions_[y].first; });
  for (size_t k = 0; k + 1 < n; ++k) {
    auto& cur = partitions_[order[k]]; const auto& nxt = partitions_[order[k + 1]];
    if (cur.first == nxt.first) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Cannot have two column partitions with the same begin value");
    if (cur.second >= nxt.first) cur.second = nxt.first - 1;
  }
}
std::pair<int64_t, int64_t> GenomicsDBImportConfig::get_column_partition(const int rank, const unsigned idx) const {
  (void)idx;
  if (!partitions_specified_ || partitions_.empty()) return {0, INT64_MAX - 1};
  return partitions_[std::min<size_t>((size_t)rank, partitions_.size() - 1)];
}
const std::string& GenomicsDBImportConfig::get_workspace(const int rank) const {
  static const std::string none;
  if (workspaces_.empty()) return none;
  return workspaces_.size() == 1 ? workspaces_[0] : workspaces_[std::min<size_t>((size_t)rank, workspaces_.size() - 1)];
}
const std::string& GenomicsDBImportConfig::get_array_name(const int rank) const {
  static const std::string none;
  if (array_names_.empty()) return none;
  return array_names_.size() == 1 ? array_names_[0] : array_names_[std::min<size_t>((size_t)rank, array_names_.size() - 1)];
}

// genomicsdb_config_base.cc:179-201: called BEFORE the query's own file is read, which then overrides what it names itself
void VariantQueryConfig::update_from_loader(const GenomicsDBImportConfig& loader_config, const int rank) {
  if (workspace_.empty()) workspace_ = loader_config.get_workspace(rank);
  if (array_name_.empty()) array_name_ = loader_config.get_array_name(rank);
  if (!vid_.is_initialized()) vid_ = loader_config.get_vid_mapper();
  if (vcf_header_filename_.empty()) vcf_header_filename_ = loader_config.get_vcf_header_filename();
  if (reference_genome_.empty()) reference_genome_ = loader_config.get_reference_genome();
  if (callset_mapping_file_.empty()) callset_mapping_file_ = loader_config.get_callset_mapping_file();
}
// genomicsdb_config_base.cc:203-223: of the query's intervals a rank keeps those that touch its column partition
void VariantQueryConfig::subset_query_column_ranges_based_on_partition(const GenomicsDBImportConfig& loader_config, const int rank) {
  if (!loader_config.is_partitioned_by_column()) return;
  const auto mine = loader_config.get_column_partition(rank);
  std::vector<std::pair<int64_t, int64_t>> kept;
  for (auto& q : intervals_) if (q.second >= mine.first && q.first <= mine.second) kept.push_back(q);
  intervals_.swap(kept);
}

// ------------------------------------------------------------------------------------------------ VCFAdapter: header + genome
VCFAdapter::VCFAdapter(bool keep_idx_fields_in_bcf_header) : keep_idx_(keep_idx_fields_in_bcf_header) {}
VCFAdapter::~VCFAdapter() {}

void VCFAdapter::initialize(const VariantQueryConfig& config) {
  if (config.get_vcf_header_filename().empty()) {   // vcf_adapter.cc:381-388
    lines_.reset(new VCFHeaderLines(true));
    lines_->append("##ALT=<ID=NON_REF,Description=\"Represents any possible alternative allele at this location\">");
    lines_->append("##INFO=<ID=END,Number=1,Type=Integer,Description=\"Stop position of the interval\">");
  } else {
    std::ifstream f(config.get_vcf_header_filename());
    if (!f) throw GenomicsDBGPUException(GDBGPU_ERR_IO, "Could not open template VCF header file " + config.get_vcf_header_filename());
    std::stringstream ss; ss << f.rdbuf();
    lines_.reset(new VCFHeaderLines(false));
    lines_->parse(ss.str(), /*sample_line_required=*/false);
  }
  const std::string& fmt = config.get_vcf_output_format();
  if (!(fmt.empty() || fmt == "bu")) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "only uncompressed BCF (\"bu\") is produced on this path");
  if (!config.get_reference_genome().empty()) {
    fasta_.reset(new IndexedFasta);
    try { fasta_->open(config.get_reference_genome()); } catch (const std::exception& e) { throw GenomicsDBGPUException(GDBGPU_ERR_IO, e.what()); }
  }
  header_printed_ = false;
}

// The reference serialises the header without IDX fields and reads it back (broad_combined_gvcf.cc:314-325: "clean up the
// header completely"), so dictionary indices are those of a fresh parse of the text.
void VCFAdapter::finalize_header(bool) {
  std::unique_ptr<VCFHeaderLines> fresh(new VCFHeaderLines(false));
  fresh->parse(lines_->format(false), false);
  lines_.swap(fresh);
  header_ = lines_->format(true);
  ids_ = lines_->ids();
  contigs_ = lines_->contigs();
  collect_info_ids();
}
void VCFAdapter::collect_info_ids() {
  info_ids_.clear();
  for (size_t pos = 0; (pos = header_.find("##INFO=<ID=", pos)) != std::string::npos;) {
    pos += 11;
    const size_t e = header_.find_first_of(",>", pos);
    info_ids_.insert(header_.substr(pos, e == std::string::npos ? std::string::npos : e - pos));
  }
}

// bcf_hdr_write / bcf_hdr_serialize (vcf.c:986-1051): magic, l_text counting the closing NUL, text
void VCFAdapter::print_header() {
  const std::string text = lines_ ? lines_->format(keep_idx_) : header_;
  const uint32_t l_text = (uint32_t)text.size() + 1u;
  std::vector<uint8_t> h(9 + l_text);
  memcpy(h.data(), "BCF\2\2", 5); memcpy(h.data() + 5, &l_text, 4); memcpy(h.data() + 9, text.data(), text.size()); h[9 + text.size()] = 0;
  append(h.data(), h.size(), false);
  header_printed_ = true;
}

char VCFAdapter::get_reference_base_at_position(const std::string& contig, int64_t pos) const {
  auto it = ref_bases_.find(contig + ":" + std::to_string(pos));
  if (it != ref_bases_.end()) return it->second;
  if (fasta_ && fasta_->is_open()) { try { return fasta_->base_at(contig, pos); } catch (const std::exception&) { return 'N'; } }
  return 'N';
}

// vcf_adapter.cc:59-248.  true: a line was added
bool add_field_to_hdr_if_missing(VCFHeaderLines& hdr, const VidMapper& id_mapper, const std::string& field_name, const std::string& vcf_name, int field_type_idx) {
  const FieldInfo* fi = id_mapper.get_field_info(field_name);
  const bool multid = fi && fi->length == "2D";
  bool exists = hdr.id2int(vcf_name) >= 0 && hdr.exists(field_type_idx, vcf_name) && hdr.hrec(field_type_idx, vcf_name) != nullptr;
  std::string description_value; bool removed = false;
  if (multid && exists) {   // such fields must be declared Number=1,Type=String: the line is taken out and put back
    const VCFHeaderRec* r = hdr.hrec(field_type_idx, vcf_name);
    const int di = r->find_key("Description");
    if (di >= 0) description_value = r->kv[di].second;
    hdr.remove(field_type_idx, vcf_name);
    exists = false; removed = true;
  }
  if (exists) return false;
  std::string line = "##";
  line += field_type_idx == HL_FLT ? "FILTER" : (field_type_idx == HL_INFO ? "INFO" : "FORMAT");
  line += "=<ID=" + vcf_name;
  if (field_type_idx != HL_FLT) {
    if (field_type_idx == HL_FMT && vcf_name == "GT") line += ",Number=1,Type=String";   // GT is weird
    else {
      if (!fi) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "field " + field_name + " is not in the vid mapping");
      std::string type = fi->type;
      for (auto& c : type) c = (char)tolower((unsigned char)c);
      const bool is_flag = type == "flag" || type == "bool" || type == "boolean";
      line += ",Number=";
      if (is_flag) line += "0";
      else if (multid) line += "1";
      else if (fi->length == "R" || fi->length == "A" || fi->length == "G") line += fi->length;
      else if (!fi->length.empty() && isdigit((unsigned char)fi->length[0])) line += fi->length;
      else if (fi->length == "VAR" || fi->length == "var" || fi->length == "." ) line += ".";
      else throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Unhandled field length descriptor " + fi->length + " of field " + field_name);
      line += ",Type=";
      if (multid) line += "String";
      else if (is_flag) line += "Flag";
      else if (type == "int" || type == "integer" || type == "int32" || type == "int64") line += "Integer";
      else if (type == "float" || type == "double") line += "Float";
      else line += "String";
    }
  }
  if (field_type_idx == HL_FMT && vcf_name == "GT") line += ",Description=\"Genotype\"";
  else if (removed && !description_value.empty()) line += ",Description=" + description_value;   // the value carries its quotes
  else line += ",Description=\"" + vcf_name + "\"";
  line += ">";
  hdr.append(line);
  return true;
}

// ------------------------------------------------------------------------------------------------ tool adapters
void VCFSerializedBufferAdapter::do_output() {
  std::vector<uint8_t>& b = get_buffer();
  if (m_do_output && m_output_fptr && !b.empty()) fwrite(b.data(), 1, b.size(), m_output_fptr);
  b.clear();
  if (m_rw_buffer) m_rw_buffer->m_num_valid_bytes = 0;
}

namespace {
double wall_now() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }
double cpu_now() { struct rusage u; getrusage(RUSAGE_SELF, &u); return u.ru_utime.tv_sec + 1e-6 * u.ru_utime.tv_usec + u.ru_stime.tv_sec + 1e-6 * u.ru_stime.tv_usec; }
}  // namespace
void Timer::start() { wall0_ = wall_now(); cpu0_ = cpu_now(); }
void Timer::stop() { wall_ = wall_now() - wall0_; cpu_ = cpu_now() - cpu0_; cum_wall_ += wall_; cum_cpu_ += cpu_; }
void Timer::print(const std::string& prefix, std::ostream& fptr) const {   // include/utils/timer.h
  fptr << "GENOMICSDB_TIMER," << prefix << ",Wall-clock time(s)," << cum_wall_ << ",Cpu time(s)," << cum_cpu_ << "\n";
}

}  // namespace genomicsdb_b200

// ------------------------------------------------------------------------------------------------ C ABI: the query stream
// One function per native method of com.intel.genomicsdb.reader.GenomicsDBQueryStream
// (src/main/jni/src/genomicsdb_GenomicsDBQueryStream.cc:29-111); the handle is the GenomicsDBBCFGenerator.
using genomicsdb_b200::GenomicsDBBCFGenerator;
namespace {
thread_local std::string g_stream_error;
int stream_fail(const std::exception& e, int code) { g_stream_error = e.what(); return code; }
}  // namespace

extern "C" {

int gdbgpu_stream_open(const char* loader_configuration_file, const char* query_configuration_file, const char* chr, int start, int end, int rank,
                       uint64_t buffer_capacity, uint64_t segment_size, int is_bcf, int produce_header_only, int use_missing_values_only_not_vector_end,
                       int keep_idx_fields_in_bcf_header, gdbgpu_stream** out) {
  if (!out || !query_configuration_file) return GDBGPU_ERR_ARG;
  *out = nullptr;
  try {
    const char* output_format = is_bcf ? "bu" : "";
    auto* g = new GenomicsDBBCFGenerator(loader_configuration_file ? loader_configuration_file : "", query_configuration_file, chr ? chr : "", start, end, rank,
                                         (size_t)buffer_capacity, (size_t)segment_size, output_format, produce_header_only != 0,
                                         is_bcf && use_missing_values_only_not_vector_end, is_bcf && keep_idx_fields_in_bcf_header);
    *out = reinterpret_cast<gdbgpu_stream*>(g);
    return GDBGPU_OK;
  } catch (const genomicsdb_b200::GenomicsDBGPUException& e) { return stream_fail(e, e.code());
  } catch (const std::exception& e) { return stream_fail(e, GDBGPU_ERR_IO); }
}
int gdbgpu_stream_close(gdbgpu_stream* s) { delete reinterpret_cast<GenomicsDBBCFGenerator*>(s); return 0; }
uint64_t gdbgpu_stream_num_bytes_available(gdbgpu_stream* s) { return s ? (uint64_t)reinterpret_cast<GenomicsDBBCFGenerator*>(s)->get_buffer_capacity() : 0; }
int gdbgpu_stream_read_next_byte(gdbgpu_stream* s) {
  if (!s) return -1;
  try { return (int)(int8_t)reinterpret_cast<GenomicsDBBCFGenerator*>(s)->read_next_byte(); } catch (const std::exception& e) { stream_fail(e, 0); return -1; }
}
// jniGenomicsDBRead: up to n bytes into dst + offset, producing batches as needed; the number of bytes read (0 at the end)
int64_t gdbgpu_stream_read(gdbgpu_stream* s, uint8_t* dst, uint64_t offset, uint64_t n) {
  if (!s) return 0;
  try { return (int64_t)reinterpret_cast<GenomicsDBBCFGenerator*>(s)->read_and_advance(dst, (size_t)offset, (size_t)n); } catch (const std::exception& e) { stream_fail(e, 0); return -1; }
}
int64_t gdbgpu_stream_skip(gdbgpu_stream* s, uint64_t n) {
  if (!s) return 0;
  try { return (int64_t)reinterpret_cast<GenomicsDBBCFGenerator*>(s)->read_and_advance(nullptr, 0, (size_t)n); } catch (const std::exception& e) { stream_fail(e, 0); return -1; }
}
const char* gdbgpu_stream_last_error(void) { return g_stream_error.c_str(); }

}  // extern "C"
