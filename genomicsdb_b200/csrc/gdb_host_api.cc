#include "gdb_host_api.h"

#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <unordered_set>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <set>
#include <tuple>
#include <sstream>

#include "gdb_json.h"
#include "gdb_vcf_header.h"

namespace genomicsdb_b200 {

using gdb::JsonValue;

// ------------------------------------------------------------------------------------------------ VidMapper
// FileBasedVidMapper (vid_mapper.cc:1209-1443): "contigs" and "fields" as a dictionary name -> info or a list of info
// dictionaries; type and length in every spelling the reference reads, reduced to one ("int" / "float" / "char" / "flag";
// "<n>" / "A" / "R" / "G" / "P" / "PP" / "VAR"); fields indexed as the reference indexes them -- a field that is both INFO and
// FORMAT is followed by its <name>_FORMAT twin, a field of tuples by one <name>_tuple_element_<j> per tuple element, and all of
// them count (FILTER cells hold these indices).  What the reference refuses is refused.
namespace {
[[noreturn]] void bad_vid(const std::string& what) { throw GenomicsDBGPUException(GDBGPU_ERR_ARG, what); }
std::string canonical_type(const std::string& t, const std::string& field) {   // vid_mapper.cc:51-89
  static const char* const ints[] = {"int", "Int", "integer", "Integer"}, * const floats[] = {"float", "Float"},
                   * const flags[] = {"bool", "Bool", "boolean", "Boolean", "flag", "Flag"}, * const chars[] = {"string", "String", "char", "Char"};
  for (const char* x : ints) if (t == x) return "int";
  for (const char* x : floats) if (t == x) return "float";
  for (const char* x : flags) if (t == x) return "flag";
  for (const char* x : chars) if (t == x) return "char";
  bad_vid("Unhandled field type " + t + " of field " + field);
}
std::string canonical_length_string(const std::string& s, const std::string& field) {   // vid_mapper.cc:33-49,802-830
  std::string u = s;
  for (char& c : u) c = (char)toupper((unsigned char)c);
  if (u == "A" || u == "BCF_VL_A") return "A";
  if (u == "R" || u == "BCF_VL_R") return "R";
  if (u == "G" || u == "BCF_VL_G") return "G";
  if (u == "P" || u == "BCF_VL_P") return "P";
  if (u == "VAR" || u == "BCF_VL_VAR") return "VAR";
  if (u == "PP" || u == "PHASED_PLOIDY") return "PP";
  if (u == "BCF_VL_FIXED") return "1";
  char* end = nullptr;
  const unsigned long long n = strtoull(s.c_str(), &end, 0);
  if (!s.empty() && (size_t)(end - s.c_str()) == s.size()) return std::to_string(n);
  std::cerr << "WARNING: unknown length descriptor " << s << " for field " << field << " ; setting to 'VAR'\n";
  return "VAR";
}
std::string canonical_length(const JsonValue& l, const std::string& field) {   // one dimension (vid_mapper.cc:1491-1528)
  if (l.is_int) return std::to_string(l.as_int());
  if (l.type == JsonValue::STR) return canonical_length_string(l.str, field);
  if (l.type == JsonValue::OBJ) {   // protobuf: { "variable_length_descriptor" : "R" } / { "fixed_length" : 2 }
    if (const JsonValue* v = l.find("variable_length_descriptor")) { if (v->type != JsonValue::STR) bad_vid("\"variable_length_descriptor\" of field " + field + " must be a string"); return canonical_length_string(v->str, field); }
    const JsonValue* f = l.find("fixed_length");
    if (!f || !f->is_int) bad_vid("the length of field " + field + " needs \"variable_length_descriptor\" or an integer \"fixed_length\"");
    return std::to_string(f->as_int());
  }
  bad_vid("the length of field " + field + " must be an integer, a string, a dictionary or a list of those");
}
// a known field the vid mapping gives no length for (known_field_info.cc:239-283)
std::string known_field_length(const std::string& name) {
  if (name == "REF" || name == "ALT" || name == "FILTER" || name == "PGT" || name == "PID") return "VAR";
  if (name == "AF" || name == "AC") return "A";
  if (name == "AD") return "R";
  if (name == "PL") return "G";
  if (name == "GT") return "PP";
  if (name == "SB") return "4";
  return "1";
}
std::string entry_name(const JsonValue& v, std::initializer_list<const char*> keys, const char* what, size_t idx) {
  std::string name; unsigned found = 0;
  for (const char* k : keys) if (const JsonValue* n = v.find(k)) { name = n->str; ++found; }
  if (found != 1) bad_vid(std::string(what) + " info dict with index " + std::to_string(idx) + (found ? " has two or more of its name keys" : " has no name key"));
  return name;
}
}  // namespace

VidMapper::VidMapper(const std::string& file) {
  JsonValue d;
  try { d = gdb::json_parse_file(file); } catch (const std::exception& e) { bad_vid("Syntax error in JSON file " + file + " (" + e.what() + ")"); }
  // callsets may sit in the vid mapping itself (vid_mapper.cc:686-711: a file name, "callsets", or protobuf's "callset_mapping")
  if (const JsonValue* c = d.find("callset_mapping_file")) { if (c->type != JsonValue::STR && c->type != JsonValue::ARR) bad_vid("\"callset_mapping_file\" must be a string or a list of strings"); read_callsets_file(c->type == JsonValue::STR ? c->str : (c->arr.empty() ? std::string() : c->arr[0].str)); }
  if (const JsonValue* c = d.find("callsets")) read_callsets(*c);
  if (const JsonValue* c = d.find("callset_mapping")) { if (c->type != JsonValue::OBJ || !c->has("callsets")) bad_vid("\"callset_mapping\" must hold \"callsets\""); read_callsets(c->at("callsets")); }
  const JsonValue* cl = d.find("contigs");
  if (!cl || (cl->type != JsonValue::OBJ && cl->type != JsonValue::ARR)) bad_vid("the vid mapping " + file + " has no \"contigs\"");
  {
    bool duplicate = false;
    auto addc = [&](const std::string& name, const JsonValue& v) {
      if (v.type != JsonValue::OBJ) bad_vid("contig " + name + ": the info must be a dictionary");
      if (get_contig_info(name)) { std::cerr << "Duplicate contig/chromosome name " << name << " found in vid file " << file << "\n"; duplicate = true; return; }
      const JsonValue* o = v.find("tiledb_column_offset"); const JsonValue* l = v.find("length");
      if (!o || !o->is_int || o->as_int() < 0) bad_vid("contig " + name + " needs an integer \"tiledb_column_offset\" >= 0");
      if (!l || !l->is_int || l->as_int() < 0) bad_vid("contig " + name + " needs an integer \"length\" >= 0");
      ContigInfo c; c.name = name; c.tiledb_column_offset = o->as_int(); c.length = l->as_int();
      contigs_.push_back(c);
    };
    if (cl->type == JsonValue::OBJ) for (auto& kv : cl->obj) addc(kv.first, kv.second);
    else for (size_t i = 0; i < cl->arr.size(); ++i) { if (cl->arr[i].type != JsonValue::OBJ) bad_vid("a contig's info must be a dictionary"); addc(entry_name(cl->arr[i], {"name", "contig_name", "chromosome_name"}, "Contig", i), cl->arr[i]); }
    if (duplicate) bad_vid("Duplicate contigs exist in vid file " + file);
    contigs_file_order_ = contigs_;
    std::stable_sort(contigs_.begin(), contigs_.end(), [](const ContigInfo& a, const ContigInfo& b) { return a.tiledb_column_offset < b.tiledb_column_offset; });
    for (size_t i = 1; i < contigs_.size(); ++i)
      if (contigs_[i].tiledb_column_offset <= contigs_[i - 1].tiledb_column_offset + contigs_[i - 1].length - 1) {
        std::cerr << "Contig/chromosome " << contigs_[i].name << " begins at TileDB column " << contigs_[i].tiledb_column_offset << " and intersects with contig/chromosome " << contigs_[i - 1].name << "\n";
        bad_vid("Overlapping contigs exist in vid file " + file);
      }
  }
  const JsonValue* fl = d.find("fields");
  if (!fl || (fl->type != JsonValue::OBJ && fl->type != JsonValue::ARR)) bad_vid("the vid mapping " + file + " has no \"fields\"");
  int field_idx = 0;
  bool duplicate_fields = false;
  auto add = [&](const std::string& name, const JsonValue& v) {
    if (v.type != JsonValue::OBJ) bad_vid("field " + name + ": the info must be a dictionary");
    if (get_field_info(name)) { std::cerr << "Duplicate field name " << name << " found in vid attribute \"fields\"\n"; duplicate_fields = true; return; }
    FieldInfo f; f.name = name;
    if (const JsonValue* c = v.find("vcf_field_class")) {
      if (c->type != JsonValue::ARR) bad_vid("\"vcf_field_class\" of field " + name + " must be a list");
      for (auto& x : c->arr) { if (x.str == "FORMAT") f.is_format = true; if (x.str == "INFO") f.is_info = true; if (x.str == "FILTER") f.is_filter = true; }
    }
    f.length = "1";
    if (const JsonValue* l = v.find("length")) {
      if (l->type == JsonValue::ARR) {
        if (l->arr.size() == 1) f.length = canonical_length(l->arr[0], name);
        else { f.length = "2D"; for (auto& x : l->arr) f.dim_lengths.push_back(canonical_length(x, name)); }
      } else f.length = canonical_length(*l, name);
    } else f.length = known_field_length(name);
    const JsonValue* t = v.find("type");
    if (!t) bad_vid("field " + name + " has no \"type\"");
    if (t->type == JsonValue::STR) f.type = canonical_type(t->str, name);
    else if (t->type == JsonValue::ARR) {
      for (auto& x : t->arr) { if (x.type != JsonValue::STR) bad_vid("\"type\" of field " + name + " must hold strings"); f.tuple_types.push_back(canonical_type(x.str, name)); }
      f.type = f.tuple_types.empty() ? "int" : f.tuple_types[0];
    } else bad_vid("Attribute 'type' of field " + name + " must be a string or an array");
    f.vcf_type = f.type;
    if (const JsonValue* vt = v.find("vcf_type")) {   // what the header line declares (vcf_adapter.cc:128,167); a multi-dimensional field is always a string there
      if (vt->type == JsonValue::STR) f.vcf_type = canonical_type(vt->str, name);
      else if (vt->type == JsonValue::ARR) { for (size_t i = 0; i < vt->arr.size(); ++i) { const std::string t = canonical_type(vt->arr[i].str, name); if (i == 0) f.vcf_type = t; } }
      else bad_vid("Attribute 'vcf_type' of field " + name + " must be a string or an array");
    }
    if (f.length == "2D")   // vid_mapper.cc:200-215
      for (auto& e : (f.tuple_types.empty() ? std::vector<std::string>(1, f.type) : f.tuple_types))
        if (e != "int" && e != "float") bad_vid("Unhandled element type for multi-D field " + name + "; only float and int multi-d fields supported");
    if (const JsonValue* c = v.find("VCF_field_combine_operation")) {
      static const char* const ops[] = {"sum", "mean", "median", "move_to_FORMAT", "element_wise_sum", "elementwise_sum", "concatenate", "histogram_sum"};
      bool known = false;
      for (const char* o : ops) if (c->type == JsonValue::STR && c->str == o) known = true;
      if (!known) bad_vid("Unknown VCF field combine operation " + c->str + " specified for field " + name);
      f.combine_op = c->str == "elementwise_sum" ? "element_wise_sum" : c->str;
      if (f.combine_op == "concatenate" && (f.length == "2D" ? f.dim_lengths[0] : f.length) != "VAR")
        bad_vid("VCF field combined operation 'concatenate' can only be used with fields whose length descriptors are 'VAR';  field " + name + " does not have 'VAR' as its length descriptor");
    } else {   // defaults of the known fields (known_field_info.cc:284-308)
      static const char* medians[] = {"BaseQRankSum", "ClippingRankSum", "MQRankSum", "ReadPosRankSum", "MQ", "MQ0", "ExcessHet"};
      for (const char* m : medians) if (name == m) f.combine_op = "median";
      if (name == "RAW_MQ") f.combine_op = "sum";
      if (name == "DP") f.combine_op = "DP";
    }
    if (const JsonValue* dl = v.find("vcf_delimiter")) {
      if (dl->type == JsonValue::ARR) { for (auto& x : dl->arr) { if (x.type != JsonValue::STR) bad_vid("\"vcf_delimiter\" of field " + name + " must hold strings"); f.vcf_delimiters.push_back(x.str.empty() ? ',' : x.str[0]); } }
      else if (dl->type == JsonValue::STR) { if (!dl->str.empty()) f.vcf_delimiters.push_back(dl->str[0]); }
      else bad_vid("\"vcf_delimiter\" of field " + name + " must be a string or a list of strings");
    }
    if (f.tuple_types.size() > 1) { if (f.length != "2D") { f.dim_lengths.assign(1, f.length); f.length = "2D"; } }   // a vector of tuples is handled like a multi-dimensional field (raw bytes, Type=String)
    while (f.length == "2D" && f.vcf_delimiters.size() < 2) f.vcf_delimiters.push_back(f.vcf_delimiters.empty() ? '|' : ',');
    f.global_idx = field_idx++;
    fields_.push_back(f);
    // vid_mapper.cc:727-788 flatten_field: the FORMAT twin, then one attribute per tuple element (of the field, then of its twin)
    const bool both = f.is_format && f.is_info;
    if (both) { FieldInfo g = f; g.name = name + "_FORMAT"; g.is_format_twin = true; g.global_idx = field_idx++; fields_.push_back(g); }
    if (f.is_composite())
      for (int j = 0; j < (both ? 2 : 1); ++j)
        for (size_t e = 0; e < f.tuple_types.size(); ++e) {
          FieldInfo g = f; g.name = name + (j ? "_FORMAT" : "") + "_tuple_element_" + std::to_string(e); g.parent = name + (j ? "_FORMAT" : ""); g.tuple_index = (int)e; g.type = f.tuple_types[e];
          g.is_format_twin = j == 1; g.global_idx = field_idx++;
          fields_.push_back(g);
        }
  };
  if (fl->type == JsonValue::OBJ) for (auto& kv : fl->obj) add(kv.first, kv.second);
  else for (size_t i = 0; i < fl->arr.size(); ++i) {
    const JsonValue& v = fl->arr[i];
    if (v.type != JsonValue::OBJ) bad_vid("a field's info must be a dictionary");
    if (v.has("name") && v.has("field_name")) bad_vid("Field information dictionary with index " + std::to_string(i) + " has both \"field_name\" or \"name\" keys - only one should be specified");
    if (!v.has("name") && !v.has("field_name")) bad_vid("Field information dictionary with index " + std::to_string(i) + " does not have a \"field_name\" or \"name\" key");
    add(v.has("name") ? v.at("name").str : v.at("field_name").str, v);
  }
  if (duplicate_fields) bad_vid("Duplicate fields exist in vid attribute \"fields\"");
  // END, REF, ALT, QUAL and FILTER follow when the file does not list them (vid_mapper.cc:611-684); the mirror answers for
  // them where they are asked about (do_query_bookkeeping) and keeps fields_ to what the file names
  num_fields_ = field_idx;
  for (const char* m : {"END", "REF", "ALT", "QUAL", "FILTER"}) if (!get_field_info(m)) ++num_fields_;
  initialized_ = true;
}
const ContigInfo* VidMapper::get_contig_location(int64_t column) const {
  const ContigInfo* best = nullptr;
  for (auto& c : contigs_) { if (c.tiledb_column_offset <= column) best = &c; else break; }
  return (best && column < best->tiledb_column_offset + best->length) ? best : nullptr;
}
const FieldInfo* VidMapper::get_field_info(const std::string& name) const {
  for (auto& f : fields_) if (f.name == name) return &f;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ VariantQueryConfig
void VariantQueryConfig::add_column_interval_to_query(int64_t b, int64_t e) {
  intervals_.push_back({b, e});
  // variant_query_config.cc:32-35,69-73: ordered by begin alone, re-sorted after every addition (intervals that share a begin
  // keep the order that leaves them in)
  std::sort(intervals_.begin(), intervals_.end(), [](const std::pair<int64_t, int64_t>& x, const std::pair<int64_t, int64_t>& y) { return x.first < y.first; });
}
void VariantQueryConfig::set_rows_to_query(const std::vector<int64_t>& rows) {
  rows_ = rows;
  std::sort(rows_.begin(), rows_.end());             // variant_query_config.cc:60-67
}
// VariantQueryConfig::read_from_file: gdb_host_config.cc (with the loader's configuration)

// ------------------------------------------------------------------------------------------------ operators
// variant_operations.cc:362-378 -- the merge itself ran on the GPU (k_union); pick the results up
void SingleVariantOperatorBase::operate(Variant& variant, const VariantQueryConfig&) {
  m_merged_reference_allele = variant.merged_ref_;
  m_merged_alt_alleles = variant.merged_alts_;
  m_NON_REF_exists = variant.nonref_;
  m_remapping_needed = variant.remapped_;
  m_is_reference_block_only = !variant.remapped_;
}
// variant_operations.cc:572-695 -- the remap ran on the GPU (k_remap_*); the variant already carries its result
void GA4GHOperator::operate(Variant& variant, const VariantQueryConfig& query_config) {
  SingleVariantOperatorBase::operate(variant, query_config);
  m_max_diploid_alt_alleles_that_can_be_genotyped = query_config.get_max_diploid_alt_alleles_that_can_be_genotyped();
  m_remapped_variant = &variant;
  if (variant.g_suppressed_) warn_too_many_alleles(query_config, variant.get_column_begin(), (unsigned)m_merged_alt_alleles.size());
}
// variant_operations.cc:613-629: one line on stderr per genotype-length field that is dropped at a site with more ALT
// alleles than can be genotyped (the kernels dropped it: GDBGPU_SITE_G_SUPPRESSED)
void GA4GHOperator::warn_too_many_alleles(const VariantQueryConfig& query_config, int64_t column, unsigned num_alt) const {
  const VidMapper& vid = query_config.get_vid_mapper();
  for (unsigned i = 0; i < query_config.get_num_remapped_fields(); ++i) {
    const FieldInfo* fi = vid.get_field_info(query_config.get_remapped_field_name(i));
    if (!fi || fi->length != "G") continue;
    const ContigInfo* contig = vid.get_contig_location(column);
    if (contig) std::cerr << "Chromosome " << contig->name << " position " << column - contig->tiledb_column_offset + 1 << " (";
    std::cerr << "TileDB column " << column;
    if (contig) std::cerr << ")";
    std::cerr << " has too many alleles in the combined VCF record : " << num_alt << " : current limit : " << m_max_diploid_alt_alleles_that_can_be_genotyped
              << ". Fields, such as  PL, with length equal to the number of genotypes will NOT be added for this location.\n";
  }
}

// ------------------------------------------------------------------------------------------------ VCFAdapter
// BCF dictionary of strings: FILTER / INFO / FORMAT IDs share one index space (IDX= when htslib wrote it, else order of first
// appearance, PASS = 0); contigs have their own (htslib vcf.c bcf_hdr_sync)
void VCFAdapter::set_vcf_header_text(const std::string& header) {
  header_ = header; ids_.clear(); contigs_.clear(); info_ids_.clear();
  int next_id = 0, next_ctg = 0;
  size_t pos = 0;
  while (pos < header.size()) {
    size_t eol = header.find('\n', pos);
    if (eol == std::string::npos) eol = header.size();
    const std::string line = header.substr(pos, eol - pos);
    pos = eol + 1;
    auto attr = [&](const char* key) -> std::string {
      size_t k = line.find(key);
      if (k == std::string::npos) return "";
      k += strlen(key);
      size_t e = line.find_first_of(",>", k);
      return line.substr(k, e == std::string::npos ? std::string::npos : e - k);
    };
    const bool dict = line.compare(0, 10, "##FILTER=<") == 0 || line.compare(0, 8, "##INFO=<") == 0 || line.compare(0, 10, "##FORMAT=<") == 0;
    if (dict) {
      const std::string id = attr("ID=");
      if (line.compare(0, 8, "##INFO=<") == 0) info_ids_.insert(id);
      bool known = false;
      for (auto& kv : ids_) if (kv.first == id) known = true;
      if (known) continue;
      const std::string idx = attr("IDX=");
      const int v = idx.empty() ? next_id : std::stoi(idx);
      ids_.push_back({id, v});
      next_id = std::max(next_id, v + 1);
    } else if (line.compare(0, 10, "##contig=<") == 0) {
      const std::string idx = attr("IDX=");
      const int v = idx.empty() ? next_ctg : std::stoi(idx);
      contigs_.push_back({attr("ID="), v});
      next_ctg = std::max(next_ctg, v + 1);
    }
  }
}
const ContigInfo* VidMapper::get_contig_info(const std::string& name) const {
  for (auto& c : contigs_) if (c.name == name) return &c;
  return nullptr;
}
int VCFAdapter::get_dictionary_id(const std::string& name) const { for (auto& kv : ids_) if (kv.first == name) return kv.second; return -1; }
int VCFAdapter::get_contig_id(const std::string& name) const { for (auto& kv : contigs_) if (kv.first == name) return kv.second; return -1; }
void VCFAdapter::append(const void* p, size_t n, bool borrowed) {
  if (!n) return;
  if (fd_ < 0) { const uint8_t* b = static_cast<const uint8_t*>(p); buffer_.insert(buffer_.end(), b, b + n); return; }
  if (!borrowed) {   // owned bytes go into chunks that never move
    if (chunk_used_ + n > chunk_cap_) { chunk_cap_ = std::max<size_t>(n, 1u << 20); chunks_.emplace_back(new uint8_t[chunk_cap_]); chunk_used_ = 0; }
    uint8_t* dst = chunks_.back().get() + chunk_used_;
    memcpy(dst, p, n);
    chunk_used_ += n;
    p = dst;
  }
  if (!pending_.empty() && pending_.back().p + pending_.back().n == p) pending_.back().n += n;
  else pending_.push_back(Range{static_cast<const uint8_t*>(p), n});
  pending_bytes_ += n;
}
void VCFAdapter::flush() {
  if (fd_ < 0 || pending_.empty()) return;
  size_t i = 0;
  std::vector<struct iovec> iov;
  while (i < pending_.size()) {
    iov.clear();
    for (size_t k = i; k < pending_.size() && iov.size() < 512; ++k) iov.push_back({const_cast<uint8_t*>(pending_[k].p), pending_[k].n});
    ssize_t w = writev(fd_, iov.data(), (int)iov.size());
    if (w < 0) { if (errno == EINTR) continue; throw GenomicsDBGPUException(GDBGPU_ERR_IO, std::string("writev: ") + strerror(errno)); }
    bytes_written_ += (uint64_t)w;
    size_t left = (size_t)w;   // a short write ends inside some range
    while (i < pending_.size() && left >= pending_[i].n) { left -= pending_[i].n; ++i; }
    if (i < pending_.size() && left) { pending_[i].p += left; pending_[i].n -= left; }
  }
  pending_.clear(); pending_bytes_ = 0;
  chunks_.clear(); chunk_used_ = chunk_cap_ = 0;
}
// ------------------------------------------------------------------------------------------------ BroadCombinedGVCFOperator
namespace {
// htslib typed values (vcf.h bcf_enc_int1 / bcf_enc_size, vcf.c bcf_enc_vchar / bcf_enc_vfloat)
enum { BT_INT8 = 1, BT_INT16 = 2, BT_INT32 = 3, BT_FLOAT = 5, BT_CHAR = 7 };
void put_bytes(std::vector<uint8_t>& b, const void* p, size_t n) { const uint8_t* s = static_cast<const uint8_t*>(p); b.insert(b.end(), s, s + n); }
template <class T> void put(std::vector<uint8_t>& b, T v) { put_bytes(b, &v, sizeof(T)); }
void enc_int1(std::vector<uint8_t>& b, int32_t x) {
  if (x == INT32_MIN + 1) { b.push_back(1 << 4 | BT_INT8); b.push_back(0x81); }
  else if (x == INT32_MIN) { b.push_back(1 << 4 | BT_INT8); b.push_back(0x80); }
  else if (x <= 127 && x >= -120) { b.push_back(1 << 4 | BT_INT8); b.push_back((uint8_t)(int8_t)x); }
  else if (x <= 32767 && x >= -32760) { b.push_back(1 << 4 | BT_INT16); put<int16_t>(b, (int16_t)x); }
  else { b.push_back(1 << 4 | BT_INT32); put<int32_t>(b, x); }
}
void enc_size(std::vector<uint8_t>& b, int size, int type) {
  if (size >= 15) { b.push_back((uint8_t)(15 << 4 | type)); enc_int1(b, size); }
  else b.push_back((uint8_t)(size << 4 | type));
}
void enc_vchar(std::vector<uint8_t>& b, const std::string& s) { enc_size(b, (int)s.size(), BT_CHAR); put_bytes(b, s.data(), s.size()); }
// bcf_enc_vint (vcf.c:1926-1972) with wsize = -1: one vector of n values
void enc_vint(std::vector<uint8_t>& b, const int32_t* a, uint32_t n) {
  if (n == 1) { enc_int1(b, a[0]); return; }
  int32_t mx = INT32_MIN + 1, mn = INT32_MAX;
  for (uint32_t i = 0; i < n; ++i) { if (a[i] == INT32_MIN || a[i] == INT32_MIN + 1) continue; mx = std::max(mx, a[i]); mn = std::min(mn, a[i]); }
  if (mx <= 127 && mn > -127) {
    enc_size(b, (int)n, BT_INT8);
    for (uint32_t i = 0; i < n; ++i) b.push_back(a[i] == INT32_MIN + 1 ? 0x81 : (a[i] == INT32_MIN ? 0x80 : (uint8_t)(int8_t)a[i]));
  } else if (mx <= 32767 && mn > -32767) {
    enc_size(b, (int)n, BT_INT16);
    for (uint32_t i = 0; i < n; ++i) put<int16_t>(b, a[i] == INT32_MIN + 1 ? (int16_t)-32767 : (a[i] == INT32_MIN ? (int16_t)-32768 : (int16_t)a[i]));
  } else { enc_size(b, (int)n, BT_INT32); put_bytes(b, a, (size_t)n * 4); }
}
}  // namespace

BroadCombinedGVCFOperator::BroadCombinedGVCFOperator(VCFAdapter& vcf_adapter, const VidMapper& id_mapper, const VariantQueryConfig& qc,
                                                     bool use_missing_values_only_not_vector_end)
    : m_vcf_adapter(&vcf_adapter), m_vid_mapper(&id_mapper), m_use_missing_values_not_vector_end(use_missing_values_only_not_vector_end) {
  if (!qc.is_bookkeeping_done()) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "BroadCombinedGVCFOperator: do_query_bookkeeping(config, true, true) must be called first");
  if (vcf_adapter.is_initialized_from_config()) build_header(vcf_adapter, id_mapper, qc);
  auto key_of = [&](const std::string& vcf_name) {
    int k = vcf_adapter.get_dictionary_id(vcf_name);
    if (k < 0) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "field " + vcf_name + " is not in the VCF header");
    return k;
  };
  const unsigned n_remapped = qc.get_num_remapped_fields();
  const auto& plain = qc.get_plain_fields();
  auto remapped_index = [&](const std::string& name) -> int { for (unsigned i = 0; i < n_remapped; ++i) if (qc.get_remapped_field_name(i) == name) return (int)i; return -1; };
  // FORMAT fields: GT first (htslib moves it to the front, vcf.c:3670-3676), then query order, DP (from DP_FORMAT) last
  // (the kernels remap GT whenever the array has it; the record only carries it when the query does -- listed, or pulled in by a
  // genotype-dependent field, query_variants.cc:617-623)
  const auto& ordered_attrs = qc.get_ordered_query_attributes();
  const bool gt_queried = std::find(ordered_attrs.begin(), ordered_attrs.end(), "GT") != ordered_attrs.end();
  if (remapped_index("GT") >= 0 && gt_queried) m_FORMAT_fields_vec.push_back({"GT", key_of("GT"), remapped_index("GT")});
  int dp_format_plain = -1;
  for (auto& name : qc.get_ordered_query_attributes()) {   // the order the reference's bookkeeping leaves (END, REF, ALT swapped to the front)
    if (name == "GT") continue;
    const FieldInfo* fi = id_mapper.get_field_info(name);
    if (!fi) continue;
    const int ri = remapped_index(name);
    const int gt_shift = remapped_index("GT") >= 0 ? 1 : 0;
    if (ri >= 0) {
      const bool twin = name.size() > 7 && name.compare(name.size() - 7, 7, "_FORMAT") == 0;
      if ((fi->is_format && (twin || !fi->is_info)) || fi->combine_op == "move_to_FORMAT") m_FORMAT_fields_vec.push_back({name, key_of(name), ri});
      else if (qc.get_element_wise_sum_mask() & (1u << (ri - gt_shift))) m_INFO_fields_vec.push_back({name, key_of(name), -1, fi->type == "float", ri - gt_shift});
      continue;
    }
    for (size_t p = 0; p < plain.size(); ++p) if (plain[p].attribute == name) {
      if (plain[p].is_qual) m_QUAL_plain = (int)p;
      else if (plain[p].is_id) m_ID_fmt_index = (int)(n_remapped + p);
      else if (plain[p].is_filter) { /* not a query-order field: below */ }
      else if (plain[p].is_raw) {
        if (name == "REF") { m_raw_REF_fmt_index = (int)(n_remapped + p); continue; }
        if (name == "ALT") { m_raw_ALT_fmt_index = (int)(n_remapped + p); continue; }
        if (fi->combine_op == "histogram_sum") {
          const FieldInfo* parent = id_mapper.get_field_info(fi->parent);
          const int parent_idx = parent ? parent->global_idx : 0;
          auto it = std::find_if(m_INFO_histogram_fields.begin(), m_INFO_histogram_fields.end(), [&](const HistogramField& h) { return h.parent_idx == parent_idx; });
          if (it == m_INFO_histogram_fields.end()) { HistogramField h; h.parent_idx = parent_idx; h.key = key_of(fi->vcf_name()); m_INFO_histogram_fields.push_back(h); it = m_INFO_histogram_fields.end() - 1; }
          const int e = fi->tuple_index == 0 ? 0 : 1;
          it->raw_fmt_index[e] = (int)(n_remapped + p); it->fi[e] = fi;
        } else {
          InfoField f; f.vcf_name = plain[p].vcf_name; f.key = key_of(plain[p].vcf_name); f.is_float = fi->type == "float";
          f.raw_fmt_index = (int)(n_remapped + p); f.fi = fi;
          m_INFO_fields_vec.push_back(f);
        }
      }
      else if (plain[p].is_concat) {
        InfoField f; f.vcf_name = plain[p].vcf_name; f.key = key_of(plain[p].vcf_name); f.is_float = plain[p].elem == GDBGPU_ELEM_FLOAT32;
        f.concat_fmt_index = (int)(n_remapped + p);
        m_INFO_fields_vec.push_back(f);
      }
      else if (plain[p].format) {
        if (name == "DP_FORMAT") dp_format_plain = (int)p;
        else m_FORMAT_fields_vec.push_back({plain[p].vcf_name, key_of(plain[p].vcf_name), (int)(n_remapped + p)});
      } else if (plain[p].info_op == GDBGPU_INFO_ELEMENT_WISE_SUM)
        m_INFO_fields_vec.push_back({plain[p].vcf_name, key_of(plain[p].vcf_name), -1, plain[p].elem == GDBGPU_ELEM_FLOAT32, (int)(n_remapped - gt_shift + p)});
      else if (plain[p].info_op)
        m_INFO_fields_vec.push_back({plain[p].vcf_name, key_of(plain[p].vcf_name), (int)p, plain[p].elem == GDBGPU_ELEM_FLOAT32, -1});
    }
  }
  if (dp_format_plain >= 0) m_FORMAT_fields_vec.push_back({"DP", key_of("DP"), (int)(n_remapped + dp_format_plain)});
  for (size_t p = 0; p < plain.size(); ++p) if (plain[p].is_filter) m_FILTER_fmt_index = (int)(n_remapped + p);
  for (size_t p = 0; p < plain.size(); ++p) if (plain[p].is_raw && plain[p].attribute == "REF") m_raw_REF_fmt_index = (int)(n_remapped + p); else if (plain[p].is_raw && plain[p].attribute == "ALT") m_raw_ALT_fmt_index = (int)(n_remapped + p);
  std::sort(m_INFO_histogram_fields.begin(), m_INFO_histogram_fields.end(), [](const HistogramField& a, const HistogramField& b) { return a.parent_idx < b.parent_idx; });
  for (auto& h : m_INFO_histogram_fields) if (h.raw_fmt_index[0] < 0 || h.raw_fmt_index[1] < 0) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "histogram_sum: both tuple elements of the field must be queried");
  for (auto& f : id_mapper.fields()) {   // broad_combined_gvcf.cc:306-313: vid field index -> header dictionary id (could be -1)
    if (f.global_idx < 0) continue;
    if ((size_t)f.global_idx >= m_global_field_idx_to_hdr_idx.size()) m_global_field_idx_to_hdr_idx.resize(f.global_idx + 1, -1);
    std::string vcf_name = f.vcf_name();   // a twin and a flattened element go by the VCF name of the field they come from
    if (f.is_format_twin && vcf_name.size() > 7 && vcf_name.compare(vcf_name.size() - 7, 7, "_FORMAT") == 0) vcf_name.resize(vcf_name.size() - 7);
    m_global_field_idx_to_hdr_idx[f.global_idx] = vcf_adapter.get_dictionary_id(vcf_name);
  }
  m_dp_rule = qc.get_dp_rule_fields()[0] != 0 || qc.get_dp_rule_fields()[2] != 0;
  if (m_dp_rule) { m_DP_key = key_of("DP"); m_dp_rule = vcf_adapter.has_info_line("DP"); }   // bcf_update_info_int32 fails without an ##INFO line: no INFO/DP then
  m_END_key = key_of("END");
  m_num_vector_slots = n_remapped - (remapped_index("GT") >= 0 ? 1u : 0u) + (unsigned)plain.size();
}

// broad_combined_gvcf.cc:168-330: the template header gets the queried INFO / FORMAT fields it lacks, the FILTER lines when
// FILTER is queried, the contigs of the vid mapping and -- unless the query is sites-only -- the samples of the queried rows;
// then it is "cleaned up completely" (written without IDX fields, read back) and goes to the head of the output.
bool add_field_to_hdr_if_missing(VCFHeaderLines& hdr, const VidMapper& id_mapper, const std::string& field_name, const std::string& vcf_name, int field_type_idx);   // gdb_host_config.cc
void BroadCombinedGVCFOperator::build_header(VCFAdapter& vcf_adapter, const VidMapper& id_mapper, const VariantQueryConfig& qc) {
  if (!id_mapper.is_callset_mapping_initialized() && !qc.sites_only_query())
    throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Callset mapping in id mapper is not initialized");
  VCFHeaderLines& hdr = *vcf_adapter.header_lines();
  const bool sites_only = qc.sites_only_query();
  bool filter_queried = false;
  for (auto& name : qc.get_ordered_query_attributes()) {
    if (name == "FILTER") { filter_queried = true; continue; }
    const FieldInfo* fi = id_mapper.get_field_info(name);
    if (!fi) continue;
    const bool twin = name.size() > 7 && name.compare(name.size() - 7, 7, "_FORMAT") == 0;
    const std::string vcf_name = twin ? name.substr(0, name.size() - 7) : fi->vcf_name();   // a flattened tuple element goes by its parent's name
    const bool is_info = fi->is_info && !twin, is_format = fi->is_format && (twin || !fi->is_info);
    const std::string& op = fi->combine_op;
    const bool dp_rule = vcf_name == "DP" && !twin && op == "DP";
    const bool add_to_INFO = is_info && name != "END" && !dp_rule && op != "move_to_FORMAT";
    const bool add_to_FORMAT = (is_format && (!sites_only || name == "DP_FORMAT" || name == "MIN_DP")) || (is_info && (dp_rule || (op == "move_to_FORMAT" && !sites_only)));
    if (add_to_INFO) {
      if (op.empty()) std::cerr << "WARNING: No valid combination operation found for INFO field " << vcf_name << " - the field will NOT be part of INFO fields in the generated VCF records\n";
      else add_field_to_hdr_if_missing(hdr, id_mapper, fi->parent.empty() ? name : fi->parent, vcf_name, HL_INFO);
    }
    if (add_to_FORMAT) {
      if (is_format || op == "move_to_FORMAT") add_field_to_hdr_if_missing(hdr, id_mapper, name, vcf_name, HL_FMT);
      else add_field_to_hdr_if_missing(hdr, id_mapper, name, "DP", HL_INFO);
    }
  }
  if (filter_queried) for (auto& f : id_mapper.fields()) if (f.is_filter) add_field_to_hdr_if_missing(hdr, id_mapper, f.name, f.name, HL_FLT);
  for (unsigned i = 0; i < id_mapper.get_num_contigs(); ++i) {
    const ContigInfo& c = id_mapper.get_contig_info(i);
    if (hdr.name2id(c.name) < 0) hdr.append("##contig=<ID=" + c.name + ",length=" + std::to_string(c.length) + ">");
  }
  if (!sites_only) {
    std::string callset_name;
    for (uint64_t i = 0; i < qc.get_num_rows_to_query(); ++i) {
      const int64_t row_idx = qc.get_array_row_idx_for_query_row_idx(i);
      if (row_idx > id_mapper.get_max_callset_row_idx()) continue;   // rows past the array's last one are not part of the query (the reference clamps them away)
      if (!id_mapper.get_callset_name(row_idx, callset_name) || callset_name.empty())
        throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "No sample/CallSet name specified in JSON file/Protobuf object for TileDB row " + std::to_string(row_idx));
      if (!hdr.add_sample(callset_name)) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Could not add sample " + callset_name + " to the combined VCF/gVCF header");
    }
  }
  vcf_adapter.finalize_header(!sites_only);
  vcf_adapter.print_header();
}

// One call's value of a 2-D field, cut into its dimension-0 entries.  The stored value (genomicsdb_multid_vector_field.h:66-83) is
// <u64 bytes of data><entry 0 data><entry 1 data>...<u64 #entries><u64 offset of entry 0, entry 1, ..., end>.  Where the reference
// hands the combine the REMAPPED variant (a site that needs remapping, a field whose first dimension follows the alleles:
// broad_combined_gvcf.cc:386-388,573-574), the entries are those of remap_allele_specific_annotations (variant_operations.cc:
// 482-555): one per merged allele (R) or merged ALT allele (A), taken from the input allele that names it, else from the
// call's <NON_REF>, else empty.  The call's input -> merged allele mapping is rebuilt from its own REF / ALT strings with the
// rules of merge_alt_alleles (variant_operations.cc:133-228): a shorter REF extends non-symbolic ALT alleles by the merged REF's tail.
bool BroadCombinedGVCFOperator::multid_entries(const gdbgpu_batch& b, uint32_t s, uint64_t q, int raw_fmt_index, const FieldInfo& fi,
                                               std::vector<std::pair<const uint8_t*, uint64_t>>& entries) const {
  entries.clear();
  const uint64_t ns = b.n_sites;
  auto raw_row = [&](int fmt_index, const uint8_t*& p, uint32_t& len) -> bool {
    const uint64_t slot = (uint64_t)fmt_index * ns + s;
    if (!b.fmt_type[slot]) return false;
    const uint32_t nps = b.fmt_nps[slot];
    const uint8_t* row = b.fmt_bytes + b.fmt_offset[slot] + q * nps;
    memcpy(&len, row, 4);
    if (len == 0xFFFFFFFFu) return false;
    if (len + 4u > nps) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "raw field row longer than its block's row");
    p = row + 4;
    return true;
  };
  const uint8_t* v = nullptr; uint32_t vlen = 0;
  if (!raw_row(raw_fmt_index, v, vlen)) return false;
  auto rd64 = [](const uint8_t* p) { uint64_t x; memcpy(&x, p, 8); return x; };
  const std::string bad = "malformed multi-dimensional value of field " + fi.name;
  if (vlen < 24) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, bad);
  const uint64_t data_bytes = rd64(v);
  if (data_bytes > vlen - 24u) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, bad);
  const uint64_t n_in = rd64(v + 8 + data_bytes);
  if (n_in > (vlen - 24u - data_bytes) / 8u) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, bad);
  const uint8_t* offs = v + 8 + data_bytes + 8;
  auto input_entry = [&](uint64_t j) {
    const uint64_t o0 = rd64(offs + 8 * j), o1 = rd64(offs + 8 * (j + 1));
    if (o0 > o1 || o1 > data_bytes) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, bad);
    return std::make_pair(v + 8 + o0, o1 - o0);
  };
  const bool site_remapped = b.site_flags[s] & GDBGPU_SITE_REMAPPED;
  if (!site_remapped || !fi.dim0_allele_dependent()) {
    for (uint64_t j = 0; j < n_in; ++j) entries.push_back(input_entry(j));
    return true;
  }
  // merged -> input allele of this call
  const uint8_t *rp = nullptr, *ap = nullptr; uint32_t rlen = 0, alen = 0;
  if (m_raw_REF_fmt_index < 0 || m_raw_ALT_fmt_index < 0 || !raw_row(m_raw_REF_fmt_index, rp, rlen) || !raw_row(m_raw_ALT_fmt_index, ap, alen))
    throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "a call with a multi-dimensional field has no REF / ALT");
  const uint32_t first = b.allele_first[s], n_merged = b.n_alleles[s];
  auto merged = [&](uint32_t m) { return std::string((const char*)b.allele_bytes + b.allele_pos[2 * (first + m)], b.allele_pos[2 * (first + m) + 1]); };
  const std::string merged_ref = merged(0);
  const bool nonref_exists = n_merged > 1 && merged(n_merged - 1) == "<NON_REF>";
  std::vector<int64_t> m2in(n_merged, -1);
  m2in[0] = 0;
  int64_t input_nonref = -1;
  {
    const std::string alt((const char*)ap, alen);
    size_t pos = 0; int64_t input_idx = 1;
    while (pos <= alt.size() && !alt.empty()) {
      size_t e = alt.find('|', pos); if (e == std::string::npos) e = alt.size();
      std::string tok = alt.substr(pos, e - pos);
      if (tok == "&") input_nonref = input_idx;
      else {
        const bool symbolic = tok == "*" || (!tok.empty() && tok.front() == '<' && tok.back() == '>') || tok.find('[') != std::string::npos || tok.find(']') != std::string::npos;
        if (rlen < merged_ref.size() && !symbolic) tok.append(merged_ref, rlen, merged_ref.size() - rlen);
        for (uint32_t m = 1; m < n_merged; ++m) if (merged(m) == tok) { m2in[m] = input_idx; break; }
      }
      ++input_idx;
      if (e == alt.size()) break;
      pos = e + 1;
    }
  }
  if (nonref_exists && input_nonref >= 0) m2in[n_merged - 1] = input_nonref;
  const bool alt_only = fi.dim_lengths[0] == "A";
  const uint32_t length = alt_only ? n_merged - 1u : n_merged;
  const int64_t input_nr = nonref_exists ? m2in[n_merged - 1] : -1;
  for (uint32_t j = 0; j < length; ++j) {
    int64_t in = m2in[alt_only ? j + 1u : j];
    if (in < 0) in = input_nr;
    if (in < 0) { entries.emplace_back(v, 0); continue; }
    const int64_t input_j = alt_only ? in - 1 : in;
    if (input_j >= 0 && (uint64_t)input_j < n_in) entries.push_back(input_entry((uint64_t)input_j));
    else entries.emplace_back(v, 0);
  }
  return true;
}

// BroadCombinedGVCFOperator::operate (broad_combined_gvcf.cc:765-901) + bcf_write (htslib vcf.c) for one site
void BroadCombinedGVCFOperator::operate_on_bcf_site(const gdbgpu_batch& b, uint32_t s, const VariantQueryConfig& qc) {
  const uint64_t nq = b.n_query_rows, ns = b.n_sites;
  const bool sites_only = qc.sites_only_query();   // no samples in the header, no FORMAT fields (broad_combined_gvcf.cc:178-182,302-318)
  const int64_t begin = b.site_begin[s], end = b.site_end[s];
  if (b.site_flags[s] & GDBGPU_SITE_G_SUPPRESSED) {
    m_max_diploid_alt_alleles_that_can_be_genotyped = qc.get_max_diploid_alt_alleles_that_can_be_genotyped();
    warn_too_many_alleles(qc, begin, b.n_alleles[s] - 1u);
  }
  const ContigInfo* contig = m_vid_mapper->get_contig_location(begin);
  if (!contig) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Unknown contig for position " + std::to_string(begin));
  const int rid = m_vcf_adapter->get_contig_id(contig->name);
  const int32_t pos = (int32_t)(begin - contig->tiledb_column_offset);
  const uint32_t first = b.allele_first[s], n_allele = b.n_alleles[s];
  auto allele = [&](uint32_t m) { return std::string((const char*)b.allele_bytes + b.allele_pos[2 * (first + m)], b.allele_pos[2 * (first + m) + 1]); };
  std::string ref = allele(0);
  if (ref.size() == 1 && ref[0] == 'N') {   // broad_combined_gvcf.cc:811-816
    char base = m_vcf_adapter->get_reference_base_at_position(contig->name, pos);
    ref[0] = (base == 'A' || base == 'C' || base == 'G' || base == 'T') ? base : 'N';
  }
  std::vector<uint8_t> shared;
  struct FormatPiece { std::vector<uint8_t> head; const uint8_t* block = nullptr; size_t block_bytes = 0; };
  std::vector<FormatPiece> pieces;
  size_t indiv_bytes = 0;
  // ID: ';'-separated tokens of every live call's ID, each once (broad_combined_gvcf.cc:729-763,801-807).  The reference
  // collects them in a std::unordered_set and emits its iteration order; the same container filled in the same (row)
  // order reproduces it.
  m_ID_value.clear();
  if (m_ID_fmt_index >= 0 && b.fmt_type[(uint64_t)m_ID_fmt_index * ns + s]) {
    const uint64_t slot = (uint64_t)m_ID_fmt_index * ns + s;
    const uint32_t nps = b.fmt_nps[slot];
    const uint8_t* blk = b.fmt_bytes + b.fmt_offset[slot];
    std::unordered_set<std::string> id_set;
    for (uint64_t q = 0; q < nq; ++q) {
      const char* e = reinterpret_cast<const char*>(blk + q * nps);
      if (e[0] == 0x07 || e[0] == 0x00) continue;   // no call, or a call without ID
      size_t len = 0;
      while (len < nps && e[len] != 0x00) ++len;
      size_t last = 0;
      for (size_t i = 0; i < len; ++i) if (e[i] == ';') { id_set.insert(std::string(e + last, i - last)); last = i + 1; }
      if (len > last) id_set.insert(std::string(e + last, len - last));
    }
    for (const auto& str : id_set) m_ID_value += str + ';';
    if (!m_ID_value.empty()) m_ID_value.pop_back();
  }
  enc_vchar(shared, m_ID_value);
  enc_vchar(shared, ref);
  for (uint32_t m = 1; m < n_allele; ++m) enc_vchar(shared, allele(m));
  // FILTER: union of the live calls' filter indices, duplicates removed (broad_combined_gvcf.cc:845-873; the reference
  // walks a std::unordered_set<int> filled in row order and keeps the slots of ids the header does not know)
  bool have_filter = false;
  if (m_FILTER_fmt_index >= 0 && b.fmt_type[(uint64_t)m_FILTER_fmt_index * ns + s]) {
    const uint64_t slot = (uint64_t)m_FILTER_fmt_index * ns + s;
    const uint32_t nps = b.fmt_nps[slot]; const uint8_t type = b.fmt_type[slot];
    const uint8_t* blk = b.fmt_bytes + b.fmt_offset[slot];
    std::unordered_set<int> filter_idx_set;
    for (uint64_t q = 0; q < nq; ++q)
      for (uint32_t k = 0; k < nps; ++k) {
        int v; bool stop;
        if (type == BT_INT8) { int8_t x = (int8_t)blk[q * nps + k]; v = x; stop = x == INT8_MIN || x == INT8_MIN + 1; }
        else if (type == BT_INT16) { int16_t x; memcpy(&x, blk + 2 * (q * nps + k), 2); v = x; stop = x == INT16_MIN || x == INT16_MIN + 1; }
        else { int32_t x; memcpy(&x, blk + 4 * (q * nps + k), 4); v = x; stop = x == INT32_MIN || x == INT32_MIN + 1; }
        if (stop) break;
        filter_idx_set.insert(v);
      }
    if (!filter_idx_set.empty()) {
      m_FILTER_idx_vec.resize(filter_idx_set.size());
      size_t idx = 0;
      for (int g : filter_idx_set) {
        const int h = (g >= 0 && (size_t)g < m_global_field_idx_to_hdr_idx.size()) ? m_global_field_idx_to_hdr_idx[g] : -1;
        if (h >= 0) m_FILTER_idx_vec[idx++] = h;
      }
      enc_vint(shared, m_FILTER_idx_vec.data(), (uint32_t)m_FILTER_idx_vec.size());
      have_filter = true;
    }
  }
  if (!have_filter) shared.push_back(0x00);   // empty typed vector
  uint32_t n_info = 0;
  int32_t rlen = (int32_t)ref.size();
  if (end > begin) {   // handle_INFO_fields: interval -> END (1-based); bcf_update_info sets rlen from END
    const int32_t vcf_end = (int32_t)(end - contig->tiledb_column_offset + 1);
    enc_int1(shared, m_END_key); enc_int1(shared, vcf_end); ++n_info;
    rlen = vcf_end - pos;
  }
  std::vector<std::pair<const uint8_t*, uint64_t>> entries, entries2;
  for (auto& f : m_INFO_fields_vec) {
    if (f.raw_fmt_index >= 0) {   // element-wise sum of a 2-D field, stringified (variant_field_handler.cc:667-741)
      uint64_t num_valid = 0;
      std::vector<std::vector<uint32_t>> result;   // float sums are kept as floats below; the words carry either type
      for (uint64_t q = 0; q < nq; ++q) {
        if (!multid_entries(b, s, q, f.raw_fmt_index, *f.fi, entries)) continue;
        if (entries.size() > result.size()) result.resize(entries.size());
        for (size_t d = 0; d < entries.size(); ++d) {
          const uint64_t n = entries[d].second / 4;
          if (n > result[d].size()) result[d].resize(n, f.is_float ? 0x7F800001u : 0x80000000u);
          for (uint64_t i = 0; i < n; ++i) {
            uint32_t w; memcpy(&w, entries[d].first + 4 * i, 4);
            const bool w_ok = f.is_float ? (w != 0x7F800001u && w != 0x7F800002u) : (w != 0x80000000u && w != 0x80000001u);
            if (!w_ok) continue;
            uint32_t& r = result[d][i];
            const bool r_ok = f.is_float ? (r != 0x7F800001u && r != 0x7F800002u) : (r != 0x80000000u && r != 0x80000001u);
            if (!r_ok) r = w;
            else if (f.is_float) { float a, c; memcpy(&a, &r, 4); memcpy(&c, &w, 4); a += c; memcpy(&r, &a, 4); }
            else r = (uint32_t)((int32_t)r + (int32_t)w);
            ++num_valid;
          }
        }
      }
      if (!num_valid) continue;
      std::stringstream str;
      for (size_t d = 0; d < result.size(); ++d) {
        if (d) str << f.fi->vcf_delimiters[0];
        for (size_t i = 0; i < result[d].size(); ++i) {
          if (i) str << f.fi->vcf_delimiters[1];
          const uint32_t r = result[d][i];
          if (f.is_float) { if (r != 0x7F800001u && r != 0x7F800002u) { float a; memcpy(&a, &r, 4); str << std::fixed << std::setprecision(3) << a; } }
          else if (r != 0x80000000u && r != 0x80000001u) str << std::fixed << std::setprecision(3) << (int32_t)r;
        }
      }
      const std::string out_str = str.str();
      if (out_str.empty()) continue;   // bcf_update_info with n == 0 removes the tag
      enc_int1(shared, f.key); enc_vchar(shared, out_str);
      ++n_info;
      continue;
    }
    if (f.concat_fmt_index >= 0) {   // concatenate: the rows of the collated block, row order, padding dropped
      const uint64_t slot = (uint64_t)f.concat_fmt_index * ns + s;
      const uint8_t type = b.fmt_type[slot];
      if (!type) continue;
      const uint32_t nps = b.fmt_nps[slot];
      const uint8_t* blk = b.fmt_bytes + b.fmt_offset[slot];
      if (type == BT_CHAR) {
        std::string cat;
        for (uint64_t q = 0; q < nq; ++q) {
          const char* e = reinterpret_cast<const char*>(blk + q * nps);
          if (e[0] == 0x07 || e[0] == 0x00) continue;
          size_t len = 0; while (len < nps && e[len] != 0x00) ++len;
          cat.append(e, len);
        }
        if (cat.empty()) continue;
        enc_int1(shared, f.key); enc_vchar(shared, cat);
      } else {
        std::vector<int32_t> cat;   // floats travel as their bit patterns
        for (uint64_t q = 0; q < nq; ++q)
          for (uint32_t k = 0; k < nps; ++k) {
            int32_t v; bool miss, vend;
            if (type == BT_INT8) { int8_t x = (int8_t)blk[q * nps + k]; miss = x == INT8_MIN; vend = x == INT8_MIN + 1; v = miss ? INT32_MIN : x; }
            else if (type == BT_INT16) { int16_t x; memcpy(&x, blk + 2 * (q * nps + k), 2); miss = x == INT16_MIN; vend = x == INT16_MIN + 1; v = miss ? INT32_MIN : x; }
            else if (type == BT_FLOAT) { uint32_t x; memcpy(&x, blk + 4 * (q * nps + k), 4); miss = x == 0x7F800001u; vend = x == 0x7F800002u; v = (int32_t)x; }
            else { int32_t x; memcpy(&x, blk + 4 * (q * nps + k), 4); miss = x == INT32_MIN; vend = x == INT32_MIN + 1; v = x; }
            if (vend || (k == 0 && miss)) break;   // end of this row's vector / a row without the field
            cat.push_back(v);
          }
        if (cat.empty()) continue;
        enc_int1(shared, f.key);
        if (f.is_float) { enc_size(shared, (int)cat.size(), BT_FLOAT); put_bytes(shared, cat.data(), cat.size() * 4); }
        else enc_vint(shared, cat.data(), (uint32_t)cat.size());
      }
      ++n_info;
      continue;
    }
    if (f.vector_slot >= 0) {   // element-wise sum: typed vector (bcf_update_info -> bcf_enc_vint / bcf_enc_vfloat)
      const uint64_t vs = (uint64_t)f.vector_slot * ns + s;
      const uint32_t n = b.vec_len[vs];
      if (!n) continue;
      const uint32_t* v = b.vec_values + b.vec_offset[vs];
      enc_int1(shared, f.key);
      if (f.is_float) { enc_size(shared, (int)n, BT_FLOAT); put_bytes(shared, v, (size_t)n * 4); }
      else enc_vint(shared, reinterpret_cast<const int32_t*>(v), n);
      ++n_info;
      continue;
    }
    const uint64_t slot = (uint64_t)f.plain_index * ns + s;
    if (!b.info_valid[slot]) continue;
    enc_int1(shared, f.key);
    if (f.is_float) { enc_size(shared, 1, BT_FLOAT); put<uint32_t>(shared, b.info_value[slot]); }
    else enc_int1(shared, (int32_t)b.info_value[slot]);
    ++n_info;
  }
  for (auto& h : m_INFO_histogram_fields) {   // compute_valid_histogram_sum_2D_vector_and_stringify (broad_combined_gvcf.cc:432-520)
    const bool bin_float = h.fi[0]->type == "float", count_float = h.fi[1]->type == "float";
    uint64_t num_calls_with_field = 0;
    // bins ordered as std::map<T1, T2> orders them; counts summed in T2
    struct Count { int32_t i = 0; float f = 0.f; };
    std::vector<std::map<float, Count>> hf; std::vector<std::map<int32_t, Count>> hi;
    for (uint64_t q = 0; q < nq; ++q) {
      const bool v0 = multid_entries(b, s, q, h.raw_fmt_index[0], *h.fi[0], entries);
      const bool v1 = multid_entries(b, s, q, h.raw_fmt_index[1], *h.fi[1], entries2);
      if (!v0 || !v1) continue;
      const size_t nd = std::min(entries.size(), entries2.size());
      if (bin_float) { if (entries.size() > hf.size()) hf.resize(entries.size()); } else if (entries.size() > hi.size()) hi.resize(entries.size());
      for (size_t d = 0; d < nd; ++d) {
        const uint64_t n = std::min(entries[d].second, entries2[d].second) / 4;
        for (uint64_t i = 0; i < n; ++i) {
          uint32_t wb, wc; memcpy(&wb, entries[d].first + 4 * i, 4); memcpy(&wc, entries2[d].first + 4 * i, 4);
          const bool b_ok = bin_float ? (wb != 0x7F800001u && wb != 0x7F800002u) : (wb != 0x80000000u && wb != 0x80000001u);
          const bool c_ok = count_float ? (wc != 0x7F800001u && wc != 0x7F800002u) : (wc != 0x80000000u && wc != 0x80000001u);
          if (!b_ok || !c_ok) continue;
          Count add; if (count_float) memcpy(&add.f, &wc, 4); else add.i = (int32_t)wc;
          Count* slot; bool fresh;
          if (bin_float) { float key; memcpy(&key, &wb, 4); auto r = hf[d].insert(std::make_pair(key, add)); slot = &r.first->second; fresh = r.second; }
          else { auto r = hi[d].insert(std::make_pair((int32_t)wb, add)); slot = &r.first->second; fresh = r.second; }
          if (!fresh) { if (count_float) slot->f += add.f; else slot->i += add.i; }
        }
      }
      ++num_calls_with_field;
    }
    if (!num_calls_with_field) continue;
    std::stringstream str;
    const char d0 = h.fi[0]->vcf_delimiters[0], d1 = h.fi[0]->vcf_delimiters[1];
    const size_t n_outer = bin_float ? hf.size() : hi.size();
    for (size_t d = 0; d < n_outer; ++d) {
      if (d) str << d0;
      bool first = true;
      auto put_pair = [&](auto key, const Count& c) {
        if (!first) str << d1;
        str << std::fixed << std::setprecision(3) << key << d1;
        if (count_float) str << c.f; else str << c.i;
        first = false;
      };
      if (bin_float) for (auto& kv : hf[d]) put_pair(kv.first, kv.second); else for (auto& kv : hi[d]) put_pair(kv.first, kv.second);
    }
    const std::string out_str = str.str();
    if (out_str.empty()) continue;   // bcf_update_info with n == 0 removes the tag
    enc_int1(shared, h.key); enc_vchar(shared, out_str);
    ++n_info;
  }
  const bool remapped = b.site_flags[s] & GDBGPU_SITE_REMAPPED;
  if (m_dp_rule && b.info_dp_sum[s] > 0 && remapped) { enc_int1(shared, m_DP_key); enc_int1(shared, b.info_dp_sum[s]); ++n_info; }   // :721-725
  uint32_t n_fmt = 0;
  for (auto& f : m_FORMAT_fields_vec) {
    if (sites_only) break;
    const uint64_t slot = (uint64_t)f.fmt_index * ns + s;
    const uint8_t type = b.fmt_type[slot];
    if (!type) continue;
    const uint32_t nps = b.fmt_nps[slot];
    const size_t w = (type == BT_INT8 || type == BT_CHAR) ? 1 : (type == BT_INT16 ? 2 : 4);
    FormatPiece pc; pc.block = b.fmt_bytes + b.fmt_offset[slot]; pc.block_bytes = (size_t)nq * nps * w;
    enc_int1(pc.head, f.key);
    enc_size(pc.head, (int)nps, type);
    indiv_bytes += pc.head.size() + pc.block_bytes;
    pieces.push_back(std::move(pc));
    ++n_fmt;
  }
  std::vector<uint8_t> out;   // the fixed part of the record
  put<uint32_t>(out, (uint32_t)(24 + shared.size()));
  put<uint32_t>(out, (uint32_t)indiv_bytes);
  put<int32_t>(out, rid); put<int32_t>(out, pos); put<int32_t>(out, rlen);
  // QUAL: missing unless the vid mapping gives QUAL a combine operation (broad_combined_gvcf.cc:796-808)
  const bool has_qual = m_QUAL_plain >= 0 && b.info_valid[(uint64_t)m_QUAL_plain * ns + s];
  put<uint32_t>(out, has_qual ? b.info_value[(uint64_t)m_QUAL_plain * ns + s] : 0x7F800001u);
  put<uint32_t>(out, n_allele << 16 | n_info);
  put<uint32_t>(out, n_fmt << 24 | (sites_only ? 0u : (uint32_t)nq));
  put_bytes(out, shared.data(), shared.size());
  m_vcf_adapter->append(out.data(), out.size(), false);
  for (auto& pc : pieces) {   // typed key + size descriptor (owned), then the block where the batch holds it (borrowed)
    m_vcf_adapter->append(pc.head.data(), pc.head.size(), false);
    m_vcf_adapter->append(pc.block, pc.block_bytes, true);
  }
  ++m_num_records;
}

// ------------------------------------------------------------------------------------------------ storage manager
VariantStorageManager::VariantStorageManager(const std::string& workspace, unsigned, int cuda_device) : workspace_(workspace) {
  int rc = gdbgpu_ctx_create(cuda_device, &ctx_);
  if (rc != GDBGPU_OK) throw GenomicsDBGPUException(rc, "VariantStorageManager: no usable CUDA device " + std::to_string(cuda_device) + " (the combine path has no CPU fallback)");
}
VariantStorageManager::~VariantStorageManager() { gdbgpu_ctx_destroy(ctx_); }

void VariantQueryProcessorScanState::reset() {
  if (scan_) gdbgpu_scan_end(scan_);
  scan_ = nullptr; have_batch_ = false; pos_in_batch_ = 0; m_done = false; sites_done_ = 0;
}

// ------------------------------------------------------------------------------------------------ query processor
VariantQueryProcessor::VariantQueryProcessor(VariantStorageManager* sm, const std::string& array_name, const VidMapper& vid)
    : sm_(sm), array_name_(array_name), vid_(vid) {}
VariantQueryProcessor::~VariantQueryProcessor() { if (arr_) gdbgpu_array_close(arr_); }

void VariantQueryProcessor::check(int rc) const {
  if (rc == GDBGPU_OK) return;
  std::string m = gdbgpu_last_error(sm_->ctx());
  if (rc == GDBGPU_ERR_OVERLAP) throw VariantQueryProcessorException(rc, m);
  throw GenomicsDBGPUException(rc, m);
}

void VariantQueryProcessor::do_query_bookkeeping(VariantQueryConfig& qc, bool alleles_required, bool stage_broad_gvcf_fields, bool stage_columnar_fields) {
  const VidMapper& vid = qc.get_vid_mapper().is_initialized() ? qc.get_vid_mapper() : vid_;
  std::vector<gdbgpu_field_desc> fd;
  unsigned esum_mask = 0;
  qc.remapped_names_.clear();
  qc.remapped_names_.push_back("GT");
  // the order of the queried attributes as the reference's bookkeeping leaves it (query_variants.cc:578-637: END appended,
  // ALT / REF / GT pulled in by allele- and genotype-dependent fields; variant_query_config.cc:161-185: END, REF, ALT swapped
  // to the front) -- what the columnar iterator's print operators go by
  std::vector<std::string> ordered = qc.get_query_attributes();
  // VidMapper::add_mandatory_fields (vid_mapper.cc:611-668): QUAL and FILTER exist even when the vid mapping does not list them
  static const FieldInfo builtin_QUAL = [] { FieldInfo f; f.name = "QUAL"; f.type = "float"; f.length = "1"; return f; }();
  static const FieldInfo builtin_FILTER = [] { FieldInfo f; f.name = "FILTER"; f.type = "int"; f.length = "VAR"; return f; }();
  auto columnar_field_info = [&](const std::string& name) -> const FieldInfo* {
    const FieldInfo* fi = vid.get_field_info(name);
    if (fi) return fi;
    return name == "QUAL" ? &builtin_QUAL : (name == "FILTER" ? &builtin_FILTER : nullptr);
  };
  {   // VariantQueryConfig::flatten_composite_fields (variant_query_config.cc:187-231): the elements of a composite field join the
      // query at its end, the composite field leaves it
    std::vector<std::string> flat, kept;
    for (auto& name : ordered) {
      const FieldInfo* fi = vid.get_field_info(name);
      if (fi && fi->is_composite()) { for (size_t j = 0; j < fi->tuple_types.size(); ++j) flat.push_back(name + "_tuple_element_" + std::to_string(j)); }
      else kept.push_back(name);
    }
    for (auto& n : flat) if (std::find(kept.begin(), kept.end(), n) == kept.end()) kept.push_back(n);
    ordered.swap(kept);
  }
  {
    auto add = [&](const char* n) { if (std::find(ordered.begin(), ordered.end(), n) == ordered.end()) ordered.push_back(n); };
    add("END");
    if (alleles_required) { add("ALT"); add("REF"); }
    for (size_t i = 0; i < ordered.size(); ++i) {
      const FieldInfo* fi = columnar_field_info(ordered[i]);
      if (!fi) {
        if (ordered[i] == "END" || ordered[i] == "REF" || ordered[i] == "ALT" || !stage_columnar_fields) continue;
        throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Field " + ordered[i] + " not found in vid mapping");
      }
      if (fi->is_allele_dependent()) { add("ALT"); add("REF"); }
      if (fi->length == "G") add("GT");
    }
    size_t first_normal = 0;
    for (const char* special : {"END", "REF", "ALT"}) {
      auto it = std::find(ordered.begin(), ordered.end(), special);
      if (it == ordered.end()) continue;
      std::swap(*it, ordered[first_normal]);
      ++first_normal;
    }
  }
  qc.ordered_attributes_ = ordered;
  for (auto& name : ordered) {
    const FieldInfo* fi = vid.get_field_info(name);
    if (!fi || !fi->is_allele_dependent()) continue;
    gdbgpu_field_desc d;
    d.name = fi->name.c_str();
    d.length_kind = fi->length == "R" ? GDBGPU_LEN_R : (fi->length == "A" ? GDBGPU_LEN_A : GDBGPU_LEN_G);
    d.elem_type = fi->type == "float" ? GDBGPU_ELEM_FLOAT32 : GDBGPU_ELEM_INT32;
    if (stage_broad_gvcf_fields && fi->is_info && !fi->is_format && fi->combine_op != "move_to_FORMAT") {
      if (fi->combine_op == "element_wise_sum") esum_mask |= 1u << fd.size();
      else if (!fi->combine_op.empty()) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "INFO combine operation " + fi->combine_op + " of allele-dependent field " + name + " is not supported on the GPU");
    }
    fd.push_back(d);
    qc.remapped_names_.push_back(fi->name);
  }
  qc.esum_remapped_mask_ = esum_mask;
  const FieldInfo* gt = vid.get_field_info("GT");
  int phased = gt && gt->length == "PP";
  // what BroadCombinedGVCFOperator's constructor sorts the queried attributes into (broad_combined_gvcf.cc:160-240)
  qc.plain_.clear(); qc.dp_fields_[0] = qc.dp_fields_[1] = qc.dp_fields_[2] = 0;
  std::vector<gdbgpu_plain_field_desc> pd;
  if (stage_broad_gvcf_fields) {
    static const char* skip[] = {"END", "REF", "ALT", "GT"};
    bool need_raw_alleles = false;
    // FILTER (vid_mapper.cc:611-668 makes it an int vector even when the vid mapping does not list it): the union over the
    // live calls goes into bcf1_t.d.flt when the query asks for it (broad_combined_gvcf.cc:845-873)
    if (qc.produce_FILTER_field() && std::find(ordered.begin(), ordered.end(), "FILTER") != ordered.end()) {
      VariantQueryConfig::PlainField pf;
      pf.attribute = pf.vcf_name = "FILTER"; pf.elem = GDBGPU_ELEM_INT32; pf.fixed_len = 0; pf.format = true; pf.is_filter = true;
      qc.plain_.push_back(pf);
    }
    for (auto& name : ordered) {
      const FieldInfo* fi = vid.get_field_info(name);
      if (!fi || fi->is_allele_dependent()) continue;
      bool skipped = false;
      for (const char* k : skip) if (name == k) skipped = true;
      if (skipped) continue;
      if (fi->is_multid()) {
        // multi-dimensional INFO fields (allele-specific annotations): element-wise sum (variant_field_handler.cc:667-715) and
        // histogram sum over the two elements of a tuple (broad_combined_gvcf.cc:432-520).  The calls' raw values travel as rows
        // of a block (no_format = 3) and are combined by the operator: one short vector per call that has the field at all.
        if (!fi->is_info || (fi->combine_op != "element_wise_sum" && fi->combine_op != "histogram_sum")) {
          if (fi->is_info && fi->combine_op.empty()) continue;   // dropped with a warning, like any INFO field without an operation
          throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "BroadCombinedGVCFOperator on the GPU: multi-dimensional field " + name + " with operation '" + fi->combine_op + "' is not supported");
        }
        if (fi->dim_lengths.size() != 2) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "multi-dimensional field " + name + ": only 2 dimensions are supported");
        if (fi->combine_op == "histogram_sum") {
          if (fi->parent.empty() || fi->tuple_types.size() != 2)
            throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Operation histogram_sum is only supported for fields whose elements are tuple with 2 constituent elements; field " + fi->vcf_name() + " does not satisfy this requirement");
          if (fi->type != "int" && fi->type != "float")
            throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Operation histogram_sum is only supported for fields whose elements are tuple with 2 constituent elements that are int or float; field " + fi->vcf_name() + " does not satisfy this requirement");
        } else if (!fi->parent.empty()) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "element-wise sum of tuple field " + fi->vcf_name() + " is not supported");
        VariantQueryConfig::PlainField pf;
        pf.attribute = name; pf.vcf_name = fi->vcf_name(); pf.elem = GDBGPU_ELEM_CHAR; pf.fixed_len = 0; pf.format = true; pf.is_raw = true;
        qc.plain_.push_back(pf);
        if (fi->dim0_allele_dependent()) need_raw_alleles = true;
        continue;
      }
      const bool twin = name.size() > 7 && name.compare(name.size() - 7, 7, "_FORMAT") == 0;
      VariantQueryConfig::PlainField pf;
      pf.attribute = name; pf.vcf_name = twin ? name.substr(0, name.size() - 7) : name;
      pf.elem = fi->type == "char" ? GDBGPU_ELEM_CHAR : (fi->type == "float" ? GDBGPU_ELEM_FLOAT32 : GDBGPU_ELEM_INT32);
      bool numeric_len = !fi->length.empty() && fi->length.find_first_not_of("0123456789") == std::string::npos;
      pf.fixed_len = numeric_len ? std::stoi(fi->length) : 0;
      const bool as_format = (fi->is_format && (twin || !fi->is_info)) || (fi->is_info && fi->combine_op == "move_to_FORMAT");
      const bool scalar_op = fi->combine_op == "median" || fi->combine_op == "sum" || fi->combine_op == "mean";
      if (name == "FILTER") continue;
      if (name == "ID") {     // bcf1_t.id: the union of the live calls' IDs (merge_ID_field, broad_combined_gvcf.cc:729-763); collated like a string FORMAT field
        pf.elem = GDBGPU_ELEM_CHAR; pf.fixed_len = 0; pf.format = true; pf.is_id = true;
        qc.plain_.push_back(pf);
        continue;
      }
      if (name == "QUAL") {   // bcf1_t.qual through the field's combine operation, else missing (broad_combined_gvcf.cc:796-808)
        if (!scalar_op) continue;
        pf.is_qual = true; pf.fixed_len = 1; pf.elem = GDBGPU_ELEM_FLOAT32;
        pf.info_op = fi->combine_op == "median" ? GDBGPU_INFO_MEDIAN : (fi->combine_op == "sum" ? GDBGPU_INFO_SUM : GDBGPU_INFO_MEAN);
        qc.plain_.push_back(pf);
        continue;
      }
      // the known field DP in the FORMAT vector (FORMAT-only in the vid mapping, or INFO combined with the DP rule) is not inserted: its
      // values are the per-call inputs of INFO/DP (handle_FORMAT_fields, broad_combined_gvcf.cc:669-673,685-727)
      if (name == "DP" && (as_format || (fi->is_info && fi->combine_op == "DP"))) { /* DP rule input only */ }
      else if (as_format) pf.format = true;
      else if (fi->is_info && fi->combine_op == "element_wise_sum") {
        if (pf.elem == GDBGPU_ELEM_CHAR) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "element-wise sum of a string field: " + name);
        pf.info_op = GDBGPU_INFO_ELEMENT_WISE_SUM;
      }
      else if (fi->is_info && scalar_op) {
        if (pf.fixed_len != 1 || pf.elem == GDBGPU_ELEM_CHAR) throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "INFO combine operation on a non-scalar field: " + name);
        pf.info_op = fi->combine_op == "median" ? GDBGPU_INFO_MEDIAN : (fi->combine_op == "sum" ? GDBGPU_INFO_SUM : GDBGPU_INFO_MEAN);
      } else if (fi->is_info && fi->combine_op == "concatenate") {
        // the calls' vectors appended in row order (VariantFieldHandler::concatenate_field, variant_field_handler.cc:743-800): the
        // device collates them like a FORMAT field, the operator strings the rows of the block together
        pf.format = true; pf.is_concat = true;
      } else if (fi->is_info && !fi->combine_op.empty())
        throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "INFO combine operation " + fi->combine_op + " of field " + name + " is not supported on the GPU");
      else continue;   // INFO field without a combine operation: the reference drops it with a warning (broad_combined_gvcf.cc:175-178)
      qc.plain_.push_back(pf);
      if (name == "DP") qc.dp_fields_[0] = (unsigned)qc.plain_.size();
      if (name == "MIN_DP") qc.dp_fields_[1] = (unsigned)qc.plain_.size();
      if (name == "DP_FORMAT") qc.dp_fields_[2] = (unsigned)qc.plain_.size();
    }
    if (need_raw_alleles)   // the allele remap of those fields (remap_allele_specific_annotations, variant_operations.cc:482-555) is redone by the
      for (const char* n : {"REF", "ALT"}) {   // operator for the few calls that carry them: their REF / ALT strings come along the same way
        VariantQueryConfig::PlainField pf;
        pf.attribute = pf.vcf_name = n; pf.elem = GDBGPU_ELEM_CHAR; pf.fixed_len = 0; pf.format = true; pf.is_raw = true;
        qc.plain_.push_back(pf);
      }
    for (auto& pf : qc.plain_)
      pd.push_back(gdbgpu_plain_field_desc{pf.attribute.c_str(), pf.elem, pf.fixed_len, pf.is_raw ? 3 : (pf.is_concat ? 4 : ((pf.is_id || pf.is_filter) ? 2 : (pf.format ? 0 : 1))), pf.info_op});
  }
  qc.query_attrs_.clear();
  if (stage_columnar_fields) {
    if (stage_broad_gvcf_fields) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "do_query_bookkeeping: stage for the Broad operator or for the columnar iterator, not both");
    int n_remapped = 0;
    for (auto& name : ordered) {
      const FieldInfo* fi = columnar_field_info(name);
      const bool special = name == "END" || name == "REF" || name == "ALT" || name == "GT";
      VariantQueryConfig::QueryAttribute qa; qa.name = name;
      if (fi) qa.elem = fi->type == "char" ? 'c' : (fi->type == "float" ? 'f' : (fi->type == "flag" ? 'b' : 'i'));   // 'b': one byte, printed as bool
      if (!special && fi->type != "char" && fi->type != "float" && fi->type != "int" && fi->type != "flag") throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "columnar iterator on the GPU: type " + fi->type + " of field " + name + " is not supported");
      if (name == "END") { qa.kind = -1; qa.fixed_len = 1; }
      else if (name == "REF") { qa.kind = GDBGPU_CELL_FIELD_REF; qa.elem = 'c'; }
      else if (name == "ALT") { qa.kind = GDBGPU_CELL_FIELD_ALT; qa.elem = 'c'; }
      else if (name == "GT") { qa.kind = GDBGPU_CELL_FIELD_GT; qa.elem = 'i'; }
      else if (fi->is_allele_dependent()) { qa.kind = GDBGPU_CELL_FIELD_REMAPPED; qa.index = n_remapped++; }
      else {
        if (fi->is_multid()) { qa.elem = 'c'; qa.multid = true; }   // a multi-dimensional value is a byte string to the print operators (they print it raw)
        const bool numeric_len = !fi->length.empty() && fi->length.find_first_not_of("0123456789") == std::string::npos;
        qa.kind = GDBGPU_CELL_FIELD_PLAIN; qa.index = (int)qc.plain_.size(); qa.fixed_len = numeric_len ? std::stoi(fi->length) : 0;
        VariantQueryConfig::PlainField pf;
        pf.attribute = pf.vcf_name = name; pf.fixed_len = qa.fixed_len;
        pf.elem = (qa.elem == 'c' || qa.elem == 'b') ? GDBGPU_ELEM_CHAR : (qa.elem == 'f' ? GDBGPU_ELEM_FLOAT32 : GDBGPU_ELEM_INT32);
        qc.plain_.push_back(pf);
      }
      qc.query_attrs_.push_back(qa);
    }
    for (auto& pf : qc.plain_) pd.push_back(gdbgpu_plain_field_desc{pf.attribute.c_str(), pf.elem, pf.fixed_len, 1, 0});
  }
  if (arr_) { gdbgpu_array_close(arr_); arr_ = nullptr; }
  check(gdbgpu_array_open_ex(sm_->ctx(), sm_->get_workspace().c_str(), array_name_.c_str(), fd.data(), (int32_t)fd.size(), pd.data(), (int32_t)pd.size(), phased, &arr_));
  gdbgpu_array_info info;
  check(gdbgpu_array_get_info(arr_, &info));
  if (!info.has_gt) qc.remapped_names_.erase(qc.remapped_names_.begin());
  n_out_ = (info.has_gt ? 1 : 0) + info.n_fields;
  coords_.resize(2 * info.n_cells); end_.resize(info.n_cells);
  if (info.n_cells) check(gdbgpu_array_copy_cells(arr_, 0, info.n_cells, coords_.data(), end_.data()));
  qc.bookkeeping_done_ = true;
}

void VariantQueryProcessor::fill_variant(const gdbgpu_batch& b, uint32_t s, Variant& v) const {
  const uint64_t nq = b.n_query_rows;
  v.begin_ = b.site_begin[s]; v.end_ = b.site_end[s];
  v.nonref_ = b.site_flags[s] & GDBGPU_SITE_NONREF; v.remapped_ = b.site_flags[s] & GDBGPU_SITE_REMAPPED; v.g_suppressed_ = b.site_flags[s] & GDBGPU_SITE_G_SUPPRESSED;
  const uint32_t first = b.allele_first[s], na = b.n_alleles[s];
  auto allele = [&](uint32_t m) { return std::string((const char*)b.allele_bytes + b.allele_pos[2 * (first + m)], b.allele_pos[2 * (first + m) + 1]); };
  v.merged_ref_ = allele(0);
  v.merged_alts_.resize(na - 1);
  for (uint32_t m = 1; m < na; ++m) v.merged_alts_[m - 1] = allele(m);
  v.calls_.resize(nq);
  for (uint64_t q = 0; q < nq; ++q) {
    VariantCall& c = v.calls_[q];
    const int32_t e = b.live[(uint64_t)s * nq + q];
    c.valid_ = e >= 0;
    if (!c.valid_) continue;
    const uint32_t cell = (uint32_t)(e & (GDBGPU_CALL_REDUCED - 1));
    c.reduced_ = (e & GDBGPU_CALL_REDUCED) != 0;
    const int64_t row = coords_[2 * (size_t)cell], col = coords_[2 * (size_t)cell + 1], END = end_[cell];
    c.row_ = (uint64_t)row;
    if (col > END) { c.begin_ = END; c.end_ = col; } else { c.begin_ = col; c.end_ = END; }
    c.fields_.resize(n_out_);
    for (int o = 0; o < n_out_; ++o) {
      const gdbgpu_field_out& f = b.fields[o];
      c.fields_[o] = {f.data + f.base[s] + q * f.stride[s], (unsigned)f.len[(uint64_t)s * nq + q]};
    }
    if (b.lut_first) { c.lut_ = b.lut_in2m + b.lut_first[(uint64_t)s * nq + q]; c.lut_len_ = b.lut_len[(uint64_t)s * nq + q]; }
    else { c.lut_ = nullptr; c.lut_len_ = 0; }
  }
}

// query_variants.cc:334-476: one call processes batches until the interval is exhausted or the operator overflows;
// with a scan_state the caller re-enters after draining its output, exactly as with the reference.
void VariantQueryProcessor::scan_and_operate(const int, const VariantQueryConfig& qc, SingleVariantOperatorBase& op, unsigned column_interval_idx,
                                             bool handle_spanning_deletions, VariantQueryProcessorScanState* scan_state, bool want_lut) const {
  if (!qc.is_bookkeeping_done()) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "scan_and_operate: do_query_bookkeeping() must be called first");
  VariantQueryProcessorScanState local;
  VariantQueryProcessorScanState& st = scan_state ? *scan_state : local;
  if (!st.scan_) {
    gdbgpu_scan_params p; memset(&p, 0, sizeof(p));
    if (qc.get_num_column_intervals() == 0) p.flags |= GDBGPU_SCAN_FULL;
    else { p.col_lo = qc.get_column_begin(column_interval_idx); p.col_hi = qc.get_column_end(column_interval_idx); }
    if (!handle_spanning_deletions) p.flags |= GDBGPU_SCAN_NO_DELETIONS;
    if (qc.produce_GT_with_min_PL_value_for_spanning_deletions()) p.flags |= GDBGPU_SCAN_MIN_PL_GT;
    if (want_lut) p.flags |= GDBGPU_SCAN_WANT_LUT;
    if (op.wants_site_summaries()) p.flags |= GDBGPU_SCAN_SITE_STATS | GDBGPU_SCAN_SUMMARY_ONLY;
    if (op.wants_bcf_blocks()) {
      p.flags |= GDBGPU_SCAN_BCF_FORMAT;
      if (qc.produce_GT_field()) p.flags |= GDBGPU_SCAN_PRODUCE_GT;
      if (op.wants_htsjdk_conventions()) p.flags |= GDBGPU_SCAN_HTSJDK_CONVENTIONS;
      p.dp_info_field = qc.dp_fields_[0]; p.dp_min_dp_field = qc.dp_fields_[1]; p.dp_format_field = qc.dp_fields_[2];
      p.esum_remapped_mask = qc.esum_remapped_mask_;
    }
    if (!qc.query_all_rows()) { p.query_rows = qc.get_rows_to_query().data(); p.n_query_rows = qc.get_rows_to_query().size(); }
    p.max_diploid_alt_alleles = qc.get_max_diploid_alt_alleles_that_can_be_genotyped();
    check(gdbgpu_scan_begin(arr_, &p, &st.scan_));
    st.have_batch_ = false; st.pos_in_batch_ = 0; st.m_done = false;
  }
  Variant variant;
  for (;;) {
    if (!st.have_batch_ || st.pos_in_batch_ >= st.batch_.n_sites) {
      op.release_batch();   // nothing may point into the batch that is about to be replaced
      check(gdbgpu_scan_next_batch(st.scan_, 0, 0, &st.batch_));
      st.have_batch_ = true; st.pos_in_batch_ = 0;
      if (st.batch_.n_sites == 0) { st.m_done = true; gdbgpu_scan_end(st.scan_); st.scan_ = nullptr; st.have_batch_ = false; return; }
    }
    while (st.pos_in_batch_ < st.batch_.n_sites) {
      if (op.overflow()) { if (!scan_state) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "operator overflow without a scan state"); return; }   // caller drains, then re-enters
      if (op.wants_bcf_blocks()) op.operate_on_bcf_site(st.batch_, st.pos_in_batch_, qc);
      else if (op.wants_site_summaries()) op.operate_on_site_summary(st.batch_, st.pos_in_batch_, qc);
      else { fill_variant(st.batch_, st.pos_in_batch_, variant); op.operate(variant, qc); }
      ++st.pos_in_batch_; ++st.sites_done_;
    }
  }
}

// ------------------------------------------------------------------------------------------------ per-site summary operators
void InterestingLocationsPrinter::operate(Variant& variant, const VariantQueryConfig&) {   // the materialised-Variant path (not used by scan_and_operate)
  uint64_t num_valid_calls = 0, num_begin_at_position = 0;
  for (auto it = variant.begin(); it != variant.end(); ++it) { ++num_valid_calls; if ((*it).get_column_begin() == variant.get_column_begin()) ++num_begin_at_position; }
  (*m_fptr) << variant.get_column_begin() << " " << num_valid_calls << " ? " << num_begin_at_position << "\n";
}
void InterestingLocationsPrinter::operate_on_site_summary(const gdbgpu_batch& b, uint32_t s, const VariantQueryConfig&) {
  (*m_fptr) << b.site_begin[s] << " " << b.n_live[s] << " " << b.site_stats[2 * s] << " " << b.site_stats[2 * s + 1] << "\n";
}
void MaxAllelesCountOperator::operate_on_site_summary(const gdbgpu_batch& b, uint32_t s, const VariantQueryConfig&) {
  ++m_total_lines;
  const uint32_t n_alt = b.n_alleles[s] - 1;
  auto more = [](const AlleleTracker& l, const AlleleTracker& r) { return l.alt.size() > r.alt.size(); };
  if (m_heap.size() < m_top_count || n_alt > m_heap.front().alt.size()) {
    if (m_heap.size() >= m_top_count) { std::pop_heap(m_heap.begin(), m_heap.end(), more); m_heap.pop_back(); }
    const uint32_t first = b.allele_first[s];
    auto allele = [&](uint32_t m) { return std::string((const char*)b.allele_bytes + b.allele_pos[2 * (first + m)], b.allele_pos[2 * (first + m) + 1]); };
    AlleleTracker t; t.ref = allele(0); t.column = b.site_begin[s];
    for (uint32_t m = 1; m <= n_alt; ++m) t.alt.push_back(allele(m));
    m_heap.push_back(std::move(t));
    std::push_heap(m_heap.begin(), m_heap.end(), more);
  }
}
void MaxAllelesCountOperator::print(std::ostream& fptr) const {
  fptr << "TOTAL " << m_total_lines << "\n";
  auto heap = m_heap;
  auto more = [](const AlleleTracker& l, const AlleleTracker& r) { return l.alt.size() > r.alt.size(); };
  while (!heap.empty()) {
    std::pop_heap(heap.begin(), heap.end(), more);
    const AlleleTracker& t = heap.back();
    fptr << t.column << "," << t.ref << "," << t.alt.size();
    for (auto& a : t.alt) fptr << "," << a;
    fptr << "\n";
    heap.pop_back();
  }
}

// ------------------------------------------------------------------------------------------------ columnar iterator surface
namespace {
const std::string g_json_indent_unit = "    ";
inline size_t elem_size(char e) { return (e == 'c' || e == 'b') ? 1u : 4u; }
template <class T> void print_list(std::ostream& f, const uint8_t* p, size_t n) {
  const T* d = reinterpret_cast<const T*>(p);
  f << "[ " << d[0];
  for (size_t i = 1; i < n; ++i) f << ", " << d[i];
  f << " ]";
}
template <class T> void print_csv_list(std::ostream& f, const uint8_t* p, size_t n, bool variable, bool valid) {
  if (variable) f << n;
  if (valid) {
    if (variable) f << ",";
    const T* d = reinterpret_cast<const T*>(p);
    f << d[0];
    for (size_t i = 1; i < n; ++i) f << "," << d[i];
  } else if (!variable)
    for (size_t i = 1; i < n; ++i) f.put(',');
}
}  // namespace

const uint8_t* GenomicsDBColumnarCell::get_field_ptr_for_query_idx(unsigned i) const {
  const auto& qa = qc_->get_query_attribute(i);
  if (qa.kind < 0) return reinterpret_cast<const uint8_t*>(&end_);
  const Column& c = (*columns_)[i];
  return c.bytes + c.off[idx_];
}
size_t GenomicsDBColumnarCell::get_field_length(unsigned i) const {
  const auto& qa = qc_->get_query_attribute(i);
  if (qa.kind < 0) return 1;
  const Column& c = (*columns_)[i];
  return (size_t)(c.off[idx_ + 1] - c.off[idx_]) / elem_size(qa.elem);
}
// genomicsdb_columnar_field.cc:358-377: variable-length fields are valid when non-empty, fixed-length ones when any element
// is not TileDB's empty value
bool GenomicsDBColumnarCell::is_valid(unsigned i) const {
  const auto& qa = qc_->get_query_attribute(i);
  if (qa.kind < 0) return true;
  const size_t n = get_field_length(i);
  if (qa.fixed_len == 0) return n > 0;
  const uint8_t* p = get_field_ptr_for_query_idx(i);
  for (size_t k = 0; k < n; ++k) {
    if (qa.elem == 'c' || qa.elem == 'b') { if (p[k] != 127) return true; }
    else { uint32_t w; memcpy(&w, p + 4 * k, 4); if (w != (qa.elem == 'f' ? 0x7F7FFFFFu : 0x7FFFFFFFu)) return true; }
  }
  return false;
}

// variant_cell.cc:119-165 + genomicsdb_columnar_field.cc:102-196,386-420
void GenomicsDBColumnarCell::print(std::ostream& fptr, const VariantQueryConfig* query_config, const std::string& indent_prefix, const VidMapper* vid_mapper) const {
  std::string indent_string = indent_prefix + g_json_indent_unit;
  fptr << indent_prefix << "{\n";
  fptr << indent_string << "\"row\": " << coords_[0] << ",\n";
  fptr << indent_string << "\"interval\": [ " << coords_[1] << ", " << end_ << " ],\n";
  if (vid_mapper) {
    const ContigInfo* contig = vid_mapper->get_contig_location(coords_[1]);
    if (contig) {
      const int64_t contig_position = coords_[1] - contig->tiledb_column_offset;
      fptr << indent_string << "\"genomic_interval\": { \"" << contig->name << "\" : [ " << contig_position + 1 << ", " << (contig_position + 1 + (end_ - coords_[1])) << " ] },\n";
    }
  }
  fptr << indent_string << "\"fields\": {\n";
  indent_string += g_json_indent_unit;
  bool first_valid_field = true;
  for (unsigned i = 1; i < query_config->get_num_queried_attributes(); ++i) {   // the first field is always END
    if (!is_valid(i)) continue;
    if (!first_valid_field) fptr << ",\n";
    const auto& qa = query_config->get_query_attribute(i);
    fptr << indent_string << "\"" << qa.name << "\": ";
    const uint8_t* p = get_field_ptr_for_query_idx(i);
    const size_t n = get_field_length(i);
    if (qa.kind == GDBGPU_CELL_FIELD_ALT) {   // '|'-separated, '&' stands for <NON_REF>
      fptr << "[ ";
      size_t t0 = 0; bool first = true;
      for (;;) {
        size_t t1 = t0;
        while (t1 < n && p[t1] != '|') ++t1;
        if (!first) fptr << ", ";
        fptr << "\"";
        if (t1 - t0 == 1 && p[t0] == '&') fptr << "<NON_REF>"; else fptr.write(reinterpret_cast<const char*>(p + t0), t1 - t0);
        fptr << "\"";
        first = false;
        if (t1 >= n) break;
        t0 = t1 + 1;
      }
      fptr << " ]";
    } else if (qa.elem == 'b') fptr << (p[0] != 0);   // BCF_HT_FLAG: GenomicsDBColumnarFieldPrintOperator<bool, false>
    else if (qa.elem == 'c') {
      if (qa.fixed_len == 1) fptr << (char)p[0];
      else { fptr << "\""; fptr.write(reinterpret_cast<const char*>(p), n); fptr << "\""; }
    } else if (qa.fixed_len == 1) {
      if (qa.elem == 'f') fptr << *reinterpret_cast<const float*>(p); else fptr << *reinterpret_cast<const int32_t*>(p);
    } else if (qa.elem == 'f') print_list<float>(fptr, p, n);
    else print_list<int32_t>(fptr, p, n);
    first_valid_field = false;
  }
  indent_string = indent_prefix + g_json_indent_unit;
  fptr << "\n" << indent_string << "}\n" << indent_prefix << "}";
}

// variant_cell.cc:167-185
void GenomicsDBColumnarCell::print_csv(std::ostream& fptr, const VariantQueryConfig* query_config) const {
  fptr << coords_[0] << "," << coords_[1] << "," << end_;
  for (unsigned i = 1; i < query_config->get_num_queried_attributes(); ++i) {
    fptr << ",";
    const auto& qa = query_config->get_query_attribute(i);
    const uint8_t* p = get_field_ptr_for_query_idx(i);
    const size_t n = get_field_length(i);
    const bool valid = is_valid(i), variable = qa.fixed_len == 0;
    if (qa.elem == 'b') { if (valid) fptr << (p[0] != 0); }
    else if (qa.elem == 'c') {
      if (qa.fixed_len == 1) { if (valid) fptr << (char)p[0]; }
      else if (valid) fptr.write(reinterpret_cast<const char*>(p), n);
    } else if (qa.fixed_len == 1) {
      if (valid) { if (qa.elem == 'f') fptr << *reinterpret_cast<const float*>(p); else fptr << *reinterpret_cast<const int32_t*>(p); }
    } else if (qa.elem == 'f') print_csv_list<float>(fptr, p, n, variable, valid);
    else print_csv_list<int32_t>(fptr, p, n, variable, valid);
  }
  fptr << "\n";
}

// variant_operations.cc:732-790
ColumnHistogramOperator::ColumnHistogramOperator(uint64_t begin, uint64_t end, uint64_t bin_size) : m_begin_column(begin), m_end_column(end), m_bin_size(bin_size) {
  if (end < begin || !bin_size) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "ColumnHistogramOperator: bad range");
  m_bin_counts_vector.assign((end - begin) / bin_size + 1, 0);
}
void ColumnHistogramOperator::operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig&, const VariantArraySchema&) {
  const uint64_t b = (uint64_t)cell.get_coordinates()[1];
  ++m_bin_counts_vector[b <= m_begin_column ? 0 : (b >= m_end_column ? m_bin_counts_vector.size() - 1 : (b - m_begin_column) / m_bin_size)];
}
void ColumnHistogramOperator::operate_on_device_interval(gdbgpu_scan* scan, unsigned, const VariantQueryConfig&) {
  m_scratch.resize(m_bin_counts_vector.size());
  if (gdbgpu_scan_cells_histogram(scan, m_begin_column, m_end_column, m_bin_size, m_scratch.data(), m_scratch.size()) != GDBGPU_OK)
    throw GenomicsDBGPUException(GDBGPU_ERR_CUDA, "ColumnHistogramOperator: the histogram kernel failed");
  for (size_t i = 0; i < m_scratch.size(); ++i) m_bin_counts_vector[i] += m_scratch[i];
}
bool ColumnHistogramOperator::equi_partition_and_print_bins(uint64_t num_bins, std::ostream& fptr) const {
  const size_t n = m_bin_counts_vector.size();
  if (num_bins >= n) { std::cerr << "Requested #equi bins is smaller than allocated bin counts vector, returning\n"; return false; }
  uint64_t total = 0;
  for (auto v : m_bin_counts_vector) total += v;
  const double per_bin = (double)total / num_bins;
  char line[128];
  snprintf(line, sizeof(line), "Total %llu #bins %llu count/bins %.1f\n", (unsigned long long)total, (unsigned long long)num_bins, per_bin);
  fptr << line;
  for (size_t i = 0; i < n;) {   // merge bins left to right until a group holds its share
    size_t j = i; uint64_t group = 0;
    while ((double)group < per_bin && j < n) group += m_bin_counts_vector[j++];
    fptr << m_begin_column + i * m_bin_size << "," << m_begin_column + j * m_bin_size - 1 << "," << group << "\n";
    if (j == i) break;
    i = j;
  }
  fptr << "\n";
  return true;
}

// variant_operations.cc:808-846
VariantCallPrintOperator::VariantCallPrintOperator(std::ostream& fptr, const std::string& indent_prefix, const VidMapper* vid_mapper)
    : m_indent_prefix(indent_prefix), m_indent_prefix_plus_one(indent_prefix + g_json_indent_unit), m_indent_prefix_plus_two(indent_prefix + g_json_indent_unit + g_json_indent_unit),
      m_fptr(&fptr), m_vid_mapper(vid_mapper && vid_mapper->is_initialized() ? vid_mapper : nullptr) {}
void VariantCallPrintOperator::operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig& query_config, const VariantArraySchema&) {
  std::ostream& f = *m_fptr;
  if (cell.at_new_query_column_interval()) {
    if (m_num_query_intervals_printed > 0) { f << "\n" << m_indent_prefix_plus_one << "]\n" << m_indent_prefix << "},\n"; }
    m_num_calls_printed = 0;
    f << m_indent_prefix << "{\n";
    const auto idx = cell.get_current_query_column_interval_idx();
    const long long begin = query_config.get_num_column_intervals() > 0 ? query_config.get_column_begin(idx) : 0ll;
    const long long end = query_config.get_num_column_intervals() > 0 ? query_config.get_column_end(idx) : INT64_MAX - 1ll;
    f << m_indent_prefix_plus_one << "\"query_interval\": [ " << begin << ", " << end << " ],\n";
    f << m_indent_prefix_plus_one << "\"variant_calls\": [\n";
  }
  if (m_num_calls_printed > 0) f << ",\n";
  cell.print(f, &query_config, m_indent_prefix_plus_two, m_vid_mapper);
  ++m_num_calls_printed;
  ++m_num_query_intervals_printed;
}
void VariantCallPrintOperator::finalize() {
  if (m_num_query_intervals_printed > 0) (*m_fptr) << "\n" << m_indent_prefix_plus_one << "]\n" << m_indent_prefix << "}";
}

// variant_operations.cc:900-1089; the counting is gdbgpu_scan_cells_allele_counts (kernels + one sort)
AlleleCountOperator::AlleleCountOperator(const VidMapper&, const VariantQueryConfig& query_config) {
  unsigned idx;
  for (const char* need : {"GT", "REF", "ALT"})
    if (!query_config.get_query_idx_for_name(need, idx)) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, std::string(need) + " field must be queried for AlleleCountOperator");
}
void AlleleCountOperator::operate_on_device_interval(gdbgpu_scan* scan, unsigned, const VariantQueryConfig&) {
  gdbgpu_allele_counts ac;
  if (gdbgpu_scan_cells_allele_counts(scan, &ac) != GDBGPU_OK) throw GenomicsDBGPUException(GDBGPU_ERR_CUDA, "AlleleCountOperator: the counting kernels failed");
  for (uint64_t k = 0; k < ac.n; ++k)
    m_records.push_back(Record{ac.column[k], std::string((const char*)ac.bytes + ac.ref_pos[2 * k], ac.ref_pos[2 * k + 1]),
                               std::string((const char*)ac.bytes + ac.alt_pos[2 * k], ac.alt_pos[2 * k + 1]), ac.count[k]});
}
void AlleleCountOperator::print_allele_counts(std::ostream& fptr) const {
  for (auto& r : m_records) fptr << r.column << " " << r.REF << " " << r.ALT << " " << r.count << "\n";
}

void VariantQueryProcessor::iterate_over_cells(const int, const VariantQueryConfig& qc, SingleCellOperatorBase& op, const bool) const {
  if (!qc.is_bookkeeping_done() || qc.get_num_queried_attributes() == 0)
    throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "iterate_over_cells: do_query_bookkeeping(config, alleles_required, false, true) must be called first");
  const unsigned n_intervals = std::max(1u, qc.get_num_column_intervals());
  for (unsigned i = 0; i < n_intervals; ++i) iterate_over_cells_of_interval(qc, op, i);
  op.finalize();
}
void VariantQueryProcessor::iterate_over_cells_of_interval(const VariantQueryConfig& qc, SingleCellOperatorBase& op, unsigned i) const {
  gdbgpu_scan_params p; memset(&p, 0, sizeof(p));
  p.flags = GDBGPU_SCAN_CELLS_ONLY;
  if (qc.get_num_column_intervals() == 0) p.flags |= GDBGPU_SCAN_FULL;
  else { p.col_lo = qc.get_column_begin(i); p.col_hi = qc.get_column_end(i); }
  if (!qc.query_all_rows()) { p.query_rows = qc.get_rows_to_query().data(); p.n_query_rows = qc.get_rows_to_query().size(); }
  gdbgpu_scan* scan = nullptr;
  check(gdbgpu_scan_begin(arr_, &p, &scan));
  struct Guard { gdbgpu_scan* s; ~Guard() { gdbgpu_scan_end(s); } } guard{scan};
  if (op.runs_on_device()) { op.operate_on_device_interval(scan, i, qc); return; }
  // the list is paged through (cells and their columns cross PCIe a window at a time, as scan_and_operate's batches do)
  static const uint64_t window = [] { const char* e = getenv("GDBGPU_CELLS_WINDOW"); uint64_t w = e ? strtoull(e, nullptr, 10) : 0; return w ? w : (1ull << 20); }();
  std::vector<GenomicsDBColumnarCell::Column> columns(qc.get_num_queried_attributes());
  GenomicsDBColumnarCell cell;
  cell.qc_ = &qc; cell.columns_ = &columns; cell.interval_idx_ = i;
  for (uint64_t first = 0;; first += window) {
    gdbgpu_cell_list list;
    check(gdbgpu_scan_cells_window(scan, first, window));
    check(gdbgpu_scan_cells(scan, &list));
    if (!list.n_cells) break;
    for (unsigned a = 0; a < qc.get_num_queried_attributes(); ++a) {
      const auto& qa = qc.get_query_attribute(a);
      if (qa.kind < 0) continue;   // END comes with the list
      check(gdbgpu_scan_cells_fetch(scan, qa.kind, qa.index, &columns[a].off, &columns[a].bytes));
    }
    for (uint64_t k = 0; k < list.n_cells; ++k) {
      cell.idx_ = k; cell.coords_[0] = list.row[k]; cell.coords_[1] = list.begin[k]; cell.end_ = list.end[k];
      cell.at_new_interval_ = first == 0 && k == 0;
      op.operate_on_columnar_cell(cell, qc, schema_);
    }
    if (first + list.n_cells >= list.n_total) break;
  }
}

// ------------------------------------------------------------------------------------------------ GA4GH range query
namespace {
// a field of a call as VariantFieldBase::print gives it (variant_field_data.h:211,257,540-556 and the vector printers): strings
// quoted, ALT as [ "A","<NON_REF>" ], numbers as [ a,b ] with floats in the stream's fixed notation
std::string ga4gh_number_list(const uint8_t* p, size_t n, char elem) {
  std::ostringstream o;
  o << std::fixed << std::setprecision(6);
  o << "[ ";
  for (size_t i = 0; i < n; ++i) {
    if (i) o << ",";
    if (elem == 'f') { float v; memcpy(&v, p + 4 * i, 4); o << v; }
    else if (elem == 'b') o << (int)(p[i] != 0);
    else { int32_t v; memcpy(&v, p + 4 * i, 4); o << v; }
  }
  o << " ]";
  return o.str();
}
std::string ga4gh_alt_list(const std::vector<std::string>& alt) {
  std::string o = "[ ";
  for (size_t i = 0; i < alt.size(); ++i) { if (i) o += ","; o += "\"" + (alt[i] == "&" ? std::string("<NON_REF>") : alt[i]) + "\""; }
  return o + " ]";
}
class GA4GHCallCollector : public SingleCellOperatorBase {
 public:
  GA4GHCallCollector(std::vector<GA4GHVariant>& variants) : m_variants(&variants) {}
  void operate_on_columnar_cell(const GenomicsDBColumnarCell& cell, const VariantQueryConfig& qc, const VariantArraySchema&) override {
    GA4GHVariantCall call;
    call.row = (uint64_t)cell.get_coordinates()[0]; call.begin = cell.get_coordinates()[1]; call.end = cell.get_end_position();
    const unsigned n = qc.get_num_queried_attributes();
    call.fields.assign(n, std::string());
    bool have_ref = false, have_alt = false;
    for (unsigned i = 0; i < n; ++i) {
      const auto& qa = qc.get_query_attribute(i);
      if (qa.kind == -1 || !cell.is_valid(i)) continue;   // END is the call's interval, not a field
      const uint8_t* p = cell.get_field_ptr_for_query_idx(i);
      const size_t len = cell.get_field_length(i);
      if (qa.kind == GDBGPU_CELL_FIELD_REF) { call.REF.assign((const char*)p, len); call.fields[i] = "\"" + call.REF + "\""; have_ref = true; }
      else if (qa.kind == GDBGPU_CELL_FIELD_ALT) {
        const std::string alt((const char*)p, len);
        for (size_t t0 = 0;;) { const size_t t1 = alt.find('|', t0); call.ALT.push_back(alt.substr(t0, t1 == std::string::npos ? std::string::npos : t1 - t0)); if (t1 == std::string::npos) break; t0 = t1 + 1; }
        call.fields[i] = ga4gh_alt_list(call.ALT); have_alt = true;
      } else if (qa.multid) {   // a vector of bytes to the reference's printer: [ 24,0,0, ... ]
        std::string o = "[ ";
        for (size_t k = 0; k < len; ++k) { if (k) o += ","; o += std::to_string((unsigned)p[k]); }
        call.fields[i] = o + " ]";
      } else if (qa.elem == 'c') call.fields[i] = "\"" + std::string((const char*)p, len) + "\"";
      else call.fields[i] = ga4gh_number_list(p, len, qa.elem);
    }
    // GA4GHCallInfoToVariantIdx::find_or_insert (variant.cc:26-70): begin, end, REF, the set of ALT alleles; a call without REF / ALT
    // in the query always opens a variant of its own
    uint64_t idx = m_variants->size();
    bool fresh = true;
    if (have_ref && have_alt) {
      Key key{call.begin, call.end, call.REF, std::set<std::string>(call.ALT.begin(), call.ALT.end())};
      auto it = m_index.find(key);
      if (it == m_index.end()) m_index.emplace(std::move(key), idx); else { fresh = false; idx = it->second; }
    }
    if (fresh) m_variants->emplace_back();
    GA4GHVariant& v = (*m_variants)[idx];
    v.begin = call.begin; v.end = call.end;
    v.calls.push_back(std::move(call));
  }
 private:
  typedef std::tuple<int64_t, int64_t, std::string, std::set<std::string>> Key;
  std::vector<GA4GHVariant>* m_variants;
  std::map<Key, uint64_t> m_index;
};
}  // namespace

void VariantQueryProcessor::gt_get_column_interval(const int, const VariantQueryConfig& qc, unsigned column_interval_idx, std::vector<GA4GHVariant>& variants) const {
  if (!qc.is_bookkeeping_done() || qc.get_num_queried_attributes() == 0)
    throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "gt_get_column_interval: do_query_bookkeeping(config, true, false, true) must be called first");
  const size_t start_variant_idx = variants.size();
  GA4GHCallCollector collector(variants);   // the grouping is per interval (call_info_2_variant is local to the reference's function)
  iterate_over_cells_of_interval(qc, collector, column_interval_idx);
  for (size_t i = start_variant_idx; i < variants.size(); ++i)
    if (variants[i].get_num_calls() > 1u) remap_ga4gh_variant(qc, variants[i]);   // possible re-arrangement of PL / AD / GT (query_variants.cc:823-834)
}

// GA4GHOperator::operate + copy_back_remapped_fields (variant_operations.cc:572-731) for one variant of several calls: its calls
// are the live calls of their rows at the variant's begin column, so the combine kernels over that one column and those rows
// give the merged REF / ALT and every call's re-indexed GT / allele-dependent fields (no handle_deletions on this path)
void VariantQueryProcessor::remap_ga4gh_variant(const VariantQueryConfig& qc, GA4GHVariant& v) const {
  std::vector<int64_t> rows;
  for (auto& c : v.calls) rows.push_back((int64_t)c.row);
  std::sort(rows.begin(), rows.end());
  rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
  gdbgpu_scan_params p; memset(&p, 0, sizeof(p));
  p.flags = GDBGPU_SCAN_NO_DELETIONS | GDBGPU_SCAN_NO_PREFETCH;
  p.col_lo = p.col_hi = v.begin;
  p.query_rows = rows.data(); p.n_query_rows = rows.size();
  p.max_diploid_alt_alleles = qc.get_max_diploid_alt_alleles_that_can_be_genotyped();
  gdbgpu_scan* scan = nullptr;
  check(gdbgpu_scan_begin(arr_, &p, &scan));
  struct Guard { gdbgpu_scan* s; ~Guard() { gdbgpu_scan_end(s); } } guard{scan};
  gdbgpu_batch b;
  check(gdbgpu_scan_next_batch(scan, 0, 0, &b));
  uint32_t s = 0;
  while (s < b.n_sites && b.site_begin[s] != v.begin) ++s;
  if (s >= b.n_sites) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "GA4GH range query: no site at column " + std::to_string(v.begin));
  const uint32_t first = b.allele_first[s], n_allele = b.n_alleles[s];
  auto allele = [&](uint32_t m) { return std::string((const char*)b.allele_bytes + b.allele_pos[2 * (first + m)], b.allele_pos[2 * (first + m) + 1]); };
  v.has_common_fields = true;
  v.common_REF = allele(0);
  v.common_ALT.clear();
  for (uint32_t m = 1; m < n_allele; ++m) v.common_ALT.push_back(allele(m));
  if (!(b.site_flags[s] & GDBGPU_SITE_REMAPPED)) return;   // reference blocks only: m_remapping_needed is false, nothing is copied back
  const uint64_t nq = b.n_query_rows;
  gdbgpu_array_info info;
  check(gdbgpu_array_get_info(arr_, &info));
  for (auto& c : v.calls) {
    const uint64_t q = (uint64_t)(std::lower_bound(rows.begin(), rows.end(), (int64_t)c.row) - rows.begin());
    if (b.live[(uint64_t)s * nq + q] < 0) continue;
    for (unsigned i = 0; i < qc.get_num_queried_attributes(); ++i) {
      const auto& qa = qc.get_query_attribute(i);
      int o;
      if (qa.kind == GDBGPU_CELL_FIELD_GT) { if (!info.has_gt) continue; o = 0; }
      else if (qa.kind == GDBGPU_CELL_FIELD_REMAPPED) o = (info.has_gt ? 1 : 0) + qa.index;
      else continue;
      if (c.fields[i].empty()) continue;   // not valid in the call: stays so
      const gdbgpu_field_out& f = b.fields[o];
      const uint32_t len = f.len[(uint64_t)s * nq + q];
      const uint32_t* w = f.data + f.base[s] + q * (uint64_t)f.stride[s];
      c.fields[i] = ga4gh_number_list(reinterpret_cast<const uint8_t*>(w), len, qa.elem == 'f' ? 'f' : 'i');
    }
  }
}

void GA4GHVariantCall::print(std::ostream& fptr, const VariantQueryConfig* qc, const std::string& indent_prefix, const VidMapper* vid_mapper) const {   // variant.cc:225-265
  auto indent_string = indent_prefix + g_json_indent_unit;
  fptr << indent_prefix << "{\n";
  fptr << indent_string << "\"row\": " << row << ",\n";
  fptr << indent_string << "\"interval\": [ " << begin << ", " << end << " ],\n";
  if (vid_mapper) {
    const ContigInfo* contig = vid_mapper->get_contig_location(begin);
    if (contig) {
      const int64_t pos = begin - contig->tiledb_column_offset;
      fptr << indent_string << "\"genomic_interval\": { \"" << contig->name << "\" : [ " << pos + 1 << ", " << (pos + 1 + (end - begin)) << " ] },\n";
    }
  }
  fptr << indent_string << "\"fields\": {\n";
  indent_string += g_json_indent_unit;
  bool first_valid_field = true;
  for (unsigned idx = 0; idx < fields.size(); ++idx) {
    if (fields[idx].empty()) continue;
    if (!first_valid_field) fptr << ",\n";
    if (qc) fptr << indent_string << "\"" << qc->get_query_attribute_name(idx) << "\": ";
    else fptr << indent_string << "\"field_" << idx << "\": ";
    fptr << fields[idx];
    first_valid_field = false;
  }
  indent_string = indent_prefix + g_json_indent_unit;
  fptr << "\n" << indent_string << "}\n" << indent_prefix << "}";
}
void GA4GHVariant::print(std::ostream& fptr, const VariantQueryConfig* qc, const std::string& indent_prefix, const VidMapper* vid_mapper) const {   // variant.cc:427-478
  fptr << indent_prefix << "{\n";
  std::string indent_string = indent_prefix + g_json_indent_unit;
  fptr << indent_string << "\"interval\": [ " << begin << ", " << end << " ],\n";
  if (vid_mapper) {
    const ContigInfo* contig = vid_mapper->get_contig_location(begin);
    if (contig) {
      const int64_t pos = begin - contig->tiledb_column_offset;
      fptr << indent_string << "\"genomic_interval\": { \"" << contig->name << "\" : [ " << pos + 1 << ", " << (pos + 1 + (end - begin)) << " ] },\n";
    }
  }
  fptr << indent_string << " \"common_fields\" : {\n";
  indent_string += g_json_indent_unit;
  if (has_common_fields) {   // the merged REF and ALT, under the names of the query's REF / ALT attributes
    unsigned ref_idx = 0, alt_idx = 0;
    const bool named = qc && qc->get_query_idx_for_name("REF", ref_idx) && qc->get_query_idx_for_name("ALT", alt_idx);
    fptr << indent_string << "\"" << (named ? qc->get_query_attribute_name(ref_idx) : std::string("field_0")) << "\": \"" << common_REF << "\",\n";
    fptr << indent_string << "\"" << (named ? qc->get_query_attribute_name(alt_idx) : std::string("field_1")) << "\": " << ga4gh_alt_list(common_ALT);
  }
  indent_string = indent_prefix + g_json_indent_unit;
  fptr << "\n" << indent_string << "},\n";
  fptr << indent_string << "\"variant_calls\": [\n";
  indent_string += g_json_indent_unit;
  for (size_t i = 0; i < calls.size(); ++i) {
    if (i) fptr << ",\n";
    calls[i].print(fptr, qc, indent_string, vid_mapper);
  }
  indent_string = indent_prefix + g_json_indent_unit;
  fptr << "\n" << indent_string << "]\n";
  fptr << indent_prefix << "}";
}
void print_variants(const std::vector<GA4GHVariant>& variants, const std::string& output_format, const VariantQueryConfig& qc, std::ostream& fptr, const VidMapper* id_mapper) {
  if (!output_format.empty() && output_format != "default")
    throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "print_variants: output format " + output_format + " is not supported (only the default JSON)");
  fptr << "{\n" << g_json_indent_unit << "\"variants\": [\n";
  const std::string indent_prefix = std::string(g_json_indent_unit) + g_json_indent_unit;
  for (size_t i = 0; i < variants.size(); ++i) {
    if (i) fptr << ",\n";
    variants[i].print(fptr, &qc, indent_prefix, id_mapper);
  }
  fptr << "\n" << g_json_indent_unit << "]\n";
  fptr << "}\n";
}

// ------------------------------------------------------------------------------------------------ GenomicsDBBCFGenerator
// genomicsdb_bcf_generator.cc:27-93.  serialized_bcf_header == nullptr: the header is generated here, from the configuration
void GenomicsDBBCFGenerator::construct(const std::string& loader_config_file, const std::string& query_config_file, const std::string* serialized_bcf_header,
                                       const char* chr, int start, int end, int my_rank, const char* output_format,
                                       bool use_missing_values_only_not_vector_end, int cuda_device, const std::string& reference_bases_tsv) {
  if (output_format && std::string(output_format) != "bu" && std::string(output_format) != "")
    throw GenomicsDBGPUException(GDBGPU_ERR_LIMIT, "GenomicsDBBCFGenerator: only uncompressed BCF (\"bu\") is produced");
  GenomicsDBImportConfig loader_config;
  if (!loader_config_file.empty()) {   // vid / callset mapping, workspace and array may come from the loader's file
    loader_config.read_from_file(loader_config_file, my_rank);
    m_query_config.update_from_loader(loader_config, my_rank);
  }
  m_query_config.read_from_file(query_config_file, my_rank);
  if (!loader_config_file.empty()) m_query_config.subset_query_column_ranges_based_on_partition(loader_config, my_rank);
  m_query_config.set_vcf_output_format("bu");
  if (serialized_bcf_header) {
    if (serialized_bcf_header->size() < 9 || serialized_bcf_header->compare(0, 5, "BCF\2\2") != 0)
      throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "GenomicsDBBCFGenerator: the header must be a serialised BCF2 header");
    uint32_t l_text; memcpy(&l_text, serialized_bcf_header->data() + 5, 4);
    m_vcf_adapter.set_vcf_header_text(std::string(serialized_bcf_header->data() + 9, strnlen(serialized_bcf_header->data() + 9, l_text)));
  } else m_vcf_adapter.initialize(m_query_config);
  m_vcf_adapter.set_buffer_capacity(m_buffer_capacity);
  if (!reference_bases_tsv.empty()) {
    std::ifstream rb(reference_bases_tsv);
    if (!rb) throw GenomicsDBGPUException(GDBGPU_ERR_IO, "cannot open " + reference_bases_tsv);
    std::string contig, base; long long pos;
    while (rb >> contig >> pos >> base) m_vcf_adapter.set_reference_base(contig, pos, base[0]);
  }
  const VidMapper& vid_mapper = m_query_config.get_vid_mapper();
  if (chr && strlen(chr) > 0u) {   // genomicsdb_bcf_generator.cc:57-66: VCF positions are 1-based
    const ContigInfo* contig = vid_mapper.get_contig_info(chr);
    if (!contig) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, std::string("Could not find TileDB column interval for contig: ") + chr);
    m_query_config.set_column_interval_to_query(contig->tiledb_column_offset + (int64_t)start - 1, contig->tiledb_column_offset + (int64_t)end - 1);
  }
  m_storage_manager.reset(new VariantStorageManager(m_query_config.get_workspace(my_rank), 10u * 1024u * 1024u, cuda_device));
  m_query_processor.reset(new VariantQueryProcessor(m_storage_manager.get(), m_query_config.get_array_name(my_rank), vid_mapper));
  m_query_processor->do_query_bookkeeping(m_query_config, true, /*stage_broad_gvcf_fields=*/true);
  if (serialized_bcf_header) { uint32_t l_text; memcpy(&l_text, serialized_bcf_header->data() + 5, 4); m_vcf_adapter.append(serialized_bcf_header->data(), 9 + (size_t)l_text, false); }   // print_header()
  m_combined_bcf_operator.reset(new BroadCombinedGVCFOperator(m_vcf_adapter, vid_mapper, m_query_config, use_missing_values_only_not_vector_end));   // generated headers are printed here
  if (!m_produce_header_only)
    m_query_processor->scan_and_operate(m_query_processor->get_array_descriptor(), m_query_config, *m_combined_bcf_operator, m_query_column_interval_idx, true, &m_scan_state);
  m_read_buffer.swap(m_vcf_adapter.get_buffer());
  m_vcf_adapter.get_buffer().clear();
}
GenomicsDBBCFGenerator::GenomicsDBBCFGenerator(const std::string& loader_config_file, const std::string& query_config_file, const std::string& serialized_bcf_header,
                                               const char* chr, const int start, const int end, int my_rank, size_t buffer_capacity, size_t,
                                               const char* output_format, const bool produce_header_only,
                                               const bool use_missing_values_only_not_vector_end, const bool keep_idx_fields_in_bcf_header, int cuda_device,
                                               const std::string& reference_bases_tsv)
    : m_produce_header_only(produce_header_only), m_buffer_capacity(buffer_capacity), m_vcf_adapter(keep_idx_fields_in_bcf_header) {
  construct(loader_config_file, query_config_file, &serialized_bcf_header, chr, start, end, my_rank, output_format, use_missing_values_only_not_vector_end, cuda_device, reference_bases_tsv);
}
GenomicsDBBCFGenerator::GenomicsDBBCFGenerator(const std::string& loader_config_file, const std::string& query_config_file,
                                               const char* chr, const int start, const int end, int my_rank, size_t buffer_capacity, size_t,
                                               const char* output_format, const bool produce_header_only,
                                               const bool use_missing_values_only_not_vector_end, const bool keep_idx_fields_in_bcf_header)
    : m_produce_header_only(produce_header_only), m_buffer_capacity(buffer_capacity), m_vcf_adapter(keep_idx_fields_in_bcf_header) {
  construct(loader_config_file, query_config_file, nullptr, chr, start, end, my_rank, output_format, use_missing_values_only_not_vector_end, 0, "");
}
GenomicsDBBCFGenerator::~GenomicsDBBCFGenerator() {
  m_scan_state.reset();               // the scan before the array, the array before the context
  m_combined_bcf_operator.reset();
  m_query_processor.reset();
  m_storage_manager.reset();
}
// genomicsdb_bcf_generator.cc:112-140
void GenomicsDBBCFGenerator::produce_next_batch() {
  if (m_done) return;
  m_read_buffer.clear(); m_next_read_idx = 0;
  while (m_read_buffer.empty()) {
    if (m_produce_header_only || m_scan_state.end()) {
      ++m_query_column_interval_idx;
      if (m_produce_header_only || m_query_column_interval_idx >= m_query_config.get_num_column_intervals()) { m_done = true; return; }
      m_scan_state.reset();
    }
    m_query_processor->scan_and_operate(m_query_processor->get_array_descriptor(), m_query_config, *m_combined_bcf_operator, m_query_column_interval_idx, true, &m_scan_state);
    m_read_buffer.swap(m_vcf_adapter.get_buffer());
    m_vcf_adapter.get_buffer().clear();
  }
}
// genomicsdb_bcf_generator.cc:142-177
size_t GenomicsDBBCFGenerator::read_and_advance(uint8_t* dst, size_t offset, size_t n) {
  size_t total = 0;
  if (n == SIZE_MAX) { produce_next_batch(); return 0; }
  while (total < n && !m_done) {
    if (m_next_read_idx >= m_read_buffer.size()) { produce_next_batch(); continue; }
    const size_t k = std::min(n - total, m_read_buffer.size() - m_next_read_idx);
    if (dst) memcpy(dst + offset + total, m_read_buffer.data() + m_next_read_idx, k);
    m_next_read_idx += k; total += k;
    if (m_next_read_idx >= m_read_buffer.size()) produce_next_batch();
  }
  return total;
}
uint8_t GenomicsDBBCFGenerator::read_next_byte() {
  uint8_t tmp;
  return read_and_advance(&tmp, 0u, 1u) > 0u ? tmp : (uint8_t)-1;
}

}  // namespace genomicsdb_b200
