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
#include <iostream>
#include <sstream>

#include "gdb_host_api.h"
#include "gdb_json.h"
#include "gdb_vcf_header.h"

namespace genomicsdb_b200 {

using gdb::JsonValue;

// ------------------------------------------------------------------------------------------------ callsets
// vid_mapper.cc:841-858,946-1055: "callsets" as a dictionary name -> info or a list of info dictionaries ("name" /
// "callset_name" / "sample_name"); "row_idx" is required; one name per row and one row per name; every row up to the largest
// must be named
void VidMapper::read_callsets_file(const std::string& file) {
  JsonValue d;
  try { d = gdb::json_parse_file(file); } catch (const std::exception& e) { throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Syntax error in JSON file " + file + " (" + e.what() + ")"); }
  const JsonValue* cs = d.find("callsets");
  if (!cs) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "no \"callsets\" in " + file);
  read_callsets(*cs);
}
void VidMapper::read_callsets(const JsonValue& cs) {
  auto bad = [](const std::string& what) { throw GenomicsDBGPUException(GDBGPU_ERR_ARG, what); };
  if (cs.type != JsonValue::OBJ && cs.type != JsonValue::ARR) bad("\"callsets\" must be a dictionary or a list");
  auto add = [&](const std::string& name, const JsonValue& v) {
    if (v.type != JsonValue::OBJ) bad("callset " + name + ": the info must be a dictionary");
    const JsonValue* r = v.find("row_idx");
    if (!r || !r->is_int || r->as_int() < 0) bad("callset " + name + " needs an integer \"row_idx\" >= 0");
    const int64_t row = r->as_int();
    if (v.has("filename") && v.has("stream_name")) bad("Cannot have both \"filename\" and \"stream_name\" as the data source for sample/CallSet " + name);
    for (size_t i = 0; i < callset_names_.size(); ++i)
      if (callset_rows_named_[i] && callset_names_[i] == name && (int64_t)i != row)
        bad("Duplicate with conflicting row index for sample/callset name " + name + " found in callsets mapping " + std::to_string(i) + ", " + std::to_string(row));
    if ((size_t)row >= callset_names_.size()) { callset_names_.resize(row + 1); callset_rows_named_.resize(row + 1, false); }
    if (callset_rows_named_[row] && callset_names_[row] != name) bad("Sample/callset " + name + " has the same row idx as " + callset_names_[row]);
    callset_names_[row] = name;
    callset_rows_named_[row] = true;
  };
  if (cs.type == JsonValue::OBJ) for (auto& kv : cs.obj) add(kv.first, kv.second);
  else for (size_t i = 0; i < cs.arr.size(); ++i) {
    const JsonValue& v = cs.arr[i];
    if (v.type != JsonValue::OBJ) bad("a callset's info must be a dictionary");
    std::string name; unsigned found = 0;
    for (const char* k : {"name", "callset_name", "sample_name"}) if (const JsonValue* n = v.find(k)) { name = n->str; ++found; }
    if (found == 0) bad("Callset info dict with index " + std::to_string(i) + " does not have any one of the keys \"name\", \"callset_name\" or \"sample_name\"");
    if (found > 1) bad("Callset info dict with index " + std::to_string(i) + " has two or more of the keys \"name\", \"callset_name\" or \"sample_name\" - at most one is allowed");
    add(name, v);
  }
  bool missing = false;
  for (size_t i = 0; i < callset_names_.size(); ++i) if (!callset_rows_named_[i]) { std::cerr << "Sample/callset information missing for row " << i << "\n"; missing = true; }
  if (missing) bad("Row indexes with missing sample/callset information found in callsets mapping");
  callsets_initialized_ = true;
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
[[noreturn]] void bad_config(const std::string& what) { throw GenomicsDBGPUException(GDBGPU_ERR_ARG, what); }
// genomicsdb_config_base.cc:36-57: 1-based contig positions -> 0-based columns; a begin past the contig is an error, an end
// past it is cut to the contig's length
std::pair<int64_t, int64_t> contig_positions_to_columns(const ContigInfo& c, int64_t begin, int64_t end) {
  std::pair<int64_t, int64_t> r(std::min(begin, end), std::max(begin, end));
  if (r.first > c.length) bad_config("Position " + std::to_string(r.first) + " queried for contig " + c.name + " which is of length " + std::to_string(c.length) + "; queried position is past end of contig");
  if (r.second > c.length) {
    std::cerr << "WARNING: position " << r.second << " queried for contig " << c.name << " which is of length " << c.length << "; queried interval is past end of contig, truncating to contig length";
    r.second = c.length;
  }
  r.first += c.tiledb_column_offset - 1;
  r.second += c.tiledb_column_offset - 1;
  return r;
}
const ContigInfo& contig_named(const VidMapper& vid, const std::string& name) {
  const ContigInfo* c = vid.get_contig_info(name);
  if (!c) bad_config("JSONConfigBase::read_from_file: Invalid contig name : " + name);
  return *c;
}
// { "chr" : pos } / { "chr" : [ b, e ] } / { "contig_position" : { "contig" : .., "position" : .. } }
// (json_config.cc:55-98 extract_contig_interval_from_object)
std::pair<int64_t, int64_t> contig_interval(const JsonValue& o, const VidMapper& vid) {
  if (o.type != JsonValue::OBJ || o.obj.size() != 1) bad_config("a contig position must be { \"<contig>\" : <position> } or { \"<contig>\" : [ <begin>, <end> ] }");
  std::string name = o.obj[0].first;
  const JsonValue* pos = &o.obj[0].second;
  if (name == "contig_position") {
    const JsonValue& d = *pos;
    if (d.type != JsonValue::OBJ || d.obj.size() != 2 || !d.has("contig") || d.at("contig").type != JsonValue::STR || !d.has("position") || !d.at("position").is_int)
      bad_config("\"contig_position\" must be { \"contig\" : <name>, \"position\" : <int> }");
    name = d.at("contig").str;
    pos = &d.at("position");
  }
  const ContigInfo& c = contig_named(vid, name);
  if (pos->type == JsonValue::ARR) {
    if (pos->arr.size() != 2 || !pos->arr[0].is_int || !pos->arr[1].is_int) bad_config("a contig interval must be [ <begin>, <end> ]");
    return contig_positions_to_columns(c, pos->arr[0].as_int(), pos->arr[1].as_int());
  }
  if (!pos->is_int) bad_config("a contig position must be an integer");
  return contig_positions_to_columns(c, pos->as_int(), pos->as_int());
}
// the protobuf-generated forms (json_config.cc:101-193 extract_interval_from_PB_struct_or_return_false)
bool pb_interval(const JsonValue& o, const VidMapper& vid, std::pair<int64_t, int64_t>& r) {
  if (o.type != JsonValue::OBJ) return false;
  if (o.obj.size() == 2 && o.has("low") && o.has("high")) { r = {o.at("low").as_int(), o.at("high").as_int()}; return true; }
  if (o.obj.size() != 1) return false;
  if (const JsonValue* ci = o.find("column_interval")) {
    if (ci->type != JsonValue::OBJ) return false;
    const JsonValue* t = ci->find("column_interval");
    if (t && t->type == JsonValue::OBJ && t->has("begin") && t->has("end")) { r = {t->at("begin").as_int(), t->at("end").as_int()}; return true; }
    const JsonValue* g = ci->find("contig_interval");
    if (g && g->type == JsonValue::OBJ && g->has("contig") && g->has("begin") && g->has("end")) {
      r = contig_positions_to_columns(contig_named(vid, g->at("contig").str), g->at("begin").as_int(), g->at("end").as_int());
      return true;
    }
    return false;
  }
  if (const JsonValue* co = o.find("column")) {
    if (co->type != JsonValue::OBJ) return false;
    const JsonValue* t = co->find("tiledb_column");
    if (t && t->is_int) { r = {t->as_int(), t->as_int()}; return true; }
    const JsonValue* g = co->find("contig_position");
    if (g && g->type == JsonValue::OBJ && g->has("contig") && g->has("position")) {
      const int64_t p = g->at("position").as_int();
      r = contig_positions_to_columns(contig_named(vid, g->at("contig").str), p, p);
      return true;
    }
  }
  return false;
}
// "column_partitions" / "row_partitions": [ { "begin" : b [, "end" : e] [, "workspace" : .., "array" : ..] }, ... ]
// (json_config.cc:341-418,472-531); a partition without a usable end stops where the next one, in begin order, starts
void read_partitions(const JsonValue& list, bool columns, const VidMapper& vid, std::vector<std::pair<int64_t, int64_t>>& parts,
                     std::vector<std::string>& workspaces, std::vector<std::string>& array_names) {
  if (list.type != JsonValue::ARR) bad_config(columns ? "\"column_partitions\" must be a list" : "\"row_partitions\" must be a list");
  const std::string ws0 = workspaces.size() == 1 ? workspaces[0] : "", an0 = array_names.size() == 1 ? array_names[0] : "";
  const size_t n = list.arr.size();
  parts.assign(n, {0, INT64_MAX - 1});
  for (size_t i = 0; i < n; ++i) {
    const JsonValue& p = list.arr[i];
    if (p.type != JsonValue::OBJ || !p.has("begin")) bad_config("a partition needs \"begin\"");
    const JsonValue& b = p.at("begin");
    if (!(b.is_int || (columns && b.type == JsonValue::OBJ))) bad_config("a partition's \"begin\" must be an integer or a contig position");
    parts[i].first = b.is_int ? b.as_int() : contig_interval(b, vid).first;
    if (const JsonValue* e = p.find("end")) {
      if (!(e->is_int || (columns && e->type == JsonValue::OBJ))) bad_config("a partition's \"end\" must be an integer or a contig position");
      parts[i].second = e->is_int ? e->as_int() : contig_interval(*e, vid).first;
    }
    if (parts[i].first > parts[i].second) std::swap(parts[i].first, parts[i].second);
    if (const JsonValue* w = p.find("workspace")) { if (workspaces.size() < n) workspaces.resize(n, ws0); workspaces[i] = w->str; }
    if (p.has("array") && p.has("array_name")) bad_config("a partition names both \"array\" and \"array_name\"");
    const JsonValue* a = p.find("array"); if (!a) a = p.find("array_name");
    if (a) { if (array_names.size() <= n) array_names.resize(n, an0); array_names[i] = a->str; }
  }
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return parts[x].first < parts[y].first; });
  for (size_t k = 0; k + 1 < n; ++k) {
    auto& cur = parts[order[k]]; const auto& nxt = parts[order[k + 1]];
    if (cur.first == nxt.first) bad_config(columns ? "Cannot have two column partitions with the same begin value" : "Cannot have two row partitions with the same begin value");
    if (cur.second >= nxt.first) cur.second = nxt.first - 1;
  }
}
// one name for all ranks, or a list with one per rank (json_config.cc:207-250,560-575)
std::string string_of_rank(const JsonValue& v, int rank, const char* key) {
  if (v.type == JsonValue::ARR) {
    if (rank < 0 || (size_t)rank >= v.arr.size() || v.arr[rank].type != JsonValue::STR) bad_config(std::string("\"") + key + "\" holds no string for rank " + std::to_string(rank));
    return v.arr[rank].str;
  }
  if (v.type != JsonValue::STR) bad_config(std::string("\"") + key + "\" must be a string or a list of strings");
  return v.str;
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
  if (!d.has("size_per_column_partition")) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "the loader configuration has no \"size_per_column_partition\"");   // json_config.cc:718-721
  if (row_based_) { if (!d.has("row_partitions")) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "row based partitioning without \"row_partitions\""); return; }
  const JsonValue* cp = d.find("column_partitions");
  if (!cp || cp->type != JsonValue::ARR) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "the loader configuration has no \"column_partitions\"");
  partitions_specified_ = true;
  read_partitions(*cp, true, vid_, partitions_, workspaces_, array_names_);
  // json_config.cc:207-250,384-402: one name serves every rank only when it is given once, as a string, outside the partitions
  single_workspace_ = d.has("workspace") && d.at("workspace").type != JsonValue::ARR;
  single_array_ = an && an->type != JsonValue::ARR;
  for (auto& p : cp->arr) { if (p.has("workspace")) single_workspace_ = false; if (p.has("array") || p.has("array_name")) single_array_ = false; }
}
std::pair<int64_t, int64_t> GenomicsDBImportConfig::get_column_partition(const int rank, const unsigned idx) const {
  (void)idx;
  if (!partitions_specified_ || partitions_.empty()) return {0, INT64_MAX - 1};
  // genomicsdb_config_base.cc:89-100: every rank must have its own partition (only a single list of query_column_ranges serves all)
  if (rank < 0 || (size_t)rank >= partitions_.size()) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "No column partition/query interval available for process with rank " + std::to_string(rank));
  return partitions_[rank];
}
const std::string& GenomicsDBImportConfig::get_workspace(const int rank) const {
  static const std::string none;
  if (workspaces_.empty()) return none;
  if (single_workspace_) return workspaces_[0];   // genomicsdb_config_base.cc:59-66
  if (rank < 0 || (size_t)rank >= workspaces_.size()) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Workspace not defined for rank " + std::to_string(rank));
  return workspaces_[rank];
}
const std::string& GenomicsDBImportConfig::get_array_name(const int rank) const {
  static const std::string none;
  if (array_names_.empty()) return none;
  if (single_array_) return array_names_[0];   // genomicsdb_config_base.cc:68-75
  if (rank < 0 || (size_t)rank >= array_names_.size()) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Could not find array for rank " + std::to_string(rank));
  return array_names_[rank];
}

// ------------------------------------------------------------------------------------------------ query configuration
// JSONConfigBase::read_from_file + VariantQueryConfig::read_from_file (json_config.cc:195-656, variant_query_config.cc:234-278):
// the same keys, forms and refusals.  Intervals: [ [ [b, e], p, "contig", { "contig" : [b, e] }, { "contig" : p }, <protobuf
// forms> ], ... ] -- one inner list for all ranks or one per rank; rows likewise.
void VariantQueryConfig::read_from_file(const std::string& filename, int rank) {
  JsonValue d;
  try { d = gdb::json_parse_file(filename); } catch (const std::exception& e) { bad_config("Syntax error in JSON file " + filename + " (" + e.what() + ")"); }
  // the mappings come first: contig names in the intervals need them (json_config.cc:204,660-706)
  std::string vid_file;
  if (const JsonValue* v = d.find("vid_mapping_file")) vid_file = string_of_rank(*v, rank, "vid_mapping_file");
  if (const JsonValue* v = d.find("callset_mapping_file")) callset_mapping_file_ = string_of_rank(*v, rank, "callset_mapping_file");
  if (!vid_file.empty()) vid_ = VidMapper(vid_file);
  if (vid_.is_initialized() && !callset_mapping_file_.empty() && !vid_.is_callset_mapping_initialized()) vid_.read_callsets_file(callset_mapping_file_);
  std::vector<std::string> workspaces, array_names;
  if (!workspace_.empty()) workspaces.push_back(workspace_);   // from the loader's configuration
  if (!array_name_.empty()) array_names.push_back(array_name_);
  if (const JsonValue* w = d.find("workspace")) {
    workspaces.clear();
    if (w->type == JsonValue::ARR) for (auto& x : w->arr) { if (x.type != JsonValue::STR) bad_config("\"workspace\" must hold strings"); workspaces.push_back(x.str); }
    else if (w->type == JsonValue::STR) workspaces.push_back(w->str);
    else bad_config("\"workspace\" must be a string or a list of strings");
  }
  if (d.has("array") && d.has("array_name")) bad_config("the configuration names both \"array\" and \"array_name\"");
  const JsonValue* an = d.find("array"); if (!an) an = d.find("array_name");
  if (an) {
    array_names.clear();
    if (an->type == JsonValue::ARR) for (auto& x : an->arr) { if (x.type != JsonValue::STR) bad_config("\"array\" must hold strings"); array_names.push_back(x.str); }
    else if (an->type == JsonValue::STR) array_names.push_back(an->str);
    else bad_config("\"array\" must be a string or a list of strings");
  }
  bool single_workspace = workspaces.size() <= 1 && !(d.has("workspace") && d.at("workspace").type == JsonValue::ARR);
  bool single_array = array_names.size() <= 1 && !(an && an->type == JsonValue::ARR);

  typedef std::pair<int64_t, int64_t> Range;
  std::vector<std::vector<Range>> column_ranges, row_ranges;
  // json_config.cc:253-260: the PRESENCE of "scan_full" (whatever its value) makes the query the whole array -- every column and
  // every row; "query_column_ranges" and "query_row_ranges" beside it are not read
  const bool scan_full = d.has("scan_full");
  if (scan_full) {
    column_ranges.assign(1, std::vector<Range>(1, Range(0, INT64_MAX - 1)));
  } else {
    if (d.has("row_partitions") && d.has("column_partitions")) bad_config("Cannot have both \"row_partitions\" and \"column_partitions\" simultaneously in the JSON file");
    if (!(d.has("query_column_ranges") || d.has("column_partitions") || d.has("query_row_ranges") || d.has("row_partitions")))
      bad_config("Must have one of \"query_column_ranges\" or \"column_partitions\" or \"query_row_ranges\" or \"row_partitions\"");
    if (d.has("query_column_ranges") && d.has("column_partitions")) bad_config("Cannot use both \"query_column_ranges\" and \"column_partitions\" simultaneously");
    if (d.has("query_row_ranges") && d.has("row_partitions")) bad_config("Cannot use both \"query_row_ranges\" and \"row_partitions\" simultaneously");
    // the list of one rank: a plain list, or the protobuf wrapper { "range_list" : [..] } / { "column_or_interval_list" : [..] }
    auto list_of = [&](const JsonValue& e, bool columns, bool& skip) -> const JsonValue& {
      skip = false;
      if (e.type == JsonValue::ARR) return e;
      if (e.type != JsonValue::OBJ) bad_config("a rank's ranges must be a list or { \"range_list\" : [ .. ] }");
      if (columns && e.obj.empty()) { skip = true; return e; }
      const JsonValue* l = e.obj.size() == 1 ? e.find("range_list") : nullptr;
      if (!l && columns && e.obj.size() == 1) l = e.find("column_or_interval_list");
      if (!l || l->type != JsonValue::ARR) bad_config("a rank's ranges must be a list or { \"range_list\" : [ .. ] }");
      return *l;
    };
    if (const JsonValue* q = d.find("query_column_ranges")) {
      if (q->type != JsonValue::ARR) bad_config("\"query_column_ranges\" must be a list");
      column_ranges.resize(q->arr.size());
      for (size_t i = 0; i < q->arr.size(); ++i) {
        bool skip;
        const JsonValue& l = list_of(q->arr[i], true, skip);
        if (skip) continue;
        for (const JsonValue& it : l.arr) {
          Range r;
          if (it.type == JsonValue::ARR) {
            if (it.arr.size() != 2 || !it.arr[0].is_int || !it.arr[1].is_int) bad_config("a column interval must be [ <begin>, <end> ]");
            r = Range(it.arr[0].as_int(), it.arr[1].as_int());
          } else if (it.is_int) r = Range(it.as_int(), it.as_int());
          else if (it.type == JsonValue::STR) {   // a whole contig
            const ContigInfo& c = contig_named(vid_, it.str);
            r = Range(c.tiledb_column_offset, c.tiledb_column_offset + c.length - 1);
          } else if (!pb_interval(it, vid_, r)) r = contig_interval(it, vid_);
          if (r.first > r.second) std::swap(r.first, r.second);
          column_ranges[i].push_back(r);
        }
      }
    } else if (const JsonValue* cp = d.find("column_partitions")) {
      std::vector<Range> parts;
      const bool had_ws_list = workspaces.size() > 1, had_an_list = array_names.size() > 1;
      const size_t nw = workspaces.size(), na = array_names.size();
      read_partitions(*cp, true, vid_, parts, workspaces, array_names);
      if (workspaces.size() != nw || had_ws_list) single_workspace = false;
      else for (auto& p : cp->arr) if (p.has("workspace")) single_workspace = false;
      if (array_names.size() != na || had_an_list) single_array = false;
      else for (auto& p : cp->arr) if (p.has("array") || p.has("array_name")) single_array = false;
      for (auto& p : parts) column_ranges.push_back(std::vector<Range>(1, p));
    }
    if (const JsonValue* q = d.find("query_row_ranges")) {
      if (q->type != JsonValue::ARR) bad_config("\"query_row_ranges\" must be a list");
      row_ranges.resize(q->arr.size());
      for (size_t i = 0; i < q->arr.size(); ++i) {
        bool skip;
        const JsonValue& l = list_of(q->arr[i], false, skip);
        for (const JsonValue& it : l.arr) {
          Range r(0, 0);
          if (it.type == JsonValue::ARR) {
            if (it.arr.size() != 2 || !it.arr[0].is_int || !it.arr[1].is_int) bad_config("a row interval must be [ <first>, <last> ]");
            r = Range(it.arr[0].as_int(), it.arr[1].as_int());
          } else if (it.is_int) r = Range(it.as_int(), it.as_int());
          else if (it.type == JsonValue::OBJ) { if (!pb_interval(it, vid_, r)) bad_config("a row interval must be [ <first>, <last> ], a row or { \"low\" : .., \"high\" : .. }"); }
          if (r.first > r.second) std::swap(r.first, r.second);
          row_ranges[i].push_back(r);
        }
      }
    } else if (const JsonValue* rp = d.find("row_partitions")) {
      std::vector<Range> parts;
      read_partitions(*rp, false, vid_, parts, workspaces, array_names);
      for (auto& p : rp->arr) { if (p.has("workspace")) single_workspace = false; if (p.has("array") || p.has("array_name")) single_array = false; }
      for (auto& p : parts) row_ranges.push_back(std::vector<Range>(1, p));
    }
  }
  if (d.has("query_attributes") && d.has("attributes")) bad_config("Query configuration cannot have both \"query_attributes\" and \"attributes\"");
  const JsonValue* attrs = d.find("query_attributes");
  if (!attrs) attrs = d.find("attributes");
  attributes_.clear();
  if (attrs) {
    if (attrs->type != JsonValue::ARR) bad_config("\"attributes\" must be a list of strings");
    for (auto& x : attrs->arr) {
      if (x.type != JsonValue::STR) bad_config("\"attributes\" must be a list of strings");
      if (std::find(attributes_.begin(), attributes_.end(), x.str) == attributes_.end()) attributes_.push_back(x.str);   // variant_query_config.cc:37-45: a name once
    }
  }
  if (const JsonValue* v = d.find("segment_size")) { if (!v->is_int) bad_config("\"segment_size\" must be an integer"); segment_size_ = (size_t)v->as_int(); }
  if (const JsonValue* v = d.find("vcf_header_filename")) vcf_header_filename_ = string_of_rank(*v, rank, "vcf_header_filename");
  if (const JsonValue* v = d.find("reference_genome")) reference_genome_ = string_of_rank(*v, rank, "reference_genome");
  vcf_output_filename_ = "-";   // stdout
  if (const JsonValue* v = d.find("vcf_output_filename")) vcf_output_filename_ = string_of_rank(*v, rank, "vcf_output_filename");
  if (const JsonValue* cp = d.find("column_partitions"))   // json_config.cc:600-617: the rank's partition may name its own
    if (cp->type == JsonValue::ARR && rank >= 0 && (size_t)rank < cp->arr.size())
      if (const JsonValue* v = cp->arr[rank].find("vcf_output_filename")) { if (v->type != JsonValue::STR) bad_config("a partition's \"vcf_output_filename\" must be a string"); vcf_output_filename_ = v->str; }
  if (const JsonValue* v = d.find("vcf_output_format")) vcf_output_format_ = v->str;
  max_alt_ = d.has("max_diploid_alt_alleles_that_can_be_genotyped") ? (unsigned)d.at("max_diploid_alt_alleles_that_can_be_genotyped").as_int() : 50u;
  min_pl_gt_ = d.has("produce_GT_with_min_PL_value_for_spanning_deletions") && d.at("produce_GT_with_min_PL_value_for_spanning_deletions").as_bool();
  produce_gt_ = d.has("produce_GT_field") && d.at("produce_GT_field").as_bool();
  sites_only_ = d.has("sites_only_query") && d.at("sites_only_query").as_bool();
  produce_filter_ = d.has("produce_FILTER_field") && d.at("produce_FILTER_field").as_bool();

  // variant_query_config.cc:240-277: this rank's workspace, array, intervals and rows
  if (!(vid_.is_initialized() && vid_.is_callset_mapping_initialized())) bad_config("the query needs a vid mapping and a callset mapping (its own, or the loader's)");
  if (workspaces.empty()) bad_config("No workspace specified");
  if (!single_workspace && (size_t)rank >= workspaces.size()) bad_config("Could not find workspace for rank " + std::to_string(rank));
  workspace_ = single_workspace ? workspaces[0] : workspaces[rank];
  if (workspace_.empty()) bad_config("Empty workspace string");
  if (array_names.empty()) bad_config("No array specified");
  if (!single_array && (size_t)rank >= array_names.size()) bad_config("Could not find array for rank " + std::to_string(rank));
  array_name_ = single_array ? array_names[0] : array_names[rank];
  if (array_name_.empty()) bad_config("Empty array name");
  if (column_ranges.empty() && !scan_full) bad_config("Query column ranges not specified");
  // json_config.cc:275-276,428-429: ONE list of query_column_ranges / query_row_ranges serves every rank; partitions do not
  const bool single_columns = scan_full || (d.has("query_column_ranges") && column_ranges.size() == 1);
  const bool single_rows = d.has("query_row_ranges") && row_ranges.size() == 1;
  if (!single_columns && (size_t)rank >= column_ranges.size()) bad_config("Rank >= query column ranges vector size");
  base_column_ranges_ = column_ranges[single_columns ? 0 : rank];
  intervals_.clear();
  if (!scan_full) for (auto& r : base_column_ranges_) add_column_interval_to_query(r.first, r.second);
  rows_.clear();
  if (!scan_full && !row_ranges.empty()) {
    if (!single_rows && (size_t)rank >= row_ranges.size()) bad_config("Rank >= query row ranges vector size");
    std::vector<int64_t> rows;
    // every row of every range is listed (variant_query_config.cc:266-274); an open-ended row partition cannot be (the reference
    // dies in its vector's resize there)
    uint64_t total = 0;
    for (auto& r : row_ranges[single_rows ? 0 : rank]) {
      total += (uint64_t)r.second - (uint64_t)r.first + 1u;
      if (total > (1ull << 31)) bad_config("the queried row ranges list more than 2^31 rows");
    }
    rows.reserve(total);
    for (auto& r : row_ranges[single_rows ? 0 : rank]) for (int64_t i = r.first; i <= r.second; ++i) rows.push_back(i);
    set_rows_to_query(rows);
  }
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
// genomicsdb_config_base.cc:203-223: of the rank's ranges AS READ FROM THE FILE the base configuration keeps those that touch
// its column partition.  Both callers (gt_mpi_gather.cc:554-557, genomicsdb_bcf_generator.cc:51-53) call it after
// read_from_file, which has already handed every range to add_column_interval_to_query -- the intervals the scan walks are
// not touched, in the reference and here: with a loader JSON every rank still produces the whole query.
void VariantQueryConfig::subset_query_column_ranges_based_on_partition(const GenomicsDBImportConfig& loader_config, const int rank) {
  if (!loader_config.is_partitioned_by_column()) return;
  const auto mine = loader_config.get_column_partition(rank);
  std::vector<std::pair<int64_t, int64_t>> kept;
  for (auto& q : base_column_ranges_) if (q.second >= mine.first && q.first <= mine.second) kept.push_back(q);
  base_column_ranges_.swap(kept);
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
    if (!lines_->parse(ss.str(), /*sample_line_required=*/false)) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Could not parse template VCF header file " + config.get_vcf_header_filename());
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
  if (exists) {
    // vcf_adapter.cc:207-259: the template's line and the vid mapping must agree on length and type (GT, strings and FILTERs are
    // not looked at)
    const int ht = hdr.id2type(field_type_idx, vcf_name);
    if (vcf_name != "GT" && ht != HT_STR && field_type_idx != HL_FLT) {
      if (!fi) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "field " + field_name + " is not in the vid mapping");
      const bool fixed = !multid && !fi->length.empty() && isdigit((unsigned char)fi->length[0]);
      const long n = fixed ? atol(fi->length.c_str()) : 0;
      const int hl = hdr.id2length(field_type_idx, vcf_name), hn = hdr.id2number(field_type_idx, vcf_name);
      const bool agree = (fixed && hl == VL_FIXED && n == hn) || (!fixed && hl != VL_FIXED) || (ht == HT_FLAG && fixed && n == 1 && hl == VL_FIXED && hn == 0);
      if (!agree) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Conflicting field length descriptors and/or field lengths in the vid JSON and VCF header for field " + field_name);
      const std::string& vt = fi->vcf_type.empty() ? fi->type : fi->vcf_type;
      const int vid_ht = vt == "flag" ? HT_FLAG : vt == "int" ? HT_INT : vt == "float" ? HT_REAL : HT_STR;
      if ((ht == HT_FLAG || ht == HT_INT || ht == HT_REAL) && vid_ht != ht)
        throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Conflicting data types in the vid JSON and VCF header for field " + field_name);
    }
    return false;
  }
  std::string line = "##";
  line += field_type_idx == HL_FLT ? "FILTER" : (field_type_idx == HL_INFO ? "INFO" : "FORMAT");
  line += "=<ID=" + vcf_name;
  if (field_type_idx != HL_FLT) {
    if (field_type_idx == HL_FMT && vcf_name == "GT") line += ",Number=1,Type=String,Description=\"Genotype\"";   // GT is weird -- and gets the general Description below as well
    else {
      if (!fi) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "field " + field_name + " is not in the vid mapping");
      const std::string& type = fi->vcf_type.empty() ? fi->type : fi->vcf_type;
      const bool is_flag = type == "flag";
      line += ",Number=";
      if (is_flag) line += "0";
      else if (multid) line += "1";
      else if (fi->length == "R" || fi->length == "A" || fi->length == "G") line += fi->length;
      else if (!fi->length.empty() && isdigit((unsigned char)fi->length[0])) line += fi->length;
      else if (fi->length == "VAR") line += ".";
      else throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "Unhandled field length descriptor " + fi->length + " of field " + field_name);
      line += ",Type=";
      if (multid) line += "String";
      else if (is_flag) line += "Flag";
      else if (type == "int") line += "Integer";
      else if (type == "float") line += "Float";
      else line += "String";
    }
  }
  if (removed && !description_value.empty()) line += ",Description=" + description_value;   // the value carries its quotes
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
