#include "gdb_host_api.h"

#include <algorithm>
#include <cstring>

#include "gdb_json.h"

namespace genomicsdb_b200 {

using gdb::JsonValue;

// ------------------------------------------------------------------------------------------------ VidMapper
VidMapper::VidMapper(const std::string& file) {
  JsonValue d = gdb::json_parse_file(file);
  const JsonValue& fl = d.at("fields");
  auto add = [&](const std::string& name, const JsonValue& v) {
    FieldInfo f; f.name = name;
    if (const JsonValue* t = v.find("type")) f.type = t->type == JsonValue::STR ? t->str : (t->type == JsonValue::ARR && !t->arr.empty() ? t->arr[0].str : "int");
    else f.type = "int";
    f.length = "1";
    if (const JsonValue* l = v.find("length")) {
      if (l->type == JsonValue::STR) f.length = l->str;
      else if (l->type == JsonValue::NUM) f.length = std::to_string(l->as_int());
      else if (l->type == JsonValue::ARR) f.length = "2D";
    }
    if (const JsonValue* c = v.find("vcf_field_class")) for (auto& x : c->arr) { if (x.str == "FORMAT") f.is_format = true; if (x.str == "INFO") f.is_info = true; }
    fields_.push_back(f);
    if (f.is_format && f.is_info) { FieldInfo g = f; g.name = name + "_FORMAT"; fields_.push_back(g); }   // vid_mapper.cc: FORMAT twin
  };
  if (fl.type == JsonValue::OBJ) for (auto& kv : fl.obj) add(kv.first, kv.second);
  else for (auto& v : fl.arr) add(v.at("name").str, v);
}
const FieldInfo* VidMapper::get_field_info(const std::string& name) const {
  for (auto& f : fields_) if (f.name == name) return &f;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ VariantQueryConfig
void VariantQueryConfig::add_column_interval_to_query(int64_t b, int64_t e) {
  intervals_.push_back({b, e});
  std::sort(intervals_.begin(), intervals_.end());   // variant_query_config.cc:69-73
}
void VariantQueryConfig::set_rows_to_query(const std::vector<int64_t>& rows) {
  rows_ = rows;
  std::sort(rows_.begin(), rows_.end());             // variant_query_config.cc:60-67
}
// keys as parsed by the reference: src/main/cpp/src/config/json_config.cc:207-656
void VariantQueryConfig::read_from_file(const std::string& filename, int rank) {
  JsonValue d = gdb::json_parse_file(filename);
  auto per_rank = [&](const JsonValue& v) -> const JsonValue& { return (v.type == JsonValue::ARR && !v.arr.empty()) ? v.arr[std::min<size_t>((size_t)rank, v.arr.size() - 1)] : v; };
  if (const JsonValue* w = d.find("workspace")) workspace_ = per_rank(*w).str;
  if (const JsonValue* a = d.find("array_name")) array_name_ = per_rank(*a).str;
  else if (const JsonValue* a2 = d.find("array")) array_name_ = per_rank(*a2).str;
  const JsonValue* attrs = d.find("attributes");
  if (!attrs) attrs = d.find("query_attributes");
  if (attrs) { attributes_.clear(); for (auto& x : attrs->arr) attributes_.push_back(x.str); }
  intervals_.clear();
  bool scan_full = d.has("scan_full") && d.at("scan_full").as_bool();
  if (!scan_full) if (const JsonValue* q = d.find("query_column_ranges")) {
    const JsonValue& mine = per_rank(*q);
    auto add_item = [&](const JsonValue& it) {
      if (it.type == JsonValue::NUM) add_column_interval_to_query(it.as_int(), it.as_int());
      else if (it.type == JsonValue::ARR && it.arr.size() == 2) add_column_interval_to_query(it.arr[0].as_int(), it.arr[1].as_int());
      else if (it.type == JsonValue::OBJ && it.has("low")) add_column_interval_to_query(it.at("low").as_int(), it.at("high").as_int());
    };
    if (mine.type == JsonValue::OBJ && mine.has("range_list")) for (auto& it : mine.at("range_list").arr) add_item(it);
    else if (mine.type == JsonValue::ARR) for (auto& it : mine.arr) add_item(it);
  }
  rows_.clear();
  if (const JsonValue* q = d.find("query_row_ranges")) {
    const JsonValue& mine = per_rank(*q);
    std::vector<int64_t> rows;
    auto add_item = [&](const JsonValue& it) {
      if (it.type == JsonValue::NUM) rows.push_back(it.as_int());
      else if (it.type == JsonValue::ARR && it.arr.size() == 2) for (int64_t r = it.arr[0].as_int(); r <= it.arr[1].as_int(); ++r) rows.push_back(r);
      else if (it.type == JsonValue::OBJ && it.has("low")) for (int64_t r = it.at("low").as_int(); r <= it.at("high").as_int(); ++r) rows.push_back(r);
    };
    if (mine.type == JsonValue::OBJ && mine.has("range_list")) for (auto& it : mine.at("range_list").arr) add_item(it);
    else if (mine.type == JsonValue::ARR) for (auto& it : mine.arr) add_item(it);
    set_rows_to_query(rows);
  }
  if (const JsonValue* v = d.find("vid_mapping_file")) vid_ = VidMapper(v->str);
  if (const JsonValue* v = d.find("max_diploid_alt_alleles_that_can_be_genotyped")) max_alt_ = (unsigned)v->as_int();
  if (const JsonValue* v = d.find("produce_GT_with_min_PL_value_for_spanning_deletions")) min_pl_gt_ = v->as_bool();
}

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

void VariantQueryProcessor::do_query_bookkeeping(VariantQueryConfig& qc, bool) {
  const VidMapper& vid = qc.get_vid_mapper().is_initialized() ? qc.get_vid_mapper() : vid_;
  std::vector<gdbgpu_field_desc> fd;
  qc.remapped_names_.clear();
  qc.remapped_names_.push_back("GT");
  for (auto& name : qc.get_query_attributes()) {
    const FieldInfo* fi = vid.get_field_info(name);
    if (!fi || !fi->is_allele_dependent()) continue;
    gdbgpu_field_desc d;
    d.name = fi->name.c_str();
    d.length_kind = fi->length == "R" ? GDBGPU_LEN_R : (fi->length == "A" ? GDBGPU_LEN_A : GDBGPU_LEN_G);
    d.elem_type = fi->type == "float" ? GDBGPU_ELEM_FLOAT32 : GDBGPU_ELEM_INT32;
    fd.push_back(d);
    qc.remapped_names_.push_back(fi->name);
  }
  const FieldInfo* gt = vid.get_field_info("GT");
  int phased = gt && gt->length == "PP";
  if (arr_) { gdbgpu_array_close(arr_); arr_ = nullptr; }
  check(gdbgpu_array_open(sm_->ctx(), sm_->get_workspace().c_str(), array_name_.c_str(), fd.data(), (int32_t)fd.size(), phased, &arr_));
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
    if (!qc.query_all_rows()) { p.query_rows = qc.get_rows_to_query().data(); p.n_query_rows = qc.get_rows_to_query().size(); }
    p.max_diploid_alt_alleles = qc.get_max_diploid_alt_alleles_that_can_be_genotyped();
    check(gdbgpu_scan_begin(arr_, &p, &st.scan_));
    st.have_batch_ = false; st.pos_in_batch_ = 0; st.m_done = false;
  }
  Variant variant;
  for (;;) {
    if (!st.have_batch_ || st.pos_in_batch_ >= st.batch_.n_sites) {
      check(gdbgpu_scan_next_batch(st.scan_, 0, 0, &st.batch_));
      st.have_batch_ = true; st.pos_in_batch_ = 0;
      if (st.batch_.n_sites == 0) { st.m_done = true; gdbgpu_scan_end(st.scan_); st.scan_ = nullptr; st.have_batch_ = false; return; }
    }
    while (st.pos_in_batch_ < st.batch_.n_sites) {
      if (op.overflow()) { if (!scan_state) throw GenomicsDBGPUException(GDBGPU_ERR_ARG, "operator overflow without a scan state"); return; }   // caller drains, then re-enters
      fill_variant(st.batch_, st.pos_in_batch_, variant);
      op.operate(variant, qc);
      ++st.pos_in_batch_; ++st.sites_done_;
    }
  }
}

}  // namespace genomicsdb_b200
