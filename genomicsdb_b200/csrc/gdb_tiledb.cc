#include "gdb_tiledb.h"

#include <dirent.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace gdb {

namespace {

enum { T_INT32 = 0, T_INT64 = 1, T_FLOAT32 = 2, T_FLOAT64 = 3, T_CHAR = 4 };
enum { CMP_NONE = 0, CMP_GZIP = 1 };

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

std::vector<uint8_t> read_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) fail("cannot open " + path);
  fseeko(f, 0, SEEK_END);
  off_t n = ftello(f);
  fseeko(f, 0, SEEK_SET);
  std::vector<uint8_t> b((size_t)n);
  if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); fail("short read on " + path); }
  fclose(f);
  return b;
}

uint64_t file_size(const std::string& path) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) fail("cannot stat " + path);
  return (uint64_t)st.st_size;
}
// bytes [off, off + len) of a file
std::vector<uint8_t> read_file_range(const std::string& path, uint64_t off, uint64_t len) {
  std::vector<uint8_t> b((size_t)len);
  if (!len) return b;
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) fail("cannot open " + path);
  uint64_t done = 0;
  while (done < len) {
    ssize_t r = pread(fd, b.data() + done, (size_t)std::min<uint64_t>(len - done, 1ull << 30), (off_t)(off + done));
    if (r <= 0) { close(fd); fail("short read on " + path); }
    done += (uint64_t)r;
  }
  close(fd);
  return b;
}

bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
bool is_file(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

// window_bits: 15 = zlib stream (tiles), 31 = gzip file (book-keeping)
void inflate_into(const uint8_t* src, size_t n, int window_bits, std::vector<uint8_t>& out) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, window_bits) != Z_OK) fail("inflateInit failed");
  zs.next_in = const_cast<Bytef*>(src);
  zs.avail_in = (uInt)n;
  size_t base = out.size();
  size_t cap = std::max<size_t>(n * 4, 1 << 12);
  int rc = Z_OK;
  size_t produced = 0;
  while (rc != Z_STREAM_END) {
    out.resize(base + cap);
    zs.next_out = out.data() + base + produced;
    zs.avail_out = (uInt)std::min<size_t>(cap - produced, 1u << 30);
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) { inflateEnd(&zs); fail("zlib inflate error"); }
    produced = zs.total_out;
    if (rc != Z_STREAM_END && produced == cap) cap *= 2;
    else if (rc == Z_BUF_ERROR && zs.avail_in == 0) { inflateEnd(&zs); fail("truncated zlib stream"); }
  }
  inflateEnd(&zs);
  out.resize(base + produced);
}

struct Cursor {
  const uint8_t* p; size_t n, off = 0;
  template <class T> T get() {
    if (off + sizeof(T) > n) fail("unexpected end of metadata buffer");
    T v; memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v;
  }
  std::string str() {
    int32_t len = get<int32_t>();
    if (len < 0 || off + (size_t)len > n) fail("bad string length in metadata");
    std::string s((const char*)p + off, (size_t)len); off += (size_t)len; return s;
  }
  template <class T> void vec(std::vector<T>& v, size_t count) {
    if (off + count * sizeof(T) > n) fail("unexpected end of metadata buffer");
    v.resize(count); if (count) memcpy(v.data(), p + off, count * sizeof(T)); off += count * sizeof(T);
  }
};

void parse_schema(const std::string& path, TdbSchema& s) {
  std::vector<uint8_t> b = read_file(path);
  Cursor c{b.data(), b.size()};
  uint32_t tag; memcpy(&tag, b.data(), 4);
  if (tag >= 0xFFFFFFF0u) { c.off = 4; (void)c.str(); }   // version tag, then workspace path
  s.array_name = c.str();
  s.dense = c.get<uint8_t>() != 0;
  s.tile_order = c.get<int8_t>();
  s.cell_order = c.get<int8_t>();
  s.capacity = c.get<int64_t>();
  int32_t nattr = c.get<int32_t>();
  for (int i = 0; i < nattr; ++i) s.attributes.push_back(c.str());
  int32_t ndim = c.get<int32_t>();
  for (int i = 0; i < ndim; ++i) (void)c.str();
  int32_t dom = c.get<int32_t>(); c.off += (size_t)dom;
  int32_t ext = c.get<int32_t>(); c.off += (size_t)ext;
  for (int i = 0; i <= nattr; ++i) s.types.push_back(c.get<int8_t>());
  for (int i = 0; i < nattr; ++i) s.cell_val_num.push_back(c.get<int32_t>());
  for (int i = 0; i <= nattr; ++i) s.compression.push_back(c.get<int8_t>());
  if (c.off != b.size()) fail("array schema: trailing bytes");
  if (s.dense || ndim != 2 || s.types[nattr] != T_INT64 || s.cell_order != 1)
    fail("array schema: expected a sparse 2-D int64 column-major array (GenomicsDB layout)");
}

void parse_book_keeping(const std::string& path, const TdbSchema& s, TdbBookKeeping& bk) {
  std::vector<uint8_t> gz = read_file(path), b;
  inflate_into(gz.data(), gz.size(), 15 + 16, b);
  Cursor c{b.data(), b.size()};
  size_t nattr = s.attributes.size();
  uint64_t dom = c.get<uint64_t>(); c.off += (size_t)dom;
  int64_t mbr_num = c.get<int64_t>(); c.off += (size_t)mbr_num * 32;
  int64_t bc_num = c.get<int64_t>();
  c.vec(bk.bounding_coords, (size_t)bc_num * 4);
  bk.tile_offsets.resize(nattr + 1);
  for (size_t i = 0; i <= nattr; ++i) { int64_t k = c.get<int64_t>(); c.vec(bk.tile_offsets[i], (size_t)k); }
  bk.tile_var_offsets.resize(nattr);
  for (size_t i = 0; i < nattr; ++i) { int64_t k = c.get<int64_t>(); c.vec(bk.tile_var_offsets[i], (size_t)k); }
  bk.tile_var_sizes.resize(nattr);
  for (size_t i = 0; i < nattr; ++i) { int64_t k = c.get<int64_t>(); c.vec(bk.tile_var_sizes[i], (size_t)k); }
  bk.last_tile_cell_num = c.get<int64_t>();
  if (c.off != b.size()) fail("book-keeping: trailing bytes");
  bk.tile_num = mbr_num;
  bk.cell_num = mbr_num ? (mbr_num - 1) * s.capacity + bk.last_tile_cell_num : 0;
}

}  // namespace

int TdbSchema::attr_id(const std::string& name) const {
  for (size_t i = 0; i < attributes.size(); ++i) if (attributes[i] == name) return (int)i;
  return -1;
}
size_t TdbSchema::type_size(int id) const {
  switch (types[id]) { case T_INT32: case T_FLOAT32: return 4; case T_INT64: case T_FLOAT64: return 8; default: return 1; }
}

TdbArray::TdbArray(const std::string& workspace, const std::string& array_name) {
  dir_ = workspace + "/" + array_name;
  if (!is_dir(dir_)) fail("array directory not found: " + dir_);
  parse_schema(dir_ + "/__array_schema.tdb", schema_);
  // fragments: visible "__*" directories holding __tiledb_fragment.tdb, ordered by the timestamp after the first '_'
  std::vector<std::pair<std::pair<int64_t, std::string>, std::string>> found;
  DIR* d = opendir(dir_.c_str());
  if (!d) fail("cannot list " + dir_);
  while (dirent* e = readdir(d)) {
    std::string name = e->d_name;
    if (name.size() < 3 || name[0] != '_' || name[1] != '_') continue;
    std::string p = dir_ + "/" + name;
    if (!is_dir(p) || !is_file(p + "/__tiledb_fragment.tdb")) continue;
    size_t us = name.find('_', 2);
    int64_t ts = (us == std::string::npos) ? 0 : strtoll(name.c_str() + us + 1, nullptr, 10);
    found.push_back({{ts, name}, p});
  }
  closedir(d);
  std::sort(found.begin(), found.end());
  for (auto& f : found) {
    TdbFragment fr; fr.dir = f.second;
    parse_book_keeping(fr.dir + "/__book_keeping.tdb.gz", schema_, fr.bk);
    fragments_.push_back(std::move(fr));
  }
  // GenomicsDB row bounds: min lb / max ub over all metadata files
  std::string md = dir_ + "/genomicsdb_meta_dir";
  if (DIR* m = opendir(md.c_str())) {
    while (dirent* e = readdir(m)) {
      std::string name = e->d_name;
      if (name.rfind("genomicsdb_meta_", 0) != 0) continue;
      std::vector<uint8_t> b = read_file(md + "/" + name);
      std::string txt((const char*)b.data(), b.size());
      auto grab = [&](const char* key, int64_t& v) {
        size_t p = txt.find(std::string("\"") + key + "\"");
        if (p == std::string::npos) return false;
        p = txt.find(':', p);
        if (p == std::string::npos) return false;
        v = strtoll(txt.c_str() + p + 1, nullptr, 10);
        return true;
      };
      int64_t v;
      if (grab("lb_row_idx", v)) { lb_ = has_lb_ ? std::min(lb_, v) : v; has_lb_ = true; }
      if (grab("max_valid_row_idx_in_array", v)) { ub_ = has_ub_ ? std::max(ub_, v) : v; has_ub_ = true; }
    }
    closedir(m);
  }
}

bool TdbArray::row_bounds(int64_t& lb, int64_t& max_valid) const {
  lb = has_lb_ ? lb_ : 0;
  max_valid = ub_;
  return has_ub_;
}

void TdbArray::tile_range(size_t fi, int64_t col_lo, int64_t col_hi, int64_t& tile_lo, int64_t& tile_hi) const {
  const TdbBookKeeping& bk = fragments_[fi].bk;
  const int64_t n = bk.tile_num;
  if ((int64_t)bk.bounding_coords.size() != 4 * n) fail("book-keeping: bounding coordinates do not match the tile count");
  // first tile whose last cell's column is >= col_lo
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (bk.bounding_coords[4 * mid + 3] < col_lo) lo = mid + 1; else hi = mid; }
  tile_lo = lo;
  // first tile whose first cell's column is > col_hi
  hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (bk.bounding_coords[4 * mid + 1] <= col_hi) lo = mid + 1; else hi = mid; }
  tile_hi = std::max(lo, tile_lo);
}
int64_t TdbArray::tile_cells(size_t fi, int64_t tile_lo, int64_t tile_hi) const {
  const TdbBookKeeping& bk = fragments_[fi].bk;
  if (tile_hi <= tile_lo) return 0;
  return (tile_hi - tile_lo) * schema_.capacity - (tile_hi == bk.tile_num ? schema_.capacity - bk.last_tile_cell_num : 0);
}

void TdbArray::read_column_raw(size_t fi, int attr_id, TdbRawColumn& out) const { read_column_raw(fi, attr_id, out, 0, fragments_[fi].bk.tile_num); }

void TdbArray::read_column_raw(size_t fi, int attr_id, TdbRawColumn& out, int64_t t0, int64_t t1, bool lazy) const {
  const TdbFragment& fr = fragments_[fi];
  const int nattr = (int)schema_.attributes.size();
  const std::string name = (attr_id == nattr) ? "__coords" : schema_.attributes[attr_id];
  if (t0 < 0 || t1 > fr.bk.tile_num || t0 > t1) fail("tile range out of bounds for " + name);
  out = TdbRawColumn();
  out.compressed = schema_.compression[attr_id] != CMP_NONE;
  if (out.compressed && schema_.compression[attr_id] != CMP_GZIP) fail("unsupported TileDB compression for " + name);
  out.is_var = attr_id < nattr && schema_.is_var(attr_id);
  const uint64_t cell = out.is_var ? 8 : ((attr_id == nattr) ? 16 : schema_.type_size(attr_id) * (uint64_t)schema_.cell_val_num[attr_id]);
  const uint64_t cap = (uint64_t)schema_.capacity;
  const bool to_end = t1 == fr.bk.tile_num;
  out.cells = (uint64_t)tile_cells(fi, t0, t1);
  out.fixed_bytes = out.cells * cell;
  const std::string fixed_path = fr.dir + "/" + name + ".tdb", var_path = fr.dir + "/" + name + "_var.tdb";
  out.fixed_path = fixed_path; out.var_path = var_path;
  if (t0 == t1) return;
  auto fetch = [&](const std::string& path, uint64_t off, uint64_t len, std::vector<uint8_t>& bytes, uint64_t& o_off, uint64_t& o_len) {
    o_off = off; o_len = len;
    if (!lazy) bytes = read_file_range(path, off, len);
  };
  if (!out.compressed) {
    // tile i of an uncompressed attribute starts at i x capacity x cell size (read_state.cc:2581)
    if (file_size(fixed_path) != (uint64_t)fr.bk.cell_num * cell) fail("column size mismatch for " + name);
    fetch(fixed_path, (uint64_t)t0 * cap * cell, out.fixed_bytes, out.fixed_file, out.fixed_off, out.fixed_len);
    if (out.is_var) {
      uint64_t v0, v1;
      { std::vector<uint8_t> first = read_file_range(fixed_path, (uint64_t)t0 * cap * 8, 8); memcpy(&v0, first.data(), 8); }
      if (to_end) v1 = file_size(var_path);
      else { std::vector<uint8_t> nx = read_file_range(fixed_path, (uint64_t)t1 * cap * 8, 8); memcpy(&v1, nx.data(), 8); }
      if (v1 < v0) fail("value offsets out of order for " + name);
      out.var_base = v0; out.var_bytes = v1 - v0;
      fetch(var_path, v0, v1 - v0, out.var_file, out.var_off, out.var_len);
    }
    return;
  }
  // compressed: tile i of the fixed file is one zlib stream at tile_offsets[i] holding capacity cells (the last tile
  // last_tile_cell_num), read_state.cc:2581-2620
  const std::vector<int64_t>& offs = fr.bk.tile_offsets[attr_id];
  if ((int64_t)offs.size() != fr.bk.tile_num) fail("tile offset count mismatch for " + name);
  const uint64_t fsize = file_size(fixed_path);
  const uint64_t f_lo = (uint64_t)offs[t0], f_hi = to_end ? fsize : (uint64_t)offs[t1];
  if (f_hi < f_lo || f_hi > fsize) fail("tile offsets out of range for " + name);
  fetch(fixed_path, f_lo, f_hi - f_lo, out.fixed_file, out.fixed_off, out.fixed_len);
  uint64_t o = 0;
  for (int64_t i = t0; i < t1; ++i) {
    const uint64_t lo = (uint64_t)offs[i], hi = (i + 1 < fr.bk.tile_num) ? (uint64_t)offs[i + 1] : fsize;
    if (hi < lo || hi > fsize) fail("tile offsets out of range for " + name);
    const uint64_t cells = (i + 1 < fr.bk.tile_num) ? cap : (uint64_t)fr.bk.last_tile_cell_num;
    out.fixed_tiles.push_back(TdbTileSpan{lo - f_lo, hi - lo, o, cells * cell});
    o += cells * cell;
  }
  if (o != out.fixed_bytes) fail("tile sizes do not add up for " + name);
  if (out.is_var) {
    const std::vector<int64_t>& voffs = fr.bk.tile_var_offsets[attr_id];
    const std::vector<uint64_t>& vsz = fr.bk.tile_var_sizes[attr_id];
    if (voffs.size() != vsz.size() || (int64_t)voffs.size() != fr.bk.tile_num) fail("var tile table mismatch for " + name);
    const uint64_t vsize = file_size(var_path);
    for (int64_t i = 0; i < t0; ++i) out.var_base += vsz[i];
    const uint64_t v_lo = (uint64_t)voffs[t0], v_hi = to_end ? vsize : (uint64_t)voffs[t1];
    if (v_hi < v_lo || v_hi > vsize) fail("var tile offsets out of range for " + name);
    fetch(var_path, v_lo, v_hi - v_lo, out.var_file, out.var_off, out.var_len);
    uint64_t vo = 0;
    for (int64_t i = t0; i < t1; ++i) {
      const uint64_t lo = (uint64_t)voffs[i], hi = (i + 1 < fr.bk.tile_num) ? (uint64_t)voffs[i + 1] : vsize;
      if (hi < lo || hi > vsize) fail("var tile offsets out of range for " + name);
      if (vsz[i] == 0) { out.var_tiles.push_back(TdbTileSpan{lo - v_lo, 0, vo, 0}); continue; }   // empty var tile: nothing was written
      out.var_tiles.push_back(TdbTileSpan{lo - v_lo, hi - lo, vo, vsz[i]});
      vo += vsz[i];
    }
    out.var_bytes = vo;
  }
}

}  // namespace gdb
