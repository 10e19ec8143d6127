// Host-side reader for the on-disk format GenomicsDB arrays use (TileDB 0.5.x sparse, column-major, read mode).
// Written from the format, not from the reference's classes: metadata parsed here, attribute files handed out whole and raw.
//   schema layout        dependencies/TileDB/core/src/array/array_schema.cc:569-594
//   book-keeping layout  dependencies/TileDB/core/src/fragment/book_keeping.cc:217-243
//   tile location/inflate dependencies/TileDB/core/src/fragment/read_state.cc:2442-2889, misc/utils.cc:466-512
//   fragment ordering    dependencies/TileDB/core/src/storage_manager/storage_manager.cc:2169-2207
//   array row bounds     src/main/cpp/src/genomicsdb/variant_storage_manager.cc:313-377
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gdb {

struct TdbSchema {
  std::string array_name;
  bool dense = false;
  int tile_order = 0, cell_order = 0;
  int64_t capacity = 0;
  std::vector<std::string> attributes;   // without __coords
  std::vector<int> types;                // attribute_num + 1 (coords last)
  std::vector<int> cell_val_num;         // attribute_num
  std::vector<int> compression;          // attribute_num + 1
  int attr_id(const std::string& name) const;
  bool is_var(int id) const { return cell_val_num[id] == 0x7FFFFFFF; }
  size_t type_size(int id) const;
};

struct TdbBookKeeping {
  int64_t tile_num = 0;
  int64_t last_tile_cell_num = 0;
  int64_t cell_num = 0;
  std::vector<int64_t> bounding_coords;               // tile_num x 4 (first row, first col, last row, last col)
  std::vector<std::vector<int64_t>> tile_offsets;      // attribute_num + 1
  std::vector<std::vector<int64_t>> tile_var_offsets;  // attribute_num
  std::vector<std::vector<uint64_t>> tile_var_sizes;   // attribute_num
};

// An attribute file as it lies on disk, plus where its tiles are: what the device-side inflate needs.  For an uncompressed
// attribute the file bytes ARE the column.
struct TdbTileSpan { uint64_t in_off, in_len, out_off, out_len; };
struct TdbRawColumn {
  bool compressed = false;
  std::vector<uint8_t> fixed_file, var_file;     // <ATTR>.tdb, <ATTR>_var.tdb as read
  std::vector<TdbTileSpan> fixed_tiles, var_tiles;   // compressed only: one zlib stream per tile
  uint64_t fixed_bytes = 0, var_bytes = 0;       // decompressed sizes
  bool is_var = false;
  // lazy reads: where the bytes are instead of the bytes (fixed_file / var_file stay empty); in_off of the tile spans is
  // relative to fixed_off / var_off as it is relative to the start of fixed_file / var_file otherwise
  std::string fixed_path, var_path;
  uint64_t fixed_off = 0, fixed_len = 0, var_off = 0, var_len = 0;
  uint64_t cells = 0;        // cells of the tiles that were read
  uint64_t var_base = 0;     // byte offset, in the fragment's whole value stream, of the first value byte that was read: the
                             // offsets in the fixed file are absolute (read_state.cc:2991-3001 rebases them the same way)
};

struct TdbFragment {
  std::string dir;
  TdbBookKeeping bk;
};

class TdbArray {
 public:
  // throws std::runtime_error on malformed input
  TdbArray(const std::string& workspace, const std::string& array_name);
  const TdbSchema& schema() const { return schema_; }
  size_t num_fragments() const { return fragments_.size(); }
  const TdbFragment& fragment(size_t i) const { return fragments_[i]; }
  // lb_row_idx / max_valid_row_idx_in_array merged over genomicsdb_meta_dir (false if no metadata present)
  bool row_bounds(int64_t& lb, int64_t& max_valid) const;
  // One attribute column of one fragment exactly as it lies on disk + its tile table from the book-keeping
  // (read_state.cc:2442-2889); attr_id == schema().attributes.size() reads __coords.  Compressed tiles are inflated on
  // the device (gdb_inflate.cuh), never here; zlib is only used for the small gzip'ed book-keeping file.
  void read_column_raw(size_t fragment_idx, int attr_id, TdbRawColumn& out) const;
  // The same for tiles [tile_lo, tile_hi) only: the files are read with pread at the tiles' offsets, nothing else is touched.
  // lazy: only say where the bytes are (TdbRawColumn::*_path, *_off, *_len); the caller streams them (gdb_engine.cu: stage_files)
  void read_column_raw(size_t fragment_idx, int attr_id, TdbRawColumn& out, int64_t tile_lo, int64_t tile_hi, bool lazy = false) const;
  // Tiles of a fragment that can hold cells of columns [col_lo, col_hi]: binary search of the book-keeping's bounding
  // coordinates (first / last cell of every tile, global order = column, then row), what
  // ReadState::compute_tile_search_range_col_or_row does (read_state.cc:1427-1527).  tile_lo == tile_hi: none.
  void tile_range(size_t fragment_idx, int64_t col_lo, int64_t col_hi, int64_t& tile_lo, int64_t& tile_hi) const;
  int64_t tile_cells(size_t fragment_idx, int64_t tile_lo, int64_t tile_hi) const;

 private:
  std::string dir_;
  TdbSchema schema_;
  std::vector<TdbFragment> fragments_;
  bool has_lb_ = false, has_ub_ = false;
  int64_t lb_ = 0, ub_ = 0;
};

}  // namespace gdb
