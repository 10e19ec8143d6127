/*
 * libgdbgpu -- C ABI of the B200-native GenomicsDB variant-combine hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, int status returns (0 = ok, <0 = error, message via
 * gdbgpu_last_error), no C++ or torch types.  Every entry point names the reference interface it stands in for
 * (paths relative to the GenomicsDB v0.10.2 source tree).  The host-side C++ mirror of the reference's plugin surface
 * (genomicsdb_b200/csrc/gdb_host_api.h: VariantQueryProcessor::scan_and_operate, SingleVariantOperatorBase::operate)
 * is written on top of exactly these calls; INTEGRATION.md shows the binding a GenomicsDB maintainer would add.
 *
 * There is NO CPU fallback behind this ABI: every call that computes needs a CUDA device (sm_100a) and fails with
 * GDBGPU_ERR_CUDA otherwise.
 */
#ifndef GDBGPU_H
#define GDBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDBGPU_OK              0
#define GDBGPU_ERR_ARG        -1
#define GDBGPU_ERR_IO         -2   /* fragment files missing / malformed */
#define GDBGPU_ERR_CUDA       -3   /* no device, launch failure, out of memory */
#define GDBGPU_ERR_OVERLAP    -4   /* "Unhandled overlapping variants" (query_variants.cc:533-535) */
#define GDBGPU_ERR_LIMIT      -5   /* a documented engine limit was exceeded (message says which) */

typedef struct gdbgpu_ctx gdbgpu_ctx;
typedef struct gdbgpu_array gdbgpu_array;
typedef struct gdbgpu_scan gdbgpu_scan;

/* Length descriptor of an allele-dependent field (known_field_info.h:95-101: BCF_VL_R / BCF_VL_A / BCF_VL_G). */
enum { GDBGPU_LEN_R = 0, GDBGPU_LEN_A = 1, GDBGPU_LEN_G = 2 };
/* Element type of a 4-byte field; decides the BCF missing / vector-end words written (vcf.h:61-134). */
enum { GDBGPU_ELEM_INT32 = 0, GDBGPU_ELEM_FLOAT32 = 1, GDBGPU_ELEM_CHAR = 2 /* plain fields only */ };

/* One remapped field = one TileDB attribute whose length depends on the allele list (PL, AD, MLEAC, ...);
 * stands in for GA4GHOperator::m_remapped_fields_query_idxs (variant_operations.cc:410-425). */
typedef struct gdbgpu_field_desc {
  const char* name;      /* TileDB attribute name, e.g. "PL" */
  int32_t length_kind;   /* GDBGPU_LEN_* */
  int32_t elem_type;     /* GDBGPU_ELEM_* */
} gdbgpu_field_desc;

/* A FORMAT field whose length does not depend on the allele list (GQ, MIN_DP, DP_FORMAT, SB, PGT, PID, ...): never
 * remapped, only collated over the query rows and BCF-encoded (GDBGPU_SCAN_BCF_FORMAT); the other entries of
 * BroadCombinedGVCFOperator::m_FORMAT_fields_vec (broad_combined_gvcf.cc:168-232,603-727). */
/* The same descriptor stages the scalar INFO fields the Broad operator combines over the valid calls of a site
 * (BroadCombinedGVCFOperator::handle_INFO_fields, broad_combined_gvcf.cc:374-420,524-545; median / sum / mean of the first
 * element: VariantFieldHandler::get_valid_median / get_valid_sum / get_valid_mean, variant_field_handler.cc:529-612). */
enum { GDBGPU_INFO_NONE = 0, GDBGPU_INFO_MEDIAN = 1, GDBGPU_INFO_SUM = 2, GDBGPU_INFO_MEAN = 3,
       GDBGPU_INFO_ELEMENT_WISE_SUM = 4 /* vector result: VariantFieldHandler::compute_valid_element_wise_sum, variant_field_handler.cc:618-660 */ };
#define GDBGPU_INFO_VECTOR_CAP 1024   /* longest element-wise result */
typedef struct gdbgpu_plain_field_desc {
  const char* name;      /* TileDB attribute name, e.g. "MIN_DP" */
  int32_t elem_type;     /* GDBGPU_ELEM_* (CHAR: a string per cell) */
  int32_t fixed_len;     /* values per cell; 0 = variable length */
  int32_t no_format;     /* 1: the field is not a FORMAT field (no block is produced for it); 2: a block is produced for a host-side
                          * combine (ID / FILTER union) and always padded with vector_end, whatever the scan's conventions; 4: as 2, and
                          * calls in spanning-deletion form have no value (INFO fields: concatenate; broad_combined_gvcf.cc:1067-1075);
                          * 3: raw rows for a host-side combine of a multi-dimensional field (variant_operations.cc:482-555,
                          * broad_combined_gvcf.cc:432-520; CHAR, variable length): a BCF_BT_CHAR block whose row is
                          * [u32 byte length, 0xFFFFFFFF = no value][the cell's bytes][zero padding], calls in spanning-deletion form
                          * have no value */
  int32_t info_op;       /* GDBGPU_INFO_*: combine the field over the valid calls of every site (4-byte fixed-length fields) */
} gdbgpu_plain_field_desc;
#define GDBGPU_MAX_PLAIN_FIELDS 24

/* Host-resident columns in TileDB's own in-file layout (global cell order = column-major (col,row)); what
 * ArrayReadState::read_sparse hands to GenomicsDB one segment at a time (array_read_state.cc:1529-1598), here whole.
 * Variable attributes: byte offsets (n_cells entries, absolute into the value stream, as in <ATTR>.tdb) + stream. */
typedef struct gdbgpu_var_column {
  const uint64_t* offsets;   /* n_cells absolute byte offsets */
  const void* values;        /* value stream */
  uint64_t values_bytes;
} gdbgpu_var_column;

typedef struct gdbgpu_columns {
  uint64_t n_cells;
  const int64_t* coords;     /* n_cells x (row, col)  == __coords.tdb */
  const int64_t* end;        /* END.tdb */
  gdbgpu_var_column ref;     /* REF.tdb / REF_var.tdb (char) */
  gdbgpu_var_column alt;     /* ALT (char, '|' separated, NON_REF stored as '&') */
  gdbgpu_var_column gt;      /* GT (int32); offsets==NULL when the array has no GT */
  int32_t gt_phased;         /* vid length "PP": [a0, phase, a1, ...] (known_field_info.h:142-156) */
  int32_t n_fields;
  const gdbgpu_field_desc* field_descs;
  const gdbgpu_var_column* fields;   /* n_fields columns of 4-byte elements */
  int64_t lb_row_idx;        /* genomicsdb_meta lb_row_idx */
  int64_t num_rows;          /* max_valid_row_idx_in_array - lb_row_idx + 1 */
  int32_t n_plain_fields;    /* optional: plain FORMAT fields (see gdbgpu_plain_field_desc) */
  const gdbgpu_plain_field_desc* plain_descs;
  const gdbgpu_var_column* plain_fields;   /* fixed-length field: offsets == NULL, values = n_cells x fixed_len elements */
} gdbgpu_columns;

typedef struct gdbgpu_array_info {
  uint64_t n_cells;          /* after the fragment merge */
  uint64_t n_begin_cells;    /* cells with col <= END */
  uint32_t n_fragments;
  int64_t lb_row_idx, num_rows;
  int64_t min_col, max_col, max_end;
  uint64_t staged_bytes;     /* fragment bytes copied host -> HBM */
  uint64_t hbm_bytes;        /* total device footprint incl. indexes */
  int32_t n_fields, has_gt, gt_phased;
  int32_t n_plain_fields;
} gdbgpu_array_info;

/* ---- context: one per host thread / CUDA device (VariantStorageManager's role, variant_storage_manager.h:160-235) */
int gdbgpu_ctx_create(int device, gdbgpu_ctx** out);
void gdbgpu_ctx_destroy(gdbgpu_ctx* ctx);
const char* gdbgpu_last_error(const gdbgpu_ctx* ctx);
/* the CUDA stream all work of this context is enqueued on (cudaStream_t as void*), for event timing by the caller */
void* gdbgpu_ctx_stream(gdbgpu_ctx* ctx);
/* number of kernel launches issued by this context so far */
uint64_t gdbgpu_ctx_launch_count(const gdbgpu_ctx* ctx);
int gdbgpu_ctx_synchronize(gdbgpu_ctx* ctx);
/* Device memory given back by closed arrays / ended scans is kept by the context and reused by later opens and scans (no
 * cudaMalloc / cudaFree in steady state).  This call returns what is currently cached to the driver. */
int gdbgpu_ctx_release_cached_memory(gdbgpu_ctx* ctx);
/* Per-kernel device time, measured with CUDA events on the launching stream (the reference's DO_PROFILING timers,
 * query_variants.h:67-124, have this role).  Off by default; stats are summed over kernels whose name contains
 * `kernel_name` (e.g. "k_remap"). */
void gdbgpu_ctx_set_profiling(gdbgpu_ctx* ctx, int enable);
void gdbgpu_ctx_reset_kernel_stats(gdbgpu_ctx* ctx);
int gdbgpu_ctx_kernel_stats(gdbgpu_ctx* ctx, const char* kernel_name, double* total_ms, uint64_t* launches);

/* ---- array: fragment bytes staged to HBM.
 * gdbgpu_array_open replaces VariantStorageManager::open_array + TileDB array/fragment/book-keeping load
 * (variant_storage_manager.cc:434-520; storage_manager.cc array_open; book_keeping.cc:368-422; read_state.cc:2442-2889):
 * it reads <workspace>/<array>/__array_schema.tdb and, per fragment dir, __book_keeping.tdb.gz, <ATTR>.tdb, <ATTR>_var.tdb,
 * copies the files' bytes to the device as they lie on disk, inflates GZIP tiles there (one zlib stream per tile and
 * attribute file) and runs the fragment merge there. */
int gdbgpu_array_open(gdbgpu_ctx* ctx, const char* workspace, const char* array_name,
                      const gdbgpu_field_desc* fields, int32_t n_fields, int32_t gt_phased, gdbgpu_array** out);
/* Same, also staging plain FORMAT fields for the FORMAT collation of GDBGPU_SCAN_BCF_FORMAT. */
int gdbgpu_array_open_ex(gdbgpu_ctx* ctx, const char* workspace, const char* array_name,
                         const gdbgpu_field_desc* fields, int32_t n_fields,
                         const gdbgpu_plain_field_desc* plain_fields, int32_t n_plain_fields,
                         int32_t gt_phased, gdbgpu_array** out);
/* Same, staging only the tiles that can hold cells of columns [col_lo, col_hi] (one rank's column partition):
 * the tile range comes from a binary search of the book-keeping's bounding coordinates, as
 * ReadState::compute_tile_search_range_col_or_row / get_next_overlapping_tile_sparse find it
 * (dependencies/TileDB/core/src/fragment/read_state.cc:1427-1527,1006-1042); only those tiles are read (pread), copied and
 * inflated.  The call live at a query's left edge is found through the row's first cell at or after that edge (its END
 * copy; gt_get_column's forward scan, query_variants.cc:845-941): the staged range grows to the right until every row shows a
 * cell at or after col_hi (or the fragments end), so every query inside [col_lo, col_hi] gives what it gives on the whole array.  INT64_MIN / INT64_MAX: the whole array. */
int gdbgpu_array_open_range(gdbgpu_ctx* ctx, const char* workspace, const char* array_name,
                            const gdbgpu_field_desc* fields, int32_t n_fields,
                            const gdbgpu_plain_field_desc* plain_fields, int32_t n_plain_fields,
                            int32_t gt_phased, int64_t col_lo, int64_t col_hi, gdbgpu_array** out);
/* Column cuts for n_parts ranks with about equal cell counts, from the fragments' book-keeping alone (tile bounding
 * coordinates and tile sizes; no attribute file is read): cuts[0] = first column, cuts[n_parts] = one past the last, rank r
 * owns [cuts[r], cuts[r + 1] - 1] -- SURVEY 8e; the reference leaves this to the loader's column_partitions
 * (genomicsdb_config_base.cc:205-223) or to ColumnHistogramOperator::equi_partition_and_print_bins.  Host only. */
int gdbgpu_array_column_cuts(const char* workspace, const char* array_name, int32_t n_parts, int64_t* cuts);

/* K0 on its own: inflates n zlib streams (RFC 1950; what TileDB's gunzip does per tile, dependencies/TileDB/core/src/misc/
 * utils.cc:466-512 via ReadState::decompress_tile_gzip, fragment/read_state.cc:1595-1702) on the device, one warp per
 * stream.  Stream i is in[in_offsets[i] .. in_offsets[i+1]) and must inflate to exactly out_offsets[i+1] - out_offsets[i]
 * bytes at out + out_offsets[i] (host buffers; n + 1 offsets each).  GDBGPU_ERR_IO when a stream is corrupt (header, data,
 * size or Adler-32); kernel_ms (optional) receives the kernel's CUDA-event time. */
int gdbgpu_inflate_streams(gdbgpu_ctx* ctx, const void* in, const uint64_t* in_offsets, void* out, const uint64_t* out_offsets,
                           uint64_t n, double* kernel_ms);
/* Same, from columns already decoded in host memory (loader hand-off, JNI buffers, synthetic generators). */
int gdbgpu_array_from_columns(gdbgpu_ctx* ctx, const gdbgpu_columns* cols, gdbgpu_array** out);
int gdbgpu_array_get_info(const gdbgpu_array* arr, gdbgpu_array_info* info);
/* Host copies of the merged coordinate / END columns (for consumers that materialise VariantCall objects). */
int gdbgpu_array_copy_cells(const gdbgpu_array* arr, uint64_t first, uint64_t count, int64_t* coords, int64_t* end);
void gdbgpu_array_close(gdbgpu_array* arr);

/* ---- scan: one query column interval == one VariantQueryProcessor::scan_and_operate call chain
 * (query_variants.h:241-243, query_variants.cc:334-476) with handle_spanning_deletions as the Broad operator asks. */
#define GDBGPU_SCAN_FULL               1u   /* query_config.get_num_column_intervals()==0 (scan_full) */
#define GDBGPU_SCAN_NO_DELETIONS       2u   /* handle_spanning_deletions=false and no handle_deletions pre-pass (GA4GHOperator only) */
#define GDBGPU_SCAN_MIN_PL_GT          4u   /* produce_GT_with_min_PL_value_for_spanning_deletions */
#define GDBGPU_SCAN_WANT_LUT           8u   /* also return the per-call input->merged allele LUT (CombineAllelesLUT) */
#define GDBGPU_SCAN_DEVICE_OUTPUT     16u   /* batch pointers are DEVICE pointers (consumer is another kernel) */
#define GDBGPU_SCAN_BCF_FORMAT        32u   /* also collate GT + the remapped fields over all query rows and encode them as BCF typed
                                             * FORMAT blocks (collect_and_extend_fields, variant_field_handler.cc:804-871; encode_GT_vector,
                                             * broad_combined_gvcf.cc:53-138; bcf_enc_vint, htslib vcf.c:1926-1972).  With host output only the
                                             * blocks, the site table and the alleles are copied back (fields[].data / .len and live are NULL) */
#define GDBGPU_SCAN_PRODUCE_GT        64u   /* query JSON "produce_GT_field": GT blocks carry the alleles, else no-calls (broad_combined_gvcf.cc:60-76) */
#define GDBGPU_SCAN_NO_PREFETCH      128u   /* do not start batch n+1 while batch n is being handed out */
#define GDBGPU_SCAN_HTSJDK_CONVENTIONS 256u /* FORMAT blocks as the JNI path needs them for htsjdk's BCF2 codec (BroadCombinedGVCFOperator's
                                             * use_missing_values_only_not_vector_end, genomicsdb_bcf_generator.cc:60-83; variant_field_handler.cc:
                                             * 845-866): rows padded with missing instead of vector_end, a row without GT is a no-call, strings
                                             * vector_end only */
#define GDBGPU_SCAN_SITE_STATS      1024u   /* batches carry site_stats (what InterestingLocationsPrinter prints, variant_operations.cc:380-400) */
#define GDBGPU_SCAN_SUMMARY_ONLY    2048u   /* host output limited to the per-site tables (sites, merged alleles, n_live, site_stats): live and
                                             * fields[].data / .len are NULL, nothing per call is copied */
#define GDBGPU_SCAN_CELLS_ONLY       512u   /* the scan serves gdbgpu_scan_cells* only (the columnar iterator neither merges calls nor minds
                                             * overlapping ones): no sweep state is built, gdbgpu_scan_next_batch is refused */

typedef struct gdbgpu_scan_params {
  int64_t col_lo, col_hi;            /* inclusive TileDB column interval; ignored with GDBGPU_SCAN_FULL */
  const int64_t* query_rows;         /* array row idxs, any order, NULL = all rows (query_row_ranges) */
  uint64_t n_query_rows;
  uint32_t flags;
  uint32_t max_diploid_alt_alleles;  /* 0 -> 50 (gt_common.h:48) */
  /* GATK-style DP (VCF_FIELD_COMBINE_OPERATION_DP, broad_combined_gvcf.cc:688-727): 1-based indices into the array's plain
   * fields of the INFO DP attribute, MIN_DP and DP_FORMAT (0 = not queried); INFO/DP of a site = sum over its rows of
   * DP, else MIN_DP, else DP_FORMAT.  Used with GDBGPU_SCAN_BCF_FORMAT. */
  uint32_t dp_info_field, dp_min_dp_field, dp_format_field;
  /* bit f set: also combine remapped field f (index into the array's allele-dependent fields, GT not counted) over the valid
   * calls of every site with an element-wise sum -- INFO fields like MLEAC / MLEAF with VCF_field_combine_operation
   * element_wise_sum.  Used with GDBGPU_SCAN_BCF_FORMAT. */
  uint32_t esum_remapped_mask;
} gdbgpu_scan_params;

typedef struct gdbgpu_scan_info {
  uint64_t n_sites;                  /* operate() invocations this interval will produce */
  uint64_t n_breakpoints;
  uint64_t n_query_rows;
  uint64_t n_forward_cells;          /* cells visited by the forward scan */
} gdbgpu_scan_info;

/* One remapped field of a batch: dense, sample-major.  Site s occupies words
 * [base[s], base[s] + n_query_rows*stride[s]) ; query row q's vector starts at base[s] + q*stride[s], holds len[s*nq+q]
 * valid words followed by BCF vector-end padding.  len==0: field invalid for that call / row not live / suppressed. */
typedef struct gdbgpu_field_out {
  const uint32_t* data;
  const uint64_t* base;      /* n_sites + 1 */
  const uint32_t* stride;    /* n_sites */
  const uint16_t* len;       /* n_sites x n_query_rows */
} gdbgpu_field_out;

#define GDBGPU_SITE_NONREF      1u   /* m_NON_REF_exists */
#define GDBGPU_SITE_REMAPPED    2u   /* m_remapping_needed */
#define GDBGPU_SITE_G_SUPPRESSED 4u  /* too many ALT alleles: G-length fields dropped (variant_operations.cc:613-629) */

#define GDBGPU_CALL_REDUCED  0x40000000   /* live[] flag (bit 30): call is in its spanning-deletion form (REF N, ALT *[,NON_REF]) */

typedef struct gdbgpu_batch {
  uint64_t first_site;       /* index of the first site of this batch within the interval */
  uint32_t n_sites;
  uint64_t n_query_rows;
  const int64_t* site_begin; /* variant.get_column_begin() */
  const int64_t* site_end;   /* variant.get_column_end() */
  const uint8_t* site_flags; /* GDBGPU_SITE_* */
  const uint32_t* n_alleles; /* 1 + merged ALT count */
  const uint32_t* n_live;    /* valid calls at the site */
  /* merged allele strings: allele m of site s is allele_bytes[allele_pos[2*(allele_first[s]+m)] .. +allele_pos[..+1]) */
  const uint32_t* allele_first;  /* n_sites */
  const uint32_t* allele_pos;    /* (offset,len) pairs */
  const uint8_t* allele_bytes;
  /* live[s*nq+q] = merged-order cell index of query row q's call at site s (bit 30: GDBGPU_CALL_REDUCED), or -1 */
  const int32_t* live;
  /* fields[0] is GT when the array has GT, then the remapped fields in array order */
  int32_t n_fields;
  gdbgpu_field_out fields[16];
  /* optional (GDBGPU_SCAN_WANT_LUT): in2m[lut_first[s*nq+q] + i] = merged idx of input allele i < lut_len[s*nq+q] (-1 missing) */
  const uint32_t* lut_first;
  const uint16_t* lut_len;   /* number of input alleles (1 + ALT count) of the call as the operator saw it */
  const int16_t* lut_in2m;
  uint64_t out_bytes;        /* algorithmic bytes of remapped output of this batch: 4 x n_query_rows x sum of the strides
                              * each field needs (alignment padding and the length arrays not counted) */
  /* GDBGPU_SCAN_BCF_FORMAT: n_fmt_fields = n_fields + the array's plain FORMAT fields (in the order they were given to
   * gdbgpu_array_open_ex).  The FORMAT block of field o (fields[] order, then the plain fields) at site s is described by entry
   * o*n_sites + s: fmt_type = 0 when the field is not emitted at the site (no valid call has it / G field suppressed),
   * else the BCF type (1 int8, 2 int16, 3 int32, 5 float, 7 char); fmt_nps values per sample; the n_query_rows*fmt_nps values
   * start at fmt_bytes + fmt_offset (16-byte aligned), sample-major, little-endian: exactly the bytes htslib puts into
   * bcf1_t.indiv after the field's typed key and (fmt_nps, fmt_type) size descriptor. */
  int32_t n_fmt_fields;
  const uint8_t* fmt_type;
  const uint32_t* fmt_nps;
  const uint64_t* fmt_offset;
  const uint8_t* fmt_bytes;
  uint64_t fmt_total_bytes;
  /* combined INFO scalars (host arrays): entry p*n_sites + s for plain field p (info_op != 0): info_valid = some valid call
   * of the site has the field, info_value = the 4-byte result (int32 or float bits); info_dp_sum[s] = the DP rule's sum */
  const uint8_t* info_valid;
  const uint32_t* info_value;
  const int32_t* info_dp_sum;
  /* element-wise sums (host arrays): slot v = remapped field f (v = f) or plain field p with GDBGPU_INFO_ELEMENT_WISE_SUM
   * (v = n_remapped_fields + p); entry v*n_sites + s: vec_len values (0: no valid value) starting at vec_values + vec_offset;
   * elements no call had a valid value for are BCF missing */
  const uint32_t* vec_len;
  const uint64_t* vec_offset;
  const uint32_t* vec_values;
  const uint32_t* site_stats; /* GDBGPU_SCAN_SITE_STATS (host memory, else NULL): [2s] live calls that are reference blocks (REF of one
                               * base, ALT = <NON_REF> alone), [2s + 1] live calls that begin at the site */
} gdbgpu_batch;

int gdbgpu_scan_begin(gdbgpu_array* arr, const gdbgpu_scan_params* params, gdbgpu_scan** out);
int gdbgpu_scan_get_info(const gdbgpu_scan* scan, gdbgpu_scan_info* info);
/* Produces the next <= max_sites sites whose remapped output fits max_bytes (0 = engine default); n_sites==0 at the end.
 * The batch stays valid until the next call on the same scan.  Unless GDBGPU_SCAN_NO_PREFETCH is set the engine starts
 * computing the following batch (same limits) before it returns, so its kernels overlap this batch's copy to the host.  Re-entrant like the reference's overflow/resume
 * protocol (query_variants.cc:305,415,460-475): the caller drains a batch, then asks for the next.
 * The interval sweep, the per-site allele union and the allele verification run once per WINDOW of consecutive sites (as many as
 * GDBGPU_FRONT_BYTES of device memory hold at 24 bytes per (site, query row); default 6 GiB), the batches are cut from it: a batch
 * never spans two windows, so it may hold fewer sites than the limits allow; allele_first indexes the window's allele pool. */
int gdbgpu_scan_next_batch(gdbgpu_scan* scan, uint32_t max_sites, uint64_t max_bytes, gdbgpu_batch* batch);
/* Rewind to site idx (resume after a consumer-side overflow in the middle of a batch). */
int gdbgpu_scan_seek(gdbgpu_scan* scan, uint64_t site_idx);
void gdbgpu_scan_end(gdbgpu_scan* scan);

/* ---- the columnar iterator ------------------------------------------------------------------------------
 * Replaces SingleCellTileDBIterator + VariantQueryProcessor::iterate_over_cells (genomicsdb_iterators.cc:181-300,425-510;
 * query_variants.cc:557-576) for one query interval: the cells the reference's iterator hands to a
 * SingleCellOperatorBase, in its order -- first the calls that began before col_lo and reach into the interval (found
 * through their END copies, ordered by begin column then row), then the interval's begin cells of the queried rows in
 * storage order.  The list is built on the device; everything below works on any scan, GDBGPU_SCAN_CELLS_ONLY makes
 * gdbgpu_scan_begin skip what only the combine needs.  Host pointers stay valid until gdbgpu_scan_end. */
typedef struct gdbgpu_cell_list {
  uint64_t n_cells, n_left;   /* entries handed out (the window); the first n_left of them are calls reaching in from the left */
  uint64_t n_total, first;    /* length of the interval's whole list; position of the window's first entry in it */
  const uint64_t* cell;       /* index in the array's merged cell order (gdbgpu_array_copy_cells); a left call is named by its END copy */
  const int64_t* row;         /* GenomicsDBColumnarCell::get_coordinates()[0] */
  const int64_t* begin;       /* ... [1]: the call's begin column (coordinates and END swapped back for an END copy) */
  const int64_t* end;
} gdbgpu_cell_list;
int gdbgpu_scan_cells(gdbgpu_scan* scan, gdbgpu_cell_list* out);
/* Page through a long list: the following gdbgpu_scan_cells / gdbgpu_scan_cells_fetch calls hand out entries
 * [first, first + count) only (count 0 = to the end; the default window is the whole list).  Pointers handed out for an
 * earlier window are invalidated.  Histogram and allele counts always cover the whole list. */
int gdbgpu_scan_cells_window(gdbgpu_scan* scan, uint64_t first, uint64_t count);

/* ColumnHistogramOperator(begin, end, bin_size) over the listed cells (variant_operations.cc:732-765): counts[n_bins],
 * n_bins = (end - begin) / bin_size + 1, host memory of the caller.  Nothing but the counts leaves the device. */
int gdbgpu_scan_cells_histogram(gdbgpu_scan* scan, uint64_t begin, uint64_t end, uint64_t bin_size, uint64_t* counts, uint64_t n_bins);

/* AlleleCountOperator over the listed cells (variant_operations.cc:900-1089): one record per (begin column, normalised
 * REF, normalised ALT) with the number of GT elements naming that ALT, ordered as print_allele_counts prints them
 * (column, REF, ALT).  Strings: bytes + ref_pos / alt_pos[2k] offset, [2k + 1] length. */
typedef struct gdbgpu_allele_counts {
  uint64_t n;
  const int64_t* column; const uint64_t* count;
  const uint32_t* ref_pos; const uint32_t* alt_pos; const uint8_t* bytes;
} gdbgpu_allele_counts;
int gdbgpu_scan_cells_allele_counts(gdbgpu_scan* scan, gdbgpu_allele_counts* out);

/* The bytes of one column for every listed cell (what GenomicsDBColumnarCell::get_field_ptr_for_query_idx / get_field_length
 * give the print operators): offsets[n_cells + 1] into bytes.  kind REF / ALT / GT, or the index-th allele-dependent
 * field (gdbgpu_field_desc order) or plain field (gdbgpu_plain_field_desc order) the array was opened with. */
#define GDBGPU_CELL_FIELD_REF      0
#define GDBGPU_CELL_FIELD_ALT      1
#define GDBGPU_CELL_FIELD_GT       2
#define GDBGPU_CELL_FIELD_REMAPPED 3
#define GDBGPU_CELL_FIELD_PLAIN    4
int gdbgpu_scan_cells_fetch(gdbgpu_scan* scan, int32_t kind, int32_t index, const uint64_t** offsets, const uint8_t** bytes);

/* ---- The query stream of the Java reader (com.intel.genomicsdb.reader.GenomicsDBQueryStream).  One entry point per native
 * method of src/main/jni/src/genomicsdb_GenomicsDBQueryStream.cc:29-111, same arguments in the same order; the handle is
 * what the JNI layer keeps in its jlong (there a GenomicsDBBCFGenerator*, src/main/cpp/src/vcf/genomicsdb_bcf_generator.cc).
 * The stream is an uncompressed BCF2 file: header (generated from the configuration: template header + queried fields +
 * contigs of the vid mapping + samples of the queried rows), then one record per combined site of the interval
 * chr:start-end (1-based; chr empty: the query file's own column ranges).  loader_configuration_file may be NULL or empty
 * when the query file names workspace, array, vid and callset mapping itself; with it, the rank's partition supplies what
 * the query file leaves out (update_from_loader).  As in the reference the scanned intervals stay the query file's own
 * (subset_query_column_ranges_based_on_partition runs after they were handed to the query and edits the base list only);
 * a rank's share of a query is the query file's per-rank list of ranges. */
typedef struct gdbgpu_stream gdbgpu_stream;
/* jniGenomicsDBInit */
int gdbgpu_stream_open(const char* loader_configuration_file, const char* query_configuration_file, const char* chr, int start, int end, int rank,
                       uint64_t buffer_capacity, uint64_t segment_size, int is_bcf, int produce_header_only, int use_missing_values_only_not_vector_end,
                       int keep_idx_fields_in_bcf_header, gdbgpu_stream** out);
/* jniGenomicsDBClose */
int gdbgpu_stream_close(gdbgpu_stream* stream);
/* jniGenomicsDBGetNumBytesAvailable: the capacity of the stream's buffer (what the reference answers) */
uint64_t gdbgpu_stream_num_bytes_available(gdbgpu_stream* stream);
/* jniGenomicsDBReadNextByte: the next byte as a signed char, -1 at the end of the stream */
int gdbgpu_stream_read_next_byte(gdbgpu_stream* stream);
/* jniGenomicsDBRead: up to n bytes into dst + offset; returns the number of bytes copied (0 at the end), -1 on error */
int64_t gdbgpu_stream_read(gdbgpu_stream* stream, uint8_t* dst, uint64_t offset, uint64_t n);
/* jniGenomicsDBSkip */
int64_t gdbgpu_stream_skip(gdbgpu_stream* stream, uint64_t n);
/* message of the last failed gdbgpu_stream_* call of this thread */
const char* gdbgpu_stream_last_error(void);

/* library identity (for the test that checks every declared symbol is exported) */
const char* gdbgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
