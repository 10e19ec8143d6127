// Device-side data model and small device functions shared by the kernels of the variant-combine path.
// Semantics restated from the reference (paths relative to src/main/cpp in GenomicsDB v0.10.2):
//   allele classes        include/utils/gt_common.h:33-47, include/vcf/known_field_info.h:209-237, src/utils/known_field_info.cc:310-319
//   field lengths         src/utils/known_field_info.cc:130-162 ; nCr include/query_operations/variant_operations.h:317-343
//   genotype index        src/genomicsdb/variant_field_handler.cc:299-321
//   sentinels             include/vcf/vcf.h:61-218 ; TileDB tiledb_constants.h:105-109
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gdb {

constexpr uint32_t BCF_INT_MISSING = 0x80000000u, BCF_INT_VEND = 0x80000001u;
constexpr uint32_t BCF_FLT_MISSING = 0x7F800001u, BCF_FLT_VEND = 0x7F800002u;
constexpr uint32_t TDB_NULL_INT = 0x7FFFFFFFu, TDB_NULL_FLT = 0x7F7FFFFFu;
constexpr int64_t I64_MIN = INT64_MIN;

enum : uint8_t { F_ENDCOPY = 1, F_REFBLOCK = 2, F_DELETION = 4, F_GTVALID = 8, F_ROWOUT = 16, F_BADEVICT = 32, F_TRIVIAL = 64 };
enum { KIND_R = 0, KIND_A = 1, KIND_G = 2 };
constexpr int MAX_FIELDS = 15;          // remapped fields besides GT
constexpr int MAX_PLOIDY = 8;           // general-ploidy enumeration keeps the genotype in registers
constexpr int MAX_CALL_ALLELES = 1024;  // alleles (REF+ALT) of one input call / of a merged site
constexpr int LIVE_REDUCED = 0x40000000;

// 16 bytes per cell, written once when the array is staged (K2 classify)
struct CellDesc {
  uint8_t flags;     // F_*
  uint8_t ploidy;    // from GT length (0 when GT invalid)
  uint16_t n_alt;    // '|'-separated non-empty ALT tokens
  uint16_t nr_idx;   // input allele index of <NON_REF> ('&'), 0 = none
  uint16_t pad;
  uint32_t ref_len;
  uint32_t gt_len;   // words
};

// 32 bytes per cell: everything the union / remap kernels need from a *small* cell in one sector.  A reference block
// (GT 0/0, PL of 3, no AD) fits entirely, so 90 % of the calls live at a site cost one 32-byte gather.
constexpr int HOT_WORDS = 5;
constexpr int HOT_FIELDS = 4;
struct __align__(16) HotRec {
  uint8_t flags;              // F_*
  uint8_t ploidy;
  uint8_t n_alt;              // saturates at 255 (then `inl` is 0 and kernels read CellDesc)
  uint8_t nr_idx;
  uint8_t gt_len;             // words; 0 when GT is invalid
  uint8_t inl;                // 1: w[] holds GT then the valid fields' words, flen[] their lengths
  uint8_t flen[HOT_FIELDS];   // input length (words) of field f (saturates at 255), 0 = absent / every element NULL; set for every cell
  uint8_t ref_len8;           // REF length, saturates at 255 (then CellDesc has it)
  uint8_t ref0;               // first byte of REF (merge_reference_allele only asks whether it is 'N')
  uint32_t w[HOT_WORDS];
};
static_assert(sizeof(HotRec) == 32, "HotRec must be one 32-byte sector");

// Warm blob: every cell that is NOT (trivial and inline) gets one contiguous, 16-byte aligned record in a pool built
// when the array is staged, so that a non-trivial call costs one more gather (hot record -> blob) instead of a chain of
// offset/value loads over five columns plus a byte-wise re-hash of its ALT string.  HotRec.w[0] = blob offset in 16-byte
// units, HotRec.w[1] = blob length in words.  Words:
//   [0] n_alt  [1] nr_idx  [2] ref_len  [3] gt_len (0: GT invalid)   [4 .. 4+NF) valid input length of field f
//   (pad to an even index) then n_alt x u64 : FNV-1a state over the token's bytes (not finished, so a REF-tail suffix can
//   still be appended at a site whose merged REF is longer; 0 for the <NON_REF> token)
//   then n_alt x u32 : token length | symbolic << 31 | non_ref << 30
//   then the tokens' bytes, one after the other, padded to a word (k_union compares them with the merged allele's bytes:
//   equal hashes alone never decide that two alleles are the same)
//   then GT words (gt_len), then the valid words of every field in field order
constexpr uint32_t WARM_TOK_SYMBOLIC = 0x80000000u, WARM_TOK_NONREF = 0x40000000u;
__host__ __device__ inline uint32_t warm_hash_word(uint32_t n_fields) { return (4u + n_fields + 1u) & ~1u; }
// word [3] of a blob: gt_len | (words taken by the token bytes) << 16
__host__ __device__ inline uint32_t warm_gt_len(uint32_t w3) { return w3 & 0xFFFFu; }
__host__ __device__ inline uint32_t warm_tok_words(uint32_t w3) { return w3 >> 16; }

// Class word of a cell (4 bytes, built when the array is staged): the per-site scan (k_site_scan) gathers this word -- from
// an array small enough to live in L2 -- instead of the 32-byte hot record of every live call:
//   00 | packed  trivial inline call: nr(1) n_alt(1) gt_len(3) ploidy(3) flen[4](3 each) -- everything the site layout needs
//   01 | blob    any other call: warm-blob offset in 16-byte units (< 2^30)
//   10 | cell    (formed in k_site_scan from the live entry) the call is in its spanning-deletion form
//   0xFFFFFFFF   (formed in k_site_scan) sample not live at the site
constexpr uint32_t CLS_GENERAL = 0x40000000u, CLS_REDUCED = 0x80000000u, CLS_DEAD = 0xFFFFFFFFu, CLS_LOW = 0x3FFFFFFFu;
__host__ __device__ inline uint32_t cls_pack(uint32_t nr, uint32_t n_alt, uint32_t gt_len, uint32_t ploidy, const uint8_t* flen) {
  return (nr ? 1u : 0u) | ((n_alt & 1u) << 1) | ((gt_len & 7u) << 2) | ((ploidy & 7u) << 5) |
         ((uint32_t)(flen[0] & 7u) << 8) | ((uint32_t)(flen[1] & 7u) << 11) | ((uint32_t)(flen[2] & 7u) << 14) | ((uint32_t)(flen[3] & 7u) << 17);
}

// one entry of a sample's cell list: everything the live-matrix walk needs from a cell in one sector
struct __align__(16) RsRec {
  int64_t col;        // column of the cell
  int64_t effend;     // effective end of the call that begins there; I64_MIN for an END copy
  uint32_t cell;      // cell index | 0x80000000 when the call contains a deletion
  uint32_t cls;       // class word of the cell (see above)
  uint32_t pad[2];
};
static_assert(sizeof(RsRec) == 32, "RsRec must be one 32-byte sector");

struct VarCol {
  const uint64_t* off;   // n_cells + 1 absolute byte offsets
  const uint8_t* val;
};

struct FieldMeta { int32_t kind, elem, is_pl; };

// the staged array as kernels see it
struct ArrayDev {
  uint64_t n;
  const int64_t* coords;   // (row, col) pairs
  const int64_t* end;
  VarCol ref, alt, gt;
  int32_t has_gt, gt_phased, n_fields;
  VarCol f[MAX_FIELDS];
  FieldMeta fm[MAX_FIELDS];
  int64_t lb_row, num_rows;
  const CellDesc* desc;
  const HotRec* hot;
  const uint32_t* warm;    // warm blob pool (see above)
  const uint32_t* cls;     // class word per cell (see above)
  const int64_t* effend;   // cell order; I64_MIN for END copies / rows outside the array
  // row-major index (per-sample cell lists, column ascending)
  uint64_t n_rs;
  const struct RsRec* rs_rec;  // the same index as one 32-byte record per entry (k_live_walk reads nothing else)
  const uint32_t* rs_cell;
  const int64_t* rs_col;
  const int64_t* rs_effend;
  const uint64_t* row_start;   // num_rows + 1
};

// spanning-deletion ("reduced") form of a deletion cell: REF N, ALT * [,<NON_REF>] (broad_combined_gvcf.cc:912-1078)
struct DelRec {
  uint32_t cell;
  uint8_t has_nr, n_alt, pad0, pad1;
  uint32_t gt_off, gt_len;
  uint32_t f_off[MAX_FIELDS];
  uint32_t f_len[MAX_FIELDS];   // 0xFFFFFFFF : field invalid
};

__host__ __device__ inline bool is_non_ref(const uint8_t* s, int n) { return n > 0 && s[0] == '&'; }
__host__ __device__ inline bool is_symbolic(const uint8_t* s, int n) {
  if (n <= 0) return false;
  if (s[0] == '&') return true;
  if (n == 1 && s[0] == '*') return true;
  if (s[0] == '<' && s[n - 1] == '>') return true;
  for (int i = 0; i < n; ++i) if (s[i] == '[' || s[i] == ']') return true;
  return false;
}
// strtok_r(.., "|") over strndup'ed bytes: NUL ends the string, empty tokens vanish (variant_field_data.h:511-537)
__host__ __device__ inline int alt_effective_len(const uint8_t* s, int n) {
  for (int i = 0; i < n; ++i) if (s[i] == 0) return i;
  return n;
}
__host__ __device__ inline bool next_token(const uint8_t* s, int n, int& pos, int& ts, int& tl) {
  while (pos < n && s[pos] == '|') ++pos;
  if (pos >= n) return false;
  ts = pos;
  while (pos < n && s[pos] != '|') ++pos;
  tl = pos - ts;
  return true;
}

__host__ __device__ inline uint64_t nCr_u64(int64_t n, int64_t r) {
  if (n < 0 || r < 0 || n < r) return 0;
  uint64_t begin_idx, end_idx;
  if (r > n - r) { begin_idx = (uint64_t)r + 1u; end_idx = (uint64_t)(n - r); }
  else { begin_idx = (uint64_t)(n - r) + 1u; end_idx = (uint64_t)r; }
  uint64_t num = 1, den = 1;
  for (uint64_t i = begin_idx; i <= (uint64_t)n; ++i) num *= i;
  for (uint64_t i = 1; i <= end_idx; ++i) den *= i;
  return num / den;
}
__host__ __device__ inline int64_t alleles2gt(int64_t a, int64_t b) { return a > b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }
__host__ __device__ inline uint32_t num_genotypes(uint32_t n_alt, uint32_t ploidy) {
  if (ploidy == 1) return n_alt + 1u;
  if (ploidy == 2) return ((n_alt + 1u) * (n_alt + 2u)) >> 1;
  return (uint32_t)nCr_u64((int64_t)ploidy + n_alt, n_alt);
}
__host__ __device__ inline uint32_t num_elements(int kind, uint32_t n_alt, uint32_t ploidy) {
  if (kind == KIND_A) return n_alt;
  if (kind == KIND_R) return n_alt + 1u;
  return num_genotypes(n_alt, ploidy);
}
__host__ __device__ inline uint32_t ploidy_of(int phased, uint32_t gt_len) { return phased ? ((gt_len + 1u) >> 1) : gt_len; }
// sorted (ascending) allele tuple -> genotype index, ploidy > 2
__host__ __device__ inline uint64_t genotype_index_sorted(const int* a, int ploidy) {
  uint64_t g = 0;
  for (int i = 0; i < ploidy; ++i) g += nCr_u64((int64_t)i + a[i], (int64_t)a[i] - 1);
  return g;
}
__host__ __device__ inline void sort_small(int* v, int n) {
  for (int i = 1; i < n; ++i) { int x = v[i], j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}
// next genotype in VCF order (highest ploidy slot is the slowest digit, slots non-decreasing); false when exhausted
__host__ __device__ inline bool next_genotype(int* vec, int ploidy, int n_alleles) {
  for (int i = 0; i < ploidy; ++i) {
    int lim = (i + 1 < ploidy) ? vec[i + 1] : n_alleles - 1;
    if (vec[i] < lim) { ++vec[i]; for (int k = 0; k < i; ++k) vec[k] = 0; return true; }
  }
  return false;
}

__host__ __device__ inline uint64_t fnv1a(uint64_t h, const uint8_t* s, int n) {
  for (int i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  return h;
}
__host__ __device__ inline uint64_t hash_finish(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h | 1ull;   // never 0 (0 marks an empty slot)
}
constexpr uint64_t FNV_BASIS = 1469598103934665603ull;

}  // namespace gdb
