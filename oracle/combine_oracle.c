/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the GenomicsDB variant-combine hot path.
 *
 * Plain-C, single-threaded, sequential restatement of the reference algorithm; it deliberately follows the
 * reference's control flow (cell-at-a-time sweep with a min-heap on END, first-seen allele union, per-call remap)
 * and NOT the data-parallel reformulation the CUDA path uses, so that the two are independent.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's library; the product
 * (genomicsdb_b200/) never does.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this file byte-for-byte against canonical dumps produced by the
 * unmodified reference (oracle/_ref/bin/gdb_ref_dump) on every non-Java scenario of the reference's own CI
 * (tests/golden/manifest.json), whose VCF output was md5-matched to the reference's golden_outputs/ files.
 *
 * Reference citations (relative to /root/reference/src/main/cpp unless noted):
 *   sweep            src/genomicsdb/query_variants.cc:296-332 (handle_gvcf_ranges), :334-476 (scan_and_operate),
 *                    :478-555 (scan_handle_cell), :845-941 (gt_get_column), :1014-1117 (gt_fill_row)
 *   field validity   include/genomicsdb/variant_field_data.h:365-384 ; ALT tokeniser :511-537
 *   REF merge        src/query_operations/variant_operations.cc:73-122
 *   ALT union + LUT  src/query_operations/variant_operations.cc:134-228 ; include/utils/lut.h:65-272
 *   operate          src/query_operations/variant_operations.cc:362-378 (base), :572-695 (GA4GHOperator)
 *   GT remap         src/query_operations/variant_operations.cc:233-263
 *   R/A remap        src/genomicsdb/variant_field_handler.cc:41-81
 *   G remap          src/genomicsdb/variant_field_handler.cc:83-191 (haploid/diploid), :198-297 (general),
 *                    :299-321 (get_genotype_index), include/query_operations/variant_operations.h:317-343 (nCr)
 *   lengths          src/utils/known_field_info.cc:130-162
 *   deletions        src/query_operations/broad_combined_gvcf.cc:912-1078 (handle_deletions), :1080-1118 (min-PL GT),
 *                    src/genomicsdb/variant_field_handler.cc:402-494
 *   sentinels        include/vcf/vcf.h:61-218 ; dependencies/TileDB/core/include/c_api/tiledb_constants.h:105-109
 *
 * Output: the canonical hot-path dump documented in oracle/ref_harness/gdb_ref_dump.cc.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define BCF_INT_MISSING   ((uint32_t)0x80000000u)
#define BCF_INT_VEND      ((uint32_t)0x80000001u)
#define BCF_FLT_MISSING   ((uint32_t)0x7F800001u)
#define BCF_FLT_VEND      ((uint32_t)0x7F800002u)
#define TDB_NULL_INT      ((uint32_t)0x7FFFFFFFu)  /* INT_MAX */
#define TDB_NULL_FLT      ((uint32_t)0x7F7FFFFFu)  /* FLT_MAX, compared bitwise (vcf.h:206-211) */

enum { KIND_R = 0, KIND_A = 1, KIND_G = 2 };
enum { ELEM_INT = 0, ELEM_FLOAT = 1 };

/* ---------------------------------------------------------------- input description (SoA, global cell order) */
typedef struct {
  const int64_t* off;   /* n_cells+1, element offsets */
  const uint32_t* val;  /* 4-byte words */
  int32_t kind;         /* KIND_R / KIND_A / KIND_G */
  int32_t elem;         /* ELEM_INT / ELEM_FLOAT */
  int32_t is_PL;        /* this field is the known field "PL" (handle_deletions uses it) */
} oracle_field_t;

typedef struct {
  int64_t n_cells;
  const int64_t* row; const int64_t* col; const int64_t* end;
  const int64_t* ref_off; const uint8_t* ref;
  const int64_t* alt_off; const uint8_t* alt;
  const int64_t* gt_off;  const int32_t* gt;      /* gt_off == NULL : GT not queried */
  int32_t gt_phased;                               /* vid length "PP": [a0, phase, a1, ...] */
  int32_t n_fields; const oracle_field_t* fields;
  /* query */
  int32_t scan_full; int64_t col_lo, col_hi;
  int64_t n_query_rows; const int64_t* query_rows; /* sorted ascending */
  int32_t max_alt_alleles_for_G;                   /* default 50 */
  int32_t handle_spanning_deletions;               /* Broad operator: 1 */
  int32_t produce_GT_with_min_PL;                  /* query flag */
} oracle_input_t;

/* ---------------------------------------------------------------- small growable byte buffer */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static void buf_put(buf_t* b, const void* src, size_t n) {
  if (b->n + n > b->cap) { b->cap = (b->cap + n) * 2 + 64; b->p = (uint8_t*)realloc(b->p, b->cap); }
  memcpy(b->p + b->n, src, n); b->n += n;
}
static void put_u8(buf_t* b, uint8_t v) { buf_put(b, &v, 1); }
static void put_u16(buf_t* b, uint16_t v) { buf_put(b, &v, 2); }
static void put_u32(buf_t* b, uint32_t v) { buf_put(b, &v, 4); }
static void put_i32(buf_t* b, int32_t v) { buf_put(b, &v, 4); }
static void put_i64(buf_t* b, int64_t v) { buf_put(b, &v, 8); }
static void put_str(buf_t* b, const char* s, uint32_t n) { put_u32(b, n); buf_put(b, s, n); }

/* ---------------------------------------------------------------- a live call (one per query row, reused) */
typedef struct { int valid; int32_t n; uint32_t* v; int32_t cap; } vec_t;
typedef struct { char* s; int32_t n; } str_t;
typedef struct {
  int initialized, valid;
  int64_t row, begin, end;
  int contains_deletion, is_reference_block;
  str_t ref;
  int32_t n_alt; str_t* alt; int32_t alt_cap;
  vec_t gt;
  vec_t* f;       /* n_fields */
  int64_t heap_pos;
} call_t;

static void vec_assign(vec_t* d, const uint32_t* src, int32_t n) {
  if (n > d->cap) { d->cap = n * 2 + 4; d->v = (uint32_t*)realloc(d->v, sizeof(uint32_t) * d->cap); }
  if (n) memcpy(d->v, src, sizeof(uint32_t) * n);
  d->n = n;
}
static void vec_resize(vec_t* d, int32_t n) {   /* std::vector::resize: zero-fill growth */
  if (n > d->cap) { d->cap = n * 2 + 4; d->v = (uint32_t*)realloc(d->v, sizeof(uint32_t) * d->cap); }
  for (int32_t i = d->n; i < n; ++i) d->v[i] = 0;
  d->n = n;
}
static void str_assign(str_t* d, const char* s, int32_t n) {
  d->s = (char*)realloc(d->s, n + 1); memcpy(d->s, s, n); d->s[n] = 0; d->n = n;
}
static int str_eq(const str_t* a, const char* s, int32_t n) { return a->n == n && memcmp(a->s, s, n) == 0; }

/* gt_common.h:36-43, known_field_info.h:209-237 */
static int is_non_ref(const char* s, int32_t n) { return n > 0 && s[0] == '&'; }
static int is_symbolic(const char* s, int32_t n) {
  if (is_non_ref(s, n)) return 1;
  if (n == 1 && s[0] == '*') return 1;
  if (n > 0) {
    if (s[0] == '<' && s[n - 1] == '>') return 1;
    for (int32_t i = 0; i < n; ++i) if (s[i] == '[' || s[i] == ']') return 1;
  }
  return 0;
}
static int is_deletion(const str_t* ref, const str_t* alt) {
  return ref->n > 1 && ((alt->n == 1 && alt->s[0] == '*') || (!is_symbolic(alt->s, alt->n) && alt->n < ref->n));
}
/* known_field_info.cc:310-319 */
static int contains_deletion(const call_t* c) {
  if (c->ref.n <= 1) return 0;
  for (int32_t i = 0; i < c->n_alt; ++i)
    if (!is_symbolic(c->alt[i].s, c->alt[i].n) && c->alt[i].n < c->ref.n) return 1;
  return 0;
}

/* variant_field_data.h:365-384 : a vector field whose every element is TileDB-NULL (or that is empty) is invalid */
static int words_all_null(const uint32_t* v, int64_t n, uint32_t nullw) {
  for (int64_t i = 0; i < n; ++i) if (v[i] != nullw) return 0;
  return 1;
}

/* ---------------------------------------------------------------- combinatorics */
/* variant_operations.h:317-343, incl. its wrap-around behaviour in uint64 */
static uint64_t nCr(int64_t n, int64_t r) {
  if (n < 0 || r < 0 || n < r) return 0;
  uint64_t begin_idx, end_idx;
  if (r > n - r) { begin_idx = (uint64_t)r + 1u; end_idx = (uint64_t)(n - r); }
  else { begin_idx = (uint64_t)(n - r) + 1u; end_idx = (uint64_t)r; }
  uint64_t num = 1, den = 1;
  for (uint64_t i = begin_idx; i <= (uint64_t)n; ++i) num *= i;
  for (uint64_t i = 1; i <= end_idx; ++i) den *= i;
  return num / den;
}
static int64_t alleles2gt(int64_t a, int64_t b) { return a > b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }
/* known_field_info.cc:130-144 */
static uint32_t num_genotypes(uint32_t n_alt, uint32_t ploidy) {
  if (ploidy == 1) return n_alt + 1u;
  if (ploidy == 2) return ((n_alt + 1u) * (n_alt + 2u)) >> 1;
  return (uint32_t)nCr((int64_t)ploidy + n_alt, n_alt);
}
/* known_field_info.cc:146-162 */
static uint32_t num_elements(int kind, uint32_t n_alt, uint32_t ploidy) {
  if (kind == KIND_A) return n_alt;
  if (kind == KIND_R) return n_alt + 1u;
  return num_genotypes(n_alt, ploidy);
}
static void sort_ints(int* v, int n) { for (int i = 1; i < n; ++i) { int x = v[i], j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; } }
/* variant_field_handler.cc:299-321 */
static uint64_t genotype_index(int* alleles, int ploidy, int is_sorted) {
  if (ploidy == 0) return 0;
  if (ploidy == 1) return (uint64_t)(int64_t)alleles[0];
  if (ploidy == 2) return (uint64_t)alleles2gt(alleles[0], alleles[1]);
  if (!is_sorted) sort_ints(alleles, ploidy);
  uint64_t g = 0;
  for (int i = 0; i < ploidy; ++i) g += nCr((int64_t)i + alleles[i], (int64_t)alleles[i] - 1);
  return g;
}
/* known_field_info.h:142-156 */
static uint32_t get_ploidy(int phased, uint32_t n) { return phased ? ((n + 1u) >> 1) : n; }

/* ---------------------------------------------------------------- allele LUT for ONE call (lut.h: both directions) */
typedef struct { int64_t* in2m; int64_t* m2in; int32_t width; } lut_t;
static void lut_reserve(lut_t* l, int32_t w) {
  if (w <= l->width) return;
  l->in2m = (int64_t*)realloc(l->in2m, sizeof(int64_t) * w);
  l->m2in = (int64_t*)realloc(l->m2in, sizeof(int64_t) * w);
  for (int32_t i = l->width; i < w; ++i) l->in2m[i] = l->m2in[i] = -1;
  l->width = w;
}
static void lut_reset(lut_t* l) { for (int32_t i = 0; i < l->width; ++i) l->in2m[i] = l->m2in[i] = -1; }
static void lut_add(lut_t* l, int64_t in, int64_t merged) {
  lut_reserve(l, (int32_t)((in > merged ? in : merged) + 1));
  l->in2m[in] = merged; l->m2in[merged] = in;      /* lut.h add_input_merged_idx_pair: later pairs overwrite */
}
static int64_t lut_m2in(const lut_t* l, int64_t m) { return (m >= 0 && m < l->width) ? l->m2in[m] : -1; }
static int64_t lut_in2m(const lut_t* l, int64_t i) { return (i >= 0 && i < l->width) ? l->in2m[i] : -1; }

/* ---------------------------------------------------------------- remap kernels (operate on one call) */
/* variant_field_handler.cc:41-81 */
static void remap_by_alleles(const uint32_t* in, int32_t n_in, const lut_t* lut, uint32_t A, int nonref, int alt_only,
                             uint32_t* out, uint32_t missing) {
  int64_t in_nr = nonref ? lut_m2in(lut, (int64_t)A - 1) : -1;
  uint32_t length = alt_only ? A - 1u : A;
  for (uint32_t j = 0; j < length; ++j) {
    uint32_t allele_j = alt_only ? j + 1u : j;
    int64_t ij = lut_m2in(lut, allele_j);
    if (ij < 0) { if (in_nr < 0) { out[j] = missing; continue; } ij = in_nr; }
    int64_t input_j = alt_only ? ij - 1 : ij;
    out[j] = (input_j < 0 || input_j >= n_in) ? missing : in[input_j];
  }
}
/* variant_field_handler.cc:83-122 */
static void remap_G_haploid(const uint32_t* in, int32_t n_in, const lut_t* lut, uint32_t A, int nonref, uint32_t* out, uint32_t missing) {
  int64_t in_nr = nonref ? lut_m2in(lut, (int64_t)A - 1) : -1;
  for (uint32_t j = 0; j < A; ++j) {
    int64_t ij = lut_m2in(lut, j);
    if (ij < 0) { if (in_nr < 0) { out[j] = missing; continue; } ij = in_nr; }
    out[j] = (ij >= n_in) ? missing : in[ij];
  }
}
/* variant_field_handler.cc:134-191 */
static void remap_G_diploid(const uint32_t* in, int32_t n_in, const lut_t* lut, uint32_t A, int nonref, uint32_t* out, uint32_t missing) {
  int64_t in_nr = nonref ? lut_m2in(lut, (int64_t)A - 1) : -1;
  for (uint32_t j = 0; j < A; ++j) {
    int64_t ij = lut_m2in(lut, j);
    if (ij < 0) {
      if (in_nr < 0) { for (uint32_t k = j; k < A; ++k) out[alleles2gt(j, k)] = missing; continue; }
      ij = in_nr;
    }
    for (uint32_t k = j; k < A; ++k) {
      int64_t g = alleles2gt(j, k);
      int64_t ik = lut_m2in(lut, k);
      if (ik < 0) { if (in_nr < 0) { out[g] = missing; continue; } ik = in_nr; }
      int64_t ig = alleles2gt(ij, ik);
      out[g] = (ig >= n_in) ? missing : in[ig];
    }
  }
}
/* variant_field_handler.cc:198-297 restated as the recursion its explicit stack implements: outermost loop over the
 * highest ploidy slot, vec[p-1] <= vec[p]; genotypes are emitted in VCF order */
typedef struct {
  const uint32_t* in; int32_t n_in; const lut_t* lut; uint32_t A; int64_t in_nr; int ploidy; uint32_t* out; uint32_t missing;
  uint64_t next_gt; int* merged_vec; int* input_vec;
  /* min-PL tracker mode (variant_field_handler.cc:402-434) */
  int track; int32_t cur_min; int found; uint64_t min_gt; int* min_vec;
} gen_t;
static void track_minimum(gen_t* g, int* vec) {
  uint64_t gi = genotype_index(vec, g->ploidy, 0);
  if (gi < (uint64_t)g->n_in) {
    uint32_t w = g->in[gi];
    if (w != BCF_INT_MISSING && w != BCF_INT_VEND && (int32_t)w < g->cur_min) {
      g->cur_min = (int32_t)w; g->min_gt = gi; g->found = 1;
      memcpy(g->min_vec, vec, sizeof(int) * g->ploidy);
    }
  }
}
static void gen_emit(gen_t* g) {
  int missing_allele = 0;
  for (int i = 0; i < g->ploidy; ++i) {
    int64_t ia = lut_m2in(g->lut, g->merged_vec[i]);
    if (ia < 0) { g->input_vec[i] = (int)g->in_nr; if (g->in_nr < 0) missing_allele = 1; }
    else g->input_vec[i] = (int)ia;
  }
  if (g->track) track_minimum(g, g->input_vec);
  else if (missing_allele) g->out[g->next_gt] = g->missing;
  else {
    uint64_t ig = genotype_index(g->input_vec, g->ploidy, 0);
    g->out[g->next_gt] = (ig >= (uint64_t)g->n_in) ? g->missing : g->in[ig];
  }
  ++g->next_gt;
}
static void gen_rec(gen_t* g, int level, int amax) {
  for (int a = 0; a <= amax; ++a) {
    g->merged_vec[level] = a;
    if (level == 0) gen_emit(g); else gen_rec(g, level - 1, a);
  }
}
static void remap_G_general(const uint32_t* in, int32_t n_in, const lut_t* lut, uint32_t A, int nonref, int ploidy, uint32_t* out, uint32_t missing) {
  if (ploidy == 0) return;
  gen_t g; memset(&g, 0, sizeof(g));
  g.in = in; g.n_in = n_in; g.lut = lut; g.A = A; g.in_nr = nonref ? lut_m2in(lut, (int64_t)A - 1) : -1;
  g.ploidy = ploidy; g.out = out; g.missing = missing;
  g.merged_vec = (int*)calloc(ploidy + 1, sizeof(int)); g.input_vec = (int*)calloc(ploidy + 1, sizeof(int));
  gen_rec(&g, ploidy - 1, (int)A - 1);
  free(g.merged_vec); free(g.input_vec);
}
/* variant_field_handler.cc:358-398 + :497-527 (remap_vector_data dispatch) */
static void remap_field(int kind, const uint32_t* in, int32_t n_in, const lut_t* lut, uint32_t A, int nonref, uint32_t ploidy,
                        uint32_t* out, uint32_t missing) {
  if (kind == KIND_G) {
    if (ploidy == 1) remap_G_haploid(in, n_in, lut, A, nonref, out, missing);
    else if (ploidy == 2) remap_G_diploid(in, n_in, lut, A, nonref, out, missing);
    else remap_G_general(in, n_in, lut, A, nonref, (int)ploidy, out, missing);
  } else remap_by_alleles(in, n_in, lut, A, nonref, kind == KIND_A, out, missing);
}
/* variant_operations.cc:233-263 */
static void remap_GT(const uint32_t* in, int32_t n, const lut_t* lut, uint32_t A, int nonref, int phased, uint32_t* out) {
  int step = phased ? 2 : 1;
  for (int32_t i = 0; i < n; i += step) {
    uint32_t w = in[i];
    if (w == TDB_NULL_INT || w == 0xFFFFFFFFu || w == BCF_INT_MISSING) out[i] = w;
    else {
      int64_t m = lut_in2m(lut, (int64_t)(int32_t)w);
      out[i] = (m < 0) ? (nonref ? A - 1u : 0xFFFFFFFFu) : (uint32_t)m;
    }
    if (phased && i + 1 < n) out[i + 1] = in[i + 1];
  }
}

/* ---------------------------------------------------------------- operator state */
typedef struct {
  const oracle_input_t* in;
  int64_t nq; call_t* calls; lut_t* luts;
  /* end-PQ (variant.h:707-711): binary min-heap of query-row indices keyed by call END */
  int64_t* heap; int64_t heap_n;
  uint64_t ndel;
  /* merged alleles */
  str_t mref; str_t* malt; int32_t n_malt, malt_cap; int nonref_exists;
  int64_t site_begin, site_end;
  buf_t* out; int64_t n_sites; int error;
  uint32_t* scratch; int32_t scratch_cap;
  int64_t* in_nr;
} op_t;

static uint32_t missing_word(const oracle_field_t* f) { return f->elem == ELEM_FLOAT ? BCF_FLT_MISSING : BCF_INT_MISSING; }

static void heap_swap(op_t* o, int64_t a, int64_t b) {
  int64_t t = o->heap[a]; o->heap[a] = o->heap[b]; o->heap[b] = t;
  o->calls[o->heap[a]].heap_pos = a; o->calls[o->heap[b]].heap_pos = b;
}
static void heap_up(op_t* o, int64_t i) { while (i > 0) { int64_t p = (i - 1) / 2; if (o->calls[o->heap[p]].end <= o->calls[o->heap[i]].end) break; heap_swap(o, p, i); i = p; } }
static void heap_down(op_t* o, int64_t i) {
  for (;;) { int64_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < o->heap_n && o->calls[o->heap[l]].end < o->calls[o->heap[m]].end) m = l;
    if (r < o->heap_n && o->calls[o->heap[r]].end < o->calls[o->heap[m]].end) m = r;
    if (m == i) break;
    heap_swap(o, m, i); i = m; }
}
static void heap_push(op_t* o, int64_t q) { o->heap[o->heap_n] = q; o->calls[q].heap_pos = o->heap_n; heap_up(o, o->heap_n++); }
static void heap_remove(op_t* o, int64_t pos) {
  --o->heap_n;
  if (pos != o->heap_n) { heap_swap(o, pos, o->heap_n); heap_down(o, pos); heap_up(o, pos); }
}

/* ---------------------------------------------------------------- handle_deletions (broad_combined_gvcf.cc:912-1078) */
static void handle_deletions(op_t* o) {
  const oracle_input_t* in = o->in;
  for (int64_t q = 0; q < o->nq; ++q) {
    call_t* c = &o->calls[q];
    if (!c->valid) continue;
    if (!(c->contains_deletion && o->site_begin > c->begin)) continue;
    if (c->n_alt > 0 && str_eq(&c->alt[0], "*", 1) && (c->n_alt == 1 || (c->n_alt == 2 && is_non_ref(c->alt[1].s, c->alt[1].n)))) continue;
    lut_t red; memset(&red, 0, sizeof(red)); lut_reserve(&red, c->n_alt + 1 > 3 ? c->n_alt + 1 : 3);
    lut_add(&red, 0, 0);
    uint32_t ploidy = 0; int gt_ok = (in->gt_off && c->gt.valid);
    if (gt_ok) ploidy = get_ploidy(in->gt_phased, (uint32_t)c->gt.n);
    int lowest = -1; int32_t lowest_pl = INT_MAX; int has_nr = 0;
    const vec_t* pl = NULL;
    for (int32_t k = 0; k < in->n_fields; ++k) if (in->fields[k].is_PL && c->f[k].valid) pl = &c->f[k];
    int* geno = (int*)calloc(ploidy + 1, sizeof(int));
    for (int32_t i = 0; i < c->n_alt; ++i) {
      int allele_idx = i + 1;
      if (is_deletion(&c->ref, &c->alt[i])) {
        if (lowest < 0) lowest = allele_idx;
        if (pl) {
          for (uint32_t p = 0; p < ploidy; ++p) geno[p] = allele_idx;
          uint64_t gi = genotype_index(geno, (int)ploidy, 1);
          if (gi < (uint64_t)pl->n && (int32_t)pl->v[gi] < lowest_pl) { lowest_pl = (int32_t)pl->v[gi]; lowest = allele_idx; }
        }
      } else if (is_non_ref(c->alt[i].s, c->alt[i].n)) { lut_add(&red, allele_idx, 2); has_nr = 1; }
    }
    free(geno);
    if (lowest < 0) { o->error = 2; free(red.in2m); free(red.m2in); continue; }
    lut_add(&red, lowest, 1);
    c->n_alt = has_nr ? 2 : 1;
    if (has_nr) str_assign(&c->alt[1], "&", 1);
    str_assign(&c->ref, "N", 1);
    str_assign(&c->alt[0], "*", 1);
    uint32_t n_red = (uint32_t)c->n_alt + 1u;
    for (int32_t k = 0; k < in->n_fields; ++k) {
      vec_t* f = &c->f[k];
      if (!f->valid) continue;
      uint32_t n_out = num_elements(in->fields[k].kind, n_red - 1u, ploidy);
      uint32_t* copy = (uint32_t*)malloc(sizeof(uint32_t) * (f->n + 1)); memcpy(copy, f->v, sizeof(uint32_t) * f->n);
      int32_t n_in = f->n;
      vec_resize(f, (int32_t)n_out);
      remap_field(in->fields[k].kind, copy, n_in, &red, n_red, has_nr, ploidy, f->v, missing_word(&in->fields[k]));
      free(copy);
    }
    if (gt_ok) {
      int remap_from_input = 1;
      /* update_GT_to_correspond_to_min_PL_value (broad_combined_gvcf.cc:1080-1118): uses the ALREADY reduced PL */
      const vec_t* pl2 = NULL;
      for (int32_t k = 0; k < in->n_fields; ++k) if (in->fields[k].is_PL && c->f[k].valid) pl2 = &c->f[k];
      if (pl2 && in->produce_GT_with_min_PL) {
        uint32_t pl_ploidy = get_ploidy(in->gt_phased, (uint32_t)c->gt.n);
        gen_t g; memset(&g, 0, sizeof(g));
        g.in = pl2->v; g.n_in = pl2->n; g.ploidy = (int)pl_ploidy; g.track = 1; g.cur_min = INT_MAX;
        g.min_vec = (int*)calloc(pl_ploidy + 1, sizeof(int));
        int* vec = (int*)calloc(pl_ploidy + 1, sizeof(int));
        uint32_t ngt = num_genotypes(n_red - 1u, pl_ploidy);
        if (pl_ploidy == 1) {
          uint32_t lim = ngt < (uint32_t)pl2->n ? ngt : (uint32_t)pl2->n;
          for (uint32_t i = 0; i < lim; ++i) { vec[0] = (int)i; track_minimum(&g, vec); }
        } else if (pl_ploidy == 2) {
          for (uint32_t i = 0; i < n_red; ++i) for (uint32_t j = i; j < n_red; ++j) { vec[0] = (int)i; vec[1] = (int)j; track_minimum(&g, vec); }
        } else if (pl_ploidy > 2) {
          lut_t idl; memset(&idl, 0, sizeof(idl)); lut_reserve(&idl, (int32_t)n_red);
          for (uint32_t i = 0; i < n_red; ++i) lut_add(&idl, i, i);
          g.lut = &idl; g.A = n_red; g.in_nr = has_nr ? (int64_t)n_red - 1 : -1;
          g.merged_vec = (int*)calloc(pl_ploidy + 1, sizeof(int)); g.input_vec = (int*)calloc(pl_ploidy + 1, sizeof(int));
          gen_rec(&g, (int)pl_ploidy - 1, (int)n_red - 1);
          free(g.merged_vec); free(g.input_vec); free(idl.in2m); free(idl.m2in);
        }
        if (g.found) {
          remap_from_input = 0;
          int step = in->gt_phased ? 2 : 1;
          for (int32_t i = 0, j = 0; i < c->gt.n; i += step, ++j) c->gt.v[i] = (uint32_t)g.min_vec[j];
        }
        free(vec); free(g.min_vec);
      }
      if (remap_from_input) {
        uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (c->gt.n + 1));
        remap_GT(c->gt.v, c->gt.n, &red, n_red, has_nr, in->gt_phased, tmp);
        memcpy(c->gt.v, tmp, sizeof(uint32_t) * c->gt.n);
        free(tmp);
      }
    }
    free(red.in2m); free(red.m2in);
  }
}

/* ---------------------------------------------------------------- merge REF / ALT (variant_operations.cc:73-228) */
static void malt_push(op_t* o, const char* s, int32_t n, const char* suffix, int32_t ns) {
  if (o->n_malt == o->malt_cap) { int32_t nc = o->malt_cap * 2 + 8; o->malt = (str_t*)realloc(o->malt, sizeof(str_t) * nc);
    memset(o->malt + o->malt_cap, 0, sizeof(str_t) * (nc - o->malt_cap)); o->malt_cap = nc; }
  str_t* d = &o->malt[o->n_malt++];
  d->s = (char*)realloc(d->s, n + ns + 1); memcpy(d->s, s, n); if (ns) memcpy(d->s + n, suffix, ns); d->s[n + ns] = 0; d->n = n + ns;
}
static void merge_alleles(op_t* o) {
  /* merge_reference_allele */
  str_assign(&o->mref, "N", 1);
  int32_t ml = 1;   /* merged_ref_length: a register of its own -- it is NOT refreshed when a shorter real REF replaces an N-leading one
                       (variant_operations.cc:108-121), so later REFs are compared with, and appended from, the stale length */
  for (int64_t q = 0; q < o->nq; ++q) {
    call_t* c = &o->calls[q];
    if (!c->valid || c->begin < o->site_begin) continue;
    if (c->ref.n > ml) {
      if (ml > 0 && o->mref.s[0] == 'N') str_assign(&o->mref, c->ref.s, c->ref.n);
      else { int32_t have = o->mref.n, add = c->ref.n - ml; o->mref.s = (char*)realloc(o->mref.s, have + add + 1); memcpy(o->mref.s + have, c->ref.s + ml, add); o->mref.n = have + add; o->mref.s[have + add] = 0; }
      ml = c->ref.n;
    } else if (o->mref.s[0] == 'N' && !(c->ref.n > 0 && c->ref.s[0] == 'N')) str_assign(&o->mref, c->ref.s, c->ref.n);
  }
  /* merge_alt_alleles: first-seen order over valid calls in row order, NON_REF last */
  o->n_malt = 0; o->nonref_exists = 0;
  for (int64_t q = 0; q < o->nq; ++q) { o->in_nr[q] = -1; if (o->calls[q].valid) lut_reset(&o->luts[q]); }
  for (int64_t q = 0; q < o->nq; ++q) {
    call_t* c = &o->calls[q];
    if (!c->valid) continue;
    int32_t ns = (c->ref.n < o->mref.n) ? o->mref.n - c->ref.n : 0;
    const char* suffix = o->mref.s + c->ref.n;
    lut_add(&o->luts[q], 0, 0);
    for (int32_t i = 0; i < c->n_alt; ++i) {
      const str_t* a = &c->alt[i];
      int input_idx = i + 1;
      if (is_non_ref(a->s, a->n)) { o->in_nr[q] = input_idx; o->nonref_exists = 1; continue; }
      int32_t use_ns = (ns > 0 && !is_symbolic(a->s, a->n)) ? ns : 0;
      int32_t found = -1;
      for (int32_t m = 0; m < o->n_malt && found < 0; ++m) {
        const str_t* x = &o->malt[m];
        if (x->n == a->n + use_ns && memcmp(x->s, a->s, a->n) == 0 && (use_ns == 0 || memcmp(x->s + a->n, suffix, use_ns) == 0)) found = m;
      }
      if (found < 0) { malt_push(o, a->s, a->n, suffix, use_ns); found = o->n_malt - 1; }
      lut_add(&o->luts[q], input_idx, found + 1);
    }
  }
  if (o->nonref_exists) {
    malt_push(o, "<NON_REF>", 9, NULL, 0);
    for (int64_t q = 0; q < o->nq; ++q) if (o->calls[q].valid && o->in_nr[q] >= 0) lut_add(&o->luts[q], o->in_nr[q], o->n_malt);
  }
}

/* ---------------------------------------------------------------- operate = handle_deletions + GA4GHOperator::operate + dump */
static uint32_t* scratch(op_t* o, int32_t n) {
  if (n > o->scratch_cap) { o->scratch_cap = n * 2 + 16; o->scratch = (uint32_t*)realloc(o->scratch, sizeof(uint32_t) * o->scratch_cap); }
  return o->scratch;
}
static void operate(op_t* o) {
  const oracle_input_t* in = o->in;
  if (in->handle_spanning_deletions) handle_deletions(o);
  merge_alleles(o);
  uint32_t A = (uint32_t)o->n_malt + 1u;
  int ref_block_only = (o->mref.n == 1 && o->n_malt == 1 && str_eq(&o->malt[0], "<NON_REF>", 9));
  int remap_needed = !ref_block_only;
  buf_t* b = o->out;
  put_i64(b, o->site_begin); put_i64(b, o->site_end);
  put_u8(b, (uint8_t)o->nonref_exists); put_u8(b, (uint8_t)remap_needed); put_u16(b, (uint16_t)o->n_malt);
  put_str(b, o->mref.s, (uint32_t)o->mref.n);
  for (int32_t m = 0; m < o->n_malt; ++m) put_str(b, o->malt[m].s, (uint32_t)o->malt[m].n);
  uint32_t L = 0;
  for (int64_t q = 0; q < o->nq; ++q) L += o->calls[q].valid ? 1u : 0u;
  put_u32(b, L);
  int suppress_G = remap_needed && (A - 1u > (uint32_t)in->max_alt_alleles_for_G);
  for (int64_t q = 0; q < o->nq; ++q) {
    call_t* c = &o->calls[q];
    if (!c->valid) continue;
    const lut_t* lut = &o->luts[q];
    put_i64(b, c->row); put_i64(b, c->begin); put_i64(b, c->end);
    put_u16(b, (uint16_t)(c->n_alt + 1));
    for (int32_t i = 0; i <= c->n_alt; ++i) put_i32(b, (int32_t)lut_in2m(lut, i));
    uint32_t ploidy = 0;
    if (in->gt_off) {
      if (!c->gt.valid) put_u32(b, 0);
      else {
        put_u32(b, (uint32_t)c->gt.n);
        if (remap_needed) {
          uint32_t* t = scratch(o, c->gt.n);
          remap_GT(c->gt.v, c->gt.n, lut, A, o->nonref_exists, in->gt_phased, t);
          buf_put(b, t, sizeof(uint32_t) * c->gt.n);
          ploidy = get_ploidy(in->gt_phased, (uint32_t)c->gt.n);
        } else buf_put(b, c->gt.v, sizeof(uint32_t) * c->gt.n);
      }
    }
    for (int32_t k = 0; k < in->n_fields; ++k) {
      const oracle_field_t* fd = &in->fields[k];
      const vec_t* f = &c->f[k];
      if (!f->valid || (fd->kind == KIND_G && suppress_G)) { put_u32(b, 0); continue; }
      if (!remap_needed) { put_u32(b, (uint32_t)f->n); buf_put(b, f->v, sizeof(uint32_t) * f->n); continue; }
      uint32_t n_out = num_elements(fd->kind, A - 1u, ploidy);
      uint32_t* t = scratch(o, (int32_t)n_out + 1);
      /* copy_field + resize (variant_operations.cc:640-652): elements the remap does not write keep the copy */
      for (uint32_t i = 0; i < n_out; ++i) t[i] = (i < (uint32_t)f->n) ? f->v[i] : 0u;
      remap_field(fd->kind, f->v, f->n, lut, A, o->nonref_exists, ploidy, t, missing_word(fd));
      put_u32(b, n_out); buf_put(b, t, sizeof(uint32_t) * n_out);
    }
  }
  ++o->n_sites;
}

/* ---------------------------------------------------------------- gt_fill_row (query_variants.cc:1014-1117) */
static int64_t query_row_idx(const oracle_input_t* in, int64_t row) {
  int64_t lo = 0, hi = in->n_query_rows - 1;
  while (lo <= hi) { int64_t mid = (lo + hi) / 2; if (in->query_rows[mid] == row) return mid; if (in->query_rows[mid] < row) lo = mid + 1; else hi = mid - 1; }
  return -1;
}
static void fill_row(op_t* o, int64_t q, int64_t cell, int64_t query_col, int traverse_end_copies) {
  const oracle_input_t* in = o->in;
  call_t* c = &o->calls[q];
  c->initialized = 1; c->contains_deletion = 0; c->is_reference_block = 0;
  int64_t column = in->col[cell], END_v = in->end[cell];
  if ((traverse_end_copies && column <= END_v && column > query_col) || (!traverse_end_copies && column > END_v)) { c->valid = 0; return; }
  c->valid = 1;
  if (column > END_v) { int64_t t = column; column = END_v; END_v = t; }
  c->row = in->row[cell]; c->begin = column; c->end = END_v;
  /* REF: string field; invalid when every char is TileDB NULL (CHAR_MAX) or empty */
  const char* r = (const char*)in->ref + in->ref_off[cell]; int32_t rn = (int32_t)(in->ref_off[cell + 1] - in->ref_off[cell]);
  str_assign(&c->ref, r, rn);
  /* ALT: strtok on '|' -- empty tokens vanish, a NUL terminates (variant_field_data.h:511-537) */
  const char* a = (const char*)in->alt + in->alt_off[cell]; int32_t an = (int32_t)(in->alt_off[cell + 1] - in->alt_off[cell]);
  for (int32_t i = 0; i < an; ++i) if (a[i] == 0) { an = i; break; }
  c->n_alt = 0;
  for (int32_t i = 0; i < an;) {
    while (i < an && a[i] == '|') ++i;
    int32_t s = i; while (i < an && a[i] != '|') ++i;
    if (i > s) {
      if (c->n_alt == c->alt_cap) { int32_t nc = c->alt_cap * 2 + 4; c->alt = (str_t*)realloc(c->alt, sizeof(str_t) * nc);
        memset(c->alt + c->alt_cap, 0, sizeof(str_t) * (nc - c->alt_cap)); c->alt_cap = nc; }
      str_assign(&c->alt[c->n_alt++], a + s, i - s);
    }
  }
  if (in->gt_off) {
    int64_t n = in->gt_off[cell + 1] - in->gt_off[cell];
    const uint32_t* g = (const uint32_t*)in->gt + in->gt_off[cell];
    c->gt.valid = !words_all_null(g, n, TDB_NULL_INT);
    vec_assign(&c->gt, g, c->gt.valid ? (int32_t)n : 0);
  }
  for (int32_t k = 0; k < in->n_fields; ++k) {
    const oracle_field_t* fd = &in->fields[k];
    int64_t n = fd->off[cell + 1] - fd->off[cell];
    const uint32_t* v = fd->val + fd->off[cell];
    c->f[k].valid = !words_all_null(v, n, fd->elem == ELEM_FLOAT ? TDB_NULL_FLT : TDB_NULL_INT);
    vec_assign(&c->f[k], v, c->f[k].valid ? (int32_t)n : 0);
  }
  c->contains_deletion = contains_deletion(c);
  c->is_reference_block = (c->ref.n == 1 && c->n_alt == 1 && is_non_ref(c->alt[0].s, c->alt[0].n));
}

/* ---------------------------------------------------------------- handle_gvcf_ranges (query_variants.cc:296-332) */
static void handle_gvcf_ranges(op_t* o, int64_t* cur, int64_t next, int is_last_call) {
  while (o->heap_n > 0 && (*cur < next || is_last_call)) {
    int64_t top_end = o->calls[o->heap[0]].end;
    int64_t min_end = (is_last_call || top_end < next - 1) ? top_end : next - 1;
    if (o->ndel) min_end = *cur;
    o->site_begin = *cur; o->site_end = min_end;
    operate(o);
    while (o->heap_n > 0 && o->calls[o->heap[0]].end == min_end) {
      call_t* c = &o->calls[o->heap[0]];
      if (c->contains_deletion) --o->ndel;
      c->valid = 0;
      heap_remove(o, 0);
    }
    *cur = min_end + 1;
  }
}

/* ---------------------------------------------------------------- scan_and_operate for ONE column interval */
static void scan_interval(op_t* o) {
  const oracle_input_t* in = o->in;
  int64_t n = in->n_cells;
  for (int64_t q = 0; q < o->nq; ++q) { o->calls[q].initialized = 0; o->calls[q].valid = 0; }
  o->heap_n = 0; o->ndel = 0;
  int64_t cur = -1, start_col = 0, pos = 0;
  if (!in->scan_full) {
    /* gt_get_column: forward iterator from col_lo; the FIRST cell seen for a row decides (query_variants.cc:845-941) */
    int64_t qcol = in->col_lo;
    while (pos < n && in->col[pos] < qcol) ++pos;
    int64_t filled = 0;
    o->site_begin = qcol;
    for (int64_t i = pos; i < n && filled < o->nq; ++i) {
      int64_t q = query_row_idx(in, in->row[i]);
      if (q < 0 || o->calls[q].initialized) continue;
      fill_row(o, q, i, qcol, 1);
      ++filled;
    }
    for (int64_t q = 0; q < o->nq; ++q) if (o->calls[q].valid) {
      heap_push(o, q);
      if (in->handle_spanning_deletions && o->calls[q].contains_deletion) ++o->ndel;
    }
    if (o->heap_n > 0) cur = qcol;
    start_col = (qcol == INT64_MAX) ? qcol : qcol + 1;
    while (pos < n && in->col[pos] < start_col) ++pos;
  }
  if (cur < 0 && pos < n) cur = in->col[pos];
  int end_loop = 0;
  for (; pos < n && !end_loop; ++pos) {
    int64_t ccol = in->col[pos];
    if (ccol > in->end[pos]) continue;                      /* END copies are ignored in the forward scan */
    if (!in->scan_full && ccol > in->col_hi) { end_loop = 1; break; }
    if (ccol != cur) { handle_gvcf_ranges(o, &cur, ccol, 0); cur = ccol; }
    int64_t q = query_row_idx(in, in->row[pos]);
    if (q < 0) continue;
    call_t* c = &o->calls[q];
    if (c->valid && c->end >= ccol) {
      heap_remove(o, c->heap_pos);
      if (!c->contains_deletion && !c->is_reference_block) o->error = 1;   /* "Unhandled overlapping variants" */
      if (c->contains_deletion) --o->ndel;
    }
    fill_row(o, q, pos, cur, 0);
    if (c->valid) { heap_push(o, q); if (in->handle_spanning_deletions && c->contains_deletion) ++o->ndel; }
  }
  if (!in->scan_full) {
    int64_t next = in->col_hi; if (next != INT64_MAX) ++next;
    handle_gvcf_ranges(o, &cur, next, 0);
  } else handle_gvcf_ranges(o, &cur, 0, 1);
}

/* ================================================================ public entry points */
/* Returns 0 on success (1: unhandled overlapping variants, 2: malformed deletion). *out is malloc'ed; free with oracle_free.
 * The dump has NO file header (the caller prepends magic + field names). */
int oracle_combine(const oracle_input_t* in, uint8_t** out, uint64_t* out_len, int64_t* n_sites) {
  op_t o; memset(&o, 0, sizeof(o));
  buf_t b; memset(&b, 0, sizeof(b));
  o.in = in; o.nq = in->n_query_rows; o.out = &b;
  o.calls = (call_t*)calloc(o.nq + 1, sizeof(call_t));
  o.luts = (lut_t*)calloc(o.nq + 1, sizeof(lut_t));
  o.heap = (int64_t*)calloc(o.nq + 1, sizeof(int64_t));
  o.in_nr = (int64_t*)calloc(o.nq + 1, sizeof(int64_t));
  for (int64_t q = 0; q < o.nq; ++q) { o.calls[q].f = (vec_t*)calloc(in->n_fields + 1, sizeof(vec_t)); lut_reserve(&o.luts[q], 10); }
  scan_interval(&o);
  for (int64_t q = 0; q < o.nq; ++q) {
    call_t* c = &o.calls[q];
    free(c->ref.s); for (int32_t i = 0; i < c->alt_cap; ++i) free(c->alt[i].s); free(c->alt); free(c->gt.v);
    for (int32_t k = 0; k < in->n_fields; ++k) free(c->f[k].v);
    free(c->f); free(o.luts[q].in2m); free(o.luts[q].m2in);
  }
  for (int32_t m = 0; m < o.malt_cap; ++m) free(o.malt[m].s);
  free(o.malt); free(o.mref.s); free(o.calls); free(o.luts); free(o.heap); free(o.in_nr); free(o.scratch);
  *out = b.p; *out_len = b.n; *n_sites = o.n_sites;
  return o.error;
}
void oracle_free(uint8_t* p) { free(p); }

/* exported for unit tests against the reference's genotype_ordering gtests (src/test/cpp/src/test_non_diploid_mapper.cc) */
uint64_t oracle_nCr(int64_t n, int64_t r) { return nCr(n, r); }
uint64_t oracle_genotype_index(const int32_t* alleles, int32_t ploidy) {
  int tmp[64]; for (int i = 0; i < ploidy && i < 64; ++i) tmp[i] = alleles[i];
  return genotype_index(tmp, ploidy, 0);
}
uint32_t oracle_num_genotypes(uint32_t n_alt, uint32_t ploidy) { return num_genotypes(n_alt, ploidy); }
/* enumerate merged genotypes in the reference's order: out[g*ploidy + i] = allele i of genotype g */
typedef struct { int32_t* out; int ploidy; uint64_t n; int* vec; } enum_t;
static void enum_rec(enum_t* e, int level, int amax) {
  for (int a = 0; a <= amax; ++a) { e->vec[level] = a;
    if (level == 0) { for (int i = 0; i < e->ploidy; ++i) e->out[e->n * e->ploidy + i] = e->vec[i]; ++e->n; }
    else enum_rec(e, level - 1, a); }
}
uint64_t oracle_enumerate_genotypes(uint32_t n_alleles, uint32_t ploidy, int32_t* out) {
  if (ploidy == 0) return 0;
  enum_t e; e.out = out; e.ploidy = (int)ploidy; e.n = 0; e.vec = (int*)calloc(ploidy + 1, sizeof(int));
  enum_rec(&e, (int)ploidy - 1, (int)n_alleles - 1);
  free(e.vec); return e.n;
}
/* remap one call's G field through an explicit merged->input LUT (the reference's RemapDataTestClass seam) */
void oracle_remap_G(const uint32_t* in, int32_t n_in, const int64_t* m2in, uint32_t A, int nonref, uint32_t ploidy, uint32_t* out) {
  lut_t l; memset(&l, 0, sizeof(l)); lut_reserve(&l, (int32_t)A);
  for (uint32_t m = 0; m < A; ++m) l.m2in[m] = m2in[m];
  remap_field(KIND_G, in, n_in, &l, A, nonref, ploidy, out, BCF_INT_MISSING);
  free(l.in2m); free(l.m2in);
}
