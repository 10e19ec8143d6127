/*
 * GDBSYN-1: deterministic synthetic gVCF cohort in GenomicsDB's cell model, written straight
 * into TileDB column layout (global column-major cell order, END copies included) -- no VCF files, no loader.
 * Tooling for bench.py and the tests; not part of the product path and not part of the oracle.
 *
 *   site i (0 <= i < n_sites) sits at column c_i = first_col + i*spacing; REF base and ALT pool are functions of i.
 *   sample s has a variant call at site i with probability p_var (counter-based hash of (seed,s,i)):
 *       k ~ U{1..min(3,a_pool)} ALTs drawn without replacement in random order + <NON_REF>, GT sorted pair from [0..k],
 *       PL ~ U[0,3000] x G(k+2), AD ~ U[0,80] x (k+2); one cell at (s, c_i), END = c_i.
 *   between its variant calls a sample is covered by reference blocks [c_h+1, c_j-1] (REF N, ALT <NON_REF>, GT 0/0,
 *       PL (0,a,b), no AD); the first block starts at c_0-2, the last ends at c_{n-1}+2.  Every block longer than one
 *       column is stored twice, at its begin and (coords.col and END swapped) at its end, exactly as the loader does
 *       (src/main/cpp/src/loader/load_operators.cc:162-186).
 * A shard is the set of cells whose column lies in the columns of sites [site_lo, site_hi) -- what one GenomicsDB
 * column partition would hold.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t seed;
  int64_t n_samples, n_sites, spacing, first_col;
  double p_var;
  int32_t a_pool;          /* 1..64: 3 SNP bases, 5 short insertions, then insertions REF + a 3-letter code */
  int64_t site_lo, site_hi; /* shard */
} gdbsyn_params;

enum { C_CELLS = 0, C_REF = 1, C_ALT = 2, C_GT = 3, C_PL = 4, C_AD = 5, C_N = 6 };

static inline uint64_t mix(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  return x;
}
static inline uint64_t h3(uint64_t seed, uint64_t s, uint64_t i, uint64_t stream) {
  return mix(seed ^ mix(s * 0x9E3779B97F4A7C15ull + 0x1234567ull) ^ mix(i * 0xC2B2AE3D27D4EB4Full + 0x89ABCDEFull) ^ (stream * 0x165667B19E3779F9ull));
}
static inline int is_variant(const gdbsyn_params* p, int64_t s, int64_t i) {
  if (i < 0 || i >= p->n_sites) return 0;
  return (double)(h3(p->seed, (uint64_t)s, (uint64_t)i, 0) >> 11) * (1.0 / 9007199254740992.0) < p->p_var;
}
static inline int64_t col_of(const gdbsyn_params* p, int64_t i) { return p->first_col + i * p->spacing; }
static inline int64_t prev_variant(const gdbsyn_params* p, int64_t s, int64_t i) { for (int64_t h = i - 1; h >= 0; --h) if (is_variant(p, s, h)) return h; return -1; }
static inline int64_t next_variant(const gdbsyn_params* p, int64_t s, int64_t i) { for (int64_t j = i + 1; j < p->n_sites; ++j) if (is_variant(p, s, j)) return j; return p->n_sites; }
static inline int64_t block_begin(const gdbsyn_params* p, int64_t h) { return h < 0 ? col_of(p, 0) - 2 : col_of(p, h) + 1; }
static inline int64_t block_end(const gdbsyn_params* p, int64_t j) { return j >= p->n_sites ? col_of(p, p->n_sites - 1) + 2 : col_of(p, j) - 1; }

static const char BASES[4] = {'A', 'C', 'G', 'T'};
static const char* INS[5] = {"A", "C", "G", "T", "AA"};

/* variant call of sample s at site i: returns k, fills pick strings */
#define SYN_MAX_POOL 64
typedef struct { int k; char alt[8][8]; int alt_len[8]; char ref; int gt[2]; } var_t;
static void make_variant(const gdbsyn_params* p, int64_t s, int64_t i, var_t* v) {
  uint64_t S = (uint64_t)s, I = (uint64_t)i;
  int r = (int)(mix(p->seed ^ (I * 0xD6E8FEB86659FD93ull)) & 3);
  v->ref = BASES[r];
  char pool[SYN_MAX_POOL][8]; int plen[SYN_MAX_POOL];
  for (int a = 0; a < 3; ++a) { pool[a][0] = BASES[(r + 1 + a) & 3]; plen[a] = 1; }
  for (int a = 0; a < 5; ++a) { pool[3 + a][0] = BASES[r]; int l = (int)strlen(INS[a]); memcpy(pool[3 + a] + 1, INS[a], l); plen[3 + a] = 1 + l; }
  for (int a = 8; a < SYN_MAX_POOL; ++a) { int c = a - 8; pool[a][0] = BASES[r]; pool[a][1] = BASES[(c >> 4) & 3]; pool[a][2] = BASES[(c >> 2) & 3]; pool[a][3] = BASES[c & 3]; plen[a] = 4; }
  int np = p->a_pool > SYN_MAX_POOL ? SYN_MAX_POOL : p->a_pool, kmax = np < 3 ? np : 3;
  v->k = 1 + (int)(h3(p->seed, S, I, 1) % (uint64_t)kmax);
  int idx[SYN_MAX_POOL]; for (int a = 0; a < np; ++a) idx[a] = a;
  for (int t = 0; t < v->k; ++t) { int j = t + (int)(h3(p->seed, S, I, 2 + t) % (uint64_t)(np - t)); int tmp = idx[t]; idx[t] = idx[j]; idx[j] = tmp; }
  for (int t = 0; t < v->k; ++t) { memcpy(v->alt[t], pool[idx[t]], plen[idx[t]]); v->alt_len[t] = plen[idx[t]]; }
  int a = (int)(h3(p->seed, S, I, 10) % (uint64_t)(v->k + 1)), b = (int)(h3(p->seed, S, I, 11) % (uint64_t)(v->k + 1));
  v->gt[0] = a < b ? a : b; v->gt[1] = a < b ? b : a;
}
static inline int variant_alt_bytes(const var_t* v) { int n = 0; for (int t = 0; t < v->k; ++t) n += v->alt_len[t] + 1; return n + 1; }

/* plan: counts per column group.  Groups: [0] leading block begins (only when site_lo==0), then 3 per site
 * (END copies at c_i-1, variants at c_i, block begins at c_i+1), then [last] trailing END copies (only site_hi==n_sites).
 * counts is (2 + 3*(site_hi-site_lo)) x C_N. */
int64_t gdbsyn_num_groups(const gdbsyn_params* p) { return 2 + 3 * (p->site_hi - p->site_lo); }

void gdbsyn_plan(const gdbsyn_params* p, uint64_t* counts) {
  int64_t ng = gdbsyn_num_groups(p);
  memset(counts, 0, sizeof(uint64_t) * C_N * (size_t)ng);
  if (p->site_lo == 0) { uint64_t* c = counts; c[C_CELLS] = p->n_samples; c[C_REF] = p->n_samples; c[C_ALT] = p->n_samples; c[C_GT] = 2 * p->n_samples; c[C_PL] = 3 * p->n_samples; }
  if (p->site_hi == p->n_sites) { uint64_t* c = counts + C_N * (ng - 1); c[C_CELLS] = p->n_samples; c[C_REF] = p->n_samples; c[C_ALT] = p->n_samples; c[C_GT] = 2 * p->n_samples; c[C_PL] = 3 * p->n_samples; }
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = p->site_lo; i < p->site_hi; ++i) {
    uint64_t nv = 0, alt = 0, pl = 0, ad = 0;
    for (int64_t s = 0; s < p->n_samples; ++s) if (is_variant(p, s, i)) {
      var_t v; make_variant(p, s, i, &v);
      ++nv; alt += variant_alt_bytes(&v); pl += (uint64_t)(v.k + 2) * (v.k + 3) / 2; ad += v.k + 2;
    }
    uint64_t* c = counts + C_N * (1 + 3 * (i - p->site_lo));
    c[C_CELLS] = nv; c[C_REF] = nv; c[C_ALT] = nv; c[C_GT] = 2 * nv; c[C_PL] = 3 * nv;                       /* END copies */
    c += C_N; c[C_CELLS] = nv; c[C_REF] = nv; c[C_ALT] = alt; c[C_GT] = 2 * nv; c[C_PL] = pl; c[C_AD] = ad; /* variants */
    c += C_N; c[C_CELLS] = nv; c[C_REF] = nv; c[C_ALT] = nv; c[C_GT] = 2 * nv; c[C_PL] = 3 * nv;             /* block begins */
  }
}

typedef struct {
  int64_t* coords; int64_t* end;
  uint64_t* ref_off; uint8_t* ref; uint64_t* alt_off; uint8_t* alt;
  uint64_t* gt_off; int32_t* gt; uint64_t* pl_off; int32_t* pl; uint64_t* ad_off; int32_t* ad;
  int32_t* dp; int32_t* gq; int32_t* min_dp;    /* fixed attributes (optional, may be NULL) */
} gdbsyn_out;

typedef struct { uint64_t cell, ref, alt, gt, pl, ad; } cur_t;

static void put_block(const gdbsyn_params* p, const gdbsyn_out* o, cur_t* c, int64_t s, int64_t h, int64_t col, int64_t end_field) {
  uint64_t S = (uint64_t)s, H = (uint64_t)(h + 1);
  o->coords[2 * c->cell] = s; o->coords[2 * c->cell + 1] = col; o->end[c->cell] = end_field;
  o->ref_off[c->cell] = c->ref; o->ref[c->ref++] = 'N';
  o->alt_off[c->cell] = c->alt; o->alt[c->alt++] = '&';
  o->gt_off[c->cell] = c->gt * 4; o->gt[c->gt++] = 0; o->gt[c->gt++] = 0;
  int a = 1 + (int)(h3(p->seed, S, H, 201) % 60), b = a + 1 + (int)(h3(p->seed, S, H, 202) % 600);
  o->pl_off[c->cell] = c->pl * 4; o->pl[c->pl++] = 0; o->pl[c->pl++] = a; o->pl[c->pl++] = b;
  o->ad_off[c->cell] = c->ad * 4;
  int dp = 5 + (int)(h3(p->seed, S, H, 203) % 60);
  if (o->dp) o->dp[c->cell] = dp;
  if (o->gq) o->gq[c->cell] = a > 99 ? 99 : a;
  if (o->min_dp) o->min_dp[c->cell] = dp - (int)(h3(p->seed, S, H, 204) % 5);
  ++c->cell;
}

/* starts: exclusive prefix sums of `counts` (same shape), in elements (cells, bytes, int32 words) */
void gdbsyn_fill(const gdbsyn_params* p, const uint64_t* starts, const gdbsyn_out* o) {
  int64_t ng = gdbsyn_num_groups(p);
  if (p->site_lo == 0) {
    const uint64_t* st = starts; cur_t c = {st[C_CELLS], st[C_REF], st[C_ALT], st[C_GT], st[C_PL], st[C_AD]};
    for (int64_t s = 0; s < p->n_samples; ++s) put_block(p, o, &c, s, -1, block_begin(p, -1), block_end(p, next_variant(p, s, -1)));
  }
  if (p->site_hi == p->n_sites) {
    const uint64_t* st = starts + C_N * (ng - 1); cur_t c = {st[C_CELLS], st[C_REF], st[C_ALT], st[C_GT], st[C_PL], st[C_AD]};
    for (int64_t s = 0; s < p->n_samples; ++s) { int64_t h = prev_variant(p, s, p->n_sites); put_block(p, o, &c, s, h, block_end(p, p->n_sites), block_begin(p, h)); }
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = p->site_lo; i < p->site_hi; ++i) {
    const uint64_t* st = starts + C_N * (1 + 3 * (i - p->site_lo));
    cur_t ce = {st[C_CELLS], st[C_REF], st[C_ALT], st[C_GT], st[C_PL], st[C_AD]}; st += C_N;
    cur_t cv = {st[C_CELLS], st[C_REF], st[C_ALT], st[C_GT], st[C_PL], st[C_AD]}; st += C_N;
    cur_t cb = {st[C_CELLS], st[C_REF], st[C_ALT], st[C_GT], st[C_PL], st[C_AD]};
    int64_t ci = col_of(p, i);
    for (int64_t s = 0; s < p->n_samples; ++s) {
      if (!is_variant(p, s, i)) continue;
      uint64_t S = (uint64_t)s, I = (uint64_t)i;
      /* END copy of the block that ends at c_i - 1 */
      int64_t h = prev_variant(p, s, i);
      put_block(p, o, &ce, s, h, ci - 1, block_begin(p, h));
      /* the variant call */
      var_t v; make_variant(p, s, i, &v);
      uint64_t c = cv.cell;
      o->coords[2 * c] = s; o->coords[2 * c + 1] = ci; o->end[c] = ci;
      o->ref_off[c] = cv.ref; o->ref[cv.ref++] = (uint8_t)v.ref;
      o->alt_off[c] = cv.alt;
      for (int t = 0; t < v.k; ++t) { memcpy(o->alt + cv.alt, v.alt[t], v.alt_len[t]); cv.alt += v.alt_len[t]; o->alt[cv.alt++] = '|'; }
      o->alt[cv.alt++] = '&';
      o->gt_off[c] = cv.gt * 4; o->gt[cv.gt++] = v.gt[0]; o->gt[cv.gt++] = v.gt[1];
      int na = v.k + 2, ng2 = na * (na + 1) / 2, dp = 0;
      o->pl_off[c] = cv.pl * 4;
      for (int g = 0; g < ng2; ++g) o->pl[cv.pl++] = (int32_t)(h3(p->seed, S, I, 20 + g) % 3001);
      o->ad_off[c] = cv.ad * 4;
      for (int a = 0; a < na; ++a) { int x = (int)(h3(p->seed, S, I, 100 + a) % 81); o->ad[cv.ad++] = x; dp += x; }
      if (o->dp) o->dp[c] = dp;
      if (o->gq) o->gq[c] = (int)(h3(p->seed, S, I, 12) % 100);
      if (o->min_dp) o->min_dp[c] = 0x7FFFFFFF;   /* TileDB NULL: MIN_DP only exists on reference blocks */
      ++cv.cell;
      /* begin cell of the block that starts at c_i + 1 */
      put_block(p, o, &cb, s, i, ci + 1, block_end(p, next_variant(p, s, i)));
    }
  }
}
