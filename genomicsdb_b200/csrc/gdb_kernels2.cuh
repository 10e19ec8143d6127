// K2b (spanning-deletion pre-pass), K4 (allele union) and K5 (remap) kernels.  See gdb_kernels.cuh for the map.
#pragma once
#include "gdb_kernels.cuh"

namespace gdb {

__device__ __constant__ uint8_t c_str_N[1] = {'N'};
__device__ __constant__ uint8_t c_str_star[1] = {'*'};
__device__ __constant__ uint8_t c_str_amp[1] = {'&'};
__device__ __constant__ uint8_t c_str_nonref[9] = {'<', 'N', 'O', 'N', '_', 'R', 'E', 'F', '>'};

__device__ inline uint32_t missing_word(int elem) { return elem ? BCF_FLT_MISSING : BCF_INT_MISSING; }
__device__ inline uint32_t vend_word(int elem) { return elem ? BCF_FLT_VEND : BCF_INT_VEND; }
__device__ inline uint32_t null_word(int elem) { return elem ? TDB_NULL_FLT : TDB_NULL_INT; }

// ------------------------------------------------------------------------------------------------ sequential remap math
// One call, one thread.  m2in(j) answers "input allele index of merged allele j" (-1 = none); in_nr is the call's own
// NON_REF index (or -1).  Restates variant_field_handler.cc:41-398 (R/A, G haploid / diploid / general).
template <class M2IN>
__device__ inline void remap_field_seq(int kind, const uint32_t* in, uint32_t n_in, M2IN m2in, uint32_t A, int in_nr, uint32_t ploidy,
                                       uint32_t* out, uint32_t missing) {
  if (kind != KIND_G || ploidy == 1) {
    bool alt_only = (kind == KIND_A);
    uint32_t length = (kind == KIND_G) ? A : (alt_only ? A - 1u : A);
    for (uint32_t j = 0; j < length; ++j) {
      int ij = m2in(alt_only ? j + 1u : j);
      if (ij < 0) ij = in_nr;
      if (ij < 0) { out[j] = missing; continue; }
      int64_t ix = alt_only ? (int64_t)ij - 1 : ij;
      out[j] = (ix < 0 || ix >= (int64_t)n_in) ? missing : in[ix];
    }
  } else if (ploidy == 2) {
    for (uint32_t k = 0; k < A; ++k) {
      int ik = m2in(k); if (ik < 0) ik = in_nr;
      for (uint32_t j = 0; j <= k; ++j) {
        int ij = m2in(j); if (ij < 0) ij = in_nr;
        uint32_t g = k * (k + 1) / 2 + j;
        if (ij < 0 || ik < 0) { out[g] = missing; continue; }
        int64_t ig = alleles2gt(ij, ik);
        out[g] = (ig >= (int64_t)n_in) ? missing : in[ig];
      }
    }
  } else if (ploidy > 2 && ploidy <= MAX_PLOIDY) {
    int vec[MAX_PLOIDY], iv[MAX_PLOIDY];
    for (uint32_t i = 0; i < ploidy; ++i) vec[i] = 0;
    uint32_t g = 0;
    do {
      bool miss = false;
      for (uint32_t i = 0; i < ploidy; ++i) { int ia = m2in((uint32_t)vec[i]); if (ia < 0) ia = in_nr; if (ia < 0) miss = true; iv[i] = ia; }
      if (miss) out[g] = missing;
      else { sort_small(iv, (int)ploidy); uint64_t ig = genotype_index_sorted(iv, (int)ploidy); out[g] = (ig >= n_in) ? missing : in[ig]; }
      ++g;
    } while (next_genotype(vec, (int)ploidy, (int)A));
  }
  // ploidy 0: the general routine returns at once; the caller pre-filled out[] with the copied input (variant_operations.cc:640-652)
}
template <class IN2M>
__device__ inline uint32_t remap_gt_word(uint32_t w, IN2M in2m, uint32_t A, bool nonref) {
  if (w == TDB_NULL_INT || w == 0xFFFFFFFFu || w == BCF_INT_MISSING) return w;
  int m = in2m((int32_t)w);
  return (m < 0) ? (nonref ? A - 1u : 0xFFFFFFFFu) : (uint32_t)m;
}
__device__ inline bool words_all_null(const uint32_t* v, uint32_t n, uint32_t nullw) {
  for (uint32_t i = 0; i < n; ++i) if (v[i] != nullw) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------ K2b deletion pre-pass
// For every cell whose call contains a deletion: the form handle_deletions gives it at positions past its begin
// (REF "N", ALT "*"[, <NON_REF>], PL/AD/GT reduced through a 3-allele LUT; broad_combined_gvcf.cc:912-1078).
// Pure function of the cell, so it is computed once per scan instead of once per site.
struct DelLut { int lowest; int nr; __device__ int operator()(uint32_t m) const { return m == 0 ? 0 : (m == 1 ? lowest : (m == 2 ? nr : -1)); } };

__device__ inline void del_analyse(const ArrayDev& A, uint64_t c, int& lowest, int& nr_idx, uint32_t& ploidy) {
  const CellDesc d = A.desc[c];
  ploidy = (d.flags & F_GTVALID) ? ploidy_of(A.gt_phased, d.gt_len) : 0u;
  const uint8_t* alt = A.alt.val + A.alt.off[c];
  int an = alt_effective_len(alt, (int)(A.alt.off[c + 1] - A.alt.off[c]));
  const uint32_t* pl = nullptr; uint32_t pl_n = 0;
  for (int f = 0; f < A.n_fields; ++f) if (A.fm[f].is_pl) {
    uint64_t o0 = A.f[f].off[c], o1 = A.f[f].off[c + 1];
    uint32_t n = (uint32_t)((o1 - o0) >> 2);
    const uint32_t* v = reinterpret_cast<const uint32_t*>(A.f[f].val + o0);
    if (n > 0 && !words_all_null(v, n, TDB_NULL_INT)) { pl = v; pl_n = n; }
  }
  lowest = -1; nr_idx = -1;
  int32_t lowest_pl = INT32_MAX;
  int pos = 0, ts, tl, idx = 0;
  while (next_token(alt, an, pos, ts, tl)) {
    ++idx;
    const uint8_t* t = alt + ts;
    bool del = d.ref_len > 1 && ((tl == 1 && t[0] == '*') || (!is_symbolic(t, tl) && (uint32_t)tl < d.ref_len));
    if (del) {
      if (lowest < 0) lowest = idx;
      if (pl) {
        uint64_t gi;
        if (ploidy == 0) gi = 0; else if (ploidy == 1) gi = (uint64_t)idx; else if (ploidy == 2) gi = (uint64_t)alleles2gt(idx, idx);
        else { gi = 0; for (uint32_t i = 0; i < ploidy; ++i) gi += nCr_u64((int64_t)i + idx, (int64_t)idx - 1); }
        if (gi < pl_n && (int32_t)pl[gi] < lowest_pl) { lowest_pl = (int32_t)pl[gi]; lowest = idx; }
      }
    } else if (is_non_ref(t, tl)) nr_idx = idx;
  }
}
// pass 1: words needed per deletion cell
__global__ void __launch_bounds__(128) k_del_size(ArrayDev A, const uint32_t* __restrict__ del_cells, uint64_t n_del, uint32_t* __restrict__ words) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_del) return;
  uint64_t c = del_cells[t];
  const CellDesc d = A.desc[c];
  uint32_t ploidy = (d.flags & F_GTVALID) ? ploidy_of(A.gt_phased, d.gt_len) : 0u;
  uint32_t n_red_alt = d.nr_idx ? 2u : 1u;
  uint32_t w = d.gt_len;
  for (int f = 0; f < A.n_fields; ++f) w += num_elements(A.fm[f].kind, n_red_alt, ploidy);
  words[t] = w;
}
// pass 2: fill
__global__ void __launch_bounds__(128) k_del_fill(ArrayDev A, int min_pl_gt, const uint32_t* __restrict__ del_cells, uint64_t n_del, const uint32_t* __restrict__ word_off,
                                                   DelRec* __restrict__ recs, uint32_t* __restrict__ pool, uint32_t* __restrict__ err) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_del) return;
  uint64_t c = del_cells[t];
  const CellDesc d = A.desc[c];
  int lowest, nr_idx; uint32_t ploidy;
  del_analyse(A, c, lowest, nr_idx, ploidy);
  if (ploidy > MAX_PLOIDY) { atomicMax(&err[1], 3u); ploidy = 0; }
  DelRec r;
  r.cell = (uint32_t)c; r.has_nr = nr_idx > 0; r.n_alt = r.has_nr ? 2 : 1; r.pad0 = r.pad1 = 0;
  uint32_t A_red = (uint32_t)r.n_alt + 1u;
  DelLut lut{lowest, nr_idx > 0 ? nr_idx : -1};
  int in_nr = nr_idx > 0 ? nr_idx : -1;
  uint32_t o = word_off[t];
  const uint32_t* pl_red = nullptr; uint32_t pl_red_n = 0;
  for (int f = 0; f < A.n_fields; ++f) {
    uint64_t o0 = A.f[f].off[c], o1 = A.f[f].off[c + 1];
    uint32_t n_in = (uint32_t)((o1 - o0) >> 2);
    const uint32_t* in = reinterpret_cast<const uint32_t*>(A.f[f].val + o0);
    uint32_t n_out = num_elements(A.fm[f].kind, A_red - 1u, ploidy);
    r.f_off[f] = o;
    if (n_in == 0 || words_all_null(in, n_in, null_word(A.fm[f].elem))) { r.f_len[f] = 0xFFFFFFFFu; o += n_out; continue; }
    for (uint32_t i = 0; i < n_out; ++i) pool[o + i] = (i < n_in) ? in[i] : 0u;
    remap_field_seq(A.fm[f].kind, in, n_in, lut, A_red, in_nr, ploidy, pool + o, missing_word(A.fm[f].elem));
    r.f_len[f] = n_out;
    if (A.fm[f].is_pl) { pl_red = pool + o; pl_red_n = n_out; }
    o += n_out;
  }
  for (int f = A.n_fields; f < MAX_FIELDS; ++f) { r.f_off[f] = 0; r.f_len[f] = 0xFFFFFFFFu; }
  r.gt_off = o; r.gt_len = d.gt_len;
  if (d.flags & F_GTVALID) {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(A.gt.val + A.gt.off[c]);
    int step = A.gt_phased ? 2 : 1;
    bool from_input = true;
    if (min_pl_gt && pl_red) {
      // update_GT_to_correspond_to_min_PL_value (broad_combined_gvcf.cc:1080-1118) over the ALREADY reduced PL
      int best[MAX_PLOIDY]; int32_t cur_min = INT32_MAX; bool found = false;
      int vec[MAX_PLOIDY];
      auto track = [&](const int* v, uint64_t gi) {
        if (gi < pl_red_n) { uint32_t w = pl_red[gi]; if (w != BCF_INT_MISSING && w != BCF_INT_VEND && (int32_t)w < cur_min) { cur_min = (int32_t)w; found = true; for (uint32_t i = 0; i < ploidy; ++i) best[i] = v[i]; } }
      };
      if (ploidy == 1) {
        uint32_t lim = num_genotypes(A_red - 1u, 1); if (pl_red_n < lim) lim = pl_red_n;
        for (uint32_t i = 0; i < lim; ++i) { vec[0] = (int)i; track(vec, i); }
      } else if (ploidy == 2) {
        for (uint32_t i = 0; i < A_red; ++i) for (uint32_t j = i; j < A_red; ++j) { vec[0] = (int)i; vec[1] = (int)j; track(vec, (uint64_t)alleles2gt(i, j)); }
      } else if (ploidy > 2) {
        for (uint32_t i = 0; i < ploidy; ++i) vec[i] = 0;
        do { track(vec, genotype_index_sorted(vec, (int)ploidy)); } while (next_genotype(vec, (int)ploidy, (int)A_red));
      }
      if (found) {
        from_input = false;
        for (uint32_t i = 0; i < d.gt_len; ++i) pool[o + i] = g[i];
        for (uint32_t i = 0, j = 0; i < d.gt_len; i += step, ++j) pool[o + i] = (uint32_t)best[j];
      }
    }
    if (from_input) {
      auto in2m = [&](int32_t a) { return a == 0 ? 0 : (a == lowest ? 1 : ((nr_idx > 0 && a == nr_idx) ? 2 : -1)); };
      for (uint32_t i = 0; i < d.gt_len; ++i) pool[o + i] = ((i % step) == 0) ? remap_gt_word(g[i], in2m, A_red, r.has_nr) : g[i];
    }
  }
  recs[t] = r;
}

// ------------------------------------------------------------------------------------------------ a call as the operator sees it
struct CallView {
  uint32_t cell; bool reduced; uint8_t flags;
  uint32_t ref_len; const uint8_t* ref;
  int n_alt, nr_idx;
  const uint8_t* alt; int alt_n;       // original ALT bytes; unused for the reduced form
  uint32_t ploidy, gt_len; const uint32_t* gt;
  const DelRec* dr;
  int64_t begin, end;
};
__device__ inline const DelRec* find_delrec(const ScanDev& S, uint32_t cell) {
  uint64_t lo = 0, hi = S.n_del;
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (S.del_cells[mid] < cell) lo = mid + 1; else hi = mid; }
  return &S.del_recs[lo];
}
__device__ inline CallView make_view(const ArrayDev& A, const ScanDev& S, int32_t live_entry) {
  CallView v;
  v.cell = (uint32_t)(live_entry & (LIVE_REDUCED - 1));
  v.reduced = (live_entry & LIVE_REDUCED) != 0;
  const CellDesc d = A.desc[v.cell];
  v.flags = d.flags;
  int64_t col = A.coords[2 * (uint64_t)v.cell + 1], END = A.end[v.cell];
  if (d.flags & F_ENDCOPY) { v.begin = END; v.end = col; } else { v.begin = col; v.end = END; }
  v.ploidy = (d.flags & F_GTVALID) ? ploidy_of(A.gt_phased, d.gt_len) : 0u;
  v.gt_len = (d.flags & F_GTVALID) ? d.gt_len : 0u;
  if (!v.reduced) {
    v.dr = nullptr;
    v.ref_len = d.ref_len; v.ref = A.ref.val + A.ref.off[v.cell];
    v.n_alt = d.n_alt; v.nr_idx = d.nr_idx;
    v.alt = A.alt.val + A.alt.off[v.cell];
    v.alt_n = alt_effective_len(v.alt, (int)(A.alt.off[v.cell + 1] - A.alt.off[v.cell]));
    v.gt = A.has_gt ? reinterpret_cast<const uint32_t*>(A.gt.val + A.gt.off[v.cell]) : nullptr;
  } else {
    v.dr = find_delrec(S, v.cell);
    v.ref_len = 1; v.ref = c_str_N;
    v.n_alt = v.dr->n_alt; v.nr_idx = v.dr->has_nr ? 2 : 0;
    v.alt = nullptr; v.alt_n = 0;
    v.gt = S.del_words + v.dr->gt_off;
  }
  return v;
}
// field f of a call: pointer, length in words, validity (variant_field_data.h:365-384)
__device__ inline bool field_of(const ArrayDev& A, const ScanDev& S, const CallView& v, int f, const uint32_t*& p, uint32_t& n) {
  if (v.reduced) {
    if (v.dr->f_len[f] == 0xFFFFFFFFu) { p = nullptr; n = 0; return false; }
    p = S.del_words + v.dr->f_off[f]; n = v.dr->f_len[f]; return true;
  }
  uint64_t o0 = A.f[f].off[v.cell], o1 = A.f[f].off[v.cell + 1];
  n = (uint32_t)((o1 - o0) >> 2);
  p = reinterpret_cast<const uint32_t*>(A.f[f].val + o0);
  return n > 0;   // caller still has to rule out "every element is TileDB NULL"
}
// hash of an ALT allele as it enters the merged list: the allele, plus the tail of the merged REF when the call's own
// REF is shorter and the allele is not symbolic (variant_operations.cc:163-190)
__device__ inline uint64_t allele_hash(const uint8_t* t, int tl, const uint8_t* mref, uint32_t mref_len, uint32_t call_ref_len, int& suffix) {
  suffix = 0;
  if (call_ref_len < mref_len && !is_symbolic(t, tl)) suffix = (int)(mref_len - call_ref_len);
  uint64_t h = fnv1a(FNV_BASIS, t, tl);
  if (suffix) h = fnv1a(h, mref + call_ref_len, suffix);
  return hash_finish(h ^ ((uint64_t)(tl + suffix) << 40));
}

// ------------------------------------------------------------------------------------------------ K4 allele union
constexpr int UNION_THREADS = 256;
constexpr int UNION_SLOTS = 1024;        // open-addressing table in shared memory
constexpr int UNION_MAX_DISTINCT = 512;  // documented engine limit on distinct ALT alleles per site (excluding NON_REF)
constexpr int REF_MAX_PIECES = 32;

// what K4 leaves behind for a site; K5 reads it
struct SiteRec {
  uint32_t n_alleles;      // A_m = 1 + merged ALT count
  uint32_t n_live;
  uint32_t flags;          // GDBGPU_SITE_*
  uint32_t allele_first;   // index into allele_pos pairs / allele_hash
  uint32_t ref_off, ref_len;            // merged REF bytes in the allele pool
  uint32_t max_gt_len, max_ain;         // max GT words / max (n_alt+1) over live calls
  uint32_t max_ploidy_all;              // max ploidy over live calls
  uint32_t n_nt;                        // live calls that are not (trivial and inline): entries of the site's nt list
  uint32_t max_ploidy[MAX_FIELDS];      // over calls whose field is present
  uint32_t max_in_len[MAX_FIELDS];
};
// per-site accumulators of k_site_scan (zeroed per batch), as words
enum { AGG_NLIVE = 0, AGG_NONREF, AGG_MAXGT, AGG_MAXAIN, AGG_MAXPLOIDY, AGG_NNT, AGG_MAXPL = 6, AGG_MAXIN = 6 + HOT_FIELDS, AGG_WORDS = 6 + 2 * HOT_FIELDS };
struct BatchDev {
  uint64_t first_site; uint32_t n_sites;
  const int32_t* live;        // n_sites x nq
  const uint32_t* lcls;       // n_sites x nq : class word of the live call (k_live_walk)
  SiteRec* sites;
  uint32_t* agg;              // n_sites x AGG_WORDS
  uint2* nt;                  // n_sites x nq : per site, its non-trivial live calls (first agg[AGG_NNT] entries, any order):
                              //   .x = query row, .y = warm-blob offset (16-byte units), or NT_REDUCED | cell for a spanning deletion
  // allele LUT of every nt-list entry, as k_verify_alleles found it (same indexing as nt): bytes 0..5 = merged allele of input
  // ALT 1..6 (0xFF: none), the upper 16 bits = slot of the call's signature in the site's table (SIG_NONE: no slot -- the remap
  // kernels then rebuild the LUT from the hashes).  Calls with one and the same (LUT, #ALT, NON_REF index, ploidy) share one
  // permutation row per site (k_sig_perm), so remapping such a call is a gather through that row (k_remap_mix).
  uint2* nt_lut;
  unsigned long long* sig_keys;   // n_sites x sig_slots, SIG_EMPTY = free (nullptr: no signature tables in this scan)
  uint32_t sig_slots;             // a power of two
  // pools (atomic bump allocation; overflow -> err[2])
  uint32_t* cursors;          // [0] allele records, [1] allele bytes, [2] lut words
  uint32_t* allele_pos; uint64_t* allele_hash; uint32_t cap_allele_recs;
  uint8_t* allele_bytes; uint32_t cap_allele_bytes;
};

constexpr uint32_t NT_REDUCED = 0x80000000u;
constexpr uint32_t SIG_MAX_SLOTS = 1024, SIG_NONE = 0xFFFFu, SIG_MAX_ALT = 6, SIG_MIN_ROWS = 4096;   // tables only for scans of >= SIG_MIN_ROWS samples
constexpr unsigned long long SIG_EMPTY = ~0ull;
constexpr uint32_t SIG_HDR_BYTES = 8;                   // a permutation row starts with the output length of every field (4 x u16)
constexpr uint32_t PERM_MISSING = 0xFFu, PERM_PAD = 0xFEu, PERM_MAX_IN = 0xFDu;   // a row's bytes: input index, or one of the two
__device__ inline uint32_t sig_slot_of(unsigned long long key) { key ^= key >> 29; key *= 0x9E3779B97F4A7C15ull; return (uint32_t)(key >> 40); }
// a non-reduced, non-(trivial and inline) call through its warm blob
struct WarmView { const uint32_t* b; uint32_t n_alt, nr_idx, ref_len, gt_len, ploidy, tok_words; };
// `at`: the blob's bytes where the caller has staged them (shared memory), nullptr = read the pool in HBM
__device__ inline WarmView warm_view(const ArrayDev& A, uint32_t blob_off, const uint32_t* at = nullptr) {
  WarmView v;
  v.b = at ? at : A.warm + (uint64_t)blob_off * 4u;
  const uint4 hd = *reinterpret_cast<const uint4*>(v.b);
  v.n_alt = hd.x; v.nr_idx = hd.y; v.ref_len = hd.z; v.gt_len = warm_gt_len(hd.w); v.tok_words = warm_tok_words(hd.w);
  v.ploidy = v.gt_len ? ploidy_of(A.gt_phased, v.gt_len) : 0u;
  return v;
}
// hash of ALT token t (0-based) as it enters the merged list at a site with merged REF `mref`; 0 for <NON_REF>
__device__ inline uint64_t warm_token_hash(const WarmView& v, uint32_t H0, uint32_t t, const uint8_t* mref, uint32_t mref_len) {
  const uint32_t tokw = v.b[H0 + 2u * v.n_alt + t];
  if (tokw & WARM_TOK_NONREF) return 0ull;
  uint64_t hs = (uint64_t)v.b[H0 + 2u * t] | ((uint64_t)v.b[H0 + 2u * t + 1] << 32);
  uint32_t tl = tokw & 0x3FFFFFFFu;
  if (v.ref_len < mref_len && !(tokw & WARM_TOK_SYMBOLIC)) { hs = fnv1a(hs, mref + v.ref_len, (int)(mref_len - v.ref_len)); tl += mref_len - v.ref_len; }
  return hash_finish(hs ^ ((uint64_t)tl << 40));
}
__device__ inline uint64_t star_hash() { const uint8_t s = '*'; return hash_finish(fnv1a(FNV_BASIS, &s, 1) ^ (1ull << 40)); }

// K4a: one streaming pass over the class words k_live_walk left per (site, sample) (layout: gdb_device.cuh).  Per site:
// number of live calls, the maxima the output layout needs (packed into the class word of every trivial inline call) and
// the list of all other live calls.  Grid = (sites, row chunks), so the whole GPU works on a batch whatever its number of
// sites.  Most calls live at a site are reference blocks with one and the same class word: a word equal to the thread's
// previous one only counts.
constexpr int SSCAN_THREADS = 256, SSCAN_PER_THREAD = 16, SSCAN_ROWS = SSCAN_THREADS * SSCAN_PER_THREAD;
__global__ void __launch_bounds__(SSCAN_THREADS) k_site_scan(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B) {
  __shared__ uint2 s_list[SSCAN_ROWS];
  __shared__ uint32_t s_red[AGG_WORDS];
  __shared__ uint32_t s_cnt, s_base;
  const uint32_t js = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t q0 = (uint64_t)blockIdx.y * SSCAN_ROWS;
  const uint32_t* lcls = B.lcls + (uint64_t)js * S.nq;
  if (tid == 0) s_cnt = 0;
  if (tid < AGG_WORDS) s_red[tid] = 0;
  __syncthreads();
  uint32_t c[SSCAN_PER_THREAD];
#pragma unroll
  for (int i = 0; i < SSCAN_PER_THREAD; ++i) { const uint64_t q = q0 + (uint64_t)i * SSCAN_THREADS + tid; c[i] = q < S.nq ? lcls[q] : CLS_DEAD; }
  uint32_t n_live = 0, nonref = 0, mgt = 0, main_ = 0, mpl = 0;
  uint32_t fpl[HOT_FIELDS] = {0, 0, 0, 0}, fin[HOT_FIELDS] = {0, 0, 0, 0};
  uint32_t last = CLS_DEAD;
#pragma unroll
  for (int i = 0; i < SSCAN_PER_THREAD; ++i) {
    const bool live = c[i] != CLS_DEAD;
    const bool general = live && (c[i] & (CLS_GENERAL | CLS_REDUCED));
    if (live) {
      ++n_live;
      if (!general && c[i] != last) {
        last = c[i];
        const uint32_t nr = c[i] & 1u, n_alt = (c[i] >> 1) & 1u, gt_len = (c[i] >> 2) & 7u, ploidy = (c[i] >> 5) & 7u;
        mgt = max(mgt, gt_len); main_ = max(main_, n_alt + 1u); mpl = max(mpl, ploidy);
        if (nr) nonref = 1;
#pragma unroll
        for (int f = 0; f < HOT_FIELDS; ++f) { const uint32_t n = (c[i] >> (8 + 3 * f)) & 7u; if (n) { fpl[f] = max(fpl[f], ploidy); fin[f] = max(fin[f], n); } }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, general);
    if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&s_cnt, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (general) s_list[base + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)(q0 + (uint64_t)i * SSCAN_THREADS + tid),
                                                                              (c[i] & CLS_REDUCED) ? (NT_REDUCED | (c[i] & CLS_LOW)) : (c[i] & CLS_LOW));
    }
  }
  uint32_t r;
  r = __reduce_add_sync(0xffffffffu, n_live); if (lane == 0 && r) atomicAdd(&s_red[AGG_NLIVE], r);
  r = __reduce_max_sync(0xffffffffu, nonref); if (lane == 0 && r) atomicMax(&s_red[AGG_NONREF], r);
  r = __reduce_max_sync(0xffffffffu, mgt); if (lane == 0 && r) atomicMax(&s_red[AGG_MAXGT], r);
  r = __reduce_max_sync(0xffffffffu, main_); if (lane == 0 && r) atomicMax(&s_red[AGG_MAXAIN], r);
  r = __reduce_max_sync(0xffffffffu, mpl); if (lane == 0 && r) atomicMax(&s_red[AGG_MAXPLOIDY], r);
#pragma unroll
  for (int f = 0; f < HOT_FIELDS; ++f) {
    r = __reduce_max_sync(0xffffffffu, fin[f]);
    if (r) {   // warp-uniform
      if (lane == 0) atomicMax(&s_red[AGG_MAXIN + f], r);
      r = __reduce_max_sync(0xffffffffu, fpl[f]); if (lane == 0 && r) atomicMax(&s_red[AGG_MAXPL + f], r);
    }
  }
  __syncthreads();
  uint32_t* agg = B.agg + (uint64_t)js * AGG_WORDS;
  if (tid == 0) { s_base = s_cnt ? atomicAdd(&agg[AGG_NNT], s_cnt) : 0u; if (s_red[AGG_NLIVE]) atomicAdd(&agg[AGG_NLIVE], s_red[AGG_NLIVE]); }
  else if (tid < AGG_WORDS && tid != AGG_NNT && s_red[tid]) atomicMax(&agg[tid], s_red[tid]);
  __syncthreads();
  uint2* dst = B.nt + (uint64_t)js * S.nq + s_base;
  for (uint32_t i = tid; i < s_cnt; i += SSCAN_THREADS) dst[i] = s_list[i];
}

// K4b: one block per site.  Merged REF from the begin cells of the site's column, merged ALT list from the site's nt list.
__global__ void __launch_bounds__(UNION_THREADS, 4) k_union(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B) {
  __shared__ unsigned long long s_hash[UNION_SLOTS];
  __shared__ unsigned long long s_order[UNION_SLOTS];
  __shared__ uint16_t s_list[UNION_MAX_DISTINCT];
  __shared__ uint16_t s_rank[UNION_MAX_DISTINCT];
  __shared__ uint32_t s_piece_cell[REF_MAX_PIECES], s_piece_from[REF_MAX_PIECES], s_piece_to[REF_MAX_PIECES], s_piece_dst[REF_MAX_PIECES];   // bytes [from, to) of the cell's REF land at dst of the merged REF
  __shared__ unsigned int s_k, s_next, s_maxN, s_maxlen, s_ndist, s_nonref, s_maxgt, s_maxain, s_maxploidy, s_npieces, s_reflen, s_refoff, s_afirst, s_overflow;
  __shared__ unsigned int s_maxpl[MAX_FIELDS], s_maxin[MAX_FIELDS];

  const uint32_t js = blockIdx.x;
  const uint64_t j = B.first_site + js;
  const int32_t* live = B.live + (uint64_t)js * S.nq;
  const uint32_t* agg = B.agg + (uint64_t)js * AGG_WORDS;
  const int tid = threadIdx.x;

  for (int i = tid; i < UNION_SLOTS; i += UNION_THREADS) { s_hash[i] = 0ull; s_order[i] = ~0ull; }
  if (tid == 0) {
    s_k = 0xFFFFFFFFu; s_next = 0xFFFFFFFFu; s_maxN = 0; s_maxlen = 0; s_ndist = 0; s_npieces = 0; s_overflow = 0;
    s_nonref = agg[AGG_NONREF]; s_maxgt = agg[AGG_MAXGT]; s_maxain = agg[AGG_MAXAIN]; s_maxploidy = agg[AGG_MAXPLOIDY];
  }
  if (tid < MAX_FIELDS) { s_maxpl[tid] = tid < HOT_FIELDS ? agg[AGG_MAXPL + tid] : 0u; s_maxin[tid] = tid < HOT_FIELDS ? agg[AGG_MAXIN + tid] : 0u; }
  __syncthreads();

  // ---- merged REF (variant_operations.cc:73-122): fold over the calls that BEGIN at this site, in row order.
  // Those are exactly the begin cells of column b (contiguous in the merged cell order; range found once per scan, k_fill_sites).
  const uint64_t c0 = S.site_c0[j], c1 = S.site_c1[j];
  // a begin cell of column b whose row is queried IS the live call of that row at this site (k_live_walk), so the 16-byte
  // header of its hot record answers everything the fold asks: flags, REF length, first REF byte
  auto candidate = [&](uint64_t c, uint32_t& rl, uint8_t& first) -> bool {
    const uint4 hh = *reinterpret_cast<const uint4*>(&A.hot[c]);
    if (hh.x & (uint32_t)(F_ENDCOPY | F_ROWOUT)) return false;
    if (S.qrow_of_row && qrow_of(A, S, A.coords[2 * c]) < 0) return false;
    rl = (hh.z >> 16) & 0xFFu; first = (uint8_t)(hh.z >> 24);
    if (rl == 255u) rl = A.desc[c].ref_len;
    return true;
  };
  {   // per-thread partial results first (the loads of successive trips do not wait for the shared-memory atomics)
    uint32_t t_maxlen = 0, t_maxN = 0, t_k = 0xFFFFFFFFu;
    for (uint64_t c = c0 + tid; c < c1; c += UNION_THREADS) {
      uint32_t rl; uint8_t first;
      if (!candidate(c, rl, first)) continue;
      t_maxlen = max(t_maxlen, rl);
      if (first != 'N') t_k = min(t_k, (uint32_t)(c - c0));
      else t_maxN = max(t_maxN, rl);
    }
    if (t_maxlen) atomicMax(&s_maxlen, t_maxlen);
    if (t_maxN) atomicMax(&s_maxN, t_maxN);
    if (t_k != 0xFFFFFFFFu) atomicMin(&s_k, t_k);
  }
  __syncthreads();
  if (s_k == 0xFFFFFFFFu) {
    if (s_maxN > 1) {   // only N-leading REFs: the first one of maximal length replaces the default "N"
      for (uint64_t c = c0 + tid; c < c1; c += UNION_THREADS) {
        uint32_t rl; uint8_t first;
        if (candidate(c, rl, first) && first == 'N' && rl == s_maxN) atomicMin(&s_next, (unsigned int)(c - c0));
      }
      __syncthreads();
      if (tid == 0) { s_piece_cell[0] = (uint32_t)(c0 + s_next); s_piece_from[0] = 0; s_piece_to[0] = s_maxN; s_piece_dst[0] = 0; s_npieces = 1; s_reflen = s_maxN; }
    } else if (tid == 0) { s_npieces = 0; s_reflen = 1; }
    __syncthreads();
  } else {
    uint32_t last = s_k, cur_len = A.desc[c0 + s_k].ref_len, str_len = cur_len;
    if (tid == 0) { s_piece_cell[0] = (uint32_t)(c0 + last); s_piece_from[0] = 0; s_piece_to[0] = cur_len; s_piece_dst[0] = 0; s_npieces = 1; }
    // The reference's fold compares every REF with a LENGTH REGISTER, not with the string it holds: an N-leading REF of several
    // bases seen before the first real REF raises the register, the real REF then replaces the string but not the register
    // (variant_operations.cc:108-121).  A later REF is appended only when it is longer than the register, and from the
    // register's offset.  (N-leading REFs longer than one base do not occur in gVCFs; s_maxN > 1 is never true there.)
    if (s_maxN > 1) {
      __syncthreads();
      if (tid == 0) s_next = 0;
      __syncthreads();
      for (uint64_t c = c0 + tid; c < c0 + s_k; c += UNION_THREADS) {
        uint32_t rl; uint8_t first;
        if (candidate(c, rl, first) && first == 'N') atomicMax(&s_next, rl);
      }
      __syncthreads();
      cur_len = max(cur_len, s_next);
    }
    while (s_maxlen > cur_len) {   // later, longer REFs append their tail ("record breakers"); none when nothing is longer
      __syncthreads();
      if (tid == 0) s_next = 0xFFFFFFFFu;
      __syncthreads();
      for (uint64_t c = c0 + last + 1 + tid; c < c1; c += UNION_THREADS) {
        uint32_t rl; uint8_t first;
        if (candidate(c, rl, first) && rl > cur_len) atomicMin(&s_next, (unsigned int)(c - c0));
      }
      __syncthreads();
      uint32_t nx = s_next;
      if (nx == 0xFFFFFFFFu) break;
      uint32_t nl = A.desc[c0 + nx].ref_len;
      if (tid == 0) {
        if (s_npieces < REF_MAX_PIECES) { s_piece_cell[s_npieces] = (uint32_t)(c0 + nx); s_piece_from[s_npieces] = cur_len; s_piece_to[s_npieces] = nl; s_piece_dst[s_npieces] = str_len; ++s_npieces; }
        else atomicMax(&S.err[1], 4u);
      }
      last = nx; str_len += nl - cur_len; cur_len = nl;
    }
    __syncthreads();
    if (tid == 0) s_reflen = str_len;
    __syncthreads();
  }
  if (tid == 0) {
    uint32_t off = atomicAdd(&B.cursors[1], s_reflen);
    if (off + s_reflen > B.cap_allele_bytes) { s_overflow = 1; off = 0; }
    s_refoff = off;
  }
  __syncthreads();
  const uint32_t mref_len = s_reflen;
  uint8_t* mref = B.allele_bytes + s_refoff;
  if (!s_overflow) {
    if (s_npieces == 0) { if (tid == 0) mref[0] = 'N'; }
    else for (uint32_t p = 0; p < s_npieces; ++p) {
      const uint8_t* src = A.ref.val + A.ref.off[s_piece_cell[p]];
      for (uint32_t i = s_piece_from[p] + tid; i < s_piece_to[p]; i += UNION_THREADS) mref[s_piece_dst[p] + (i - s_piece_from[p])] = src[i];
    }
  }
  __syncthreads();   // mref is read below by other threads of this block (global memory, same block: needs the barrier only)

  // ---- merged ALT list (variant_operations.cc:134-228): first-seen order over live calls in row order.  Trivial calls
  // add no allele, so only the site's nt list is visited; the order key (row, token index) makes the result independent
  // of the order of that list.
  uint32_t loc_maxgt = 0, loc_maxain = 0, loc_maxploidy = 0;
  uint32_t loc_maxpl[MAX_FIELDS], loc_maxin[MAX_FIELDS];
#pragma unroll
  for (int f = 0; f < MAX_FIELDS; ++f) { loc_maxpl[f] = 0; loc_maxin[f] = 0; }
  bool loc_nonref = false, any = false;
  auto insert = [&](uint64_t h, uint64_t q, uint32_t idx) {
    unsigned long long order = ((unsigned long long)q << 12) | idx;
    uint32_t slot = (uint32_t)(h % UNION_SLOTS);
    for (int probe = 0; probe < UNION_SLOTS; ++probe) {
      unsigned long long old = atomicCAS(&s_hash[slot], 0ull, (unsigned long long)h);
      if (old == 0ull) { if (atomicAdd(&s_ndist, 1u) >= UNION_MAX_DISTINCT) atomicMax(&S.err[1], 2u); }
      if (old == 0ull || old == h) { atomicMin(&s_order[slot], order); return; }
      slot = (slot + 1) % UNION_SLOTS;
    }
    atomicMax(&S.err[1], 2u);
  };
  const uint32_t n_nt = agg[AGG_NNT];
  const uint2* nt = B.nt + (uint64_t)js * S.nq;
  const uint32_t H0 = warm_hash_word((uint32_t)A.n_fields);
  // UNION_ILP entries per thread and trip: their list entries, then the first 32 bytes of their blobs (header + field
  // lengths) are loaded together, and the blobs' second sector (token hashes / lengths of a call with up to 3 ALT alleles)
  // is on its way while the headers are decoded -- one memory round trip per step of the dependency chain and trip
  // instead of one per entry
  constexpr int UNION_ILP = 2;
  if (!s_overflow) for (uint32_t i0 = tid; i0 < n_nt; i0 += UNION_ILP * UNION_THREADS) {
    uint2 ent[UNION_ILP]; uint4 hd[UNION_ILP], fl[UNION_ILP];
#pragma unroll
    for (int u = 0; u < UNION_ILP; ++u) { const uint32_t i = i0 + u * UNION_THREADS; ent[u] = i < n_nt ? nt[i] : make_uint2(0xFFFFFFFFu, NT_REDUCED); }
#pragma unroll
    for (int u = 0; u < UNION_ILP; ++u) {
      hd[u] = make_uint4(0, 0, 0, 0); fl[u] = hd[u];
      if (!(ent[u].y & NT_REDUCED)) {
        const uint4* p = reinterpret_cast<const uint4*>(A.warm + (uint64_t)ent[u].y * 4u);
        hd[u] = p[0]; fl[u] = p[1];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 2));
      }
    }
#pragma unroll
    for (int u = 0; u < UNION_ILP; ++u) {
      if (ent[u].x == 0xFFFFFFFFu) continue;
      const uint32_t q = ent[u].x;
      any = true;
      if (ent[u].y & NT_REDUCED) {
        CallView v = make_view(A, S, (int32_t)((ent[u].y & ~NT_REDUCED) | (uint32_t)LIVE_REDUCED));
        loc_maxgt = max(loc_maxgt, v.gt_len); loc_maxain = max(loc_maxain, (uint32_t)v.n_alt + 1u); loc_maxploidy = max(loc_maxploidy, v.ploidy);
#pragma unroll
        for (int f = 0; f < MAX_FIELDS; ++f) if (f < A.n_fields) {
          const uint32_t* p; uint32_t n;
          if (field_of(A, S, v, f, p, n)) { loc_maxpl[f] = max(loc_maxpl[f], v.ploidy); loc_maxin[f] = max(loc_maxin[f], n); }
        }
        if (v.nr_idx) loc_nonref = true;
        insert(star_hash(), q, 1);
        continue;
      }
      WarmView v;
      v.b = A.warm + (uint64_t)ent[u].y * 4u;
      v.n_alt = hd[u].x; v.nr_idx = hd[u].y; v.ref_len = hd[u].z; v.gt_len = warm_gt_len(hd[u].w); v.tok_words = warm_tok_words(hd[u].w);
      v.ploidy = v.gt_len ? ploidy_of(A.gt_phased, v.gt_len) : 0u;
      loc_maxgt = max(loc_maxgt, v.gt_len); loc_maxain = max(loc_maxain, v.n_alt + 1u); loc_maxploidy = max(loc_maxploidy, v.ploidy);
      const uint32_t fl4[4] = {fl[u].x, fl[u].y, fl[u].z, fl[u].w};
#pragma unroll
      for (int f = 0; f < MAX_FIELDS; ++f) if (f < A.n_fields) {
        const uint32_t n = f < 4 ? fl4[f] : v.b[4 + f];
        if (n) { loc_maxpl[f] = max(loc_maxpl[f], v.ploidy); loc_maxin[f] = max(loc_maxin[f], n); }
      }
      if (v.nr_idx) loc_nonref = true;
      if (v.n_alt > (uint32_t)MAX_CALL_ALLELES - 1u) { atomicMax(&S.err[1], 1u); continue; }
      for (uint32_t t = 0; t < v.n_alt; ++t) {
        const uint64_t h = warm_token_hash(v, H0, t, mref, mref_len);
        if (h) insert(h, q, t + 1u);
      }
    }
  }
  if (any) {
    atomicMax(&s_maxgt, loc_maxgt); atomicMax(&s_maxain, loc_maxain); atomicMax(&s_maxploidy, loc_maxploidy);
    if (loc_nonref) s_nonref = 1;
#pragma unroll
    for (int f = 0; f < MAX_FIELDS; ++f) if (f < A.n_fields) { if (loc_maxin[f]) { atomicMax(&s_maxpl[f], loc_maxpl[f]); atomicMax(&s_maxin[f], loc_maxin[f]); } }
  }
  __syncthreads();
  const uint32_t ndist = s_ndist < UNION_MAX_DISTINCT ? s_ndist : UNION_MAX_DISTINCT;
  // compact occupied slots, rank them by first occurrence
  if (tid == 0) s_next = 0;
  __syncthreads();
  for (int i = tid; i < UNION_SLOTS; i += UNION_THREADS) if (s_hash[i] != 0ull) { uint32_t p = atomicAdd(&s_next, 1u); if (p < UNION_MAX_DISTINCT) s_list[p] = (uint16_t)i; }
  __syncthreads();
  for (uint32_t i = tid; i < ndist; i += UNION_THREADS) {
    unsigned long long o = s_order[s_list[i]]; uint32_t r = 0;
    for (uint32_t k = 0; k < ndist; ++k) r += s_order[s_list[k]] < o;
    s_rank[i] = (uint16_t)r;
  }
  const uint32_t A_m = 1u + ndist + (s_nonref ? 1u : 0u);
  if (tid == 0) {
    uint32_t first = atomicAdd(&B.cursors[0], A_m);
    if (first + A_m > B.cap_allele_recs) { s_overflow = 1; first = 0; }
    s_afirst = first;
  }
  __syncthreads();
  if (!s_overflow) {
    uint32_t first = s_afirst;
    if (tid == 0) {
      B.allele_pos[2 * first] = s_refoff; B.allele_pos[2 * first + 1] = mref_len; B.allele_hash[first] = 0;
      if (s_nonref) {
        uint32_t off = atomicAdd(&B.cursors[1], 9u);
        if (off + 9u > B.cap_allele_bytes) s_overflow = 1;
        else { for (int i = 0; i < 9; ++i) B.allele_bytes[off + i] = c_str_nonref[i]; B.allele_pos[2 * (first + A_m - 1)] = off; B.allele_pos[2 * (first + A_m - 1) + 1] = 9; B.allele_hash[first + A_m - 1] = 0; }
      }
    }
    // one thread per distinct allele writes its string (representative = first occurrence)
    for (uint32_t i = tid; i < ndist; i += UNION_THREADS) {
      uint32_t slot = s_list[i], m = 1u + s_rank[i];
      unsigned long long o = s_order[slot];
      uint64_t q = o >> 12; uint32_t idx = (uint32_t)(o & 0xFFFu);
      CallView v = make_view(A, S, live[q]);
      const uint8_t* t = c_str_star; int tl = 1; uint32_t crl = 1;
      if (!v.reduced) {
        int pos = 0, ts = 0, l = 0; uint32_t k = 0;
        while (next_token(v.alt, v.alt_n, pos, ts, l)) { if (++k == idx) break; }
        t = v.alt + ts; tl = l; crl = v.ref_len;
      }
      int sfx = (crl < mref_len && !is_symbolic(t, tl)) ? (int)(mref_len - crl) : 0;
      uint32_t len = (uint32_t)(tl + sfx);
      uint32_t off = atomicAdd(&B.cursors[1], len);
      if (off + len > B.cap_allele_bytes) { s_overflow = 1; continue; }
      for (int k2 = 0; k2 < tl; ++k2) B.allele_bytes[off + k2] = t[k2];
      for (int k2 = 0; k2 < sfx; ++k2) B.allele_bytes[off + tl + k2] = mref[crl + k2];
      B.allele_pos[2 * (first + m)] = off; B.allele_pos[2 * (first + m) + 1] = len;
      B.allele_hash[first + m] = s_hash[slot];
    }
  }
  __syncthreads();
  __syncthreads();
  if (tid == 0) {
    if (s_overflow) atomicOr(&S.err[2], 1u);
    SiteRec r;
    r.n_alleles = A_m; r.n_live = agg[AGG_NLIVE]; r.allele_first = s_afirst; r.ref_off = s_refoff; r.ref_len = mref_len; r.n_nt = n_nt;
    bool ref_block_only = (mref_len == 1 && ndist == 0 && s_nonref);
    r.flags = (s_nonref ? 1u : 0u) | (ref_block_only ? 0u : 2u);
    if (!ref_block_only && A_m - 1u > S.max_alt) r.flags |= 4u;
    r.max_gt_len = s_maxgt; r.max_ain = s_maxain; r.max_ploidy_all = s_maxploidy;
    for (int f = 0; f < MAX_FIELDS; ++f) { r.max_ploidy[f] = s_maxpl[f]; r.max_in_len[f] = s_maxin[f]; }
    B.sites[js] = r;
  }
}

// K4c: equal hashes never decide alone.  k_union keys its table by a 64-bit hash of the allele string and the remap kernels
// find a call's alleles in the merged list by that hash; this pass compares, for every ALT token of every non-trivial call
// of the batch, the token's bytes (+ the tail of the merged REF it was extended by) with the bytes of the merged allele its
// hash leads to -- what the reference's unordered_map<string> does implicitly (variant_operations.cc:163-190).  A mismatch
// is a 64-bit collision between two different alleles of one site: reported (GDBGPU_ERR_LIMIT), never merged silently.
// One thread per nt-list entry, grid = (row chunks, sites): a throughput pass, off the remap kernels' dependency chains.
constexpr int VERIFY_THREADS = 256;
__global__ void __launch_bounds__(VERIFY_THREADS) k_verify_alleles(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B) {
  const uint32_t js = blockIdx.y;
  const SiteRec* __restrict__ site = &B.sites[js];
  const uint32_t i = blockIdx.x * VERIFY_THREADS + threadIdx.x;
  if (i >= site->n_nt || site->n_alleles < 2u) return;
  const uint32_t A_m = site->n_alleles, afirst = site->allele_first, mref_len = site->ref_len;
  const uint8_t* mref = B.allele_bytes + site->ref_off;
  const uint64_t at = (uint64_t)js * S.nq + i;
  const uint2 ent = B.nt[at];
  auto merged_of = [&](uint64_t h) -> uint32_t { for (uint32_t m = 1; m < A_m; ++m) if (B.allele_hash[afirst + m] == h) return afirst + m; return 0xFFFFFFFFu; };
  unsigned long long lut = 0x0000FFFFFFFFFFFFull;   // bytes 0..5: merged allele of input ALT 1..6
  uint32_t n_alt = 0, nr_idx = 0, ploidy = 0;
  auto set_lut = [&](uint32_t t, uint32_t m) { if (t < SIG_MAX_ALT) lut = (lut & ~(0xFFull << (8 * t))) | ((unsigned long long)(m & 0xFFu) << (8 * t)); };
  if (ent.y & NT_REDUCED) {   // a spanning deletion contributes "*"
    const uint32_t m = merged_of(star_hash());
    if (m == 0xFFFFFFFFu || B.allele_pos[2 * m + 1] != 1u || B.allele_bytes[B.allele_pos[2 * m]] != (uint8_t)'*') atomicMax(&S.err[1], 6u);
    const uint32_t cell = ent.y & ~NT_REDUCED;
    const DelRec* dr = find_delrec(S, cell);
    const CellDesc d = A.desc[cell];
    n_alt = dr->n_alt; nr_idx = dr->has_nr ? 2u : 0u;
    const uint32_t gt_len = (d.flags & F_GTVALID) ? d.gt_len : 0u; ploidy = gt_len ? ploidy_of(A.gt_phased, gt_len) : 0u;
    if (m != 0xFFFFFFFFu) set_lut(0, m - afirst);
    if (nr_idx) set_lut(nr_idx - 1u, A_m - 1u);
  } else {
    const WarmView v = warm_view(A, ent.y);
    const uint32_t H0 = warm_hash_word((uint32_t)A.n_fields);
    const uint32_t* tokw = v.b + H0 + 2u * v.n_alt;
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(tokw + v.n_alt);
    n_alt = v.n_alt; nr_idx = v.nr_idx; ploidy = v.ploidy;
    for (uint32_t t = 0; t < v.n_alt && t < (uint32_t)MAX_CALL_ALLELES; ++t) {
      const uint32_t w = tokw[t], tl = w & 0x3FFFFFFFu;
      if (!(w & WARM_TOK_NONREF)) {
        const uint32_t m = merged_of(warm_token_hash(v, H0, t, mref, mref_len));
        bool same = m != 0xFFFFFFFFu;
        if (same) {
          const uint8_t* as = B.allele_bytes + B.allele_pos[2 * m]; const uint32_t al = B.allele_pos[2 * m + 1];
          const uint32_t sfx = (v.ref_len < mref_len && !(w & WARM_TOK_SYMBOLIC)) ? mref_len - v.ref_len : 0u;
          same = al == tl + sfx;
          for (uint32_t x = 0; same && x < tl; ++x) same = as[x] == tb[x];
          for (uint32_t x = 0; same && x < sfx; ++x) same = as[tl + x] == mref[v.ref_len + x];
        }
        if (!same) atomicMax(&S.err[1], 6u);
        else set_lut(t, m - afirst);
      }
      tb += tl;
    }
    if (nr_idx) set_lut(nr_idx - 1u, A_m - 1u);   // after the tokens: the remap kernels give <NON_REF> the last merged index whatever its hash
  }
  // ---- the call's signature -> its slot in the site's table (open addressing, first come first served)
  uint32_t slot = SIG_NONE;
  if (B.sig_keys && n_alt <= SIG_MAX_ALT && ploidy <= 2u && A_m <= 0xFDu && (site->flags & 2u)) {
    const unsigned long long key = lut | ((unsigned long long)(n_alt | (nr_idx << 3) | (ploidy << 6)) << 48);
    unsigned long long* tab = B.sig_keys + (uint64_t)js * B.sig_slots;
    uint32_t p = sig_slot_of(key) & (B.sig_slots - 1u);
    for (uint32_t probe = 0; probe < 32u; ++probe, p = (p + 1u) & (B.sig_slots - 1u)) {   // a full neighbourhood: the call keeps its own LUT
      unsigned long long cur = tab[p];
      if (cur == SIG_EMPTY) cur = atomicCAS(&tab[p], SIG_EMPTY, key);
      if (cur == SIG_EMPTY || cur == key) { slot = p; break; }
    }
  }
  B.nt_lut[at] = make_uint2((uint32_t)lut, (uint32_t)(lut >> 32) | (slot << 16));
}

// ------------------------------------------------------------------------------------------------ K5 remap
constexpr int REMAP_THREADS = 256;
constexpr int REMAP_WARPS = REMAP_THREADS / 32;
constexpr int REMAP_ROWS = 64;            // rows of one site handled by one block

struct FieldOutDev { uint32_t* data; const uint64_t* base; const uint32_t* stride; uint16_t* len; };
struct RemapOut {
  int32_t n_out;                 // GT (if present) + remapped fields
  FieldOutDev f[MAX_FIELDS + 1];
  uint32_t* lut_first; uint16_t* lut_len; int16_t* lut_in2m; uint32_t cap_lut;   // optional
};

// One warp per call.  The allele LUT of the call is rebuilt in shared memory from the site's merged-allele hashes
// (K4), then every output word is an indexed read of the call's input vector.
__global__ void __launch_bounds__(REMAP_THREADS) k_remap_generic(ArrayDev A, ScanDev S, BatchDev B, RemapOut O, const uint32_t* __restrict__ site_list) {
  __shared__ unsigned long long s_ahash[UNION_MAX_DISTINCT + 2];
  __shared__ int16_t s_m2in[REMAP_WARPS][MAX_CALL_ALLELES];
  __shared__ int16_t s_in2m[REMAP_WARPS][MAX_CALL_ALLELES];
  const uint32_t tiles_per_site = (uint32_t)((S.nq + REMAP_ROWS - 1) / REMAP_ROWS);
  const uint32_t js = site_list[blockIdx.x / tiles_per_site], tile = blockIdx.x % tiles_per_site;
  const SiteRec site = B.sites[js];
  const int32_t* live = B.live + (uint64_t)js * S.nq;
  const uint32_t A_m = site.n_alleles;
  const bool nonref = site.flags & 1u, remap = site.flags & 2u, g_suppressed = site.flags & 4u;
  const uint8_t* mref = B.allele_bytes + site.ref_off;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < A_m && i < UNION_MAX_DISTINCT + 2; i += REMAP_THREADS) s_ahash[i] = B.allele_hash[site.allele_first + i];
  __syncthreads();
  int16_t* m2in = s_m2in[w]; int16_t* in2m = s_in2m[w];
  const uint64_t q_end = min((uint64_t)(tile + 1) * REMAP_ROWS, S.nq);
  for (uint64_t q = (uint64_t)tile * REMAP_ROWS + w; q < q_end; q += REMAP_WARPS) {
    const int32_t e = live[q];
    const uint64_t sq = (uint64_t)js * S.nq + q;
    if (e < 0) {
      for (int o = 0; o < O.n_out; ++o) {
        const FieldOutDev fo = O.f[o];
        uint32_t stride = fo.stride[js]; uint32_t* dst = fo.data + fo.base[js] + q * stride;
        int fi = o - (A.has_gt ? 1 : 0);
        uint32_t pad = (fi >= 0) ? vend_word(A.fm[fi].elem) : BCF_INT_VEND;
        for (uint32_t i = lane; i < stride; i += 32) dst[i] = pad;
        if (lane == 0) fo.len[sq] = 0;
      }
      if (O.lut_first && lane == 0) { O.lut_first[sq] = 0xFFFFFFFFu; O.lut_len[sq] = 0; }
      continue;
    }
    const CallView v = make_view(A, S, e);
    const uint32_t A_in = (uint32_t)v.n_alt + 1u;
    const bool trivial = !v.reduced && (v.flags & F_TRIVIAL);
    int in_nr = (nonref && v.nr_idx) ? v.nr_idx : -1;
    // ---- allele LUT (merge_alt_alleles' add_input_merged_idx_pair calls for this call, variant_operations.cc:160-227)
    if (remap && !trivial) {
      for (uint32_t i = lane; i < A_m; i += 32) m2in[i] = -1;
      for (uint32_t i = lane; i < A_in; i += 32) in2m[i] = -1;
      __syncwarp();
      if (lane == 0) {
        m2in[0] = 0; in2m[0] = 0;
        if (v.reduced) {
          int sfx; uint64_t h = allele_hash(c_str_star, 1, mref, site.ref_len, 1, sfx);
          for (uint32_t m = 1; m < A_m; ++m) if (s_ahash[m] == h) { in2m[1] = (int16_t)m; m2in[m] = 1; break; }
        } else {
          int pos = 0, ts, tl; uint32_t idx = 0;
          while (next_token(v.alt, v.alt_n, pos, ts, tl)) {
            ++idx;
            if (is_non_ref(v.alt + ts, tl)) continue;
            int sfx; uint64_t h = allele_hash(v.alt + ts, tl, mref, site.ref_len, v.ref_len, sfx);
            for (uint32_t m = 1; m < A_m; ++m) if (s_ahash[m] == h) { in2m[idx] = (int16_t)m; m2in[m] = (int16_t)idx; break; }
          }
        }
        if (v.nr_idx) { in2m[v.nr_idx] = (int16_t)(A_m - 1); m2in[A_m - 1] = (int16_t)v.nr_idx; }
      }
      __syncwarp();
    }
    auto M2IN = [&](uint32_t m) -> int {
      if (trivial) return m == 0 ? 0 : ((v.nr_idx && m == A_m - 1) ? v.nr_idx : -1);
      return m < A_m ? (int)m2in[m] : -1;
    };
    auto IN2M = [&](int32_t a) -> int {
      if (a < 0 || (uint32_t)a >= A_in) return -1;
      if (trivial) return a == 0 ? 0 : ((v.nr_idx && a == v.nr_idx) ? (int)A_m - 1 : -1);
      return (int)in2m[a];
    };
    if (O.lut_first) {
      uint32_t first = 0;
      if (lane == 0) { first = atomicAdd(&B.cursors[2], A_in); if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; } O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in; }
      first = __shfl_sync(0xffffffffu, first, 0);
      if (first != 0xFFFFFFFFu) for (uint32_t i = lane; i < A_in; i += 32) O.lut_in2m[first + i] = remap ? (int16_t)IN2M((int32_t)i) : (int16_t)(i == 0 ? 0 : ((v.nr_idx && (int)i == v.nr_idx) ? (int)A_m - 1 : -1));
    }
    int o = 0;
    // ---- GT (variant_operations.cc:233-263)
    if (A.has_gt) {
      const FieldOutDev fo = O.f[o++];
      uint32_t stride = fo.stride[js]; uint32_t* dst = fo.data + fo.base[js] + q * stride;
      int step = A.gt_phased ? 2 : 1;
      for (uint32_t i = lane; i < stride; i += 32) {
        uint32_t val = BCF_INT_VEND;
        if (i < v.gt_len) { uint32_t g = v.gt[i]; val = (remap && (i % step) == 0) ? remap_gt_word(g, IN2M, A_m, nonref) : g; }
        dst[i] = val;
      }
      if (lane == 0) fo.len[sq] = (uint16_t)v.gt_len;
    }
    // ---- allele-dependent fields
    for (int f = 0; f < A.n_fields; ++f) {
      const FieldOutDev fo = O.f[o++];
      const uint32_t stride = fo.stride[js]; uint32_t* dst = fo.data + fo.base[js] + q * stride;
      const int kind = A.fm[f].kind, elem = A.fm[f].elem;
      const uint32_t pad = vend_word(elem), miss = missing_word(elem);
      const uint32_t* in; uint32_t n_in;
      bool valid = field_of(A, S, v, f, in, n_in);
      if (valid && !v.reduced) {   // every element TileDB-NULL -> invalid
        bool any = false;
        for (uint32_t i = lane; i < n_in; i += 32) any |= (in[i] != null_word(elem));
        valid = __any_sync(0xffffffffu, any);
      }
      if (!valid || (kind == KIND_G && g_suppressed)) {
        for (uint32_t i = lane; i < stride; i += 32) dst[i] = pad;
        if (lane == 0) fo.len[sq] = 0;
        continue;
      }
      if (!remap) {
        for (uint32_t i = lane; i < stride; i += 32) dst[i] = (i < n_in) ? in[i] : pad;
        if (lane == 0) fo.len[sq] = (uint16_t)n_in;
        continue;
      }
      const uint32_t n_out = num_elements(kind, A_m - 1u, v.ploidy);
      if (kind != KIND_G || v.ploidy == 1) {
        const bool alt_only = (kind == KIND_A);
        for (uint32_t j = lane; j < stride; j += 32) {
          uint32_t val = pad;
          if (j < n_out) {
            int ij = M2IN(alt_only ? j + 1u : j);
            if (ij < 0) ij = in_nr;
            if (ij < 0) val = miss;
            else { int64_t ix = alt_only ? (int64_t)ij - 1 : ij; val = (ix < 0 || ix >= (int64_t)n_in) ? miss : in[ix]; }
          }
          dst[j] = val;
        }
      } else if (v.ploidy == 2) {
        for (uint32_t g = lane; g < stride; g += 32) {
          uint32_t val = pad;
          if (g < n_out) {
            // g = k(k+1)/2 + j, j <= k
            uint32_t k = (uint32_t)((sqrtf(8.0f * (float)g + 1.0f) - 1.0f) * 0.5f);
            while (k * (k + 1) / 2 > g) --k;
            while ((k + 1) * (k + 2) / 2 <= g) ++k;
            uint32_t jj = g - k * (k + 1) / 2;
            int ij = M2IN(jj), ik = M2IN(k);
            if (ij < 0) ij = in_nr;
            if (ik < 0) ik = in_nr;
            if (ij < 0 || ik < 0) val = miss;
            else { int64_t ig = alleles2gt(ij, ik); val = (ig >= (int64_t)n_in) ? miss : in[ig]; }
          }
          dst[g] = val;
        }
      } else {
        // ploidy 0 (copy of the first element) or general ploidy: sequential on one lane (rare)
        if (v.ploidy > MAX_PLOIDY) { if (lane == 0) atomicMax(&S.err[1], 3u); }
        else if (lane == 0) {
          for (uint32_t i = 0; i < n_out && i < stride; ++i) dst[i] = (i < n_in) ? in[i] : 0u;
          if (n_out <= stride) remap_field_seq(kind, in, n_in, M2IN, A_m, in_nr, v.ploidy, dst, miss);
        }
        __syncwarp();
        for (uint32_t i = n_out + lane; i < stride; i += 32) dst[i] = pad;
      }
      if (lane == 0) fo.len[sq] = (uint16_t)n_out;
    }
  }
}

}  // namespace gdb

// ================================================================================================ K5 tile kernels
// The fast path of the remap (sites with <= 64 merged alleles, ploidy <= 2, short GT, <= 4 remapped fields -- everything
// but pathological sites, which k_remap_generic keeps handling) is two kernels:
//   k_remap_tile  one block = 256 samples of one site; handles the dead rows and the trivial inline calls (reference
//                 block or <NON_REF>-only call whose GT/PL/AD fit the 32-byte hot record: ~90 % of the calls live at a
//                 site) and is a pure store stream:
//       phase A   one thread per sample: ONE 32-byte gather describes the call; the thread remaps GT, writes the lengths
//                 and leaves, per field, the (at most) three distinct values such a call contributes: genotype (0,0),
//                 (0,k), (j>0,k) -- plus the padding word.
//       phase B   per field, each lane owns one 16-byte quad of the dense sample-major output; a per-site class table
//                 says which of the four words (v0,v1,v2,pad) every output word takes -> coalesced STG.128.
//   k_remap_nt    one warp per entry of the sites' nt lists (every other live call): one more gather (hot record -> warm
//                 blob) brings the pre-hashed ALT tokens and the GT/field words; lanes look the tokens up in the site's
//                 allele hashes (merged<->input LUT in shared memory, NON_REF fallback folded in) and write the row's
//                 vectors lane-parallel.
// Field strides of remapped sites are multiples of 4 words so that every sample's vector is 16-byte aligned.
// Output bytes dominate (26.8 MB per site at 100 k samples x 10 alleles): the bound is HBM write bandwidth; no tensor
// cores, no TMA (inputs are 256 scattered 32-byte sectors per block, not a tile).
namespace gdb {

constexpr int TILE_ROWS = 256;
constexpr int TILE_MAX_A = 44;            // merged alleles (incl. REF) of a tile-path site: G(44) = 990 genotypes <= 256 quads
constexpr int TILE_MAX_GT = 8;
constexpr int TILE_MAX_FIELDS = 4;
constexpr int TILE_MAX_G = TILE_MAX_A * (TILE_MAX_A + 1) / 2;
constexpr uint32_t INFO_FAST = 1u;

// g = k(k+1)/2 + j, j <= k
__device__ inline void tri_unrank(uint32_t g, uint32_t& j, uint32_t& k) {
  k = (uint32_t)((sqrtf(8.0f * (float)g + 1.0f) - 1.0f) * 0.5f);
  while (k * (k + 1) / 2 > g) --k;
  while ((k + 1) * (k + 2) / 2 <= g) ++k;
  j = g - k * (k + 1) / 2;
}
// class of output word g of a trivial call's vector: 0 -> v0, 1 -> v1, 2 -> v2 (see phase A)
__device__ inline uint32_t trivial_class(int kind, uint32_t ploidy, uint32_t g) {
  if (kind == KIND_A) return 1;
  if (kind == KIND_G && ploidy == 2) { uint32_t j, k; tri_unrank(g, j, k); return k == 0 ? 0 : (j == 0 ? 1 : 2); }
  return g == 0 ? 0 : 1;   // R fields, haploid G fields; ploidy 0 G fields have one element
}
// num_elements for ploidy <= 2 (the only ploidies a tile-path site holds)
__device__ inline uint32_t num_elements_le2(int kind, uint32_t n_alt, uint32_t ploidy) {
  if (kind == KIND_A) return n_alt;
  if (kind == KIND_R || ploidy == 1) return n_alt + 1u;
  return ploidy == 2 ? ((n_alt + 1u) * (n_alt + 2u)) >> 1 : 1u;
}

// n / d for n, d <= 65535 without the integer-division sequence
__device__ inline uint32_t small_div(uint32_t n, uint32_t d) {
  uint32_t qt = (uint32_t)(__uint2float_rz(n) * __frcp_rn(__uint2float_rz(d)));
  if (qt * d > n) --qt;
  if ((qt + 1u) * d <= n) ++qt;
  return qt;
}
// Universal tables (built once per context, gdb_engine.cu): for genotype index g < 1024 of a diploid vector,
//   cls_tab[g / 4] byte g % 4 = class of word g (0: (0,0)  1: (0,k)  2: (j>0,k)),  jk_tab[g] = k << 8 | j.
struct TileTabs { const uint32_t* cls; const uint16_t* jk; };

// k_remap_nt: NT_GL lanes per nt-list entry (4: a diploid 10-allele PL row is 14 quads = 4 trips at 87 % lane use, AD / tokens / GT
// one trip; measured 3.9 -> 2.7 ms per step against 8 lanes)
// k_remap_nt: NT_GL lanes per nt-list entry (non-trivial, non-inline or spanning-deletion call), NT_GROUPS entries of ONE
// site per block (block -> (site, first entry) comes from a map built by k_nt_blocks), so a warp keeps four independent
// gather chains in flight: entry -> blob header -> tokens / GT / field words.
#ifndef GDB_NT_GL
#define GDB_NT_GL 4
#endif
constexpr int NT_THREADS = 256, NT_GL = GDB_NT_GL, NT_GROUPS = NT_THREADS / NT_GL;
__global__ void __launch_bounds__(256) k_nt_blocks(const __grid_constant__ BatchDev B, const uint32_t* __restrict__ site_list, const uint32_t* __restrict__ blk_prefix, uint2* __restrict__ blk_map) {
  const uint32_t si = blockIdx.x, js = site_list[si], b0 = blk_prefix[si], nb = blk_prefix[si + 1] - b0;
  for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) blk_map[b0 + b] = make_uint2(js, b * NT_GROUPS);
}
// everything an nt-list entry needs to know about its site
struct NtSite { uint32_t js, A_m, sflags, mref_len; const uint8_t* mref; };
// where nt_entry puts a row's field vectors: the site's block of the global output (tile[] unused), or the block's tile
// in shared memory (k_remap_tma: tile[f] = row q0's vector of field f; the tile leaves the SM as one bulk copy)
// blob_win / win_lo: a window of the warm-blob pool staged in shared memory, starting at 16-byte unit win_lo (blobs that begin
// at or after it lie entirely inside it; nullptr = none)
struct NtDst { bool smem; uint32_t q0; uint32_t* tile[4]; const uint32_t* blob_win; uint32_t win_lo; };

// One nt-list entry (non-trivial, non-inline or spanning-deletion call) handled by NT_GL consecutive lanes of a warp: gather of
// the warm blob (pre-hashed ALT tokens, GT / field words), allele LUT against the site's hashes, GT / field vectors written
// as 16-byte quads.  s_jk / s_ah: the site's genotype (j,k) table and merged-allele hashes in shared memory; m2in_row /
// in2m_row: this lane group's scratch rows.  Called with all lanes of the warp (inactive groups only take part in the
// warp barriers).
template <int NF, bool GT>
__device__ __forceinline__ void nt_entry(const ArrayDev& A, const ScanDev& S, const BatchDev& B, const RemapOut& O, const NtSite& st,
                                         const uint16_t* __restrict__ s_jk, const unsigned long long* __restrict__ s_ah, int* m2in_row, int8_t* in2m_row,
                                         const bool active, const uint2 ent, const int gl, const NtDst nd = NtDst{false, 0u, {nullptr, nullptr, nullptr, nullptr}, nullptr, 0u}) {
  const uint32_t js = st.js, A_m = st.A_m, mref_len = st.mref_len;
  const uint8_t* mref = st.mref;
  const bool nonref = st.sflags & 1u, remap = st.sflags & 2u, g_suppressed = st.sflags & 4u;
  constexpr int o_f0 = GT ? 1 : 0;
  for (uint32_t m = gl; m < A_m; m += NT_GL) m2in_row[m] = m == 0 ? 0 : -1;

  // ---- the call
  const uint32_t q = ent.x;
  const uint64_t sq = (uint64_t)js * S.nq + q;
  const bool reduced = (ent.y & NT_REDUCED) != 0;
  uint32_t n_alt = 0, nr_idx = 0, gt_len = 0, ploidy = 0;
  const uint32_t* gt = nullptr; const uint32_t* fin[NF]; uint32_t fn[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) { fin[f] = nullptr; fn[f] = 0; }
  WarmView wv{};
  const uint32_t H0 = warm_hash_word((uint32_t)A.n_fields);
  if (active) {
    if (reduced) {
      const uint32_t cell = ent.y & ~NT_REDUCED;
      const DelRec* dr = find_delrec(S, cell);
      const CellDesc d = A.desc[cell];
      n_alt = dr->n_alt; nr_idx = dr->has_nr ? 2u : 0u;
      gt_len = (d.flags & F_GTVALID) ? d.gt_len : 0u; ploidy = gt_len ? ploidy_of(A.gt_phased, gt_len) : 0u;
      gt = S.del_words + dr->gt_off;
#pragma unroll
      for (int f = 0; f < NF; ++f) { const uint32_t n = dr->f_len[f]; fn[f] = n == 0xFFFFFFFFu ? 0u : n; fin[f] = S.del_words + dr->f_off[f]; }
    } else {
      wv = warm_view(A, ent.y, (nd.blob_win && ent.y >= nd.win_lo) ? nd.blob_win + (size_t)(ent.y - nd.win_lo) * 4u : nullptr);
      n_alt = wv.n_alt; nr_idx = wv.nr_idx; gt_len = wv.gt_len; ploidy = wv.ploidy;
      gt = wv.b + H0 + 3u * n_alt + wv.tok_words;
      const uint32_t* p = gt + gt_len;
#pragma unroll
      for (int f = 0; f < NF; ++f) { fn[f] = wv.b[4 + f]; fin[f] = p; p += fn[f]; }
    }
  }
  const uint32_t A_in = n_alt + 1u;
  __syncwarp();   // m2in_row is initialised; the caller made the site's tables (s_ah, s_jk) visible to the block
  // ---- allele LUT (merge_alt_alleles' add_input_merged_idx_pair calls for this call, variant_operations.cc:160-227)
  if (active) {
    if (gl == 0) in2m_row[0] = 0;
    for (uint32_t t = gl; t < n_alt && t + 1u < (uint32_t)TILE_MAX_A; t += NT_GL) {
      const uint32_t idx = t + 1u;
      int m = -1;
      if (remap) {
        uint64_t h = 0;
        if (reduced) { if (idx == 1) h = star_hash(); }
        else h = warm_token_hash(wv, H0, t, mref, mref_len);
        if (h) for (uint32_t mm = 1; mm < A_m; ++mm) if (s_ah[mm] == h) { m = (int)mm; break; }
      }
      if (nr_idx && idx == nr_idx) m = (int)A_m - 1;
      in2m_row[idx] = (int8_t)m;
      if (remap && m >= 0) atomicMax(&m2in_row[m], (int)idx);   // a merged allele named twice by one call: the last one wins
    }
  }
  __syncwarp();
  if (active && remap) {
    const int in_nr = (nonref && nr_idx) ? (int)nr_idx : -1;   // NON_REF fallback folded in
    for (uint32_t m = gl; m < A_m; m += NT_GL) {
      int a = m2in_row[m];
      if (a < 0) a = in_nr;
      m2in_row[m] = a < 0 ? -1 : (a | ((a * (a + 1) / 2) << 8));
    }
  }
  __syncwarp();
  if (!active) return;
  const int* m2in = m2in_row;
  if (O.lut_first) {
    uint32_t first = 0;
    if (gl == 0) { first = atomicAdd(&B.cursors[2], A_in); if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; } O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in; }
    first = __shfl_sync(__activemask(), first, 0, NT_GL);
    if (first != 0xFFFFFFFFu) for (uint32_t i = gl; i < A_in; i += NT_GL) O.lut_in2m[first + i] = (int16_t)in2m_row[i];
  }
  // ---- GT (variant_operations.cc:233-263)
  if (GT) {
    const FieldOutDev fo = O.f[0];
    const uint32_t gs = fo.stride[js]; uint32_t* dst = fo.data + fo.base[js] + (uint64_t)q * gs;
    const uint32_t step = A.gt_phased ? 2u : 1u;
    for (uint32_t i = gl; i < gs; i += NT_GL) {
      uint32_t val = BCF_INT_VEND;
      if (i < gt_len) {
        const uint32_t g = gt[i]; val = g;
        if (remap && (i % step) == 0 && !(g == TDB_NULL_INT || g == 0xFFFFFFFFu || g == BCF_INT_MISSING)) {
          const int m = g < A_in ? (int)in2m_row[g] : -1;
          val = m < 0 ? (nonref ? A_m - 1u : 0xFFFFFFFFu) : (uint32_t)m;
        }
      }
      dst[i] = val;
    }
    if (gl == 0) fo.len[sq] = (uint16_t)gt_len;
  }
  // ---- allele-dependent fields (variant_field_handler.cc:41-191)
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const FieldOutDev fo = O.f[o_f0 + f];
    const int kind = A.fm[f].kind, elem = A.fm[f].elem;
    const uint32_t pad = vend_word(elem), miss = missing_word(elem);
    const uint32_t n = fn[f]; const uint32_t* in = fin[f];
    uint32_t n_out = 0;
    if (n && !(kind == KIND_G && g_suppressed)) n_out = remap ? num_elements(kind, A_m - 1u, ploidy) : n;
    const uint32_t stride = fo.stride[js];
    uint32_t* dst = nd.smem ? nd.tile[f] + (size_t)(q - nd.q0) * stride : fo.data + fo.base[js] + (uint64_t)q * stride;
    if (gl == 0) fo.len[sq] = (uint16_t)n_out;
    // s_m2in[m] = input allele of merged allele m (NON_REF fallback folded in), or -1; the diploid loop wants it packed
    // with its triangular number: a | a(a+1)/2 << 8 (still negative when there is no input allele)
    if (remap && (stride & 3u) == 0 && kind == KIND_G && ploidy == 2) {
      uint4* dst4 = reinterpret_cast<uint4*>(dst);
      for (uint32_t qd = gl; qd < (stride >> 2); qd += NT_GL) {
        const uint2 jk4 = *reinterpret_cast<const uint2*>(&s_jk[4 * qd]);   // (k << 8 | j) of words 4qd .. 4qd+3
        uint32_t w[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const uint32_t jk = ((x < 2 ? jk4.x : jk4.y) >> (16 * (x & 1))) & 0xFFFFu;
          const int pj = m2in[jk & 0xFF], pk = m2in[jk >> 8];
          const uint32_t aj = (uint32_t)pj & 0xFFu, ak = (uint32_t)pk & 0xFFu;
          const uint32_t ig = aj > ak ? ((uint32_t)pj >> 8) + ak : ((uint32_t)pk >> 8) + aj;
          uint32_t v = miss;
          if ((pj | pk) >= 0 && ig < n) v = in[ig];
          w[x] = (4 * qd + x < n_out) ? v : pad;
        }
        dst4[qd] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else if (remap && (stride & 3u) == 0 && kind == KIND_R) {
      // R field (AD): out[j] = in[input allele of merged allele j]
      uint4* dst4 = reinterpret_cast<uint4*>(dst);
      for (uint32_t qd = gl; qd < (stride >> 2); qd += NT_GL) {
        uint32_t w[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const uint32_t j = 4 * qd + x;
          uint32_t v = pad;
          if (j < n_out) { const int pj = m2in[j]; const uint32_t a = (uint32_t)pj & 0xFFu; v = (pj >= 0 && a < n) ? in[a] : miss; }
          w[x] = v;
        }
        dst4[qd] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      const int mode = !remap ? 0 : ((kind == KIND_G && ploidy == 2) ? 1 : ((kind == KIND_G && ploidy == 0) ? 2 : (kind == KIND_A ? 3 : 4)));
      auto word = [&](uint32_t g) -> uint32_t {
        if (g >= n_out) return pad;
        if (mode == 0) return in[g];
        if (mode == 1) {
          const uint32_t jk = s_jk[g];
          const int pj = m2in[jk & 0xFF], pk = m2in[jk >> 8];
          if ((pj | pk) < 0) return miss;
          const uint32_t aj = (uint32_t)pj & 0xFFu, ak = (uint32_t)pk & 0xFFu;
          const uint32_t ig = aj > ak ? ((uint32_t)pj >> 8) + ak : ((uint32_t)pk >> 8) + aj;
          return ig < n ? in[ig] : miss;
        }
        if (mode == 2) return in[0];
        const int pj = m2in[mode == 3 ? g + 1u : g];
        const int ij = pj < 0 ? -1 : (pj & 0xFF);
        const int ix = mode == 3 ? ij - 1 : ij;
        return (ij < 0 || ix < 0 || (uint32_t)ix >= n) ? miss : in[ix];
      };
      if (remap && (stride & 3u) == 0) {   // 16-byte aligned vector: one quad per lane
        uint4* dst4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t qd = gl; qd < (stride >> 2); qd += NT_GL) dst4[qd] = make_uint4(word(4 * qd), word(4 * qd + 1), word(4 * qd + 2), word(4 * qd + 3));
      } else {
        for (uint32_t g = gl; g < stride; g += NT_GL) dst[g] = word(g);
      }
    }
  }
}

// Sites that need no remapping (merged REF of length 1 and <NON_REF> the only ALT: the reference-block intervals between
// variant sites): every call's vectors go through unchanged.  One thread per sample; grid = (tiles, sites).
template <int NF, bool GT>
__global__ void __launch_bounds__(TILE_ROWS) k_copy_tile(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B,
                                                        const __grid_constant__ RemapOut O, const uint32_t* __restrict__ site_list) {
  __shared__ uint32_t s_hw[HOT_WORDS * TILE_ROWS];   // hot words as an indexable register file, word i of thread t at [i * TILE_ROWS + t]
  constexpr int o_f0 = GT ? 1 : 0;
  const uint32_t js = site_list[blockIdx.y], tid = threadIdx.x, nq = (uint32_t)S.nq;
  const uint32_t q = blockIdx.x * TILE_ROWS + tid;
  if (q >= nq) return;
  const uint64_t sq = (uint64_t)js * nq + q;
  const int32_t e = B.live[sq];
  uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
  if (e >= 0 && !(e & LIVE_REDUCED)) { const uint4* hp = reinterpret_cast<const uint4*>(&A.hot[e]); h0 = hp[0]; h1 = hp[1]; }
  const uint32_t A_m = B.sites[js].n_alleles;
  uint32_t* const hw = s_hw + tid;
  hw[0] = h0.w; hw[TILE_ROWS] = h1.x; hw[2 * TILE_ROWS] = h1.y; hw[3 * TILE_ROWS] = h1.z; hw[4 * TILE_ROWS] = h1.w;
  const uint32_t hflags = h0.x & 0xFF, n_alt = (h0.x >> 16) & 0xFF, nr = h0.x >> 24, gt_len = h0.y & 0xFF;
  const bool inl = ((h0.y >> 8) & 0xFF) != 0;
  if (e >= 0 && ((e & LIVE_REDUCED) || !inl || !(hflags & F_TRIVIAL))) return;   // on the site's nt list: k_remap_nt writes this row
  const bool dead = e < 0;
  if (GT) {
    const uint32_t gs = O.f[0].stride[js]; uint32_t* d = O.f[0].data + O.f[0].base[js] + (uint64_t)q * gs;
    if (gs == 2) *reinterpret_cast<uint2*>(d) = make_uint2(gt_len > 0 ? hw[0] : BCF_INT_VEND, gt_len > 1 ? hw[TILE_ROWS] : BCF_INT_VEND);
    else for (uint32_t i = 0; i < gs; ++i) d[i] = i < gt_len ? hw[i * TILE_ROWS] : BCF_INT_VEND;
    O.f[0].len[sq] = (uint16_t)gt_len;
  }
  if (O.lut_first) {
    if (dead) { O.lut_first[sq] = 0xFFFFFFFFu; O.lut_len[sq] = 0; }
    else {
      const uint32_t A_in = n_alt + 1u;
      uint32_t first = atomicAdd(&B.cursors[2], A_in);
      if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; }
      O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in;
      if (first != 0xFFFFFFFFu) for (uint32_t i = 0; i < A_in; ++i) O.lut_in2m[first + i] = (int16_t)(i == 0 ? 0 : ((nr && i == nr) ? (int)A_m - 1 : -1));
    }
  }
  const uint32_t fl4 = (h0.y >> 16) | (h0.z << 16);
  uint32_t k = gt_len;   // fields follow GT in the record
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const uint32_t n = (fl4 >> (8 * f)) & 0xFF;   // 0 for a dead row
    const uint32_t padw = vend_word(A.fm[f].elem);
    const FieldOutDev fo = O.f[o_f0 + f];
    const uint32_t stride = fo.stride[js];
    fo.len[sq] = (uint16_t)n;
    const uint32_t* fw = hw + k * TILE_ROWS;
    uint32_t* d = fo.data + fo.base[js] + (uint64_t)q * stride;
    for (uint32_t x = 0; x < stride; ++x) d[x] = x < n ? fw[x * TILE_ROWS] : padw;
    k += n;
  }
}

// Remapped sites.  One block = 256 samples of one site; grid = (tiles, sites).
#ifndef GDB_TILE_BLOCKS
#define GDB_TILE_BLOCKS 5
#endif
template <int NF, bool GT>
__global__ void __launch_bounds__(TILE_ROWS, GDB_TILE_BLOCKS) k_remap_tile(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B,
                                                            const __grid_constant__ RemapOut O, const uint32_t* __restrict__ site_list, const TileTabs tabs) {
  __shared__ __align__(16) uint4 s_tri[TILE_ROWS * NF];     // per (sample, field): v0, v1, v2, pad
  __shared__ uint32_t s_info[TILE_ROWS * NF];               // per (sample, field): INFO_FAST or 0
  __shared__ uint32_t s_hw[HOT_WORDS * TILE_ROWS];          // the samples' hot words, word i of thread t at [i * TILE_ROWS + t]:
                                                            // an indexable register file (each thread reads only its own column)
  // phase C (the tile's non-trivial calls, see nt_entry): their list and the per-group scratch
  __shared__ uint2 s_ntl[TILE_ROWS];
  __shared__ uint32_t s_nnt;
  __shared__ __align__(8) uint16_t s_jk[TILE_MAX_G + 6];
  __shared__ unsigned long long s_ah[TILE_MAX_A];
  __shared__ int s_m2in[NT_GROUPS][TILE_MAX_A];
  __shared__ int8_t s_in2m[NT_GROUPS][TILE_MAX_A];
  constexpr int o_f0 = GT ? 1 : 0;
  const uint32_t js = site_list[blockIdx.y], tile = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t nq = (uint32_t)S.nq;
  const uint32_t q = tile * TILE_ROWS + tid;
  const uint64_t sq = (uint64_t)js * nq + q;
  if (tid == 0) s_nnt = 0;
  __syncthreads();
  // issue the two dependent gathers of this thread's sample first; everything below overlaps their latency
  int32_t e = -1;
  uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
  if (q < nq) {
    e = B.live[sq];
    if (e >= 0 && !(e & LIVE_REDUCED)) { const uint4* hp = reinterpret_cast<const uint4*>(&A.hot[e]); h0 = hp[0]; h1 = hp[1]; }
  }
  const SiteRec* __restrict__ site_g = &B.sites[js];
  const uint32_t A_m = site_g->n_alleles, sflags = site_g->flags;
  const bool nonref = sflags & 1u, g_suppressed = sflags & 4u;
  uint32_t fstride[NF];
  uint32_t* fdst[NF];   // this site's block of field f
#pragma unroll
  for (int f = 0; f < NF; ++f) { fstride[f] = O.f[o_f0 + f].stride[js]; fdst[f] = O.f[o_f0 + f].data + O.f[o_f0 + f].base[js]; }
  uint32_t* const hw = s_hw + tid;
  hw[0] = h0.w; hw[TILE_ROWS] = h1.x; hw[2 * TILE_ROWS] = h1.y; hw[3 * TILE_ROWS] = h1.z; hw[4 * TILE_ROWS] = h1.w;

  // ---------------------------------------------------------------- phase A: one thread per sample
  if (q < nq) {
    if (e < 0) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const uint32_t padw = vend_word(A.fm[f].elem);
        s_tri[tid * NF + f] = make_uint4(padw, padw, padw, padw);
        s_info[tid * NF + f] = INFO_FAST;
        O.f[o_f0 + f].len[sq] = 0;
      }
      if (GT) { const uint32_t gs = O.f[0].stride[js]; uint32_t* d = O.f[0].data + O.f[0].base[js] + (uint64_t)q * gs; for (uint32_t i = 0; i < gs; ++i) d[i] = BCF_INT_VEND; O.f[0].len[sq] = 0; }
      if (O.lut_first) { O.lut_first[sq] = 0xFFFFFFFFu; O.lut_len[sq] = 0; }
    } else {
      const uint32_t hflags = h0.x & 0xFF, ploidy = (h0.x >> 8) & 0xFF, n_alt = (h0.x >> 16) & 0xFF, nr = h0.x >> 24, gt_len = h0.y & 0xFF;
      const bool inl = ((h0.y >> 8) & 0xFF) != 0;
      if ((e & LIVE_REDUCED) || !inl || !(hflags & F_TRIVIAL)) {
        // a non-trivial call: phase C of this block writes the row
#pragma unroll
        for (int f = 0; f < NF; ++f) s_info[tid * NF + f] = 0;
        s_ntl[atomicAdd(&s_nnt, 1u)] = make_uint2(q, (e & LIVE_REDUCED) ? (NT_REDUCED | (uint32_t)(e & (LIVE_REDUCED - 1))) : h0.w);
        if (!(e & LIVE_REDUCED)) {   // pull the call's warm blob (h1.x words at 16-byte unit h0.w) into L2 while phase B streams
          const char* blob = reinterpret_cast<const char*>(A.warm) + (uint64_t)h0.w * 16u;
          for (uint32_t off = 0; off < h1.x * 4u; off += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(blob + off));
        }
      } else {
        // ---- trivial call, fully described by its hot record
        const uint32_t fl4 = (h0.y >> 16) | (h0.z << 16);
        const int in_nr = (nonref && nr) ? (int)nr : -1;
        if (GT) {
          const uint32_t gs = O.f[0].stride[js]; uint32_t* d = O.f[0].data + O.f[0].base[js] + (uint64_t)q * gs;
          const uint32_t step = A.gt_phased ? 2u : 1u;
          auto gt_word = [&](uint32_t i) -> uint32_t {
            if (i >= gt_len) return BCF_INT_VEND;
            uint32_t g = hw[i * TILE_ROWS];
            if ((i % step) == 0 && !(g == TDB_NULL_INT || g == 0xFFFFFFFFu || g == BCF_INT_MISSING))
              g = (g == 0) ? 0u : ((nr && g == nr) ? A_m - 1u : (nonref ? A_m - 1u : 0xFFFFFFFFu));
            return g;
          };
          if (gs == 2) *reinterpret_cast<uint2*>(d) = make_uint2(gt_word(0), gt_word(1));   // site blocks are 16-byte aligned
          else for (uint32_t i = 0; i < gs; ++i) d[i] = gt_word(i);
          O.f[0].len[sq] = (uint16_t)gt_len;
        }
        if (O.lut_first) {
          const uint32_t A_in = n_alt + 1u;
          uint32_t first = atomicAdd(&B.cursors[2], A_in);
          if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; }
          O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in;
          if (first != 0xFFFFFFFFu) for (uint32_t i = 0; i < A_in; ++i) O.lut_in2m[first + i] = (int16_t)(i == 0 ? 0 : ((nr && i == nr) ? (int)A_m - 1 : -1));
        }
        uint32_t k = gt_len;   // fields follow GT in the record
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const uint32_t n = (fl4 >> (8 * f)) & 0xFF;
          const int kind = A.fm[f].kind;
          const uint32_t miss = missing_word(A.fm[f].elem), padw = vend_word(A.fm[f].elem);
          uint32_t n_out = 0;
          if (n && !(kind == KIND_G && g_suppressed)) n_out = num_elements_le2(kind, A_m - 1u, ploidy);
          O.f[o_f0 + f].len[sq] = (uint16_t)n_out;
          const uint32_t* fw = hw + k * TILE_ROWS;   // this field's inline words
          if (!n_out) {
            s_tri[tid * NF + f] = make_uint4(padw, padw, padw, padw);
            s_info[tid * NF + f] = INFO_FAST;
          } else {
            uint32_t v0 = fw[0], v1 = miss, v2 = miss;
            if (in_nr >= 0) {
              if (kind == KIND_A) { const uint32_t i1 = (uint32_t)in_nr - 1u; if (i1 < n) v1 = fw[i1 * TILE_ROWS]; }
              else if (kind == KIND_R || ploidy != 2) { if ((uint32_t)in_nr < n) v1 = fw[in_nr * TILE_ROWS]; }
              else {
                const uint32_t g1 = (uint32_t)in_nr * ((uint32_t)in_nr + 1u) / 2u, g2 = g1 + (uint32_t)in_nr;
                if (g1 < n) v1 = fw[g1 * TILE_ROWS];
                if (g2 < n) v2 = fw[g2 * TILE_ROWS];
              }
            }
            // the class pattern of phase B is the one of the site's ploidy for this field; a sample of another ploidy
            // (a haploid call among diploid ones ...) writes its own, shorter vector from this thread
            const uint32_t fploidy = site_g->max_ploidy[f];
            const bool fast = (kind != KIND_G) || (ploidy == fploidy);
            s_tri[tid * NF + f] = make_uint4(v0, v1, v2, padw);
            s_info[tid * NF + f] = fast ? INFO_FAST : 0u;
            if (!fast) {
              uint32_t* d = fdst[f] + (uint64_t)q * fstride[f];
              for (uint32_t x = 0; x < fstride[f]; ++x) {
                uint32_t w = padw;
                if (x < n_out) { const uint32_t c = trivial_class(kind, ploidy, x); w = c == 0 ? v0 : (c == 1 ? v1 : v2); }
                d[x] = w;
              }
            }
          }
          k += n;
        }
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase B: the block streams the tile out
  // Field f of the tile is rows x lpr quads, contiguous.  Thread t keeps ONE quad column (t % lpr) and walks the rows
  // (t / lpr, + 256 / lpr per step), so which of (v0, v1, v2, pad) each of its four words takes is loop-invariant, and
  // a warp's stores cover 512 contiguous bytes.  Two rows per trip, loads first, stores predicated.
  const uint32_t rows = min((uint32_t)TILE_ROWS, nq - tile * TILE_ROWS);
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const uint32_t stride = fstride[f];
    if (stride == 0) continue;
    const uint32_t lpr = stride >> 2;   // quads per sample (stride is a multiple of 4 at remapped sites), <= 256
    const uint32_t rpi = small_div((uint32_t)TILE_ROWS, lpr);   // rows per step
    const uint32_t r_off = small_div(tid, lpr), col = tid - r_off * lpr;
    if (r_off >= rpi) continue;
    const int kind = A.fm[f].kind;
    const uint32_t fploidy = site_g->max_ploidy[f];
    const uint32_t nfull = num_elements_le2(kind, A_m - 1u, fploidy);   // length of a full vector of this field at this site
    const uint32_t cw = (kind == KIND_A) ? 0x01010101u : ((kind == KIND_G && fploidy == 2) ? tabs.cls[col] : (col == 0 ? 0x01010100u : 0x01010101u));
    uint32_t c[4];
#pragma unroll
    for (uint32_t x = 0; x < 4; ++x) c[x] = (4 * col + x < nfull) ? ((cw >> (8 * x)) & 0xFFu) : 3u;
    uint4* out4 = reinterpret_cast<uint4*>(fdst[f] + (uint64_t)tile * TILE_ROWS * stride) + col;
    const uint32_t* tri = reinterpret_cast<const uint32_t*>(s_tri);
    for (uint32_t r = r_off; r < rows; r += 2 * rpi) {
      const uint32_t rb = min(r + rpi, (uint32_t)TILE_ROWS - 1u);
      const uint32_t sa = r * NF + f, sb = rb * NF + f;
      const uint32_t ia = s_info[sa], ib = s_info[sb];
      const uint32_t* ta = tri + 4 * sa; const uint32_t* tb = tri + 4 * sb;
      const uint4 va = make_uint4(ta[c[0]], ta[c[1]], ta[c[2]], ta[c[3]]);
      const uint4 vb = make_uint4(tb[c[0]], tb[c[1]], tb[c[2]], tb[c[3]]);
      if (ia & INFO_FAST) out4[(uint64_t)r * lpr] = va;
      if ((ib & INFO_FAST) && r + rpi < rows) out4[(uint64_t)(r + rpi) * lpr] = vb;
    }
  }

  // ---------------------------------------------------------------- phase C: the tile's non-trivial calls
  // ~10 % of the rows: instruction-heavy (LUT construction, genotype index arithmetic) where phases A/B are a store stream.
  // Doing them inside the same block lets their instructions fill the issue slots the store stream leaves idle on this
  // SM (other resident blocks are in their phase B) instead of running as a separate, issue-bound kernel afterwards.
  const uint32_t n_ntl = s_nnt;   // final since the barrier after phase A
  if (n_ntl) {
    if (tid < A_m) s_ah[tid] = B.allele_hash[site_g->allele_first + tid];
    const uint32_t jk_n = (A_m * (A_m + 1u)) >> 1;
    for (uint32_t g = tid; g < ((jk_n + 3u) & ~3u) && g < (uint32_t)TILE_MAX_G + 6u; g += TILE_ROWS) s_jk[g] = tabs.jk[g];
    __syncthreads();
    const NtSite st{js, A_m, sflags, site_g->ref_len, B.allele_bytes + site_g->ref_off};
    const int grp = (int)tid / NT_GL, gl = (int)tid % NT_GL;
    for (uint32_t base = 0; base < n_ntl; base += NT_GROUPS) {
      if (base + (tid & ~31u) / NT_GL >= n_ntl) break;   // none of this warp's groups has an entry (warp-uniform)
      const bool active = base + grp < n_ntl;
      const uint2 ent = active ? s_ntl[base + grp] : make_uint2(0u, 0u);
      nt_entry<NF, GT>(A, S, B, O, st, s_jk, s_ah, s_m2in[grp], s_in2m[grp], active, ent, gl);
      __syncwarp();
    }
  }
}

template <int NF, bool GT>
__global__ void __launch_bounds__(NT_THREADS) k_remap_nt(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B,
                                                        const __grid_constant__ RemapOut O, const uint2* __restrict__ blk_map, const uint32_t jk_n, const TileTabs tabs) {
  __shared__ __align__(8) uint16_t s_jk[TILE_MAX_G + 6];     // g -> (k << 8 | j), the same for every site
  __shared__ unsigned long long s_ah[TILE_MAX_A];            // the site's merged-allele hashes
  __shared__ int s_m2in[NT_GROUPS][TILE_MAX_A];
  __shared__ int8_t s_in2m[NT_GROUPS][TILE_MAX_A];
  const int tid = threadIdx.x, grp = tid / NT_GL, gl = tid % NT_GL;
  const uint2 bm = blk_map[blockIdx.x];
  const SiteRec* __restrict__ site_g = &B.sites[bm.x];
  const NtSite st{bm.x, site_g->n_alleles, site_g->flags, site_g->ref_len, B.allele_bytes + site_g->ref_off};

  const uint32_t ei = bm.y + grp;
  const bool active = ei < site_g->n_nt;
  uint2 ent = make_uint2(0, 0);
  if (active) ent = B.nt[(uint64_t)st.js * S.nq + ei];
  if ((uint32_t)tid < st.A_m) s_ah[tid] = B.allele_hash[site_g->allele_first + tid];
  for (uint32_t g = tid; g < ((jk_n + 3u) & ~3u) && g < (uint32_t)TILE_MAX_G + 6u; g += NT_THREADS) s_jk[g] = tabs.jk[g];   // whole quads: the tail is read, then discarded
  __syncthreads();
  nt_entry<NF, GT>(A, S, B, O, st, s_jk, s_ah, s_m2in[grp], s_in2m[grp], active, ent, gl);
}

}  // namespace gdb
