// CUDA kernels of the variant-combine hot path (sm_100a).  One kernel (family) per stage of SURVEY.md section 2:
//   K1 fragment merge      k_merge_*            (array_read_state.cc:1836-2132, newest fragment wins)
//   K2 classify / filter   k_classify, k_effend (genomicsdb_iterators.cc:535-636, query_variants.cc:422-428,1014-1117)
//   K2b deletion pre-pass  k_del_*              (broad_combined_gvcf.cc:912-1118)
//   K3 interval sweep      k_left_sweep, k_mark_events, k_extract_bp, k_deltas, k_site_counts, k_fill_sites, k_live_walk
//                                               (query_variants.cc:296-332,334-555,845-941)
//   K4 allele union        k_union              (variant_operations.cc:73-228,362-378)
//   K5 remap               k_remap              (variant_operations.cc:233-263,572-695; variant_field_handler.cc:41-398)
// All of it is integer indexing and byte moves; the roofline that bounds it is HBM bandwidth.
#pragma once
#include "gdb_device.cuh"

namespace gdb {

// ------------------------------------------------------------------------------------------------ scan primitive
// Three-launch exclusive scan (block scan, scan of block sums, add-back).  T is uint32_t / int32_t / uint64_t.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T>
__device__ inline T warp_incl_scan(T v) {
  for (int d = 1; d < 32; d <<= 1) { T o = __shfl_up_sync(0xffffffffu, v, d); if ((threadIdx.x & 31) >= d) v += o; }
  return v;
}
template <class TI, class TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const TI* __restrict__ in, TO* __restrict__ out, TO* __restrict__ tile_sums, uint64_t n) {
  __shared__ TO warp_tot[SCAN_THREADS / 32];
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  TO v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? (TO)in[base + i] : (TO)0; sum += v[i]; }
  TO incl = warp_incl_scan(sum);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) { TO t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : (TO)0; TO ti = warp_incl_scan(t); if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t; }
  __syncthreads();
  TO run = warp_tot[w] + incl - sum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}
template <class TO>
__global__ void __launch_bounds__(1024) k_scan_sums(TO* __restrict__ tile_sums, uint64_t n_tiles, TO* __restrict__ total) {
  __shared__ TO warp_tot[32];
  __shared__ TO carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_tiles; base += 1024) {
    uint64_t i = base + threadIdx.x;
    TO v = (i < n_tiles) ? tile_sums[i] : (TO)0;
    TO incl = warp_incl_scan(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) { TO t = warp_tot[lane]; TO ti = warp_incl_scan(t); warp_tot[lane] = ti - t; }
    __syncthreads();
    TO excl = carry + warp_tot[w] + incl - v;
    if (i < n_tiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
template <class TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(TO* __restrict__ out, const TO* __restrict__ tile_sums, uint64_t n) {
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  TO add = tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) out[base + i] += add;
}

// ------------------------------------------------------------------------------------------------ K2 classify
// One thread per cell: what gt_fill_row derives from a cell (query_variants.cc:1014-1117) plus END-copy detection
// (query_variants.cc:422-428), computed once per staged array.
__global__ void __launch_bounds__(256) k_classify(ArrayDev A, CellDesc* __restrict__ desc, HotRec* __restrict__ hot, uint32_t* __restrict__ warm_units) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int64_t row = A.coords[2 * i], col = A.coords[2 * i + 1], END = A.end[i];
  CellDesc d;
  d.flags = 0; d.pad = 0;
  if (col > END) d.flags |= F_ENDCOPY;
  if (row < A.lb_row || row >= A.lb_row + A.num_rows) d.flags |= F_ROWOUT;
  uint64_t r0 = A.ref.off[i], r1 = A.ref.off[i + 1];
  d.ref_len = (uint32_t)(r1 - r0);
  const uint8_t* alt = A.alt.val + A.alt.off[i];
  int an = alt_effective_len(alt, (int)(A.alt.off[i + 1] - A.alt.off[i]));
  int pos = 0, ts, tl, n_alt = 0, nr = 0;
  uint32_t tok_bytes = 0;
  bool deletion = false, first_is_nr = false;
  while (next_token(alt, an, pos, ts, tl)) {
    ++n_alt; tok_bytes += (uint32_t)tl;
    if (is_non_ref(alt + ts, tl)) { nr = n_alt; if (n_alt == 1) first_is_nr = true; }
    else if (d.ref_len > 1 && !is_symbolic(alt + ts, tl) && (uint32_t)tl < d.ref_len) deletion = true;
  }
  d.n_alt = (uint16_t)(n_alt > 0xFFFF ? 0xFFFF : n_alt);
  d.nr_idx = (uint16_t)nr;
  if (deletion) d.flags |= F_DELETION;
  if (d.ref_len == 1 && n_alt == 1 && first_is_nr) d.flags |= F_REFBLOCK;
  if (n_alt == 0 || (n_alt == 1 && first_is_nr)) d.flags |= F_TRIVIAL;   // allele LUT is implied: REF->REF, NON_REF->last
  d.gt_len = 0; d.ploidy = 0;
  if (A.has_gt) {
    uint64_t g0 = A.gt.off[i], g1 = A.gt.off[i + 1];
    uint32_t n = (uint32_t)((g1 - g0) >> 2);
    const uint32_t* g = reinterpret_cast<const uint32_t*>(A.gt.val + g0);
    bool all_null = true;
    for (uint32_t k = 0; k < n; ++k) if (g[k] != TDB_NULL_INT) { all_null = false; break; }
    if (!all_null) { d.flags |= F_GTVALID; d.gt_len = n; uint32_t p = ploidy_of(A.gt_phased, n); d.ploidy = (uint8_t)(p > 255 ? 255 : p); }
  }
  desc[i] = d;
  // hot record
  HotRec h;
  h.flags = d.flags; h.ploidy = d.ploidy; h.n_alt = (uint8_t)(n_alt > 255 ? 255 : n_alt); h.nr_idx = (uint8_t)(nr > 255 ? 255 : nr);
  h.gt_len = (uint8_t)(d.gt_len > 255 ? 255 : d.gt_len); h.inl = 0;
  h.ref_len8 = (uint8_t)(d.ref_len > 255 ? 255 : d.ref_len); h.ref0 = d.ref_len ? A.ref.val[r0] : (uint8_t)0;
  for (int f = 0; f < HOT_FIELDS; ++f) h.flen[f] = 0;
  for (int k = 0; k < HOT_WORDS; ++k) h.w[k] = 0;
  bool fits = (n_alt < 255) && (d.gt_len <= 255) && (A.n_fields <= HOT_FIELDS);
  uint32_t used = d.gt_len, field_words = 0;
  uint32_t flen[HOT_FIELDS] = {0, 0, 0, 0};
  for (int f = 0; f < A.n_fields; ++f) {
    uint64_t o0 = A.f[f].off[i], o1 = A.f[f].off[i + 1];
    uint32_t n = (uint32_t)((o1 - o0) >> 2);
    const uint32_t* v = reinterpret_cast<const uint32_t*>(A.f[f].val + o0);
    uint32_t nullw = A.fm[f].elem ? TDB_NULL_FLT : TDB_NULL_INT;
    bool all_null = true;
    for (uint32_t k = 0; k < n; ++k) if (v[k] != nullw) { all_null = false; break; }
    if (all_null) n = 0;     // invalid field (variant_field_data.h:365-384)
    if (f < HOT_FIELDS) flen[f] = n;
    used += n; field_words += n;
  }
  // which fields the call carries is recorded for every cell (saturating; the words themselves only when they fit): the remap
  // kernels know the output lengths of a call from this record alone
  for (int f = 0; f < A.n_fields && f < HOT_FIELDS; ++f) h.flen[f] = (uint8_t)(flen[f] > 255u ? 255u : flen[f]);
  if (fits && used <= HOT_WORDS) {
    h.inl = 1;
    uint32_t k = 0;
    if (d.gt_len) { const uint32_t* g = reinterpret_cast<const uint32_t*>(A.gt.val + A.gt.off[i]); for (uint32_t x = 0; x < d.gt_len; ++x) h.w[k++] = g[x]; }
    for (int f = 0; f < A.n_fields; ++f) {
      h.flen[f] = (uint8_t)flen[f];
      const uint32_t* v = reinterpret_cast<const uint32_t*>(A.f[f].val + A.f[f].off[i]);
      for (uint32_t x = 0; x < flen[f]; ++x) h.w[k++] = v[x];
    }
  }
  // every other cell gets a warm blob (filled by k_warm_fill once the pool offsets are known); 16-byte units
  uint32_t units = 0;
  if (!(h.inl && (d.flags & F_TRIVIAL))) {
    uint64_t words = (uint64_t)warm_hash_word((uint32_t)A.n_fields) + 3ull * (uint64_t)n_alt + ((tok_bytes + 3u) >> 2) + d.gt_len + field_words;
    units = (uint32_t)((words + 3) >> 2);
    h.w[1] = (uint32_t)words;
  }
  warm_units[i] = units;
  hot[i] = h;
}

// fills the warm blobs (layout: gdb_device.cuh); warm_off = exclusive scan of warm_units
__global__ void __launch_bounds__(256) k_warm_fill(ArrayDev A, const CellDesc* __restrict__ desc, HotRec* __restrict__ hot, const uint32_t* __restrict__ warm_units,
                                                    const uint32_t* __restrict__ warm_off, uint32_t* __restrict__ pool, uint32_t* __restrict__ err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n || warm_units[i] == 0) return;
  const CellDesc d = desc[i];
  uint32_t* b = pool + (uint64_t)warm_off[i] * 4u;
  hot[i].w[0] = warm_off[i];
  const uint8_t* alt = A.alt.val + A.alt.off[i];
  int an = alt_effective_len(alt, (int)(A.alt.off[i + 1] - A.alt.off[i]));
  uint32_t n_alt = 0;
  { int pos = 0, ts, tl; while (next_token(alt, an, pos, ts, tl)) ++n_alt; }
  b[0] = n_alt; b[1] = d.nr_idx; b[2] = d.ref_len;
  const uint32_t H0 = warm_hash_word((uint32_t)A.n_fields);
  if ((4u + (uint32_t)A.n_fields) < H0) b[H0 - 1] = 0;
  uint32_t* tok = b + H0 + 2u * n_alt;
  {
    int pos = 0, ts, tl; uint32_t t = 0;
    while (next_token(alt, an, pos, ts, tl)) {
      const bool nrt = is_non_ref(alt + ts, tl);
      uint64_t hs = nrt ? 0ull : fnv1a(FNV_BASIS, alt + ts, tl);
      b[H0 + 2u * t] = (uint32_t)hs; b[H0 + 2u * t + 1] = (uint32_t)(hs >> 32);
      tok[t] = (uint32_t)tl | (is_symbolic(alt + ts, tl) ? WARM_TOK_SYMBOLIC : 0u) | (nrt ? WARM_TOK_NONREF : 0u);
      ++t;
    }
  }
  uint32_t* w = tok + n_alt;
  {   // token bytes, packed
    uint8_t* tb = reinterpret_cast<uint8_t*>(w);
    int pos = 0, ts, tl; uint32_t nb = 0;
    while (next_token(alt, an, pos, ts, tl)) { for (int x = 0; x < tl; ++x) tb[nb + x] = alt[ts + x]; nb += (uint32_t)tl; }
    const uint32_t tw = (nb + 3u) >> 2;
    for (uint32_t x = nb; x < 4u * tw; ++x) tb[x] = 0;
    if (d.gt_len > 0xFFFFu || tw > 0xFFFFu) atomicMax(err, 1u);   // engine limit: GT vector / ALT string of one cell below 2^16 words
    b[3] = (d.gt_len & 0xFFFFu) | (tw << 16);
    w += tw;
  }
  if (d.gt_len) { const uint32_t* g = reinterpret_cast<const uint32_t*>(A.gt.val + A.gt.off[i]); for (uint32_t x = 0; x < d.gt_len; ++x) *w++ = g[x]; }
  for (int f = 0; f < A.n_fields; ++f) {
    uint64_t o0 = A.f[f].off[i], o1 = A.f[f].off[i + 1];
    uint32_t n = (uint32_t)((o1 - o0) >> 2);
    const uint32_t* v = reinterpret_cast<const uint32_t*>(A.f[f].val + o0);
    uint32_t nullw = A.fm[f].elem ? TDB_NULL_FLT : TDB_NULL_INT;
    bool all_null = true;
    for (uint32_t k = 0; k < n; ++k) if (v[k] != nullw) { all_null = false; break; }
    if (all_null) n = 0;
    b[4 + f] = n;
    for (uint32_t x = 0; x < n; ++x) *w++ = v[x];
  }
  const uint32_t words = hot[i].w[1];
  for (uint32_t* e = b + (((uint64_t)words + 3) & ~3ull); w < e; ++w) *w = 0;
}

// class word per cell (layout: gdb_device.cuh), once the hot records are final
__global__ void __launch_bounds__(256) k_cls_fill(const HotRec* __restrict__ hot, uint64_t n, uint32_t* __restrict__ cls) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HotRec h = hot[i];
  cls[i] = (h.inl && (h.flags & F_TRIVIAL)) ? cls_pack(h.nr_idx, h.n_alt, h.gt_len, h.ploidy, h.flen) : (CLS_GENERAL | (h.w[0] & CLS_LOW));
}

__global__ void __launch_bounds__(256) k_row_keys(ArrayDev A, const CellDesc* __restrict__ desc, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  keys[i] = (desc[i].flags & F_ROWOUT) ? (uint32_t)A.num_rows : (uint32_t)(A.coords[2 * i] - A.lb_row);
  vals[i] = (uint32_t)i;
}
// row_start[r] = first position in the row-sorted list whose key >= r
__global__ void __launch_bounds__(256) k_row_bounds(const uint32_t* __restrict__ sorted_keys, uint64_t n, int64_t num_rows, uint64_t* __restrict__ row_start) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > num_rows) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sorted_keys[mid] < (uint32_t)r) lo = mid + 1; else hi = mid; }
  row_start[r] = lo;
}
// rs_cell entries carry the cell's "contains a deletion" flag in bit 31 (cells < 2^30), so the walk needs no CellDesc gather
constexpr uint32_t RS_CELL_MASK = 0x7FFFFFFFu;
__global__ void __launch_bounds__(256) k_rs_flag_del(const CellDesc* __restrict__ desc, uint32_t* __restrict__ rs_cell, uint64_t n_rs) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_rs && (desc[rs_cell[k]].flags & F_DELETION)) rs_cell[k] |= 0x80000000u;
}
__global__ void __launch_bounds__(256) k_rs_cols(ArrayDev A, const uint32_t* __restrict__ rs_cell, uint64_t n_rs, int64_t* __restrict__ rs_col) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_rs) rs_col[k] = A.coords[2 * (uint64_t)rs_cell[k] + 1];
}
// Effective end of every begin cell: a later cell of the same sample that begins while the call is still live evicts it
// (query_variants.cc:519-543).  stats: [0] #begin cells, [1] #bad evictions, [2] max END (as int64), [3] #deletion cells
__global__ void __launch_bounds__(256) k_effend(ArrayDev A, CellDesc* __restrict__ desc, const uint32_t* __restrict__ rs_cell, const int64_t* __restrict__ rs_col,
                                                 const uint32_t* __restrict__ sorted_keys, uint64_t n_rs, int64_t* __restrict__ rs_effend,
                                                 int64_t* __restrict__ effend, unsigned long long* __restrict__ stats,
                                                 uint32_t* __restrict__ bad_cell, int64_t* __restrict__ bad_pred_col, uint32_t bad_cap) {
  __shared__ unsigned int s_begin, s_del;
  __shared__ long long s_max;
  if (threadIdx.x == 0) { s_begin = 0; s_del = 0; s_max = INT64_MIN; }
  __syncthreads();
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_rs) {
    uint32_t c = rs_cell[k];
    uint8_t fl = desc[c].flags;
    if (fl & F_ENDCOPY) { rs_effend[k] = I64_MIN; effend[c] = I64_MIN; atomicMax(&s_max, (long long)rs_col[k]); }
    else {
      int64_t END = A.end[c];
      int64_t eff = END;
      uint32_t key = sorted_keys[k];
      for (uint64_t k2 = k + 1; k2 < n_rs && sorted_keys[k2] == key; ++k2) {
        uint32_t c2 = rs_cell[k2];
        if (desc[c2].flags & F_ENDCOPY) continue;
        int64_t nb = rs_col[k2];
        if (nb <= END) {
          eff = nb - 1;
          if (!(fl & (F_DELETION | F_REFBLOCK))) {
            unsigned long long slot = atomicAdd(&stats[1], 1ull);
            if (slot < bad_cap) { bad_cell[slot] = c2; bad_pred_col[slot] = rs_col[k]; }
          }
        }
        break;
      }
      rs_effend[k] = eff; effend[c] = eff;
      atomicAdd(&s_begin, 1u);
      atomicMax(&s_max, (long long)END);
      if (fl & F_DELETION) atomicAdd(&s_del, 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_begin) atomicAdd(&stats[0], (unsigned long long)s_begin);
    if (s_del) atomicAdd(&stats[3], (unsigned long long)s_del);
    if (s_max != INT64_MIN) atomicMax((long long*)&stats[2], s_max);
  }
}
__global__ void __launch_bounds__(256) k_effend_rowout(ArrayDev A, const CellDesc* __restrict__ desc, int64_t* __restrict__ effend) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.n && (desc[i].flags & F_ROWOUT)) effend[i] = I64_MIN;
}

// ------------------------------------------------------------------------------------------------ K1 fragment merge
// Every fragment is a run of cells already in global (col, row) order; the array is their k-way merge, and of two cells with
// the same coordinates the one of the NEWER fragment wins (ArrayReadState::read / sort_fragment_cell_pos_ranges + PQ order,
// array_read_state.cc:1836-2132, comparator :2369-2402).  Here: log2(k) rounds of pairwise merges of neighbouring runs (run
// 2i = the older fragments, run 2i + 1 = the newer ones; on equal keys the newer run's cell goes first), each round one
// streaming pass: a block produces MERGE_TILE consecutive outputs of one pair -- two binary searches along the merge path
// find which pieces of the two runs they come from, the pieces are loaded into shared memory, every element finds its rank
// among the other run's piece by a binary search there.  Keys travel with the cell index as three streams (col, row, cell).
constexpr int MERGE_THREADS = 256, MERGE_VT = 8, MERGE_TILE = MERGE_THREADS * MERGE_VT;
struct MergeKey { int64_t col, row; };
__device__ inline bool key_less(int64_t ac, int64_t ar, int64_t bc, int64_t br) { return ac < bc || (ac == bc && ar < br); }

__global__ void __launch_bounds__(256) k_merge_init(const int64_t* __restrict__ coords, uint64_t n, int64_t* __restrict__ col, int64_t* __restrict__ row, uint32_t* __restrict__ cell) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  row[i] = coords[2 * i]; col[i] = coords[2 * i + 1]; cell[i] = (uint32_t)i;
}
// tile_pair[t] = (pair index, tile index within the pair); run_start[0 .. n_runs] = where the runs of this round begin;
// pair p merges runs 2p (A, older) and 2p + 1 (B, newer; empty when 2p + 1 == n_runs) into [run_start[2p], run_start[2p + 2])
__global__ void __launch_bounds__(MERGE_THREADS) k_merge_path(const int64_t* __restrict__ col_in, const int64_t* __restrict__ row_in, const uint32_t* __restrict__ cell_in,
                                                              int64_t* __restrict__ col_out, int64_t* __restrict__ row_out, uint32_t* __restrict__ cell_out,
                                                              const uint64_t* __restrict__ run_start, const uint32_t n_runs, const uint2* __restrict__ tile_pair) {
  __shared__ int64_t s_col[MERGE_TILE], s_row[MERGE_TILE];
  __shared__ uint32_t s_cell[MERGE_TILE];
  __shared__ uint64_t s_split[2];
  const uint32_t tid = threadIdx.x;
  const uint2 tp = tile_pair[blockIdx.x];
  const uint64_t a_lo = run_start[2 * tp.x], b_lo = run_start[min(2 * tp.x + 1, n_runs)], b_hi = run_start[min(2 * tp.x + 2, n_runs)];
  const uint64_t nA = b_lo - a_lo, nB = b_hi - b_lo;
  const uint64_t d0 = (uint64_t)tp.y * MERGE_TILE, d1 = min(d0 + (uint64_t)MERGE_TILE, nA + nB);
  // merge path: how many of the first d outputs come from A -- the first i with !(A[i] < B[d - 1 - i]) (B goes first on equal keys)
  if (tid < 2) {
    const uint64_t d = tid == 0 ? d0 : d1;
    uint64_t lo = d > nB ? d - nB : 0, hi = min(d, nA);
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1, ia = a_lo + mid, ib = b_lo + (d - 1 - mid);
      if (key_less(col_in[ia], row_in[ia], col_in[ib], row_in[ib])) lo = mid + 1; else hi = mid;
    }
    s_split[tid] = lo;
  }
  __syncthreads();
  const uint64_t a0 = s_split[0], a1 = s_split[1], b0 = d0 - a0, b1 = d1 - a1;
  const uint32_t na = (uint32_t)(a1 - a0), nb = (uint32_t)(b1 - b0);   // na + nb <= MERGE_TILE; A's piece at [0, na), B's at [na, na + nb)
  for (uint32_t i = tid; i < na + nb; i += MERGE_THREADS) {
    const uint64_t g = i < na ? a_lo + a0 + i : b_lo + b0 + (i - na);
    s_col[i] = col_in[g]; s_row[i] = row_in[g]; s_cell[i] = cell_in[g];
  }
  __syncthreads();
  const uint64_t out0 = a_lo + d0;
  for (uint32_t i = tid; i < na + nb; i += MERGE_THREADS) {
    const int64_t c = s_col[i], r = s_row[i];
    uint32_t lo, hi, rank;
    if (i < na) {   // an A element: after the B elements with key <= its own
      lo = na; hi = na + nb;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (!key_less(c, r, s_col[mid], s_row[mid])) lo = mid + 1; else hi = mid; }
      rank = i + (lo - na);
    } else {        // a B element: after the A elements with key < its own
      lo = 0; hi = na;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (key_less(s_col[mid], s_row[mid], c, r)) lo = mid + 1; else hi = mid; }
      rank = (i - na) + lo;
    }
    col_out[out0 + rank] = c; row_out[out0 + rank] = r; cell_out[out0 + rank] = s_cell[i];
  }
}
__global__ void __launch_bounds__(256) k_merge_keep(const int64_t* __restrict__ coords, const uint32_t* __restrict__ perm, uint64_t n, uint32_t* __restrict__ keep) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = 1;
  if (i > 0) {
    uint64_t a = perm[i], b = perm[i - 1];
    if (coords[2 * a] == coords[2 * b] && coords[2 * a + 1] == coords[2 * b + 1]) k = 0;   // older fragment, same coordinates
  }
  keep[i] = k;
}
__global__ void __launch_bounds__(256) k_merge_compact(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint64_t n,
                                                        uint32_t* __restrict__ src) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) src[pos[i]] = perm[i];
}
__global__ void __launch_bounds__(256) k_gather_fixed(const uint8_t* __restrict__ in, const uint32_t* __restrict__ src, uint64_t n_out, uint32_t cell_bytes, uint8_t* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t words = cell_bytes / 8;   // coords (16) and END (8) only
  if (t >= n_out * words) return;
  uint64_t i = t / words, w = t % words;
  reinterpret_cast<uint64_t*>(out)[t] = reinterpret_cast<const uint64_t*>(in)[(uint64_t)src[i] * words + w];
}
__global__ void __launch_bounds__(256) k_gather_len(const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ src, uint64_t n_out, uint64_t* __restrict__ len) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out) { uint64_t s = src[i]; len[i] = in_off[s + 1] - in_off[s]; }
}
__global__ void __launch_bounds__(256) k_gather_var(const uint64_t* __restrict__ in_off, const uint8_t* __restrict__ in_val, const uint32_t* __restrict__ src,
                                                     const uint64_t* __restrict__ out_off, uint64_t n_out, uint8_t* __restrict__ out_val) {
  // one warp per cell
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_out) return;
  uint64_t s = src[i], so = in_off[s], n = in_off[s + 1] - so, d = out_off[i];
  for (uint64_t k = threadIdx.x & 31; k < n; k += 32) out_val[d + k] = in_val[so + k];
}

// ------------------------------------------------------------------------------------------------ per-query state
struct ScanDev {
  int32_t scan_full, handle_del, min_pl_gt, want_lut;
  int64_t q_lo, q_hi;            // query interval (clamped hi)
  uint64_t nq;                   // queried rows
  const int32_t* qrow_of_row;    // num_rows entries, -1 = not queried (nullptr: identity)
  const int64_t* row_of_qrow;    // nq
  uint32_t max_alt;              // G-length fields dropped above this many ALT alleles
  // left sweep result
  int32_t* init_cell;            // nq : cell of the call live at q_lo (-1 none)
  int64_t* init_effend;
  // forward cell range in merged order: [iL, i1) has col >= q_lo, [i0, i1) has col > q_lo
  uint64_t iL, i0, i1;
  // breakpoints
  int64_t base; uint64_t span_bits, n_words;
  unsigned long long* bitmap; uint32_t* word_rank;
  uint64_t m;                    // number of breakpoints
  int64_t* bp; uint32_t* bp_cell0;
  int32_t* dlive; int32_t* ddel;
  uint64_t* site_start;          // m + 1
  uint64_t n_sites;
  int64_t* site_begin; int64_t* site_end;
  uint32_t* site_c0; uint32_t* site_c1;   // cells of the forward range whose column is the site's begin: [c0, c1) (empty: c0 == c1)
  // deletion pre-pass
  uint64_t n_del; const uint32_t* del_cells; const DelRec* del_recs; const uint32_t* del_words;
  const int64_t* coords;         // the staged array's (row, col) pairs
  // error flags: [0] bad eviction hit, [1] limit exceeded (code), [2] pool overflow
  uint32_t* err;
};

__device__ inline int32_t qrow_of(const ArrayDev& A, const ScanDev& S, int64_t row) {
  int64_t r = row - A.lb_row;
  if (r < 0 || r >= A.num_rows) return -1;
  return S.qrow_of_row ? S.qrow_of_row[r] : (int32_t)r;
}
// Bitmap of breakpoints: bit 0 stands for q_lo, bit 1 + (p - base) for position p >= base, where base = q_lo clamped to just
// below the first column of the staged data -- a query that begins far to the left of the data ([0, 1e9] over one contig)
// does not pay for the empty stretch.  (When q_lo >= that column, base = q_lo and bit 0 is never set.)
__device__ inline uint64_t bp_bit(const ScanDev& S, int64_t pos) { return pos >= S.base ? (uint64_t)(pos - S.base) + 1ull : 0ull; }
__device__ inline int64_t bp_pos(const ScanDev& S, uint64_t bit) { return bit == 0 ? S.q_lo : S.base + (int64_t)(bit - 1ull); }
__device__ inline uint64_t bp_rank(const ScanDev& S, int64_t pos) {
  uint64_t bit = bp_bit(S, pos);
  uint64_t w = bit >> 6;
  unsigned long long word = S.bitmap[w];
  unsigned long long below = word & ((1ull << (bit & 63)) - 1ull);
  return (uint64_t)S.word_rank[w] + (uint64_t)__popcll(below);
}
__device__ inline void set_bit(const ScanDev& S, int64_t pos) {
  uint64_t bit = bp_bit(S, pos);
  unsigned long long m = 1ull << (bit & 63);
  if (!(S.bitmap[bit >> 6] & m)) atomicOr(&S.bitmap[bit >> 6], m);   // thousands of cells share a column: test first
}
__device__ inline int64_t clip_end(const ScanDev& S, int64_t e) { return (!S.scan_full && e > S.q_hi) ? S.q_hi : e; }

// gt_get_column (query_variants.cc:845-941): per queried sample, the first cell at or after q_lo decides.
__global__ void __launch_bounds__(256) k_left_sweep(ArrayDev A, ScanDev S) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S.nq) return;
  int32_t cell = -1; int64_t eff = 0;
  if (!S.scan_full) {
    int64_t r = S.row_of_qrow[q] - A.lb_row;
    uint64_t lo = A.row_start[r], hi = A.row_start[r + 1], e = hi;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (A.rs_col[mid] < S.q_lo) lo = mid + 1; else hi = mid; }
    if (lo < e) {
      int64_t c = A.rs_col[lo], ef = A.rs_effend[lo];
      if (ef == I64_MIN) { cell = (int32_t)(A.rs_cell[lo] & RS_CELL_MASK); eff = c; }          // END copy: call [END field, col]
      else if (c == S.q_lo) { cell = (int32_t)(A.rs_cell[lo] & RS_CELL_MASK); eff = ef; }     // begin cell exactly at q_lo
    }
  }
  S.init_cell[q] = cell; S.init_effend[q] = eff;
  if (cell >= 0) { set_bit(S, S.q_lo); set_bit(S, clip_end(S, eff) + 1); }
}

// breakpoints = begin columns of every cell the forward scan visits + (effective end + 1) of every queried call.
// Thousands of cells share a column (and most calls end where another sample's call begins), so lanes holding the same
// position elect one leader per warp before touching the bitmap.
__device__ inline void set_bit_warp(const ScanDev& S, bool want, int64_t pos) {
  const unsigned act = __ballot_sync(0xffffffffu, want);
  if (!want) return;
  const unsigned same = __match_any_sync(act, pos);
  if ((int)(threadIdx.x & 31) == __ffs(same) - 1) set_bit(S, pos);
}
__global__ void __launch_bounds__(256) k_mark_events(ArrayDev A, ScanDev S) {
  uint64_t i = S.i0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool begin = false, queried = false;
  int64_t col = 0, end1 = 0;
  if (i < S.i1) {
    uint8_t fl = A.desc[i].flags;
    if (!(fl & (F_ENDCOPY | F_ROWOUT))) {
      begin = true;
      col = A.coords[2 * i + 1];
      if (qrow_of(A, S, A.coords[2 * i]) >= 0) { queried = true; end1 = clip_end(S, A.effend[i]) + 1; }
    }
  }
  set_bit_warp(S, begin, col);
  set_bit_warp(S, queried, end1);
}
// "Unhandled overlapping variants" (query_variants.cc:519-535): an eviction of a call that is neither a deletion nor a
// reference block throws, but only when the evicted call is live, i.e. it began inside this query.
__global__ void k_check_evictions(ArrayDev A, ScanDev S, const uint32_t* __restrict__ bad_cell, const int64_t* __restrict__ bad_pred_col, uint32_t n_bad) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_bad) return;
  uint64_t c = bad_cell[t];
  int64_t col = A.coords[2 * c + 1];
  if (qrow_of(A, S, A.coords[2 * c]) < 0) return;
  if (S.scan_full || (col > S.q_lo && col <= S.q_hi && bad_pred_col[t] >= S.q_lo)) atomicOr(&S.err[0], 1u);
}
__global__ void __launch_bounds__(256) k_popc(const unsigned long long* __restrict__ bitmap, uint64_t n_words, uint32_t* __restrict__ cnt) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n_words) cnt[w] = (uint32_t)__popcll(bitmap[w]);
}
__global__ void __launch_bounds__(256) k_extract_bp(ScanDev S) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= S.n_words) return;
  unsigned long long word = S.bitmap[w];
  uint64_t r = S.word_rank[w];
  while (word) { int b = __ffsll((long long)word) - 1; S.bp[r] = bp_pos(S, w * 64 + b); S.bp_cell0[r] = 0xFFFFFFFFu; ++r; word &= word - 1; }
}
// +1/-1 live-call deltas per breakpoint (and the same for calls containing a deletion); first cell of each column.
// Cells are sorted by column, so the breakpoint ranks a block touches start at the rank of its first cell's column and
// mostly fall within the next few thousand: they are accumulated in shared-memory bins and flushed once per block.
constexpr int DELTA_BINS = 2048;
__global__ void __launch_bounds__(256) k_deltas(ArrayDev A, ScanDev S) {
  __shared__ int s_live[DELTA_BINS];
  __shared__ int s_del[DELTA_BINS];
  __shared__ unsigned long long s_r0;
  const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x, t = t0 + threadIdx.x;
  const uint64_t n_cells = S.i1 - S.iL;
  for (int i = threadIdx.x; i < DELTA_BINS; i += 256) { s_live[i] = 0; s_del[i] = 0; }
  if (threadIdx.x == 0) {
    // lowest rank this block can touch: its first cell's column, or q_lo for the left-sweep calls
    int64_t c = (t0 < n_cells) ? A.coords[2 * (S.iL + t0) + 1] : S.q_lo;
    if (c < S.q_lo) c = S.q_lo;
    s_r0 = bp_rank(S, c);
  }
  __syncthreads();
  const uint64_t r0 = s_r0;
  auto add = [&](int* bins, int32_t* glob, uint64_t r, int v) {
    if (r >= r0 && r - r0 < (uint64_t)DELTA_BINS) atomicAdd(&bins[r - r0], v); else atomicAdd(&glob[r], v);
  };
  if (t < n_cells) {
    uint64_t i = S.iL + t;
    uint8_t fl = A.desc[i].flags;
    int64_t col = A.coords[2 * i + 1];
    bool first_of_col = (i == 0) || (A.coords[2 * (i - 1) + 1] != col);
    bool forward = i >= S.i0;
    if (first_of_col) {
      // record where column `col` starts if `col` is a breakpoint (it is whenever a call begins there)
      uint64_t bit = bp_bit(S, col);
      if (col >= S.q_lo && bit < S.span_bits && ((S.bitmap[bit >> 6] >> (bit & 63)) & 1ull)) S.bp_cell0[bp_rank(S, col)] = (uint32_t)i;
    }
    if (!(fl & (F_ROWOUT | F_ENDCOPY)) && forward && qrow_of(A, S, A.coords[2 * i]) >= 0) {
      uint64_t rs = bp_rank(S, col), re = bp_rank(S, clip_end(S, A.effend[i]) + 1);
      add(s_live, S.dlive, rs, 1); add(s_live, S.dlive, re, -1);
      if (S.handle_del && (fl & F_DELETION)) { add(s_del, S.ddel, rs, 1); add(s_del, S.ddel, re, -1); }
    }
  } else {
    uint64_t q = t - n_cells;
    if (q < S.nq) {
      int32_t c = S.init_cell[q];
      if (c >= 0) {
        uint64_t rs = bp_rank(S, S.q_lo), re = bp_rank(S, clip_end(S, S.init_effend[q]) + 1);
        add(s_live, S.dlive, rs, 1); add(s_live, S.dlive, re, -1);
        if (S.handle_del && (A.desc[c].flags & F_DELETION)) { add(s_del, S.ddel, rs, 1); add(s_del, S.ddel, re, -1); }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DELTA_BINS; i += 256) {
    const int v = s_live[i], d = s_del[i];
    if (v) atomicAdd(&S.dlive[r0 + i], v);
    if (d) atomicAdd(&S.ddel[r0 + i], d);
  }
}
// after the inclusive scans: number of operate() calls per breakpoint interval (1 bp stepping while a deletion is live,
// query_variants.cc:310)
__global__ void __launch_bounds__(256) k_site_counts(ScanDev S, const int32_t* __restrict__ live_excl, const int32_t* __restrict__ del_excl, uint64_t* __restrict__ cnt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.m) return;
  int32_t live = live_excl[i] + S.dlive[i], del = del_excl[i] + S.ddel[i];
  uint64_t c = 0;
  if (live > 0 && i + 1 < S.m) c = (del > 0) ? (uint64_t)(S.bp[i + 1] - S.bp[i]) : 1ull;
  cnt[i] = c;
}
__global__ void __launch_bounds__(256) k_fill_sites(ScanDev S, const int32_t* __restrict__ del_excl) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S.n_sites) return;
  uint64_t lo = 0, hi = S.m;   // last i with site_start[i] <= j
  while (lo + 1 < hi) { uint64_t mid = (lo + hi) >> 1; if (S.site_start[mid] <= j) lo = mid; else hi = mid; }
  uint64_t i = lo, k = j - S.site_start[i];
  bool stepping = (del_excl[i] + S.ddel[i]) > 0;
  int64_t b = S.bp[i] + (stepping ? (int64_t)k : 0);
  S.site_begin[j] = b;
  S.site_end[j] = stepping ? b : S.bp[i + 1] - 1;
  // the begin cells of column b (merge_reference_allele folds over exactly these, k_union)
  uint32_t c0 = 0, c1 = 0;
  if (k == 0 && S.bp_cell0[i] != 0xFFFFFFFFu) {
    c0 = S.bp_cell0[i];
    uint64_t l2 = c0 + 1ull, h2 = S.i1;
    while (l2 < h2) { uint64_t mid = (l2 + h2) >> 1; if (S.coords[2 * mid + 1] == b) l2 = mid + 1; else h2 = mid; }
    c1 = (uint32_t)l2;
  }
  S.site_c0[j] = c0; S.site_c1[j] = c1;
}

// ------------------------------------------------------------------------------------------------ K3 live matrix
__global__ void __launch_bounds__(256) k_rs_pack(ArrayDev A, uint64_t n_rs, RsRec* __restrict__ rec) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_rs) return;
  const uint32_t rc = A.rs_cell[k];
  RsRec r; r.col = A.rs_col[k]; r.effend = A.rs_effend[k]; r.cell = rc; r.cls = A.cls[rc & RS_CELL_MASK]; r.pad[0] = r.pad[1] = 0;
  rec[k] = r;
}
// One thread per queried sample walks that sample's cell list against `chunk` consecutive sites of the batch ("k-way merge"
// of the sample lists with the site list); adjacent threads write adjacent entries of each site's row -> coalesced.  Per
// (site, sample) it leaves the live entry and the class word of the live call (gdb_device.cuh), so the per-site scan that
// follows (k_site_scan) streams instead of gathering.  The list is read as 32-byte records, the next one kept in registers.
// (A scatter formulation -- every begin cell fills its range of sites in a shared-memory tile -- was measured at 3.9 ms per
// step against 1.8 ms for this walk: the fill loops diverge as much as the walk does, profiles/r02_k34_experiments.json.)
__global__ void __launch_bounds__(256) k_live_walk(ArrayDev A, ScanDev S, uint64_t first_site, uint32_t n_sites, uint32_t chunk, int32_t* __restrict__ live, uint32_t* __restrict__ lcls) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S.nq) return;
  const uint32_t s0 = blockIdx.y * chunk, s1 = min(n_sites, s0 + chunk);
  int64_t r = S.row_of_qrow[q] - A.lb_row;
  uint64_t lo = A.row_start[r], e = A.row_start[r + 1];
  int64_t b0 = S.site_begin[first_site + s0];
  // first entry with col > b0
  uint64_t k = lo, hi = e;
  while (k < hi) { uint64_t mid = (k + hi) >> 1; if (A.rs_col[mid] <= b0) k = mid + 1; else hi = mid; }
  // current call: last begin cell with q_lo < col <= b0 (or any col for a full scan), else the left-sweep call
  int32_t cur = S.init_cell[q]; int64_t cur_eff = S.init_effend[q]; int64_t cur_begin = 0; bool cur_del = false; uint32_t cur_cls = CLS_DEAD;
  if (cur >= 0) { uint8_t fl = A.desc[cur].flags; cur_del = fl & F_DELETION; cur_begin = (fl & F_ENDCOPY) ? A.end[cur] : A.coords[2 * (uint64_t)cur + 1]; cur_cls = A.cls[cur]; }
  for (uint64_t kk = k; kk > lo;) {
    --kk;
    const uint4 ra = reinterpret_cast<const uint4*>(A.rs_rec + kk)[0];
    const int64_t c = (int64_t)(((uint64_t)ra.y << 32) | ra.x), ef = (int64_t)(((uint64_t)ra.w << 32) | ra.z);
    if (!S.scan_full && c <= S.q_lo) break;
    if (ef != I64_MIN) { const uint2 rb = reinterpret_cast<const uint2*>(A.rs_rec + kk)[2]; cur = (int32_t)(rb.x & RS_CELL_MASK); cur_eff = ef; cur_begin = c; cur_del = rb.x >> 31; cur_cls = rb.y; break; }
  }
  // the sample's next TWO cells stay in registers (most sites of the chunk do not advance the list; when one does, the record
  // after it is already on its way: the walk is a dependent chain of L2 / HBM latencies otherwise)
  int64_t next_col = INT64_MAX, next_ef = I64_MIN; uint32_t next_rc = 0, next_cls = 0;
  uint4 pa = make_uint4(0xFFFFFFFFu, 0x7FFFFFFFu, 0, 0x80000000u); uint2 pb = make_uint2(0, 0);   // the record after `next` (col = INT64_MAX past the end)
  auto load2 = [&](uint64_t kk) {
    if (kk < e) { pa = reinterpret_cast<const uint4*>(A.rs_rec + kk)[0]; pb = reinterpret_cast<const uint2*>(A.rs_rec + kk)[2]; }
    else pa = make_uint4(0xFFFFFFFFu, 0x7FFFFFFFu, 0, 0x80000000u);
  };
  auto shift = [&]() {
    next_col = (int64_t)(((uint64_t)pa.y << 32) | pa.x); next_ef = (int64_t)(((uint64_t)pa.w << 32) | pa.z); next_rc = pb.x; next_cls = pb.y;
  };
  load2(k); shift(); load2(k + 1);
  for (uint32_t s = s0; s < s1; ++s) {
    const int64_t b = S.site_begin[first_site + s];
    while (next_col <= b) {
      if (next_ef != I64_MIN) { cur = (int32_t)(next_rc & RS_CELL_MASK); cur_eff = next_ef; cur_begin = next_col; cur_del = next_rc >> 31; cur_cls = next_cls; }
      ++k;
      shift(); load2(k + 1);
    }
    int32_t v = -1; uint32_t c = CLS_DEAD;
    if (cur >= 0 && cur_eff >= b) {
      v = cur; c = cur_cls;
      if (S.handle_del && cur_del && b > cur_begin) { v |= LIVE_REDUCED; c = CLS_REDUCED | (uint32_t)cur; }
    }
    live[(uint64_t)s * S.nq + q] = v;
    lcls[(uint64_t)s * S.nq + q] = c;
  }
}

}  // namespace gdb
