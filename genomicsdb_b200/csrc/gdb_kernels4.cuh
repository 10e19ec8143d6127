// K7: the columnar iterator and its single-cell operators.
//
// The reference walks a query interval cell by cell (SingleCellTileDBIterator, genomicsdb_iterators.cc:181-300,425-510):
// first the calls that began before the interval and reach into it -- found through their END copies, handed out
// ordered by (begin column, row) -- then every begin cell of the interval in storage (column-major) order, rows outside
// the query skipped.  Here that walk is a LIST of cell indices built with two flag/scan/compact passes, and the
// operators the reference plugs into iterate_over_cells (query_variants.cc:557-576) are kernels over the list:
//   ColumnHistogramOperator  (variant_operations.cc:732-765)   k_cells_hist
//   AlleleCountOperator      (variant_operations.cc:900-1062)  k_ac_count / k_ac_fill (+ radix sort, k_ac_heads)
// The print operators need the cells' bytes on the host: k_cells_fill + the engine's column gathers.
#pragma once
#include "gdb_kernels3.cuh"

namespace gdb {

// left phase: per queried row, the first cell at or after q_lo (k_left_sweep found it) counts when it is the END copy of
// a call that began before q_lo (is_duplicate_cell_at_end_position_that_begins_before_query_interval)
__global__ void k_cells_left_flag(ArrayDev A, ScanDev S, uint32_t* __restrict__ flag, uint64_t* __restrict__ key) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S.nq) return;
  int32_t c = S.init_cell[q];
  bool keep = false; int64_t begin = 0;
  if (c >= 0 && (A.desc[c].flags & F_ENDCOPY)) { begin = A.end[c]; keep = begin < S.q_lo; }
  flag[q] = keep ? 1u : 0u;
  key[q] = (uint64_t)begin;
}
__global__ void k_cells_left_compact(ScanDev S, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, const uint64_t* __restrict__ key,
                                     uint32_t* __restrict__ out_cell, uint64_t* __restrict__ out_key) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S.nq || !flag[q]) return;
  out_cell[pos[q]] = (uint32_t)S.init_cell[q];
  out_key[pos[q]] = key[q];
}
// simple traversal: begin cells of queried rows in [i0, i1)
__global__ void k_cells_fwd_flag(ArrayDev A, ScanDev S, uint32_t* __restrict__ flag) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = S.i0 + t;
  if (i >= S.i1) return;
  flag[t] = (!(A.desc[i].flags & (F_ENDCOPY | F_ROWOUT)) && qrow_of(A, S, A.coords[2 * i]) >= 0) ? 1u : 0u;
}
__global__ void k_cells_fwd_compact(ScanDev S, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t* __restrict__ out_cell) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (S.i0 + t >= S.i1 || !flag[t]) return;
  out_cell[pos[t]] = (uint32_t)(S.i0 + t);
}
// row / interval of every listed cell; a left cell is an END copy, its coordinates and END swap places
__global__ void k_cells_fill(ArrayDev A, const uint32_t* __restrict__ list, uint64_t n, uint64_t* __restrict__ cell, int64_t* __restrict__ row,
                             int64_t* __restrict__ begin, int64_t* __restrict__ end) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t c = list[i];
  const int64_t col = A.coords[2 * c + 1], e = A.end[c];
  cell[i] = c; row[i] = A.coords[2 * c];
  begin[i] = col > e ? e : col; end[i] = col > e ? col : e;
}

// ColumnHistogramOperator::operate_on_columnar_cell (variant_operations.cc:754-765)
__device__ inline uint64_t hist_bin(const ArrayDev& A, uint64_t c, uint64_t h_begin, uint64_t h_end, uint64_t bin_size, uint64_t n_bins) {
  const int64_t col = A.coords[2 * c + 1], e = A.end[c];
  const uint64_t b = (uint64_t)(col > e ? e : col);
  return b <= h_begin ? 0ull : (b >= h_end ? n_bins - 1 : (b - h_begin) / bin_size);
}
// the bins the listed cells touch: range[0] = lowest, range[1] = highest (the reference's tool spans 40 M bins, a query a few)
__global__ void k_cells_bin_range(ArrayDev A, const uint32_t* __restrict__ list, uint64_t n, uint64_t h_begin, uint64_t h_end, uint64_t bin_size, uint64_t n_bins,
                                  unsigned long long* __restrict__ range) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long lo = ~0ull, hi = 0ull;
  if (i < n) lo = hi = hist_bin(A, list[i], h_begin, h_end, bin_size, n_bins);
  for (int d = 16; d; d >>= 1) {
    const unsigned long long lo2 = __shfl_xor_sync(0xffffffffu, lo, d), hi2 = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = lo2 < lo ? lo2 : lo; hi = hi2 > hi ? hi2 : hi;
  }
  // one atomic per block, and only when it would change the answer (every warp hitting two words serialises)
  __shared__ unsigned long long s_lo, s_hi;
  if (threadIdx.x == 0) { s_lo = ~0ull; s_hi = 0ull; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && lo != ~0ull) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
  __syncthreads();
  if (threadIdx.x == 0 && s_lo != ~0ull) {
    if (s_lo < *(volatile unsigned long long*)&range[0]) atomicMin(&range[0], s_lo);
    if (s_hi > *(volatile unsigned long long*)&range[1]) atomicMax(&range[1], s_hi);
  }
}
// thousands of cells share a column, so a warp adds once per bin; counts[] starts at bin `first`
__global__ void k_cells_hist(ArrayDev A, const uint32_t* __restrict__ list, uint64_t n, uint64_t h_begin, uint64_t h_end, uint64_t bin_size, uint64_t n_bins,
                             uint64_t first, unsigned long long* __restrict__ counts) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n;
  uint64_t bin = 0;
  if (on) bin = hist_bin(A, list[i], h_begin, h_end, bin_size, n_bins) - first;
  const unsigned act = __ballot_sync(0xffffffffu, on);
  if (!on) return;
  const unsigned same = __match_any_sync(act, bin);
  if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&counts[bin], (unsigned long long)__popc(same));
}

// InterestingLocationsPrinter::operate (variant_operations.cc:380-400) over one batch of the combine's sites: per site, the
// live calls that are reference blocks (REF of one base, ALT = <NON_REF> alone) and the live calls that begin at the site.
// One block per site walks its row of the live matrix; stats[2 s] = reference blocks, stats[2 s + 1] = begin here.
__global__ void __launch_bounds__(256) k_site_stats(ArrayDev A, ScanDev S, const int32_t* __restrict__ live, uint64_t first_site, uint32_t n_sites,
                                                    uint32_t* __restrict__ stats) {
  const uint32_t s = blockIdx.x;
  if (s >= n_sites) return;
  const int64_t begin = S.site_begin[first_site + s];
  uint32_t n_ref = 0, n_beg = 0;
  for (uint64_t q = threadIdx.x; q < S.nq; q += blockDim.x) {
    const int32_t e = live[(uint64_t)s * S.nq + q];
    if (e < 0) continue;
    const uint64_t c = (uint32_t)e & 0x3FFFFFFFu;
    const CellDesc d = A.desc[c];
    n_ref += (d.ref_len == 1 && d.n_alt == 1 && d.nr_idx == 1) ? 1u : 0u;
    const int64_t col = A.coords[2 * c + 1], en = A.end[c];
    n_beg += ((col > en ? en : col) == begin) ? 1u : 0u;
  }
  __shared__ uint32_t s_ref, s_beg;
  if (threadIdx.x == 0) { s_ref = 0; s_beg = 0; }
  __syncthreads();
  for (int d = 16; d; d >>= 1) { n_ref += __shfl_xor_sync(0xffffffffu, n_ref, d); n_beg += __shfl_xor_sync(0xffffffffu, n_beg, d); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_ref, n_ref); atomicAdd(&s_beg, n_beg); }
  __syncthreads();
  if (threadIdx.x == 0) { stats[2 * s] = s_ref; stats[2 * s + 1] = s_beg; }
}

// --- AlleleCountOperator --------------------------------------------------------------------------------------------
__device__ inline bool ac_symbolic(const uint8_t* a, uint32_t n) {   // VariantUtils::is_symbolic_allele (known_field_info.h:221-232)
  if (n == 0) return false;
  if (n == 1 && (a[0] == '&' || a[0] == '*')) return true;
  if (a[0] == '<' && a[n - 1] == '>') return true;
  for (uint32_t k = 0; k < n; ++k) if (a[k] == '[' || a[k] == ']') return true;
  return false;
}
// AlleleCountOperator::normalize_REF_ALT_pair (variant_operations.cc:1007-1048): lengths of the REF / ALT prefixes kept
__device__ inline void ac_normalize(const uint8_t* alt, uint32_t ref_len, uint32_t alt_len, uint32_t& ref_keep, uint32_t& alt_keep) {
  ref_keep = ref_len; alt_keep = alt_len;
  if (ref_len > 1 && alt_len) {
    if (ac_symbolic(alt, alt_len)) { ref_keep = 1; return; }
    uint32_t suffix = 0;
    if (alt_len >= ref_len) suffix = ref_len - 1;      // SNV written over a deletion's REF, or insertion
    else if (alt_len > 1) suffix = alt_len - 1;       // deletion
    ref_keep = ref_len - suffix; alt_keep = alt_len - suffix;
  }
}
__device__ inline bool ac_gt_counts(int32_t v) { return v != INT32_MIN && v != INT32_MIN + 1 && v > 0; }   // is_bcf_valid_value && not REF

struct AcCell { const uint8_t* ref; uint32_t ref_len; const uint8_t* alt; uint32_t alt_len; const int32_t* gt; uint32_t gt_len; int64_t column; };
__device__ inline bool ac_cell(const ArrayDev& A, uint64_t c, AcCell& x) {
  const uint64_t r0 = A.ref.off[c], r1 = A.ref.off[c + 1], a0 = A.alt.off[c], a1 = A.alt.off[c + 1], g0 = A.gt.off[c], g1 = A.gt.off[c + 1];
  x.ref = (const uint8_t*)A.ref.val + r0; x.ref_len = (uint32_t)(r1 - r0);
  x.alt = (const uint8_t*)A.alt.val + a0; x.alt_len = (uint32_t)(a1 - a0);
  x.gt = (const int32_t*)((const uint8_t*)A.gt.val + g0); x.gt_len = (uint32_t)((g1 - g0) / 4);
  const int64_t col = A.coords[2 * c + 1], e = A.end[c];
  x.column = col > e ? e : col;
  return x.ref_len && x.alt_len && x.gt_len;   // any invalid field: the operator skips the cell
}
__global__ void k_ac_count(ArrayDev A, const uint32_t* __restrict__ list, uint64_t n, uint32_t gt_step, uint32_t* __restrict__ cnt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AcCell x; uint32_t k = 0;
  if (A.has_gt && ac_cell(A, list[i], x)) for (uint32_t g = 0; g < x.gt_len; g += gt_step) k += ac_gt_counts(x.gt[g]) ? 1u : 0u;
  cnt[i] = k;
}
// one entry per counted GT element: key = FNV-1a(column, normalised REF, normalised ALT); the entry remembers (list index, ALT index)
__global__ void k_ac_fill(ArrayDev A, const uint32_t* __restrict__ list, uint64_t n, uint32_t gt_step, const uint32_t* __restrict__ pos,
                          uint64_t* __restrict__ key, uint32_t* __restrict__ ent_cell, uint32_t* __restrict__ ent_alt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AcCell x;
  if (!A.has_gt || !ac_cell(A, list[i], x)) return;
  uint32_t out = pos[i];
  for (uint32_t g = 0; g < x.gt_len; g += gt_step) {
    const int32_t v = x.gt[g];
    if (!ac_gt_counts(v)) continue;
    // the (v-1)-th '|'-separated ALT token
    uint32_t t0 = 0, tok = 0;
    while (tok < (uint32_t)(v - 1) && t0 < x.alt_len) { if (x.alt[t0] == '|') ++tok; ++t0; }
    uint32_t t1 = t0;
    while (t1 < x.alt_len && x.alt[t1] != '|') ++t1;
    uint32_t rk, ak;
    ac_normalize(x.alt + t0, x.ref_len, t1 - t0, rk, ak);
    uint64_t h = 1469598103934665603ull;
    for (int b = 0; b < 8; ++b) { h ^= (uint64_t)((x.column >> (8 * b)) & 0xFF); h *= 1099511628211ull; }
    for (uint32_t b = 0; b < rk; ++b) { h ^= x.ref[b]; h *= 1099511628211ull; }
    h ^= 0xFFull; h *= 1099511628211ull;
    for (uint32_t b = 0; b < ak; ++b) { h ^= x.alt[t0 + b]; h *= 1099511628211ull; }
    key[out] = h; ent_cell[out] = (uint32_t)i; ent_alt[out] = (uint32_t)(v - 1);
    ++out;
  }
}
__global__ void k_ac_heads(const uint64_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}
__global__ void k_ac_compact(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ent_sorted, uint64_t n,
                             uint32_t* __restrict__ head_ent, uint32_t* __restrict__ head_start) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) { head_ent[pos[i]] = ent_sorted[i]; head_start[pos[i]] = (uint32_t)i; }
}
__global__ void k_gather_u32(const uint32_t* __restrict__ in, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

}  // namespace gdb
