// K6: FORMAT collation + BCF typed encoding of the remapped fields (first slice).
//
// What the reference does per site after the remap (paths relative to GenomicsDB v0.10.2):
//   VariantFieldHandler<T>::collect_and_extend_fields   src/main/cpp/src/genomicsdb/variant_field_handler.cc:804-871
//       every query row gets `nps` values (nps = longest valid vector at the site): its vector, or [missing]
//       ([vector_end] for GT) when the row has no valid field, padded with vector_end
//   encode_GT_vector                                    src/main/cpp/src/query_operations/broad_combined_gvcf.cc:53-138,646-658
//   bcf_update_format -> bcf_enc_vint                   dependencies/htslib/vcf.c:3607-3700,1926-1972
//       one scan for min / max (missing and vector_end excluded) picks int8 / int16 / int32 for the whole block
// K5's output already is the collated layout (dense, sample-major, vector_end padded) in 32-bit words, so K6 is
//   k_fmt_range   min / max per (site, field) over the words the block will hold          (HBM read, 4 B / word)
//   (host)        type + values-per-sample + 16-byte aligned byte offset of every block
//   k_fmt_pack    narrows the words into the block bytes htslib would have produced        (HBM read 4 B, write 1-4 B / word)
// The blocks are what bcf1_t.indiv holds after the field's typed key and size descriptor; the host emits those few
// header bytes (gdb_host_api).  Integer / byte work, bound by HBM bandwidth.
#pragma once
#include "gdb_kernels2.cuh"

namespace gdb {

enum : uint8_t { BT_NONE = 0, BT_INT8 = 1, BT_INT16 = 2, BT_INT32 = 3, BT_FLOAT = 5, BT_CHAR = 7 };
constexpr int FMT_THREADS = 256;

struct FmtPlanDev {          // [o * n_sites + site]
  const uint8_t* type;       // BT_*; BT_NONE: the field is not emitted at this site
  const uint32_t* nps;       // values per sample
  const uint64_t* off;       // byte offset of the block in `bytes`, multiple of 16
  uint8_t* bytes;
};

// allele value -> BCF genotype int (encode_GT_element); missing / vector_end pass through
__device__ inline uint32_t gt_encode(uint32_t v, uint32_t phased, int produce_gt) {
  if (v == BCF_INT_MISSING || v == BCF_INT_VEND) return v;
  return produce_gt ? (((v + 1u) << 1) | phased) : phased;
}
// element i (< nps) of query row q's GT vector as the FORMAT block holds it
__device__ inline uint32_t gt_element(const uint32_t* __restrict__ row, uint32_t i, int layout_phased, int produce_gt) {
  if (!layout_phased || i == 0) return gt_encode(row[i], 0u, produce_gt);
  return gt_encode(row[2 * i], (int32_t)row[2 * i - 1] > 0 ? 1u : 0u, produce_gt);
}

// min / max of the values of every (site, field) block; mm[(o * n_sites + site) * 2 + {0: min, 1: max}] starts at
// {INT32_MAX, INT32_MIN + 1} like bcf_enc_vint's accumulators.  grid = (chunks, sites)
__global__ void __launch_bounds__(FMT_THREADS) k_fmt_range(const __grid_constant__ ArrayDev A, const __grid_constant__ RemapOut O, const uint32_t n_sites, const uint32_t nq,
                                                           const int produce_gt, const int java, int32_t* __restrict__ mm) {
  __shared__ int32_t s_min[FMT_THREADS / 32], s_max[FMT_THREADS / 32];
  const uint32_t js = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int o = 0; o < O.n_out; ++o) {
    const FieldOutDev fo = O.f[o];
    const uint32_t stride = fo.stride[js];
    if (stride == 0) continue;
    const bool is_gt = A.has_gt && o == 0;
    const int fi = o - (A.has_gt ? 1 : 0);
    if (!is_gt && A.fm[fi].elem) continue;   // float blocks are never narrowed
    const uint32_t* __restrict__ src = fo.data + fo.base[js];
    int32_t lo = INT32_MAX, hi = INT32_MIN + 1;
    if (is_gt) {
      const uint32_t nps = A.gt_phased ? (stride + 1u) >> 1 : stride;
      for (uint32_t q = blockIdx.x * FMT_THREADS + tid; q < nq; q += gridDim.x * FMT_THREADS) {
        const uint32_t* row = src + (uint64_t)q * stride;
        if (java && fo.len[(uint64_t)js * nq + q] == 0) { lo = min(lo, 0); hi = max(hi, 0); continue; }   // JNI conventions: a row without GT is a no-call
        for (uint32_t i = 0; i < nps; ++i) {
          const uint32_t v = gt_element(row, i, A.gt_phased, produce_gt);
          if (v != BCF_INT_MISSING && v != BCF_INT_VEND) { lo = min(lo, (int32_t)v); hi = max(hi, (int32_t)v); }
        }
      }
    } else {
      const uint64_t n_words = (uint64_t)nq * stride, n_quads = n_words >> 2;
      const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(src);   // site blocks start on 16-byte boundaries
      for (uint64_t g = (uint64_t)blockIdx.x * FMT_THREADS + tid; g < n_quads; g += (uint64_t)gridDim.x * FMT_THREADS) {
        const uint4 v = src4[g];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int x = 0; x < 4; ++x) if (w[x] != BCF_INT_MISSING && w[x] != BCF_INT_VEND) { lo = min(lo, (int32_t)w[x]); hi = max(hi, (int32_t)w[x]); }
      }
      if (blockIdx.x == 0 && tid < (uint32_t)(n_words & 3u)) {
        const uint32_t w = src[(n_quads << 2) + tid];
        if (w != BCF_INT_MISSING && w != BCF_INT_VEND) { lo = min(lo, (int32_t)w); hi = max(hi, (int32_t)w); }
      }
    }
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { s_min[wid] = lo; s_max[wid] = hi; }
    __syncthreads();
    if (tid == 0) {
      for (int k = 1; k < FMT_THREADS / 32; ++k) { lo = min(lo, s_min[k]); hi = max(hi, s_max[k]); }
      int32_t* p = mm + ((uint64_t)o * n_sites + js) * 2;
      if (lo != INT32_MAX) atomicMin(p, lo);
      if (hi != INT32_MIN + 1) atomicMax(p + 1, hi);
    }
    __syncthreads();
  }
}

// Which kernel narrows a block: k_fmt_pack_tile stages PACK_ROWS rows of the block in shared memory (coalesced 16-byte loads
// of K5's padded rows in, 16-byte stores of the dense block out), which needs PACK_ROWS * nps * width bytes; GT (encoded
// on the fly) and very long vectors stay with k_fmt_pack.
constexpr uint32_t PACK_ROWS = 256, PACK_SMEM = 40 * 1024;   // rows of a site per block; bytes of the staging buffer
__host__ __device__ inline uint32_t bt_width(uint32_t type) { return (type == BT_INT8 || type == BT_CHAR) ? 1u : (type == BT_INT16 ? 2u : 4u); }
// rows staged at a time: a multiple of 16 (so every slice of a block starts on a 16-byte boundary whatever nps and the width are)
__host__ __device__ inline uint32_t pack_tile_rows(uint32_t nps, uint32_t type) { const uint32_t r = (PACK_SMEM / (nps * bt_width(type))) & ~15u; return r > PACK_ROWS ? PACK_ROWS : r; }
__host__ __device__ inline bool pack_tile_takes(uint32_t nps, uint32_t type) { return pack_tile_rows(nps, type) >= 16u; }

// one value into the staging buffer, narrowed to the block's type (the type is a compile-time constant of the loop)
template <uint32_t TYPE>
__device__ inline void pack_put(uint8_t* buf, uint32_t k, uint32_t v) {
  if (TYPE == BT_INT8) buf[k] = v == BCF_INT_VEND ? 0x81u : (v == BCF_INT_MISSING ? 0x80u : (uint8_t)v);
  else if (TYPE == BT_INT16) reinterpret_cast<uint16_t*>(buf)[k] = v == BCF_INT_VEND ? 0x8001u : (v == BCF_INT_MISSING ? 0x8000u : (uint16_t)v);
  else reinterpret_cast<uint32_t*>(buf)[k] = v;
}
// `rows` rows of whole quads (stride = 4 * qpr words) -> the dense slice in `buf`.  Thread t takes quads t, t + 256, ...;
// (row, quad-in-row) advance without a division; ILP independent 16-byte loads are in flight per thread.
template <uint32_t TYPE>
__device__ inline void pack_quads(uint8_t* buf, const uint4* __restrict__ src4, const uint16_t* __restrict__ len, uint32_t rows, uint32_t qpr, uint32_t nps,
                                  uint32_t miss, uint32_t vend, uint32_t tid, const bool java) {
  constexpr int ILP = 4;
  const uint32_t n_quads = rows * qpr, dr = FMT_THREADS / qpr, dj = FMT_THREADS - dr * qpr;
  uint32_t r = tid / qpr, j = tid - r * qpr;
  for (uint32_t g0 = tid; g0 < n_quads; g0 += ILP * FMT_THREADS) {
    uint4 v[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) { const uint32_t g = g0 + u * FMT_THREADS; v[u] = g < n_quads ? src4[g] : make_uint4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (g0 + u * FMT_THREADS < n_quads) {
        uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        // a row without the field is all vector_end (K5); its first value becomes missing.  Only such rows look at len[].
        if (j == 0 && x[0] == vend && len[r] == 0) x[0] = miss;
        if (java) {   // JNI conventions: rows are padded with missing, not vector_end
#pragma unroll
          for (int t = 0; t < 4; ++t) if (x[t] == vend) x[t] = miss;
        }
        const uint32_t k = r * nps + 4 * j;
#pragma unroll
        for (uint32_t t = 0; t < 4; ++t) if (4 * j + t < nps) pack_put<TYPE>(buf, k + t, x[t]);
      }
      r += dr; j += dj;
      if (j >= qpr) { j -= qpr; ++r; }
    }
  }
}

// grid = (row groups of PACK_ROWS, sites); a block walks its rows in slices that fit the staging buffer.
__global__ void __launch_bounds__(FMT_THREADS) k_fmt_pack_tile(const __grid_constant__ ArrayDev A, const __grid_constant__ RemapOut O, const __grid_constant__ FmtPlanDev P,
                                                               const uint32_t n_sites, const uint32_t nq, const int java) {
  __shared__ __align__(16) uint8_t s_buf[PACK_SMEM];
  const uint32_t js = blockIdx.y, tid = threadIdx.x;
  const uint32_t qb = blockIdx.x * PACK_ROWS, qe = min(qb + PACK_ROWS, nq);
  const int o0 = A.has_gt ? 1 : 0;
  for (int o = o0; o < O.n_out; ++o) {
    const uint64_t pi = (uint64_t)o * n_sites + js;
    const uint32_t type = P.type[pi];
    if (type == BT_NONE) continue;
    const uint32_t nps = P.nps[pi];
    if (!pack_tile_takes(nps, type)) continue;
    const uint32_t slice = pack_tile_rows(nps, type);
    const FieldOutDev fo = O.f[o];
    const uint32_t stride = fo.stride[js], w = bt_width(type);
    const uint32_t miss = missing_word(A.fm[o - o0].elem);
    auto put = [&](uint32_t r, uint32_t i, uint32_t v) {   // value i of local row r -> its place in the dense block slice
      const uint32_t k = r * nps + i;
      if (type == BT_INT8) s_buf[k] = v == BCF_INT_VEND ? 0x81u : (v == BCF_INT_MISSING ? 0x80u : (uint8_t)v);
      else if (type == BT_INT16) reinterpret_cast<uint16_t*>(s_buf)[k] = v == BCF_INT_VEND ? 0x8001u : (v == BCF_INT_MISSING ? 0x8000u : (uint16_t)v);
      else reinterpret_cast<uint32_t*>(s_buf)[k] = v;
    };
    for (uint32_t q0 = qb; q0 < qe; q0 += slice) {
      const uint32_t rows = min(slice, qe - q0);
      const uint32_t* __restrict__ src = fo.data + fo.base[js] + (uint64_t)q0 * stride;
      const uint16_t* __restrict__ len = fo.len + (uint64_t)js * nq + q0;
      if ((stride & 3u) == 0) {   // rows are whole quads (every remapped site)
        const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(src);
        const uint32_t vend = vend_word(A.fm[o - o0].elem);
        if (type == BT_INT8) pack_quads<BT_INT8>(s_buf, src4, len, rows, stride >> 2, nps, miss, vend, tid, java != 0);
        else if (type == BT_INT16) pack_quads<BT_INT16>(s_buf, src4, len, rows, stride >> 2, nps, miss, vend, tid, java != 0);
        else pack_quads<BT_INT32>(s_buf, src4, len, rows, stride >> 2, nps, miss, vend, tid, java != 0);
      } else {
        const uint32_t n_words = rows * stride;
        for (uint32_t g = tid; g < n_words; g += FMT_THREADS) {
          const uint32_t r = g / stride, i = g - r * stride;
          if (i >= nps) continue;
          uint32_t v = src[g];
          if (i == 0 && len[r] == 0) v = miss;
          if (java && v == vend_word(A.fm[o - o0].elem)) v = miss;
          put(r, i, v);
        }
      }
      __syncthreads();
      const uint32_t bytes = rows * nps * w;
      uint8_t* __restrict__ dst = P.bytes + P.off[pi] + (uint64_t)q0 * nps * w;
      const uint32_t n16 = bytes >> 4;
      for (uint32_t g = tid; g < n16; g += FMT_THREADS) reinterpret_cast<uint4*>(dst)[g] = reinterpret_cast<const uint4*>(s_buf)[g];
      for (uint32_t b = (n16 << 4) + tid; b < bytes; b += FMT_THREADS) dst[b] = s_buf[b];
      __syncthreads();
    }
  }
}

// narrows the words of every emitted (site, field) into its block.  grid = (chunks, sites); a thread makes 4 consecutive
// values of the block per trip (one 4 / 8 / 16-byte store).
__global__ void __launch_bounds__(FMT_THREADS) k_fmt_pack(const __grid_constant__ ArrayDev A, const __grid_constant__ RemapOut O, const __grid_constant__ FmtPlanDev P,
                                                          const uint32_t n_sites, const uint32_t nq, const int produce_gt, const int java) {
  const uint32_t js = blockIdx.y, tid = threadIdx.x;
  for (int o = 0; o < O.n_out; ++o) {
    const uint64_t pi = (uint64_t)o * n_sites + js;
    const uint32_t type = P.type[pi];
    if (type == BT_NONE) continue;
    const FieldOutDev fo = O.f[o];
    const uint32_t stride = fo.stride[js], nps = P.nps[pi];
    const bool is_gt = A.has_gt && o == 0;
    if (!is_gt && pack_tile_takes(nps, type)) continue;   // k_fmt_pack_tile writes this block
    const int fi = o - (A.has_gt ? 1 : 0);
    const uint32_t miss = is_gt ? BCF_INT_VEND : missing_word(A.fm[fi].elem);   // first value of a row without the field
    const uint32_t* __restrict__ src = fo.data + fo.base[js];
    const uint16_t* __restrict__ len = fo.len + (uint64_t)js * nq;
    uint8_t* __restrict__ dst = P.bytes + P.off[pi];
    const uint64_t n_el = (uint64_t)nq * nps, n_grp = (n_el + 3) >> 2;
    const bool flat = !is_gt && stride == nps;
    for (uint64_t g = (uint64_t)blockIdx.x * FMT_THREADS + tid; g < n_grp; g += (uint64_t)gridDim.x * FMT_THREADS) {
      const uint64_t e = g << 2;
      uint32_t q = (uint32_t)(e / nps), i = (uint32_t)(e - (uint64_t)q * nps);
      uint32_t w[4];
      const uint32_t cnt = (uint32_t)min((uint64_t)4, n_el - e);
      if (flat && cnt == 4) { const uint4 v = *reinterpret_cast<const uint4*>(src + e); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
#pragma unroll
      for (uint32_t x = 0; x < 4; ++x) {
        if (x < cnt) {
          const uint32_t* row = src + (uint64_t)q * stride;
          if (is_gt) {
            w[x] = gt_element(row, i, A.gt_phased, produce_gt);
            if (java) w[x] = len[q] == 0 ? (i == 0 ? 0u : BCF_INT_MISSING) : (w[x] == BCF_INT_VEND ? BCF_INT_MISSING : w[x]);   // no-call, missing padding
          } else {
            if (!(flat && cnt == 4)) w[x] = row[i];
            if (i == 0 && len[q] == 0) w[x] = miss;
            if (java && w[x] == vend_word(A.fm[fi].elem)) w[x] = missing_word(A.fm[fi].elem);
          }
          if (++i == nps) { i = 0; ++q; }
        } else w[x] = 0;
      }
      if (type == BT_INT8) {
        uint32_t pk = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) pk |= (w[x] == BCF_INT_VEND ? 0x81u : (w[x] == BCF_INT_MISSING ? 0x80u : (w[x] & 0xFFu))) << (8 * x);
        if (cnt == 4) *reinterpret_cast<uint32_t*>(dst + e) = pk;
        else for (uint32_t x = 0; x < cnt; ++x) dst[e + x] = (uint8_t)(pk >> (8 * x));
      } else if (type == BT_INT16) {
        uint32_t h[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) h[x] = w[x] == BCF_INT_VEND ? 0x8001u : (w[x] == BCF_INT_MISSING ? 0x8000u : (w[x] & 0xFFFFu));
        if (cnt == 4) *reinterpret_cast<uint2*>(dst + 2 * e) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        else for (uint32_t x = 0; x < cnt; ++x) *reinterpret_cast<uint16_t*>(dst + 2 * (e + x)) = (uint16_t)h[x];
      } else {   // int32 / float: the words as they are
        if (cnt == 4) *reinterpret_cast<uint4*>(dst + 4 * e) = make_uint4(w[0], w[1], w[2], w[3]);
        else for (uint32_t x = 0; x < cnt; ++x) *reinterpret_cast<uint32_t*>(dst + 4 * (e + x)) = w[x];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ plain FORMAT fields
// FORMAT fields whose length does not depend on the alleles (GQ, MIN_DP, DP_FORMAT, SB, PGT, PID ...) are not remapped:
// their collation is a gather through the live matrix (the call of query row q at site s -> its cell -> the cell's
// values), with the same row rules as above (collect_and_extend_fields, variant_field_handler.cc:804-871) and, for
// strings, bcf_str_missing / bcf_str_vector_end (htslib vcf.h) and BCF_BT_CHAR blocks (vcf.c:3653-3657).  A field of a
// call is valid unless every element is the TileDB null of its type (variant_field_data.h:236-251,365-384).
constexpr int MAX_PLAIN = 24;
enum { ELEM_INT32 = 0, ELEM_FLOAT32 = 1, ELEM_CHAR = 2 };
constexpr uint8_t TDB_NULL_CHAR = 127, BCF_STR_MISSING = 0x07, BCF_STR_VEND = 0x00;
// host_combine: the block feeds a host-side combine, always padded the htslib way -- HC_BLOCK (ID, FILTER union, concatenate);
// | HC_SKIP_REDUCED: calls in spanning-deletion form have no value (handle_deletions invalidates their INFO fields,
// broad_combined_gvcf.cc:1067-1075); | HC_RAW: every row is [u32 byte length, 0xFFFFFFFF = no value][the cell's bytes][zero pad] --
// multi-dimensional fields (genomicsdb_multid_vector_field.h:66-83: sizes and offsets inside the value) and the REF / ALT
// strings their allele remap needs, binary-safe
enum { HC_BLOCK = 1, HC_RAW = 2, HC_SKIP_REDUCED = 4 };
struct PlainDev { const uint64_t* off; const uint8_t* val; int32_t elem, fixed_len; int32_t host_combine; };
struct PlainSet { int32_t n; PlainDev f[MAX_PLAIN]; };
enum { PS_MIN = 0, PS_MAX, PS_MAXLEN, PS_ANY, PS_WORDS };   // per (plain field, site) statistics, int32 each

// values of plain field F in `cell`: pointer, number of elements, validity
__device__ inline bool plain_fetch(const PlainDev& F, uint32_t cell, const uint8_t*& p, uint32_t& n) {
  const uint32_t es = F.elem == ELEM_CHAR ? 1u : 4u;
  if (F.fixed_len) { n = (uint32_t)F.fixed_len; p = F.val + (uint64_t)cell * n * es; }
  else { const uint64_t o0 = F.off[cell], o1 = F.off[cell + 1]; p = F.val + o0; n = (uint32_t)((o1 - o0) / es); }
  if (F.elem == ELEM_CHAR) { for (uint32_t i = 0; i < n; ++i) if (p[i] != TDB_NULL_CHAR) return true; return false; }
  const uint32_t nullw = F.elem == ELEM_FLOAT32 ? TDB_NULL_FLT : TDB_NULL_INT;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  for (uint32_t i = 0; i < n; ++i) if (w[i] != nullw) return true;
  return false;
}

// pass 1: per (field, site): is any row's field valid, the longest valid vector, min / max of the integer values.
// grid = (row chunks, sites), one thread per query row.
__global__ void __launch_bounds__(FMT_THREADS) k_plain_scan(const __grid_constant__ PlainSet P, const int32_t* __restrict__ live, const uint32_t n_sites, const uint32_t nq,
                                                            int32_t* __restrict__ stat) {
  __shared__ int32_t s_st[MAX_PLAIN][PS_WORDS];
  const uint32_t js = blockIdx.y, tid = threadIdx.x, q = blockIdx.x * FMT_THREADS + tid;
  if (tid < (uint32_t)P.n * PS_WORDS) { const int w = tid % PS_WORDS; s_st[tid / PS_WORDS][w] = w == PS_MIN ? INT32_MAX : (w == PS_MAX ? INT32_MIN + 1 : 0); }
  __syncthreads();
  const int32_t e = q < nq ? live[(uint64_t)js * nq + q] : -1;
  const uint32_t cell = e >= 0 ? (uint32_t)(e & (LIVE_REDUCED - 1)) : 0u;
  const uint32_t lane = tid & 31;
  for (int f = 0; f < P.n; ++f) {   // every lane takes part: the statistics are reduced over the warp before they touch shared memory
    const uint8_t* p = nullptr; uint32_t n = 0;
    bool valid = e >= 0 && !((P.f[f].host_combine & HC_SKIP_REDUCED) && (e & LIVE_REDUCED)) && plain_fetch(P.f[f], cell, p, n);
    if (valid && (P.f[f].host_combine & HC_RAW)) n += 4u;   // the length word in front of the bytes
    int32_t lo = INT32_MAX, hi = INT32_MIN + 1;
    if (valid && P.f[f].elem == ELEM_INT32) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
      for (uint32_t i = 0; i < n; ++i) if (w[i] != BCF_INT_MISSING && w[i] != BCF_INT_VEND) { lo = min(lo, (int32_t)w[i]); hi = max(hi, (int32_t)w[i]); }
    }
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, valid ? n : 0u);
    if (!__any_sync(0xffffffffu, valid)) continue;
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) {
      atomicMax(&s_st[f][PS_MAXLEN], (int32_t)maxlen);
      s_st[f][PS_ANY] = 1;
      if (lo != INT32_MAX) { atomicMin(&s_st[f][PS_MIN], lo); atomicMax(&s_st[f][PS_MAX], hi); }
    }
  }
  __syncthreads();
  if (tid < (uint32_t)P.n * PS_WORDS) {
    const int f = tid / PS_WORDS, w = tid % PS_WORDS;
    int32_t* g = stat + ((uint64_t)f * n_sites + js) * PS_WORDS + w;
    const int32_t v = s_st[f][w];
    if (w == PS_MIN) { if (v < *g) atomicMin(g, v); }
    else if (v > *g) atomicMax(g, v);
  }
}

// pass 2: writes every emitted (field, site) block; the plan entries of the plain fields follow those of the remapped
// fields (index (n_out + f) * n_sites + site).  One thread per (site, query row).
__global__ void __launch_bounds__(FMT_THREADS) k_plain_pack(const __grid_constant__ PlainSet P, const __grid_constant__ FmtPlanDev plan, const int n_out,
                                                            const int32_t* __restrict__ live, const uint32_t n_sites, const uint32_t nq, const int java) {
  const uint32_t js = blockIdx.y, q = blockIdx.x * FMT_THREADS + threadIdx.x;
  if (q >= nq) return;
  const int32_t e = live[(uint64_t)js * nq + q];
  const uint32_t cell = e >= 0 ? (uint32_t)(e & (LIVE_REDUCED - 1)) : 0u;
  for (int f = 0; f < P.n; ++f) {
    const uint64_t pi = (uint64_t)(n_out + f) * n_sites + js;
    const uint32_t type = plan.type[pi];
    if (type == BT_NONE) continue;
    const uint32_t nps = plan.nps[pi];
    const uint8_t* p = nullptr; uint32_t n = 0;
    const bool valid = e >= 0 && !((P.f[f].host_combine & HC_SKIP_REDUCED) && (e & LIVE_REDUCED)) && plain_fetch(P.f[f], cell, p, n);
    if (!valid) n = 0;
    uint8_t* dst = plan.bytes + plan.off[pi];
    const bool jv = java && !P.f[f].host_combine;
    if (type == BT_CHAR) {
      uint8_t* d = dst + (uint64_t)q * nps;
      if (P.f[f].host_combine & HC_RAW) {
        const uint32_t len = valid ? n : 0xFFFFFFFFu;
        for (uint32_t i = 0; i < 4u; ++i) d[i] = (uint8_t)(len >> (8u * i));
        for (uint32_t i = 4u; i < nps; ++i) d[i] = i - 4u < n ? p[i - 4u] : 0u;
        continue;
      }
      for (uint32_t i = 0; i < nps; ++i) d[i] = i < n ? p[i] : ((i == 0 && n == 0 && !jv) ? BCF_STR_MISSING : BCF_STR_VEND);   // JNI: vector_end only
      continue;
    }
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    const uint32_t miss = missing_word(P.f[f].elem == ELEM_FLOAT32), vend = vend_word(P.f[f].elem == ELEM_FLOAT32);
    for (uint32_t i = 0; i < nps; ++i) {
      const uint32_t v = i < n ? w[i] : (((i == 0 && n == 0) || jv) ? miss : vend);
      const uint64_t k = (uint64_t)q * nps + i;
      if (type == BT_INT8) dst[k] = v == BCF_INT_VEND ? 0x81u : (v == BCF_INT_MISSING ? 0x80u : (uint8_t)v);
      else if (type == BT_INT16) reinterpret_cast<uint16_t*>(dst)[k] = v == BCF_INT_VEND ? 0x8001u : (v == BCF_INT_MISSING ? 0x8000u : (uint16_t)v);
      else reinterpret_cast<uint32_t*>(dst)[k] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ INFO scalars
// BroadCombinedGVCFOperator::handle_INFO_fields (broad_combined_gvcf.cc:524-545) combines every scalar INFO field over the
// valid calls of the site: median = element n/2 of the sorted valid values (std::nth_element, variant_field_handler.cc:
// 529-560), sum = left-to-right sum in row order (:562-593; float addition is not associative, so the order is kept), mean
// = sum / n (:596-607).  A value counts when the call is live, is not in its spanning-deletion form (handle_deletions
// invalidates the INFO fields of such calls, broad_combined_gvcf.cc:1067-1075), the field is not the TileDB null and the
// value is neither BCF missing nor vector_end.  INFO/DP follows the GATK rule of broad_combined_gvcf.cc:688-727.
struct InfoSet { int32_t n; int32_t field[MAX_PLAIN]; int32_t op[MAX_PLAIN]; int32_t dp_info, dp_min_dp, dp_format; };   // field: index into PlainSet; dp_*: index or -1
enum { INFO_MEDIAN = 1, INFO_SUM = 2, INFO_MEAN = 3 };

__device__ inline bool info_value_ok(const PlainDev& F, uint32_t cell, uint32_t& v) {
  v = reinterpret_cast<const uint32_t*>(F.val)[(uint64_t)cell * (uint32_t)F.fixed_len];
  if (F.elem == ELEM_FLOAT32) return v != TDB_NULL_FLT && v != BCF_FLT_MISSING && v != BCF_FLT_VEND;
  return v != TDB_NULL_INT && v != BCF_INT_MISSING && v != BCF_INT_VEND;
}
// first element of plain field F of a call, or BCF missing when the field is invalid there (collect_and_extend_fields)
__device__ inline uint32_t scalar_or_missing(const PlainDev& F, bool call_valid, uint32_t cell) {
  if (!call_valid) return BCF_INT_MISSING;
  const uint32_t v = reinterpret_cast<const uint32_t*>(F.val)[(uint64_t)cell * (uint32_t)F.fixed_len];
  return v == TDB_NULL_INT ? BCF_INT_MISSING : v;
}

// one block per site: the valid values of every INFO field, compacted in row order, + the DP sum
__global__ void __launch_bounds__(FMT_THREADS) k_info_collect(const __grid_constant__ PlainSet P, const __grid_constant__ InfoSet I, const int32_t* __restrict__ live,
                                                              const uint32_t n_sites, const uint32_t nq, uint32_t* __restrict__ vals, uint32_t* __restrict__ cnt,
                                                              int32_t* __restrict__ dp_sum) {
  __shared__ uint32_t s_warp[FMT_THREADS / 32];
  __shared__ uint32_t s_base[MAX_PLAIN];
  __shared__ int32_t s_dp;
  const uint32_t js = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < MAX_PLAIN) s_base[tid] = 0;
  if (tid == 0) s_dp = 0;
  __syncthreads();
  int32_t dp = 0;
  for (uint32_t q0 = 0; q0 < nq; q0 += FMT_THREADS) {
    const uint32_t q = q0 + tid;
    const int32_t e = q < nq ? live[(uint64_t)js * nq + q] : -1;
    const bool call = e >= 0, plain_call = call && !(e & LIVE_REDUCED);
    const uint32_t cell = call ? (uint32_t)(e & (LIVE_REDUCED - 1)) : 0u;
    for (int k = 0; k < I.n; ++k) {
      uint32_t v = 0;
      const bool ok = plain_call && info_value_ok(P.f[I.field[k]], cell, v);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_warp[wid] = __popc(m);
      __syncthreads();
      uint32_t before = s_base[k], total = 0;
      for (uint32_t w = 0; w < FMT_THREADS / 32; ++w) { if (w < wid) before += s_warp[w]; total += s_warp[w]; }
      if (ok) vals[((uint64_t)k * n_sites + js) * nq + before + __popc(m & ((1u << lane) - 1u))] = v;
      __syncthreads();
      if (tid == 0) s_base[k] += total;
    }
    if (call && (I.dp_info >= 0 || I.dp_format >= 0)) {
      auto ok32 = [](uint32_t v) { return v != BCF_INT_MISSING && v != BCF_INT_VEND; };
      uint32_t a = I.dp_info >= 0 ? scalar_or_missing(P.f[I.dp_info], true, cell) : BCF_INT_MISSING;
      if (!ok32(a)) {
        const uint32_t mdp = I.dp_min_dp >= 0 ? scalar_or_missing(P.f[I.dp_min_dp], true, cell) : BCF_INT_MISSING;
        a = ok32(mdp) ? mdp : (I.dp_format >= 0 ? scalar_or_missing(P.f[I.dp_format], true, cell) : BCF_INT_MISSING);
      }
      if (ok32(a)) dp += (int32_t)a;
    }
  }
  dp = __reduce_add_sync(0xffffffffu, dp);
  if (lane == 0 && dp) atomicAdd(&s_dp, dp);
  __syncthreads();
  if (tid < (uint32_t)I.n) cnt[(uint64_t)tid * n_sites + js] = s_base[tid];
  if (tid == 0) dp_sum[js] = s_dp;
}

// order-preserving key of a 4-byte value: floats by value, ints by value
__device__ inline uint32_t info_key(uint32_t v, bool is_float) { return is_float ? ((v & 0x80000000u) ? ~v : (v | 0x80000000u)) : (v ^ 0x80000000u); }

// one block per (site, INFO field): median by radix select, sum / mean by one thread in row order
__global__ void __launch_bounds__(FMT_THREADS) k_info_reduce(const __grid_constant__ PlainSet P, const __grid_constant__ InfoSet I, const uint32_t n_sites, const uint32_t nq,
                                                             const uint32_t* __restrict__ vals, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ out_value,
                                                             uint8_t* __restrict__ out_valid) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_rank;
  const uint32_t js = blockIdx.x, k = blockIdx.y, tid = threadIdx.x;
  const uint64_t slot = (uint64_t)k * n_sites + js;
  const uint32_t n = cnt[slot];
  const uint32_t* v = vals + slot * nq;
  const bool is_float = P.f[I.field[k]].elem == ELEM_FLOAT32;
  if (n == 0) { if (tid == 0) { out_valid[slot] = 0; out_value[slot] = 0; } return; }
  if (I.op[k] != INFO_MEDIAN) {
    if (tid == 0) {
      uint32_t r;
      if (is_float) { float s = 0.0f; for (uint32_t i = 0; i < n; ++i) s = __fadd_rn(s, __uint_as_float(v[i])); if (I.op[k] == INFO_MEAN) s = __fdiv_rn(s, (float)n); r = __float_as_uint(s); }
      else { int32_t s = 0; for (uint32_t i = 0; i < n; ++i) s += (int32_t)v[i]; if (I.op[k] == INFO_MEAN) s = (int32_t)((uint32_t)s / n); r = (uint32_t)s; }
      out_value[slot] = r; out_valid[slot] = 1;
    }
    return;
  }
  // the key of rank n/2 (ascending), one byte per pass from the top
  if (tid == 0) { s_prefix = 0; s_rank = n / 2u; }
  for (int pass = 3; pass >= 0; --pass) {
    s_hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
    for (uint32_t i = tid; i < n; i += FMT_THREADS) {
      const uint32_t key = info_key(v[i], is_float);
      if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> (8 * pass)) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t r = s_rank, b = 0;
      for (; b < 256; ++b) { if (r < s_hist[b]) break; r -= s_hist[b]; }
      s_rank = r; s_prefix = prefix | (b << (8 * pass));
    }
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t key = s_prefix;
    out_value[slot] = is_float ? ((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key) : (key ^ 0x80000000u);
    out_valid[slot] = 1;
  }
}

// Element-wise sum of a vector field over the valid calls of a site (compute_valid_element_wise_sum, variant_field_handler.cc:
// 618-660): result[i] = sum, in row order, of the valid i-th values; missing where no call has one; as long as the last
// index any call has a valid value at.  Calls in spanning-deletion form do not count (their INFO fields were invalidated).
// One block per (site, slot); thread t owns elements t, t + 64, ... and walks the rows, so a float sum adds in row order.
constexpr int ESUM_THREADS = 64, ESUM_CAP = 1024;
struct EsumSet { int32_t n; int32_t remapped[MAX_FIELDS + MAX_PLAIN]; int32_t index[MAX_FIELDS + MAX_PLAIN]; };   // remapped ? field f : plain field p
__global__ void __launch_bounds__(ESUM_THREADS) k_info_esum(const __grid_constant__ ArrayDev A, const __grid_constant__ PlainSet P, const __grid_constant__ RemapOut O,
                                                            const __grid_constant__ EsumSet E, const int32_t* __restrict__ live, const int32_t* __restrict__ plain_stat,
                                                            const uint32_t n_sites, const uint32_t nq, uint32_t* __restrict__ values, uint32_t* __restrict__ lens,
                                                            uint32_t* __restrict__ err) {
  __shared__ uint32_t s_nvalid;
  const uint32_t js = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
  if (tid == 0) s_nvalid = 0;
  __syncthreads();
  const bool remapped = E.remapped[v] != 0;
  const int idx = E.index[v];
  const bool is_float = remapped ? A.fm[idx].elem != 0 : P.f[idx].elem == ELEM_FLOAT32;
  const uint32_t miss = is_float ? BCF_FLT_MISSING : BCF_INT_MISSING, vend = is_float ? BCF_FLT_VEND : BCF_INT_VEND;
  const int o = idx + (A.has_gt ? 1 : 0);
  const uint32_t stride = remapped ? O.f[o].stride[js] : 0u;
  const uint32_t* __restrict__ data = remapped ? O.f[o].data + O.f[o].base[js] : nullptr;
  const uint16_t* __restrict__ len = remapped ? O.f[o].len + (uint64_t)js * nq : nullptr;
  uint32_t* out = values + ((uint64_t)v * n_sites + js) * ESUM_CAP;
  // longest vector at the site: the field's stride (remapped) / the maximum k_plain_scan found (plain)
  uint32_t n_el = remapped ? stride : (uint32_t)plain_stat[((uint64_t)idx * n_sites + js) * PS_WORDS + PS_MAXLEN];
  if (n_el > (uint32_t)ESUM_CAP) { if (tid == 0) atomicMax(&err[1], 5u); n_el = ESUM_CAP; }
  for (uint32_t i = tid; i < n_el; i += ESUM_THREADS) {
    bool any = false; uint32_t acc = 0;
    for (uint32_t q = 0; q < nq; ++q) {
      const int32_t e = live[(uint64_t)js * nq + q];
      if (e < 0 || (e & LIVE_REDUCED)) continue;
      uint32_t x;
      if (remapped) { if (i >= len[q]) continue; x = data[(uint64_t)q * stride + i]; }
      else {
        const uint8_t* p; uint32_t n;
        if (!plain_fetch(P.f[idx], (uint32_t)e, p, n) || i >= n) continue;
        x = reinterpret_cast<const uint32_t*>(p)[i];
      }
      if (x == miss || x == vend) continue;
      if (!any) { acc = x; any = true; }
      else acc = is_float ? __float_as_uint(__fadd_rn(__uint_as_float(acc), __uint_as_float(x))) : acc + x;
    }
    out[i] = any ? acc : miss;
    if (any) atomicMax(&s_nvalid, i + 1u);
  }
  __syncthreads();
  if (tid == 0) lens[(uint64_t)v * n_sites + js] = s_nvalid;
}

}  // namespace gdb
