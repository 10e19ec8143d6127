// K5 on Blackwell's bulk-copy engine: k_remap_tma.
//
// The remapped output of one (site, field) is dense and sample-major: rows x stride words, contiguous.  A block owns
// TMA_ROWS consecutive samples of one site, builds THE SAME bytes in shared memory (one tile per field) and hands every
// tile to the copy engine with one cp.async.bulk.global.shared::cta (SASS UBLKCP): the threads never compute a global
// address for the 36 GB store stream and never wait for it -- the engine drains the tile while the SM's other resident
// blocks are in their gather / remap phases (tools/store_bench.cu: a block per SM sustains 6.4 TB/s this way).
//
//   phase A   one thread per sample: live entry -> 32-byte hot record (a reference block or <NON_REF>-only call is fully
//             described by it).  The thread owns its sample's ROW of every tile: a trivial call contributes at most three
//             distinct values per field -- genotype classes (0,0), (0,k), (j>0,k) -- and which class output word g takes
//             is a compile-time constant after unrolling, so a quad costs four selects and one STS.128.
//   phase C   the tile's other calls (nt_entry, gdb_kernels2.cuh: warm blob -> allele LUT -> vectors) write their rows
//             into the same tiles.
//   store     fence.proxy.async, barrier, one lane issues the bulk copies (cp.async.bulk.wait_group.read before the tiles
//             are written again).
//
// Tiles in shared memory cost occupancy (35 KB per 128 samples at 10 alleles: 4 blocks per SM), and what bounds this kernel
// is the latency of its gathers, not the store engine (profiles/r02_k5_tma_experiments.json), so a row is shared by
// TMA_TPR = 2 threads -- 8 warps per block, 32 per SM -- which split the row's quads; both read the same live entry and
// hot record (one sector, the second read hits L1).  The warm blobs of the tile's non-trivial calls arrive in shared
// memory through ONE bulk copy (cp.async.bulk.shared.global + mbarrier, SASS UBLKCP.S.G / SYNCS) issued before the fill:
// the calls that begin at a site's column are neighbours in cell order, so their blobs are one contiguous window of the
// pool -- the fragment bytes of the tile, TMA-staged.  Blobs outside the window (calls that began upstream) are read from HBM.
//
// Sites whose vectors do not fit (a field above 64 words, or more than TMA_SMEM_WORDS words per sample) keep going through
// k_remap_tile (gdb_kernels2.cuh).  Reference semantics: variant_operations.cc:233-263,572-695,
// variant_field_handler.cc:41-398 -- unchanged, the arithmetic is shared with k_remap_tile.
#pragma once
#include <utility>

#include "gdb_kernels4.cuh"

namespace gdb {

constexpr int TMA_ROWS = 128;
constexpr int TMA_TPR = 2;               // threads per row
constexpr int TMA_THREADS = TMA_ROWS * TMA_TPR;
constexpr int TMA_GROUPS = TMA_ROWS / NT_GL;   // phase C runs on the block's first TMA_ROWS threads (scratch rows cost shared memory)
constexpr int TMA_MAXQ = 16;            // quads (16 bytes) of one field vector: stride <= 64 words
constexpr int TMA_SMEM_WORDS = 128;     // words per sample over all fields: at most 64 KB of tiles per block
constexpr int TMA_JK = 4 * TMA_MAXQ + 8;
constexpr int TMA_WIN_UNITS = 256;      // warm-blob window staged per tile: 4 KB (about 25 calls of 3 ALT alleles)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store_g2s_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// class of word g of a diploid genotype-indexed vector: 0 -> (0,0), 1 -> (0,k), 2 -> (j>0,k)
__host__ __device__ constexpr uint32_t g2_class(uint32_t g) {
  uint32_t k = 0;
  while ((k + 1) * (k + 2) / 2 <= g) ++k;
  return k == 0 ? 0u : (g - k * (k + 1) / 2 == 0 ? 1u : 2u);
}
template <uint32_t C> __device__ __forceinline__ uint32_t pick3(uint32_t v0, uint32_t v1, uint32_t v2) { return C == 0 ? v0 : (C == 1 ? v1 : v2); }
template <int QD>
__device__ __forceinline__ bool g2_quad(uint4* srow, uint32_t lpr, uint32_t nfull, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t padw) {
  if ((uint32_t)QD >= lpr) return false;
  constexpr uint32_t g = 4u * QD;
  srow[QD] = make_uint4(g < nfull ? pick3<g2_class(g)>(v0, v1, v2) : padw, g + 1 < nfull ? pick3<g2_class(g + 1)>(v0, v1, v2) : padw,
                        g + 2 < nfull ? pick3<g2_class(g + 2)>(v0, v1, v2) : padw, g + 3 < nfull ? pick3<g2_class(g + 3)>(v0, v1, v2) : padw);
  return true;
}
// quads FIRST + 0 .. FIRST + sizeof...(QD) - 1 of a row
template <int FIRST, int... QD>
__device__ __forceinline__ void g2_row(std::integer_sequence<int, QD...>, uint4* srow, uint32_t lpr, uint32_t nfull, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t padw) {
  (void)(g2_quad<FIRST + QD>(srow, lpr, nfull, v0, v1, v2, padw) && ...);
}

template <int NF, bool GT>
__global__ void __launch_bounds__(TMA_THREADS, 4) k_remap_tma(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B,
                                                              const __grid_constant__ RemapOut O, const uint32_t* __restrict__ site_list, const TileTabs tabs) {
  extern __shared__ __align__(128) uint8_t s_tiles[];       // field f's tile at word offset TMA_ROWS * (stride[0] + .. + stride[f-1])
  __shared__ __align__(16) uint32_t s_blob[TMA_WIN_UNITS * 4];   // the tile's window of the warm-blob pool
  __shared__ __align__(8) unsigned long long s_bar;         // its mbarrier
  __shared__ uint32_t s_win_hi;                             // end of the window, in 16-byte units of the pool (0: none)
  __shared__ uint32_t s_hw[HOT_WORDS * TMA_ROWS];           // the samples' hot words, word i of row r at [i * TMA_ROWS + r]
  __shared__ uint2 s_ntl[TMA_ROWS];
  __shared__ uint32_t s_nnt;
  __shared__ __align__(8) uint16_t s_jk[TMA_JK];
  __shared__ unsigned long long s_ah[TILE_MAX_A];
  __shared__ int s_m2in[TMA_GROUPS][TILE_MAX_A];
  __shared__ int8_t s_in2m[TMA_GROUPS][TILE_MAX_A];
  constexpr int o_f0 = GT ? 1 : 0;
  const uint32_t js = site_list[blockIdx.y], tile = blockIdx.x;
  const uint32_t tid = threadIdx.x, row = tid % TMA_ROWS, half = tid / TMA_ROWS;
  const uint32_t nq = (uint32_t)S.nq;
  const uint32_t q0 = tile * TMA_ROWS, q = q0 + row;
  const uint64_t sq = (uint64_t)js * nq + q;
  if (tid == 0) { s_nnt = 0; s_win_hi = 0; mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // the two dependent gathers of this thread's sample first; everything up to their use overlaps their latency
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
  uint32_t* ftile[NF];   // this block's tile of field f (shared memory)
  {
    uint32_t words = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) { fstride[f] = O.f[o_f0 + f].stride[js]; ftile[f] = reinterpret_cast<uint32_t*>(s_tiles) + (size_t)TMA_ROWS * words; words += fstride[f]; }
  }
  for (uint32_t g = tid; g < (uint32_t)TMA_JK; g += TMA_THREADS) s_jk[g] = tabs.jk[g];   // g -> (k << 8 | j), the same for every site
  if (tid < A_m) s_ah[tid] = B.allele_hash[site_g->allele_first + tid];
  __syncthreads();   // s_nnt, s_win_hi, the barrier

  const bool dead = e < 0;
  const bool nt = !dead && ((e & LIVE_REDUCED) || ((h0.y >> 8) & 0xFF) == 0 || !(h0.x & F_TRIVIAL));
  uint32_t* const hw = s_hw + row;
  if (half == 0) {
    hw[0] = h0.w; hw[TMA_ROWS] = h1.x; hw[2 * TMA_ROWS] = h1.y; hw[3 * TMA_ROWS] = h1.z; hw[4 * TMA_ROWS] = h1.w;
    if (nt) {   // phase C of this block writes the row; its blob (h1.x words at 16-byte unit h0.w) joins the window
      s_ntl[atomicAdd(&s_nnt, 1u)] = make_uint2(q, (e & LIVE_REDUCED) ? (NT_REDUCED | (uint32_t)(e & (LIVE_REDUCED - 1))) : h0.w);
      if (!(e & LIVE_REDUCED)) atomicMax(&s_win_hi, h0.w + ((h1.x + 3u) >> 2));
    }
  }
  __syncthreads();
  const uint32_t win_hi = s_win_hi, win_lo = win_hi > (uint32_t)TMA_WIN_UNITS ? win_hi - (uint32_t)TMA_WIN_UNITS : 0u;
  if (tid == 0 && win_hi) { mbar_expect_tx(&s_bar, (win_hi - win_lo) * 16u); bulk_load(s_blob, A.warm + (size_t)win_lo * 4u, (win_hi - win_lo) * 16u, &s_bar); }

  // ---------------------------------------------------------------- phase A: two threads per sample, one row of every tile
  if (q < nq && !nt) {
    const uint32_t ploidy = (h0.x >> 8) & 0xFF, n_alt = (h0.x >> 16) & 0xFF, nr = h0.x >> 24, gt_len = h0.y & 0xFF;
    const uint32_t fl4 = (h0.y >> 16) | (h0.z << 16);
    const int in_nr = (nonref && nr) ? (int)nr : -1;
    if (half == 0) {
      if (GT) {
        const uint32_t gs = O.f[0].stride[js]; uint32_t* d = O.f[0].data + O.f[0].base[js] + (uint64_t)q * gs;
        const uint32_t step = A.gt_phased ? 2u : 1u;
        auto gt_word = [&](uint32_t i) -> uint32_t {
          if (dead || i >= gt_len) return BCF_INT_VEND;
          uint32_t g = hw[i * TMA_ROWS];
          if ((i % step) == 0 && !(g == TDB_NULL_INT || g == 0xFFFFFFFFu || g == BCF_INT_MISSING))
            g = (g == 0) ? 0u : ((nr && g == nr) ? A_m - 1u : (nonref ? A_m - 1u : 0xFFFFFFFFu));
          return g;
        };
        if (gs == 2) *reinterpret_cast<uint2*>(d) = make_uint2(gt_word(0), gt_word(1));   // site blocks are 16-byte aligned
        else for (uint32_t i = 0; i < gs; ++i) d[i] = gt_word(i);
        O.f[0].len[sq] = dead ? (uint16_t)0 : (uint16_t)gt_len;
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
    }
    uint32_t k = gt_len;   // fields follow GT in the record
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const uint32_t n = dead ? 0u : ((fl4 >> (8 * f)) & 0xFF);
      const int kind = A.fm[f].kind;
      const uint32_t miss = missing_word(A.fm[f].elem), padw = vend_word(A.fm[f].elem);
      uint32_t n_out = 0;
      if (n && !(kind == KIND_G && g_suppressed)) n_out = num_elements_le2(kind, A_m - 1u, ploidy);
      if (half == 0) O.f[o_f0 + f].len[sq] = (uint16_t)n_out;
      const uint32_t stride = fstride[f], lpr = stride >> 2;
      const uint32_t* fw = hw + k * TMA_ROWS;   // this field's inline words
      k += n;
      if (!stride) continue;
      uint32_t v0 = padw, v1 = padw, v2 = padw;
      if (n_out) {
        v0 = fw[0]; v1 = miss; v2 = miss;
        if (in_nr >= 0) {
          if (kind == KIND_A) { const uint32_t i1 = (uint32_t)in_nr - 1u; if (i1 < n) v1 = fw[i1 * TMA_ROWS]; }
          else if (kind == KIND_R || ploidy != 2) { if ((uint32_t)in_nr < n) v1 = fw[in_nr * TMA_ROWS]; }
          else {
            const uint32_t g1 = (uint32_t)in_nr * ((uint32_t)in_nr + 1u) / 2u, g2 = g1 + (uint32_t)in_nr;
            if (g1 < n) v1 = fw[g1 * TMA_ROWS];
            if (g2 < n) v2 = fw[g2 * TMA_ROWS];
          }
        }
      }
      uint4* srow = reinterpret_cast<uint4*>(ftile[f]) + (size_t)row * lpr;
      // the class pattern is the one of the site's ploidy for this field; a sample of another ploidy (a haploid call
      // among diploid ones ...) has a shorter vector with its own pattern
      const uint32_t fploidy = site_g->max_ploidy[f];
      if (n_out && kind == KIND_G && ploidy != fploidy) {
        if (half == 0) {
          uint32_t* d = reinterpret_cast<uint32_t*>(srow);
          for (uint32_t x = 0; x < stride; ++x) {
            uint32_t wv = padw;
            if (x < n_out) { const uint32_t c = trivial_class(kind, ploidy, x); wv = c == 0 ? v0 : (c == 1 ? v1 : v2); }
            d[x] = wv;
          }
        }
      } else {
        const uint32_t nfull = num_elements_le2(kind, A_m - 1u, fploidy);   // length of a full vector of this field at this site
        if (kind == KIND_G && fploidy == 2) {
          if (half == 0) g2_row<0>(std::make_integer_sequence<int, TMA_MAXQ / 2>{}, srow, lpr, nfull, v0, v1, v2, padw);
          else g2_row<TMA_MAXQ / 2>(std::make_integer_sequence<int, TMA_MAXQ / 2>{}, srow, lpr, nfull, v0, v1, v2, padw);
        } else {
          const uint32_t first = kind == KIND_A ? v1 : v0;   // R fields, haploid / ploidy-0 G fields: word 0 is REF's
          for (uint32_t qd = half; qd < lpr; qd += TMA_TPR) {
            const uint32_t g = 4u * qd;
            srow[qd] = make_uint4(g < nfull ? (g ? v1 : first) : padw, g + 1 < nfull ? v1 : padw, g + 2 < nfull ? v1 : padw, g + 3 < nfull ? v1 : padw);
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------- phase C: the tile's non-trivial calls
  const uint32_t n_ntl = s_nnt;   // final since the barrier before phase A
#ifdef GDB_TMA_NO_C
  if (false) {
#else
  if (n_ntl && tid < (uint32_t)TMA_ROWS && tid / 32u * (32u / NT_GL) < n_ntl) {   // warps none of whose groups has an entry go straight to the store barrier
#endif
    if (win_hi) mbar_wait(&s_bar, 0u);
    const NtSite nts{js, A_m, sflags, site_g->ref_len, B.allele_bytes + site_g->ref_off};
    NtDst nd{true, q0, {nullptr, nullptr, nullptr, nullptr}, win_hi ? s_blob : nullptr, win_lo};
#pragma unroll
    for (int f = 0; f < NF; ++f) nd.tile[f] = ftile[f];
    const int grp = (int)tid / NT_GL, gl = (int)tid % NT_GL;
    for (uint32_t base = 0; base < n_ntl; base += TMA_GROUPS) {
      if (base + (tid & ~31u) / NT_GL >= n_ntl) break;   // (warp-uniform)
      const bool active = base + grp < n_ntl;
      const uint2 ent = active ? s_ntl[base + grp] : make_uint2(0u, 0u);
      nt_entry<NF, GT>(A, S, B, O, nts, s_jk, s_ah, s_m2in[grp], s_in2m[grp], active, ent, gl, nd);
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- the tiles leave through the copy engine
  bulk_store_g2s_fence();   // this thread's shared-memory writes become visible to the async proxy
  __syncthreads();
  if (tid == 0) {
    const uint32_t rows = min((uint32_t)TMA_ROWS, nq - q0);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      if (!fstride[f]) continue;
      uint32_t* g = O.f[o_f0 + f].data + O.f[o_f0 + f].base[js] + (uint64_t)q0 * fstride[f];
      bulk_store(g, ftile[f], rows * fstride[f] * 4u);
    }
    bulk_commit();
    bulk_wait_read_all();   // shared memory must outlive the engine's reads
  }
}

}  // namespace gdb
