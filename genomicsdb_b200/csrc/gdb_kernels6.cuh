// K5, third form: k_remap_mix -- ONE launch whose blocks play one of two roles.
//
//   tile blocks   256 samples of one remapped site: phases A (one thread per sample: live entry -> 32-byte hot record ->
//                 the three distinct values a trivial call contributes per field) and B (the block streams the dense
//                 sample-major tile out, one 16-byte column per thread) of k_remap_tile.  A pure store stream.
//   nt blocks     64 non-trivial calls of one site (remapped or copy-through), four lanes per call: nt_entry, as in
//                 k_remap_nt.  A latency-bound gather chain with little traffic of its own.
//
// k_remap_tile ran the tile's non-trivial calls as a third phase of the same block (ncu, profiles/r02_k5_tile_source.txt): a
// quarter of the kernel's instructions, issued by 3 of the block's 8 warps while the other 5 had already left -- the SM's
// slots sat half empty (sm__warps_active 46 %) behind a dependent gather chain.  Here the two kinds of work are different
// BLOCKS of one grid, interleaved in block order in proportion to their numbers, so that every SM holds store-stream
// blocks and gather-chain blocks side by side all the time: the chain's latency hides behind the stores, every warp of an
// nt block has calls to work on, and a tile block is over as soon as its stores are issued.  The nt lists come from
// k_site_scan (B.nt), so the two roles do not depend on each other; a tile block skips the rows on its site's nt list.
//
// Phase B is also rewritten around its instruction count (the old loop spent 25 instructions per 16-byte store on index
// arithmetic): each thread keeps four shared-memory cursors (one per output word of its column: the class of a word is
// loop-invariant, a padding word's cursor points at one fixed word and never moves) and one global cursor, all advanced by
// loop-invariant steps; the row's "write me" flag travels in the fourth word of its value quad.
//
// Reference semantics: variant_operations.cc:233-263,572-695, variant_field_handler.cc:41-398 -- the arithmetic is the
// one of k_remap_tile / nt_entry (gdb_kernels2.cuh).
#pragma once
#include "gdb_kernels5.cuh"

namespace gdb {

// K4d: one permutation row per signature of a site.  Row (4-byte aligned, row_words words at sig_off[site] + slot * row_words):
// [4 x u16 output length of field f for a call that has the field] then, per field, `stride` bytes: for output word w the
// index of the input word that lands there, PERM_MISSING (no input allele: the missing value) or PERM_PAD (beyond the
// vector: BCF vector end).  The arithmetic is nt_entry's (variant_field_handler.cc:41-191), done once per signature instead
// of once per call.  One warp per occupied slot.
constexpr int SIGPERM_THREADS = 256;
__global__ void __launch_bounds__(SIGPERM_THREADS) k_sig_perm(const __grid_constant__ ArrayDev A, const __grid_constant__ BatchDev B, const __grid_constant__ RemapOut O,
                                                             const uint32_t* __restrict__ sig_off, uint32_t* __restrict__ sig_perm) {
  __shared__ int s_m2in[SIGPERM_THREADS / 32][256];
  const uint32_t js = blockIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t off = sig_off[js];
  if (off == 0xFFFFFFFFu) return;
  const SiteRec* __restrict__ site = &B.sites[js];
  const uint32_t A_m = site->n_alleles, sflags = site->flags;
  const bool nonref = sflags & 1u, g_suppressed = sflags & 4u;
  const int o_f0 = A.has_gt ? 1 : 0;
  uint32_t row_words = SIG_HDR_BYTES / 4;
  for (int f = 0; f < A.n_fields; ++f) row_words += O.f[o_f0 + f].stride[js] >> 2;
  int* m2in = s_m2in[warp];
  // one warp per occupied slot: merged -> input allele (the last input allele naming a merged allele wins, <NON_REF> stands in
  // for the rest), then one lane per output word
  for (uint32_t slot0 = warp * 32u; slot0 < B.sig_slots; slot0 += SIGPERM_THREADS) {   // 32 slots per load, then the occupied ones in turn
    const unsigned long long my_key = slot0 + lane < B.sig_slots ? B.sig_keys[(uint64_t)js * B.sig_slots + slot0 + lane] : SIG_EMPTY;
    uint32_t occ = __ballot_sync(0xFFFFFFFFu, my_key != SIG_EMPTY);
    while (occ) {
    const uint32_t src = __ffs(occ) - 1u; occ &= occ - 1u;
    const unsigned long long key = __shfl_sync(0xFFFFFFFFu, my_key, src);
    const uint32_t slot = slot0 + src;
    const uint32_t meta = (uint32_t)(key >> 48), n_alt = meta & 7u, nr_idx = (meta >> 3) & 7u, ploidy = meta >> 6;
    uint8_t* row = reinterpret_cast<uint8_t*>(sig_perm + off + (uint64_t)slot * row_words);
    __syncwarp();
    for (uint32_t m = lane; m < A_m; m += 32u) m2in[m] = m == 0 ? 0 : -1;
    __syncwarp();
    if (lane == 0) for (uint32_t t = 0; t < n_alt; ++t) { const uint32_t m = (uint32_t)(key >> (8 * t)) & 0xFFu; if (m != 0xFFu && m < A_m) m2in[m] = max(m2in[m], (int)t + 1); }
    __syncwarp();
    const int in_nr = (nonref && nr_idx) ? (int)nr_idx : -1;
    for (uint32_t m = lane; m < A_m; m += 32u) if (m2in[m] < 0) m2in[m] = in_nr;
    __syncwarp();
    uint8_t* pb = row + SIG_HDR_BYTES;
    for (int f = 0; f < A.n_fields && f < 4; ++f) {
      const int kind = A.fm[f].kind;
      const uint32_t stride = O.f[o_f0 + f].stride[js];
      const uint32_t n_out = (kind == KIND_G && g_suppressed) ? 0u : num_elements(kind, A_m - 1u, ploidy);
      if (lane == 0) reinterpret_cast<uint16_t*>(row)[f] = (uint16_t)n_out;
      for (uint32_t w = lane; w < stride; w += 32u) {
        uint32_t pv = PERM_PAD;
        if (w < n_out) {
          int ix = -1;
          if (kind == KIND_G && ploidy == 2) {
            uint32_t j, k; tri_unrank(w, j, k);
            const int aj = m2in[j], ak = m2in[k];
            if (aj >= 0 && ak >= 0) ix = aj > ak ? aj * (aj + 1) / 2 + ak : ak * (ak + 1) / 2 + aj;
          } else if (kind == KIND_G && ploidy == 0) ix = 0;
          else if (kind == KIND_A) { const int ij = m2in[w + 1u]; ix = ij < 0 ? -1 : ij - 1; }
          else ix = m2in[w];
          pv = (ix < 0 || ix > (int)PERM_MAX_IN) ? PERM_MISSING : (uint32_t)ix;
        }
        pb[w] = (uint8_t)pv;
      }
      pb += stride;
    }
    }
  }
}

// (Tried and dropped: band blocks writing their samples' signature-row calls themselves, as a third phase per site -- the calls'
// gather chain then sits between two block barriers of every site: 18.1 ms per step against 9.5 ms with the calls in nt blocks.)
#ifndef GDB_MIX_TRI_BUFS
#define GDB_MIX_TRI_BUFS 1   // value-quad buffers of a band block: 2 saves one barrier per site, 1 saves 4 KB x fields of shared memory (occupancy)
#endif
#ifndef GDB_MIX_BLOCKS
#define GDB_MIX_BLOCKS 5
#endif
#ifndef GDB_MIX_SPB
#define GDB_MIX_SPB 8
#endif
constexpr int MIX_SPB = GDB_MIX_SPB;   // consecutive sites a band block walks

// per-thread, per-field column plan of phase B (computed once per block, the same for every tile of the site)
constexpr uint32_t CFG_VALID = 1u << 16;   // r_off(8) | class of word 0..3 (2 bits each, 3 = padding) << 8 | valid << 16 | iterations of a full tile << 17

#ifndef GDB_MIX_DEPTH
#define GDB_MIX_DEPTH 2
#endif
constexpr int MIX_DEPTH = GDB_MIX_DEPTH;   // hot records in flight: tiles t .. t + MIX_DEPTH - 1

template <int NF>
struct MixSmem {
  // tile role
  static constexpr int TRI_WORDS = 4 * TILE_ROWS * NF;        // per (sample, field): v0, v1, v2, flag -- two buffers
  static constexpr int HW_WORDS = HOT_WORDS * TILE_ROWS;
  static constexpr int CFG_WORDS = TILE_ROWS * NF;
  static constexpr int FC_WORDS = 16 * NF;                    // two buffers; per field: lpr, rpi, pad word of the site's plan
  static constexpr int SITE_WORDS = 8 + 4 * NF;               // per site of the band: js, A_m, flags, GT stride / dst, sig_off, -, per field stride / ploidy / dst
  static constexpr int LIVE_WORDS = MIX_SPB * TILE_ROWS;      // the band's live entries, one bulk copy per site
  static constexpr int RING_WORDS = MIX_DEPTH * 2 * 4 * TILE_ROWS;   // hot records: [stage][half][sample] x 16 bytes
  static constexpr int TILE_BYTES = (RING_WORDS + LIVE_WORDS + GDB_MIX_TRI_BUFS * TRI_WORDS + HW_WORDS + CFG_WORDS + FC_WORDS + SITE_WORDS * MIX_SPB) * 4 + 16;   // + the mbarrier
  // nt role
  static constexpr int JK_BYTES = ((TILE_MAX_G + 6) * 2 + 15) & ~15;
  static constexpr int AH_BYTES = TILE_MAX_A * 8;
  static constexpr int M2IN_BYTES = NT_GROUPS * TILE_MAX_A * 4;
  static constexpr int IN2M_BYTES = NT_GROUPS * TILE_MAX_A;
  static constexpr int NT_BYTES = JK_BYTES + AH_BYTES + M2IN_BYTES + IN2M_BYTES;
  static constexpr int BYTES = TILE_BYTES > NT_BYTES ? TILE_BYTES : NT_BYTES;
};

__device__ __forceinline__ void ldgsts16(uint32_t sdst, const void* gsrc) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory"); }
__device__ __forceinline__ void ldgsts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ldgsts_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A call whose signature has a permutation row (k_sig_perm): the row's GT through the call's allele LUT (`rec`, gdb_kernels2.cuh:
// nt_lut), every field vector a gather through the row.  NT_GL lanes per call (gl = lane of the group); ent_y = warm-blob
// offset or NT_REDUCED | cell.  Same results as nt_entry (variant_operations.cc:233-263, variant_field_handler.cc:41-191).
struct SigSite { uint32_t js, A_m, sflags; const uint32_t* rows; };   // rows: the site's permutation rows (sig_perm + sig_off[js])
template <int NF, bool GT>
__device__ __forceinline__ void nt_apply(const ArrayDev& A, const ScanDev& S, const RemapOut& O, const SigSite& st, const uint32_t q, const uint32_t ent_y,
                                         const uint2 rec, const int gl) {
  constexpr int o_f0 = GT ? 1 : 0;
  const uint32_t js = st.js, A_m = st.A_m, slot = rec.y >> 16;
  const uint64_t sq = (uint64_t)js * S.nq + q;
  const bool nonref = st.sflags & 1u;
  uint32_t n_alt = 0, gt_len = 0;
  const uint32_t* gt = nullptr; const uint32_t* fin[NF]; uint32_t fn[NF];
  if (ent_y & NT_REDUCED) {
    const uint32_t cell = ent_y & ~NT_REDUCED;
    const DelRec* dr = find_delrec(S, cell);
    const CellDesc d = A.desc[cell];
    n_alt = dr->n_alt;
    gt_len = (d.flags & F_GTVALID) ? d.gt_len : 0u;
    gt = S.del_words + dr->gt_off;
#pragma unroll
    for (int f = 0; f < NF; ++f) { const uint32_t n = dr->f_len[f]; fn[f] = n == 0xFFFFFFFFu ? 0u : n; fin[f] = S.del_words + dr->f_off[f]; }
  } else {
    const WarmView wv = warm_view(A, ent_y);
    n_alt = wv.n_alt; gt_len = wv.gt_len;
    gt = wv.b + warm_hash_word((uint32_t)A.n_fields) + 3u * n_alt + wv.tok_words;
    const uint32_t* p = gt + gt_len;
#pragma unroll
    for (int f = 0; f < NF; ++f) { fn[f] = wv.b[4 + f]; fin[f] = p; p += fn[f]; }
  }
  if (GT) {   // variant_operations.cc:233-263
    const FieldOutDev fo = O.f[0];
    const uint32_t gs = fo.stride[js]; uint32_t* dst = fo.data + fo.base[js] + (uint64_t)q * gs;
    const uint32_t step = A.gt_phased ? 2u : 1u;
    const unsigned long long lut = ((unsigned long long)(rec.y & 0xFFFFu) << 32) | rec.x;
    for (uint32_t i = gl; i < gs; i += NT_GL) {
      uint32_t val = BCF_INT_VEND;
      if (i < gt_len) {
        const uint32_t g = gt[i]; val = g;
        if ((i % step) == 0 && !(g == TDB_NULL_INT || g == 0xFFFFFFFFu || g == BCF_INT_MISSING)) {
          uint32_t m = g == 0 ? 0u : ((g <= n_alt && g <= SIG_MAX_ALT) ? (uint32_t)(lut >> (8 * (g - 1u))) & 0xFFu : 0xFFu);
          val = m == 0xFFu ? (nonref ? A_m - 1u : 0xFFFFFFFFu) : m;
        }
      }
      dst[i] = val;
    }
    if (gl == 0 && (ent_y & NT_REDUCED)) fo.len[sq] = (uint16_t)gt_len;   // other calls: the band block wrote the lengths with the neighbours'
  }
  uint32_t row_words = SIG_HDR_BYTES / 4;
#pragma unroll
  for (int f = 0; f < NF; ++f) row_words += O.f[o_f0 + f].stride[js] >> 2;
  const uint32_t* row = st.rows + (uint64_t)slot * row_words;
  const uint2 hdr = *reinterpret_cast<const uint2*>(row);
  const uint32_t* prm = row + SIG_HDR_BYTES / 4;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const FieldOutDev fo = O.f[o_f0 + f];
    const uint32_t pad = vend_word(A.fm[f].elem), miss = missing_word(A.fm[f].elem);
    const uint32_t n = fn[f]; const uint32_t* in = fin[f];
    const uint32_t stride = fo.stride[js], lpr = stride >> 2;
    const uint32_t n_out = n ? (((f < 2 ? hdr.x : hdr.y) >> (16 * (f & 1))) & 0xFFFFu) : 0u;
    if (gl == 0 && (ent_y & NT_REDUCED)) fo.len[sq] = (uint16_t)n_out;
    uint4* dst4 = reinterpret_cast<uint4*>(fo.data + fo.base[js] + (uint64_t)q * stride);
    for (uint32_t qd = gl; qd < lpr; qd += NT_GL) {
      const uint32_t p4 = n ? prm[qd] : 0xFEFEFEFEu;
      uint32_t w[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const uint32_t pv = (p4 >> (8 * x)) & 0xFFu;
        uint32_t v = pv == PERM_PAD ? pad : miss;
        if (pv < n) v = in[pv];   // n <= PERM_MAX_IN at these sites
        w[x] = v;
      }
      dst4[qd] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    prm += lpr;
  }
}


// site_list / n_fast: the tile-path sites of the batch in site order; band blocks = tiles x ceil(n_fast / MIX_SPB), band-major
// (has_bands = 0: the grid holds nt blocks only)
template <int NF, bool GT>
__global__ void __launch_bounds__(TILE_ROWS, GDB_MIX_BLOCKS) k_remap_mix(const __grid_constant__ ArrayDev A, const __grid_constant__ ScanDev S, const __grid_constant__ BatchDev B,
                                                                       const __grid_constant__ RemapOut O, const uint32_t* __restrict__ site_list, const uint32_t n_fast, const uint32_t tiles,
                                                                       const uint32_t has_bands, const uint2* __restrict__ blk_map, const uint32_t n_nt_blocks,
                                                                       const uint32_t nt_ratio, const TileTabs tabs, const uint32_t* __restrict__ sig_off, const uint32_t* __restrict__ sig_perm) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int o_f0 = GT ? 1 : 0;
  const uint32_t tid = threadIdx.x;
  // block -> role: nt blocks are spread evenly over the grid; f(b) = b * nt_ratio / 2^32 grows by at most one per block, block b
  // is the f(b)-th nt block iff f(b + 1) > f(b) (nt_ratio = ceil(2^32 N / G), so every nt block gets its turn before the grid ends)
  const uint32_t f0 = has_bands ? __umulhi(blockIdx.x, nt_ratio) : blockIdx.x, f1 = has_bands ? __umulhi(blockIdx.x + 1u, nt_ratio) : blockIdx.x + 1u;
  const uint32_t nt_before = min(n_nt_blocks, f0);   // nt blocks among the blocks before this one
  if (f0 < n_nt_blocks && f1 > f0) {
    // ------------------------------------------------------------------------------------------------ nt block
    uint16_t* s_jk = reinterpret_cast<uint16_t*>(smem);
    unsigned long long* s_ah = reinterpret_cast<unsigned long long*>(smem + MixSmem<NF>::JK_BYTES);
    int (*s_m2in)[TILE_MAX_A] = reinterpret_cast<int (*)[TILE_MAX_A]>(smem + MixSmem<NF>::JK_BYTES + MixSmem<NF>::AH_BYTES);
    int8_t (*s_in2m)[TILE_MAX_A] = reinterpret_cast<int8_t (*)[TILE_MAX_A]>(smem + MixSmem<NF>::JK_BYTES + MixSmem<NF>::AH_BYTES + MixSmem<NF>::M2IN_BYTES);
    const int grp = (int)tid / NT_GL, gl = (int)tid % NT_GL;
    const uint2 bm = blk_map[nt_before];
    const SiteRec* __restrict__ site_g = &B.sites[bm.x];
    const NtSite st{bm.x, site_g->n_alleles, site_g->flags, site_g->ref_len, B.allele_bytes + site_g->ref_off};
    const uint32_t ei = bm.y + grp;
    const bool active = ei < site_g->n_nt;
    uint2 ent = make_uint2(0, 0);
    if (active) ent = B.nt[(uint64_t)st.js * S.nq + ei];
    if (tid < st.A_m) s_ah[tid] = B.allele_hash[site_g->allele_first + tid];
    const uint32_t jk_n = (st.A_m * (st.A_m + 1u)) >> 1;
    for (uint32_t g = tid; g < ((jk_n + 3u) & ~3u) && g < (uint32_t)TILE_MAX_G + 6u; g += TILE_ROWS) s_jk[g] = tabs.jk[g];   // whole quads: the tail is read, then discarded
    // ---- calls whose signature has a permutation row (k_sig_perm): a gather through the row instead of a LUT of their own
    // (nt_apply); the others -- copy-through sites, long ALT lists, sites without a signature table -- rebuild theirs (nt_entry)
    const uint32_t soff = sig_off ? sig_off[st.js] : 0xFFFFFFFFu;
    uint2 rec = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    if (active && soff != 0xFFFFFFFFu && !O.lut_first) rec = B.nt_lut[(uint64_t)st.js * S.nq + ei];
    const bool fast = active && soff != 0xFFFFFFFFu && !O.lut_first && (rec.y >> 16) != SIG_NONE;
    __syncthreads();
    if (__any_sync(0xFFFFFFFFu, active && !fast)) nt_entry<NF, GT>(A, S, B, O, st, s_jk, s_ah, s_m2in[grp], s_in2m[grp], active && !fast, ent, gl);
    if (fast) { const SigSite ss{st.js, st.A_m, st.sflags, sig_perm + soff}; nt_apply<NF, GT>(A, S, O, ss, ent.x, ent.y, rec, gl); }
    return;
  }

  // -------------------------------------------------------------------------------------------------- band block
  // 256 samples x MIX_SPB consecutive sites of the tile path (remapped and copy-through alike, in site order).  A sample's
  // live call is the same cell at most neighbouring sites (a reference block spans many of them), so walking the SITES
  // inside the block turns the per-(site, sample) gather of the 32-byte hot record into one gather per distinct cell.
  uint4* const s_ring = reinterpret_cast<uint4*>(smem);                      // [(stage * 2 + half) * TILE_ROWS + sample]
  int32_t* const s_live = reinterpret_cast<int32_t*>(smem) + MixSmem<NF>::RING_WORDS;      // [site * TILE_ROWS + sample]
  uint32_t* const s_tri = reinterpret_cast<uint32_t*>(s_live) + MixSmem<NF>::LIVE_WORDS;   // buffer b: [b * TRI_WORDS + 4 * (row * NF + f) + {v0, v1, v2, flag}]
  uint32_t* const s_hw = s_tri + GDB_MIX_TRI_BUFS * MixSmem<NF>::TRI_WORDS;                 // the samples' hot words, word i of thread t at [i * TILE_ROWS + t]
  uint32_t* const s_cfg = s_hw + MixSmem<NF>::HW_WORDS;                      // [f * TILE_ROWS + t]
  uint32_t* const s_fc = s_cfg + MixSmem<NF>::CFG_WORDS;                     // [8 * NF * buf + 8 * f + {lpr, rpi, pad}] of the site's plan
  uint32_t* const s_site = s_fc + MixSmem<NF>::FC_WORDS;                     // [SITE_WORDS * k + ...] parameters of the band's sites
  unsigned long long* const s_bar = reinterpret_cast<unsigned long long*>(s_site + MixSmem<NF>::SITE_WORDS * MIX_SPB);
  constexpr int SW = MixSmem<NF>::SITE_WORDS;   // js, A_m, flags, gt stride, gt dst (2), sig_off, -, then per field: stride, site ploidy, dst (2)
  const uint32_t tb = blockIdx.x - nt_before;
  const uint32_t band = tb / tiles, tile = tb - band * tiles;
  const uint32_t k0 = band * MIX_SPB, n_k = min((uint32_t)MIX_SPB, n_fast - k0);
  const uint32_t nq = (uint32_t)S.nq;
  const uint32_t q = tile * TILE_ROWS + tid;
  const uint32_t rows = min((uint32_t)TILE_ROWS, nq - tile * TILE_ROWS);
  // ---- the band's live entries: one bulk copy (TMA engine) per site of the contiguous piece of its live-matrix row, all
  // signalled on one mbarrier; rows that are not 16-byte aligned (and the last entries) come through plain loads.  When every
  // site of the batch is on the tile path (site_list == nullptr: site k of the list IS site k) the copies start at once,
  // side by side with the loads of the sites' parameters.
  auto js_of = [&](uint32_t k) -> uint32_t { return site_list ? site_list[k0 + k] : k0 + k; };
  const bool bulk_ok = (((uintptr_t)B.live & 15u) == 0) && ((nq & 3u) == 0);
  const uint32_t bulk_words = bulk_ok ? (rows & ~3u) : 0u;
  if (tid == 0) {
    mbar_init(s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bulk_words) {
      mbar_expect_tx(s_bar, bulk_words * 4u * n_k);
      for (uint32_t k = 0; k < n_k; ++k) bulk_load(s_live + k * TILE_ROWS, B.live + (uint64_t)js_of(k) * nq + (uint64_t)tile * TILE_ROWS, bulk_words * 4u, s_bar);
    }
  }
  if (tid >= 32u && tid < 32u + n_k) {   // a lane of the second warp per site (the first one is busy issuing the copies)
    const uint32_t k = tid - 32u, js = js_of(k);
    const SiteRec* __restrict__ sg = &B.sites[js];
    uint32_t* w = s_site + SW * k;
    w[0] = js; w[1] = sg->n_alleles; w[2] = sg->flags; w[6] = (sig_off && !O.lut_first) ? sig_off[js] : 0xFFFFFFFFu;
    if (GT) { uint32_t* d = O.f[0].data + O.f[0].base[js]; w[3] = O.f[0].stride[js]; w[4] = (uint32_t)(uintptr_t)d; w[5] = (uint32_t)((uintptr_t)d >> 32); }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      uint32_t* d = O.f[o_f0 + f].data + O.f[o_f0 + f].base[js];
      w[8 + 4 * f] = O.f[o_f0 + f].stride[js]; w[9 + 4 * f] = sg->max_ploidy[f]; w[10 + 4 * f] = (uint32_t)(uintptr_t)d; w[11 + 4 * f] = (uint32_t)((uintptr_t)d >> 32);
    }
  }
  if (bulk_words < (uint32_t)TILE_ROWS)
    for (uint32_t k = 0; k < n_k; ++k)
      for (uint32_t i = bulk_words + tid; i < (uint32_t)TILE_ROWS; i += TILE_ROWS) s_live[k * TILE_ROWS + i] = i < rows ? B.live[(uint64_t)js_of(k) * nq + (uint64_t)tile * TILE_ROWS + i] : -1;
  uint32_t* const hw = s_hw + tid;
  const uint32_t sbase = smem_addr(s_tri);
  __syncthreads();   // the mbarrier is initialised; s_site and the plainly loaded live entries are in place
  if (bulk_words) mbar_wait(s_bar, 0);
  // ---- hot records: asynchronous 16-byte copies into a ring (LDGSTS), one group per site, MIX_DEPTH sites ahead -- and only
  // when the sample's live cell differs from the one of the site before
  const uint32_t ring0 = smem_addr(s_ring) + 16u * tid;
  auto live_at = [&](uint32_t k) -> int32_t { return s_live[k * TILE_ROWS + tid]; };
  auto fetch_hot = [&](uint32_t k) {
    if (k < n_k) {
      const int32_t ee = live_at(k);
      if (ee >= 0 && !(ee & LIVE_REDUCED) && (k == 0 || ee != live_at(k - 1))) {
        const uint32_t st = ring0 + (k % MIX_DEPTH) * (2u * 16u * TILE_ROWS);
        const char* src = reinterpret_cast<const char*>(&A.hot[ee]);
        ldgsts16(st, src); ldgsts16(st + 16u * TILE_ROWS, src + 16);
      }
    }
    ldgsts_commit();
  };
#pragma unroll
  for (int d = 0; d < MIX_DEPTH; ++d) fetch_hot(d);
  uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
  int32_t e_prev = -1;
  uint32_t buf = 0, plan_key[NF];
  bool any_remapped = false;
#pragma unroll
  for (int f = 0; f < NF; ++f) plan_key[f] = 0xFFFFFFFFu;

  for (uint32_t k = 0; k < n_k; ++k) {
    const uint32_t* sw = s_site + SW * k;
    const uint32_t js = sw[0], A_m = sw[1], sflags = sw[2];
    const bool nonref = sflags & 1u, g_suppressed = sflags & 4u, remapped = sflags & 2u;
    const uint64_t sq = (uint64_t)js * nq + q;
    ldgsts_wait<MIX_DEPTH - 1>();   // this site's group has landed (each thread reads only what it copied itself)
    const int32_t e = live_at(k);
    if (e < 0 || (e & LIVE_REDUCED)) { h0 = make_uint4(0, 0, 0, 0); h1 = make_uint4(0, 0, 0, 0); }
    else if (k == 0 || e != e_prev) {
      const uint4* st = s_ring + (k % MIX_DEPTH) * (2 * TILE_ROWS) + tid;
      h0 = st[0]; h1 = st[TILE_ROWS];
      hw[0] = h0.w; hw[TILE_ROWS] = h1.x; hw[2 * TILE_ROWS] = h1.y; hw[3 * TILE_ROWS] = h1.z; hw[4 * TILE_ROWS] = h1.w;
    }
    e_prev = e;
    fetch_hot(k + MIX_DEPTH);    // the stage is free again: its record sits in h0 / h1
    const uint32_t gt_stride = GT ? sw[3] : 0u;
    uint32_t* const gt_dst = GT ? reinterpret_cast<uint32_t*>((uintptr_t)sw[4] | ((uintptr_t)sw[5] << 32)) : nullptr;
    const uint32_t hflags = h0.x & 0xFF, ploidy = (h0.x >> 8) & 0xFF, n_alt = (h0.x >> 16) & 0xFF, nr = h0.x >> 24, gt_len = h0.y & 0xFF;
    const bool inl = ((h0.y >> 8) & 0xFF) != 0;
    const bool on_nt_list = e >= 0 && ((e & LIVE_REDUCED) || !inl || !(hflags & F_TRIVIAL));   // an nt block writes the row
    const uint32_t fl4 = (h0.y >> 16) | (h0.z << 16);

    if (!remapped) {
      // ---------------------------------------------------------------- copy-through site (merged REF of length 1, <NON_REF> the only ALT)
      if (q < nq && !on_nt_list) {
        if (GT) {
          uint32_t* d = gt_dst + (uint64_t)q * gt_stride;
          if (gt_stride == 2) *reinterpret_cast<uint2*>(d) = make_uint2(gt_len > 0 ? hw[0] : BCF_INT_VEND, gt_len > 1 ? hw[TILE_ROWS] : BCF_INT_VEND);
          else for (uint32_t i = 0; i < gt_stride; ++i) d[i] = i < gt_len ? hw[i * TILE_ROWS] : BCF_INT_VEND;
          O.f[0].len[sq] = (uint16_t)gt_len;
        }
        if (O.lut_first) {
          if (e < 0) { O.lut_first[sq] = 0xFFFFFFFFu; O.lut_len[sq] = 0; }
          else {
            const uint32_t A_in = n_alt + 1u;
            uint32_t first = atomicAdd(&B.cursors[2], A_in);
            if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; }
            O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in;
            if (first != 0xFFFFFFFFu) for (uint32_t i = 0; i < A_in; ++i) O.lut_in2m[first + i] = (int16_t)(i == 0 ? 0 : ((nr && i == nr) ? (int)A_m - 1 : -1));
          }
        }
        uint32_t kk = gt_len;   // fields follow GT in the record
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const uint32_t n = (fl4 >> (8 * f)) & 0xFF;   // 0 for a dead row
          const uint32_t padw = vend_word(A.fm[f].elem);
          const uint32_t stride = sw[8 + 4 * f];
          O.f[o_f0 + f].len[sq] = (uint16_t)n;
          const uint32_t* fw = hw + kk * TILE_ROWS;
          uint32_t* d = reinterpret_cast<uint32_t*>((uintptr_t)sw[10 + 4 * f] | ((uintptr_t)sw[11 + 4 * f] << 32)) + (uint64_t)q * stride;
          for (uint32_t x = 0; x < stride; ++x) d[x] = x < n ? fw[x * TILE_ROWS] : padw;
          kk += n;
        }
      }
      continue;
    }

    // ------------------------------------------------------------------ remapped site
    if (GDB_MIX_TRI_BUFS == 1 && any_remapped) __syncthreads();   // every thread is done streaming the remapped site before out of the one buffer
    any_remapped = true;
    // ---- per-field plan of phase B; kept while the sites' layouts stay the same
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const uint32_t stride = sw[8 + 4 * f], fploidy = sw[9 + 4 * f];
      const uint32_t key = stride | (A_m << 12) | (fploidy << 20);
      if (key != plan_key[f]) {   // block-uniform
        plan_key[f] = key;
        uint32_t cfg = 0;
        uint32_t lpr = stride >> 2, rpi = 0;   // quads per sample (stride is a multiple of 4 at remapped sites), rows per step
        if (stride) {
          rpi = (uint32_t)TILE_ROWS / lpr;
          const uint32_t r_off = tid / lpr, col = tid - r_off * lpr;
          if (r_off < rpi) {
            const int kind = A.fm[f].kind;
            const uint32_t nfull = num_elements_le2(kind, A_m - 1u, fploidy);   // length of a full vector of this field at this site
            const uint32_t cw = (kind == KIND_A) ? 0x01010101u : ((kind == KIND_G && fploidy == 2) ? tabs.cls[col] : (col == 0 ? 0x01010100u : 0x01010101u));
            cfg = r_off | CFG_VALID | (((TILE_ROWS - r_off + rpi - 1u) / rpi) << 17);
#pragma unroll
            for (uint32_t x = 0; x < 4; ++x) cfg |= ((4 * col + x < nfull) ? ((cw >> (8 * x)) & 3u) : 3u) << (8 + 2 * x);
          }
        }
        s_cfg[f * TILE_ROWS + tid] = cfg;   // read back by this thread only
      }
      if (tid == 0) {   // read by the block after the site's barrier (double-buffered like the value quads)
        const uint32_t lpr = stride >> 2;
        uint32_t* fc = s_fc + 8 * NF * buf + 8 * f;
        fc[0] = lpr; fc[1] = lpr ? (uint32_t)TILE_ROWS / lpr : 0u; fc[2] = vend_word(A.fm[f].elem);
      }
    }
    // ---------------------------------------------------------------- phase A: one thread per sample
    uint4* const tri4 = reinterpret_cast<uint4*>(s_tri + buf * MixSmem<NF>::TRI_WORDS) + tid * NF;
    if (q < nq) {
      if (e < 0) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const uint32_t padw = vend_word(A.fm[f].elem);
          tri4[f] = make_uint4(padw, padw, padw, INFO_FAST);
          O.f[o_f0 + f].len[sq] = 0;
        }
        if (GT) { uint32_t* d = gt_dst + (uint64_t)q * gt_stride; for (uint32_t i = 0; i < gt_stride; ++i) d[i] = BCF_INT_VEND; O.f[0].len[sq] = 0; }
        if (O.lut_first) { O.lut_first[sq] = 0xFFFFFFFFu; O.lut_len[sq] = 0; }
      } else if (on_nt_list) {
#pragma unroll
        for (int f = 0; f < NF; ++f) tri4[f] = make_uint4(0u, 0u, 0u, 0u);
        if (!(e & LIVE_REDUCED)) {
          // the row's vectors come from an nt block, but its lengths are known here (the hot record has the call's ploidy and which
          // fields it carries): written with the neighbours' they leave L2 as whole sectors, where 2-byte stores from an nt block
          // would make every one of them a read-modify-write in DRAM
          if (GT) O.f[0].len[sq] = (uint16_t)gt_len;
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            const uint32_t n = (fl4 >> (8 * f)) & 0xFF;
            const int kind = A.fm[f].kind;
            O.f[o_f0 + f].len[sq] = (uint16_t)((n && !(kind == KIND_G && g_suppressed)) ? num_elements_le2(kind, A_m - 1u, ploidy) : 0u);
          }
        }
      } else {
        // ---- trivial call, fully described by its hot record
        const int in_nr = (nonref && nr) ? (int)nr : -1;
        if (GT) {
          uint32_t* d = gt_dst + (uint64_t)q * gt_stride;
          const uint32_t step = A.gt_phased ? 2u : 1u;
          auto gt_word = [&](uint32_t i) -> uint32_t {
            if (i >= gt_len) return BCF_INT_VEND;
            uint32_t g = hw[i * TILE_ROWS];
            if ((i % step) == 0 && !(g == TDB_NULL_INT || g == 0xFFFFFFFFu || g == BCF_INT_MISSING))
              g = (g == 0) ? 0u : ((nr && g == nr) ? A_m - 1u : (nonref ? A_m - 1u : 0xFFFFFFFFu));
            return g;
          };
          if (gt_stride == 2) *reinterpret_cast<uint2*>(d) = make_uint2(gt_word(0), gt_word(1));   // site blocks are 16-byte aligned
          else for (uint32_t i = 0; i < gt_stride; ++i) d[i] = gt_word(i);
          O.f[0].len[sq] = (uint16_t)gt_len;
        }
        if (O.lut_first) {
          const uint32_t A_in = n_alt + 1u;
          uint32_t first = atomicAdd(&B.cursors[2], A_in);
          if (first + A_in > O.cap_lut) { atomicOr(&S.err[2], 2u); first = 0xFFFFFFFFu; }
          O.lut_first[sq] = first; O.lut_len[sq] = (uint16_t)A_in;
          if (first != 0xFFFFFFFFu) for (uint32_t i = 0; i < A_in; ++i) O.lut_in2m[first + i] = (int16_t)(i == 0 ? 0 : ((nr && i == nr) ? (int)A_m - 1 : -1));
        }
        uint32_t kk = gt_len;   // fields follow GT in the record
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const uint32_t n = (fl4 >> (8 * f)) & 0xFF;
          const int kind = A.fm[f].kind;
          const uint32_t miss = missing_word(A.fm[f].elem), padw = vend_word(A.fm[f].elem);
          uint32_t n_out = 0;
          if (n && !(kind == KIND_G && g_suppressed)) n_out = num_elements_le2(kind, A_m - 1u, ploidy);
          O.f[o_f0 + f].len[sq] = (uint16_t)n_out;
          const uint32_t* fw = hw + kk * TILE_ROWS;   // this field's inline words
          if (!n_out) {
            tri4[f] = make_uint4(padw, padw, padw, INFO_FAST);
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
            const bool fast = (kind != KIND_G) || (ploidy == sw[9 + 4 * f]);
            tri4[f] = make_uint4(v0, v1, v2, fast ? INFO_FAST : 0u);
            if (!fast) {
              const uint32_t fstride = sw[8 + 4 * f];
              uint32_t* d = reinterpret_cast<uint32_t*>((uintptr_t)sw[10 + 4 * f] | ((uintptr_t)sw[11 + 4 * f] << 32)) + (uint64_t)q * fstride;
              for (uint32_t x = 0; x < fstride; ++x) {
                uint32_t w = padw;
                if (x < n_out) { const uint32_t c = trivial_class(kind, ploidy, x); w = c == 0 ? v0 : (c == 1 ? v1 : v2); }
                d[x] = w;
              }
            }
          }
          kk += n;
        }
      }
    }
    __syncthreads();   // the site's value quads and s_fc are complete

    // ---------------------------------------------------------------- phase B: the block streams the tile out
    // Field f of the tile is rows x lpr quads, contiguous.  Thread t keeps ONE quad column (t % lpr) and walks the rows
    // (t / lpr, + 256 / lpr per step), so which of (v0, v1, v2, pad) each of its four words takes is loop-invariant, and
    // a warp's stores cover 512 contiguous bytes.
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const uint32_t cfg = s_cfg[f * TILE_ROWS + tid];
      if (!(cfg & CFG_VALID)) continue;
      const uint32_t* fc = s_fc + 8 * NF * buf + 8 * f;
      const uint32_t lpr = fc[0], rpi = fc[1];
      const uint32_t r_off = cfg & 0xFFu, col = tid - r_off * lpr;
      const uint32_t step_b = 16u * NF * rpi;
      uint32_t row_at = sbase + 4u * (buf * MixSmem<NF>::TRI_WORDS) + 16u * (r_off * NF + f);   // this thread's first row: (v0, v1, v2, flag)
      const uint32_t padw = fc[2];
      // class of each of the column's four words (0: v0, 1: v1, 2: v2, 3: padding), loop-invariant
      const bool l0 = cfg & (1u << 8), h0c = cfg & (1u << 9), l1 = cfg & (1u << 10), h1c = cfg & (1u << 11);
      const bool l2 = cfg & (1u << 12), h2c = cfg & (1u << 13), l3 = cfg & (1u << 14), h3c = cfg & (1u << 15);
      uint4* out = reinterpret_cast<uint4*>((uintptr_t)sw[10 + 4 * f] | ((uintptr_t)sw[11 + 4 * f] << 32)) + ((uint64_t)(tile * TILE_ROWS + r_off) * lpr + col);
      const uint32_t ostep = rpi * lpr;
      uint32_t n_it = cfg >> 17;
      if (rows < (uint32_t)TILE_ROWS) n_it = r_off < rows ? (rows - r_off + rpi - 1u) / rpi : 0u;   // rows r_off, r_off + rpi, ... < rows
      // one 16-byte shared load per 16-byte store, the words picked by selects
#pragma unroll 4
      for (uint32_t it = 0; it < n_it; ++it) {
        uint32_t tx, ty, tz, tw;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(tx), "=r"(ty), "=r"(tz), "=r"(tw) : "r"(row_at));
        const uint32_t w0 = l0 ? (h0c ? padw : ty) : (h0c ? tz : tx);
        const uint32_t w1 = l1 ? (h1c ? padw : ty) : (h1c ? tz : tx);
        const uint32_t w2 = l2 ? (h2c ? padw : ty) : (h2c ? tz : tx);
        const uint32_t w3 = l3 ? (h3c ? padw : ty) : (h3c ? tz : tx);
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p st.global.v4.u32 [%0], {%1, %2, %3, %4}; }" ::"l"(out), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(tw & INFO_FAST) : "memory");
        row_at += step_b; out += ostep;
      }
    }
    if (GDB_MIX_TRI_BUFS == 2) buf ^= 1u;
  }
}

}  // namespace gdb
