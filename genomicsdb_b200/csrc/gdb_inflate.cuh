// K0: DEFLATE on the device, one WARP per zlib stream.  GenomicsDB arrays are usually written with "compress_tiledb_array":
// every tile of every attribute file is one zlib stream (ReadState::decompress_tile_gzip -> gunzip, dependencies/TileDB/core/
// src/fragment/read_state.cc:1595-1702, misc/utils.cc:466-512).  A fragment has (cells / capacity) tiles per attribute --
// hundreds of thousands of independent streams of very different sizes (a coordinates tile inflates to 16 KB, a PL tile of a
// variant-dense stretch to hundreds of KB), so the time of a launch is the time of its longest stream: the decoder has to be
// fast WITHIN a stream, not only across streams.
//
//   * The 32 lanes run the symbol decoder in lock step on the same bit buffer (uniform control flow, no divergence, no
//     broadcasts): every table lookup is one shared-memory broadcast read.  Tables per warp in shared memory: a 10-bit primary
//     table for the literal/length code, an 8-bit one for the distance code, and the canonical count[] / symbol[] arrays
//     (RFC 1951 3.2.2; the structure of zlib's contrib/puff/puff.c, a public-domain reference decoder that is not part of
//     /root/reference) for the rare longer codes.
//   * The compressed bytes are staged through shared memory 256 bytes at a time by coalesced loads, one block ahead.
//   * The most recent 2 KB of output (GDB_INFLATE_WIN; 8 KB and 4 KB rings were slower: 20.0 / 16.4 / 14.2 ms on the open benchmark,
//     occupancy 16 / 24 / 32 warps per SM) live in a shared-memory ring.  An LZ77 match is copied by the whole warp (258 bytes = 9
//     strided trips, overlapping matches by index arithmetic), from the ring when its source is still there, from the bytes
//     already flushed to HBM otherwise; the ring leaves for HBM in 512-byte pieces, 16 bytes per lane, aligned to the output
//     address (the ring index is skewed by the output pointer's misalignment so that both sides are 16-byte aligned).
//   * Adler-32 (RFC 1950) is folded in at flush time: a chunk of n bytes adds (sum d_i) to a and n a + sum (n - i) d_i to b, two
//     warp reductions per chunk.
// The launch takes the streams in descending order of their inflated size, so the long ones start first.
// The host never sees the decompressed bytes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gdb {

struct InflateJob { const uint8_t* in; uint64_t in_len; uint8_t* out; uint64_t out_len; };   // one zlib stream -> one tile (device pointers)
enum : uint32_t { INF_OK = 0, INF_ERR_HEADER = 1, INF_ERR_DATA = 2, INF_ERR_SIZE = 3, INF_ERR_ADLER = 4 };

namespace inflate_detail {

constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, FIXLCODES = 288;
constexpr int LUT_BITS = 10, DLUT_BITS = 8;
#ifndef GDB_INFLATE_WIN
#define GDB_INFLATE_WIN 2048
#endif
constexpr uint32_t WIN = GDB_INFLATE_WIN, WIN_MASK = WIN - 1;   // ring of recent output (bytes)
constexpr uint32_t RING_REACH = WIN - (WIN >= 8192 ? 1024 : 512);           // a match this far back or nearer is still in the ring whatever is being written
constexpr uint32_t FLUSH_AT = WIN >= 8192 ? 2048 : 1024;    // unflushed bytes that trigger a flush
constexpr int STAGE_WORDS = 128;                      // staged input: two blocks of 64 words

struct alignas(16) WarpMem {
  uint8_t win[WIN];
  uint32_t stage[STAGE_WORDS];
  uint16_t lut[1 << LUT_BITS];     // next LUT_BITS bits -> symbol << 4 | code length (0: longer code)
  uint16_t dlut[1 << DLUT_BITS];
  int16_t lcount[MAXBITS + 1], lsymbol[FIXLCODES], dcount[MAXBITS + 1], dsymbol[MAXDCODES];
  int16_t lengths[MAXLCODES + MAXDCODES + 4];
  uint16_t code[FIXLCODES];
};
#ifndef GDB_INFLATE_WARPS
#define GDB_INFLATE_WARPS 4
#endif
constexpr int WARPS_PER_BLOCK = GDB_INFLATE_WARPS;

__device__ __constant__ uint16_t c_lens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ __constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ __constant__ uint16_t c_dists[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ __constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ __constant__ uint8_t c_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// The warp's view of one stream.  Every lane holds the same values (the lane index aside).
struct Stream {
  WarpMem* m; uint32_t lane;
  // input: 32-bit words from the 4-byte aligned address at or below the first deflate byte
  const uint32_t* gwords; uint32_t n_words, next; uint64_t buf; int cnt; bool bad;
  // output
  uint8_t* out; uint32_t out_len, pos, flushed, check; uint32_t skew; uint32_t adler_a, adler_b;   // (a tile is far below 4 GB)

  __device__ void stage_block(uint32_t b) {   // words [64 b, 64 b + 64) -> their half of the staging buffer
    __syncwarp();   // every lane is done reading the block this one replaces
    const uint32_t w0 = 64u * b + 2u * lane;
    m->stage[w0 & (STAGE_WORDS - 1)] = w0 < n_words ? gwords[w0] : 0u;
    m->stage[(w0 + 1u) & (STAGE_WORDS - 1)] = w0 + 1u < n_words ? gwords[w0 + 1u] : 0u;
  }
  __device__ void refill() {   // at least 33 valid bits afterwards (zero bits past the end of the stream: `bad` is set two words later)
    if (cnt <= 32) {
      if ((next & 63u) == 0u) { stage_block(next / 64u + 1u); __syncwarp(); }
      if (next >= n_words + 2u) bad = true;
      buf |= (uint64_t)m->stage[next & (STAGE_WORDS - 1)] << cnt;
      cnt += 32; ++next;
    }
  }
  __device__ uint32_t bits(int need) {   // need <= 24
    if (need == 0) return 0u;
    refill();
    const uint32_t v = (uint32_t)buf & ((1u << need) - 1u);
    buf >>= need; cnt -= need;
    return v;
  }
  // canonical walk over the code lengths for a code the primary table does not hold
  __device__ int decode_slow(const int16_t* count, const int16_t* symbol) {
    uint64_t bitbuf = buf;
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= MAXBITS; ++l) {
      code |= (int)(bitbuf & 1u); bitbuf >>= 1;
      const int c = count[l];
      if (code - c < first) { buf = bitbuf; cnt -= l; return symbol[index + (code - first)]; }
      index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
  }
  __device__ int decode_len() {
    refill();
    const uint32_t e = m->lut[(uint32_t)buf & ((1u << LUT_BITS) - 1u)];
    if (e) { buf >>= (e & 15u); cnt -= (int)(e & 15u); return (int)(e >> 4); }
    return decode_slow(m->lcount, m->lsymbol);
  }
  __device__ int decode_dist() {
    refill();
    const uint32_t e = m->dlut[(uint32_t)buf & ((1u << DLUT_BITS) - 1u)];
    if (e) { buf >>= (e & 15u); cnt -= (int)(e & 15u); return (int)(e >> 4); }
    return decode_slow(m->dcount, m->dsymbol);
  }

  // ---- output
  __device__ void adler_fold(uint32_t s1, uint32_t s2, uint32_t n) {   // lane sums of d and of (n - i) d over a chunk of n <= 512 bytes
    s1 = __reduce_add_sync(0xFFFFFFFFu, s1); s2 = __reduce_add_sync(0xFFFFFFFFu, s2);
    adler_b = (adler_b + n * adler_a + s2) % 65521u;
    adler_a = (adler_a + s1) % 65521u;
  }
  __device__ void flush_bytes(uint32_t n) {   // n <= 512 bytes, byte stores
    uint32_t s1 = 0, s2 = 0;
    for (uint32_t i = lane; i < n; i += 32u) {
      const uint32_t d = m->win[(flushed + i + skew) & WIN_MASK];
      out[flushed + i] = (uint8_t)d;
      s1 += d; s2 += (n - i) * d;
    }
    adler_fold(s1, s2, n);
    flushed += n;
  }
  __device__ void flush(bool final) {
    __syncwarp();
    const uint32_t a = (uint32_t)((uintptr_t)(out + flushed) & 511u);
    if (a) {   // up to the next 512-byte boundary of the output
      const uint32_t h = 512u - a;
      if (pos - flushed >= h) flush_bytes(h);
      else if (final) flush_bytes(pos - flushed);
      else { check = min(out_len, flushed + FLUSH_AT); return; }
    }
    while (pos - flushed >= 512u) {
      const uint4 v = *reinterpret_cast<const uint4*>(&m->win[(flushed + 16u * lane + skew) & WIN_MASK]);
      *reinterpret_cast<uint4*>(out + flushed + 16u * lane) = v;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t s1 = 0, s2 = 0;
#pragma unroll
      for (uint32_t j = 0; j < 16u; ++j) { const uint32_t d = (w[j >> 2] >> (8u * (j & 3u))) & 0xFFu; s1 += d; s2 += (512u - 16u * lane - j) * d; }
      adler_fold(s1, s2, 512u);
      flushed += 512u;
    }
    if (final) while (pos > flushed) flush_bytes(pos - flushed < 512u ? pos - flushed : 512u);
    check = min(out_len, flushed + FLUSH_AT);   // the next position at which the decoder has to look up from its loop
  }
  __device__ bool literal(uint32_t sym) {   // one compare covers "output full" and "time to flush"
    if (pos >= check) { if (pos >= out_len) return false; flush(false); }
    if (lane == 0) m->win[(pos + skew) & WIN_MASK] = (uint8_t)sym;
    ++pos;
    return true;
  }
  __device__ bool match(uint32_t len, uint32_t dist) {
    if (dist > pos || len > out_len - pos) return false;
    __syncwarp();   // the literals and matches before this one are in the ring for every lane
    const uint32_t from = pos - dist + skew, to = pos + skew;
    if (dist <= RING_REACH) {
      if (dist >= len) { for (uint32_t i = lane; i < len; i += 32u) m->win[(to + i) & WIN_MASK] = m->win[(from + i) & WIN_MASK]; }
      else for (uint32_t i = lane; i < len; i += 32u) m->win[(to + i) & WIN_MASK] = m->win[(from + i % dist) & WIN_MASK];   // overlapping: the pattern repeats
    } else {   // the source left the ring long ago: it is in HBM (dist > len here, no overlap)
      if (pos - dist + len > flushed) flush(true);   // (only a small ring can hold unflushed bytes that far back)
      for (uint32_t i = lane; i < len; i += 32u) m->win[(to + i) & WIN_MASK] = __ldcg(out + (pos - dist) + i);
    }
    pos += len;
    if (pos >= check && pos < out_len) flush(false);
    return true;
  }
};

// canonical Huffman code from code lengths -- sequential, done by lane 0 (the tables are shared; a read-modify-write by 32 lanes
// would lean on lock step); returns how many codes of the longest length are left (0 complete, < 0 over-subscribed)
__device__ inline int construct(int16_t* count, int16_t* symbol, const int16_t* length, int n, uint32_t lane) {
  __syncwarp();
  int left = 1;
  if (lane == 0) {
    for (int l = 0; l <= MAXBITS; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) count[length[s]]++;
    if (count[0] == n) left = 0;
    else {
      for (int l = 1; l <= MAXBITS; ++l) { left <<= 1; left -= count[l]; if (left < 0) break; }
      if (left >= 0) {
        int16_t offs[MAXBITS + 1];
        offs[1] = 0;
        for (int l = 1; l < MAXBITS; ++l) offs[l + 1] = offs[l] + count[l];
        for (int s = 0; s < n; ++s) if (length[s] != 0) symbol[offs[length[s]]++] = (int16_t)s;
      }
    }
  }
  left = __shfl_sync(0xFFFFFFFFu, left, 0);
  __syncwarp();
  return left;
}
// primary table: the codes in order (sequential, lane 0), then the table entries lane-parallel
__device__ inline void build_lut(uint16_t* lut, int lut_bits, uint16_t* code_of, const int16_t* length, int n, uint32_t lane) {
  __syncwarp();
  if (lane == 0) {
    uint32_t next_code[MAXBITS + 2], bl_count[MAXBITS + 1];
    for (int l = 0; l <= MAXBITS; ++l) bl_count[l] = 0;
    for (int s = 0; s < n; ++s) bl_count[length[s]]++;
    bl_count[0] = 0;
    uint32_t code = 0;
    for (int l = 1; l <= MAXBITS; ++l) { code = (code + bl_count[l - 1]) << 1; next_code[l] = code; }
    for (int s = 0; s < n; ++s) { const int l = length[s]; if (l) code_of[s] = (uint16_t)next_code[l]++; }
  }
  for (uint32_t i = lane; i < (1u << lut_bits); i += 32u) lut[i] = 0;
  __syncwarp();
  for (int s = (int)lane; s < n; s += 32) {
    const int l = length[s];
    if (l == 0 || l > lut_bits) continue;
    const uint32_t rev = __brev((uint32_t)code_of[s]) >> (32 - l);   // the stream carries Huffman codes most significant bit first
    for (uint32_t k = rev; k < (1u << lut_bits); k += 1u << l) lut[k] = (uint16_t)((s << 4) | l);
  }
  __syncwarp();
}

__device__ inline bool codes(Stream& s) {
  for (;;) {
    int sym = s.decode_len();
    if (sym < 0 || s.bad) return false;
    if (sym < 256) { if (!s.literal((uint32_t)sym)) return false; continue; }
    if (sym == 256) return true;
    sym -= 257;
    if (sym >= 29) return false;
    const uint32_t len = c_lens[sym] + s.bits(c_lext[sym]);
    const int ds = s.decode_dist();
    if (ds < 0 || ds >= 30) return false;
    const uint32_t dist = c_dists[ds] + s.bits(c_dext[ds]);
    if (s.bad || !s.match(len, dist)) return false;
  }
}

// one zlib stream, the whole warp; returns INF_*
__device__ inline uint32_t inflate_stream(WarpMem* m, uint32_t lane, const uint8_t* in, uint64_t in_len, uint8_t* out, uint64_t out_len) {
  if (in_len < 6) return INF_ERR_HEADER;
  const uint32_t cmf = in[0], flg = in[1];
  if ((cmf & 0x0F) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return INF_ERR_HEADER;
  Stream s;
  s.m = m; s.lane = lane;
  const uint8_t* d0 = in + 2;                                        // first deflate byte; the last four bytes are the Adler-32
  const uint32_t lead = (uint32_t)((uintptr_t)d0 & 3u);
  s.gwords = reinterpret_cast<const uint32_t*>(d0 - lead);
  s.n_words = (uint32_t)((lead + (in_len - 6) + 3u) / 4u);
  s.next = 0; s.buf = 0; s.cnt = 0; s.bad = false;
  if (out_len >= (1ull << 32)) return INF_ERR_SIZE;
  s.out = out; s.out_len = (uint32_t)out_len; s.pos = 0; s.flushed = 0; s.check = min((uint32_t)out_len, FLUSH_AT); s.skew = (uint32_t)((uintptr_t)out & 15u); s.adler_a = 1; s.adler_b = 0;
  __syncwarp();
  s.stage_block(0);
  __syncwarp();
  if (lead) s.bits(8 * (int)lead);
  int last;
  do {
    last = (int)s.bits(1);
    const uint32_t type = s.bits(2);
    if (s.bad) return INF_ERR_DATA;
    if (type == 0) {   // stored: back to the byte boundary, LEN / NLEN, then LEN bytes
      s.bits(s.cnt & 7);
      const uint32_t n = s.bits(16), nn = s.bits(16);
      if (s.bad || n != (~nn & 0xFFFFu) || n > s.out_len - s.pos) return INF_ERR_DATA;
      for (uint32_t i = 0; i < n; ++i) s.literal(s.bits(8));
      if (s.bad) return INF_ERR_DATA;
    } else if (type == 1 || type == 2) {
      int nlen = FIXLCODES, ndist = MAXDCODES;
      int16_t* lengths = m->lengths;
      if (type == 1) {   // fixed codes
        __syncwarp();
        for (int sym = (int)lane; sym < FIXLCODES; sym += 32) lengths[sym] = (int16_t)(sym < 144 ? 8 : (sym < 256 ? 9 : (sym < 280 ? 7 : 8)));
        for (int sym = (int)lane; sym < MAXDCODES; sym += 32) lengths[FIXLCODES + sym] = 5;
        __syncwarp();
        construct(m->lcount, m->lsymbol, lengths, FIXLCODES, lane);
        construct(m->dcount, m->dsymbol, lengths + FIXLCODES, MAXDCODES, lane);
      } else {           // dynamic codes
        nlen = (int)s.bits(5) + 257; ndist = (int)s.bits(5) + 1;
        const int ncode = (int)s.bits(4) + 4;
        if (s.bad || nlen > MAXLCODES || ndist > MAXDCODES) return INF_ERR_DATA;
        __syncwarp();
        int index = 0;
        for (; index < ncode; ++index) lengths[c_order[index]] = (int16_t)s.bits(3);
        for (; index < 19; ++index) lengths[c_order[index]] = 0;
        if (construct(m->lcount, m->lsymbol, lengths, 19, lane) != 0) return INF_ERR_DATA;
        build_lut(m->lut, 7, m->code, lengths, 19, lane);   // code-length codes are at most 7 bits long: the table holds them all
        __syncwarp();
        index = 0;
        while (index < nlen + ndist) {
          s.refill();
          const uint32_t e = m->lut[(uint32_t)s.buf & 127u];
          if (!e) return INF_ERR_DATA;
          s.buf >>= (e & 15u); s.cnt -= (int)(e & 15u);
          const int sym = (int)(e >> 4);
          if (sym < 16) lengths[index++] = (int16_t)sym;
          else {
            int len = 0, rep;
            if (sym == 16) { if (index == 0) return INF_ERR_DATA; __syncwarp(); len = lengths[index - 1]; rep = 3 + (int)s.bits(2); }
            else if (sym == 17) rep = 3 + (int)s.bits(3);
            else rep = 11 + (int)s.bits(7);
            if (s.bad || index + rep > nlen + ndist) return INF_ERR_DATA;
            while (rep--) lengths[index++] = (int16_t)len;
          }
        }
        __syncwarp();
        if (lengths[256] == 0) return INF_ERR_DATA;
        int err = construct(m->lcount, m->lsymbol, lengths, nlen, lane);
        if (err < 0 || (err > 0 && nlen != m->lcount[0] + m->lcount[1])) return INF_ERR_DATA;   // incomplete only when one 1-bit code
        err = construct(m->dcount, m->dsymbol, lengths + nlen, ndist, lane);
        if (err < 0 || (err > 0 && ndist != m->dcount[0] + m->dcount[1])) return INF_ERR_DATA;
      }
      build_lut(m->lut, LUT_BITS, m->code, lengths, nlen, lane);
      build_lut(m->dlut, DLUT_BITS, m->code, lengths + nlen, ndist, lane);
      if (!codes(s)) return INF_ERR_DATA;
    } else return INF_ERR_DATA;
    if (s.bad) return INF_ERR_DATA;
  } while (!last);
  // the decoder may have buffered whole words past the last block, but it must not have used bits the stream does not have
  if ((uint64_t)s.next * 32u - (uint64_t)s.cnt > 8ull * (lead + (in_len - 6))) return INF_ERR_DATA;
  if (s.pos != s.out_len) return INF_ERR_SIZE;
  s.flush(true);
  const uint8_t* t = in + in_len - 4;
  const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
  return ((s.adler_b << 16) | s.adler_a) == want ? INF_OK : INF_ERR_ADLER;
}

}  // namespace inflate_detail

constexpr int INFLATE_THREADS = 32 * inflate_detail::WARPS_PER_BLOCK;
constexpr size_t INFLATE_SMEM = sizeof(inflate_detail::WarpMem) * inflate_detail::WARPS_PER_BLOCK;

// one warp per tile, `order[k]` = the job the k-th warp takes (longest first); err[0] = first INF_ERR_* seen, err[1] = the job it was seen in
__global__ void __launch_bounds__(INFLATE_THREADS) k_inflate(const InflateJob* __restrict__ jobs, const uint32_t* __restrict__ order, uint64_t n_jobs, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char inflate_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint64_t k = (uint64_t)blockIdx.x * inflate_detail::WARPS_PER_BLOCK + warp;
  if (k >= n_jobs) return;
  const uint32_t j = order ? order[k] : (uint32_t)k;
  const InflateJob job = jobs[j];
  inflate_detail::WarpMem* m = reinterpret_cast<inflate_detail::WarpMem*>(inflate_smem) + warp;
  const uint32_t rc = job.out_len == 0 && job.in_len == 0 ? (uint32_t)INF_OK : inflate_detail::inflate_stream(m, lane, job.in, job.in_len, job.out, job.out_len);
  if (rc != INF_OK && lane == 0 && atomicCAS(&err[0], 0u, rc) == 0u) err[1] = j;
}

}  // namespace gdb
