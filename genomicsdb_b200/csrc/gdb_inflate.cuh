// K0: DEFLATE on the device.  GenomicsDB arrays are usually written with "compress_tiledb_array": every tile of every
// attribute file is one zlib stream (ReadState::decompress_tile_gzip -> gunzip, dependencies/TileDB/core/src/fragment/
// read_state.cc:1595-1702, misc/utils.cc:466-512).  A fragment has (cells / capacity) tiles per attribute -- hundreds of
// thousands of independent streams -- so the parallelism is across streams: one thread inflates one tile (RFC 1950 / 1951:
// stored, fixed and dynamic Huffman blocks; canonical codes decoded bit by bit from per-length counts), straight from the
// compressed file bytes staged in HBM into the column's place in HBM, and checks the Adler-32 trailer.  The host never sees
// the decompressed bytes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gdb {

struct InflateJob { const uint8_t* in; uint64_t in_len; uint8_t* out; uint64_t out_len; };   // one zlib stream -> one tile (device pointers)
enum : uint32_t { INF_OK = 0, INF_ERR_HEADER = 1, INF_ERR_DATA = 2, INF_ERR_SIZE = 3, INF_ERR_ADLER = 4 };

namespace inflate_detail {

constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, FIXLCODES = 288;

struct BitIn {
  const uint8_t* in; uint64_t len, pos; uint32_t buf; int cnt; bool bad;
  __device__ uint32_t bits(int need) {
    uint32_t v = buf;
    while (cnt < need) {
      if (pos >= len) { bad = true; return 0; }
      v |= (uint32_t)in[pos++] << cnt;
      cnt += 8;
    }
    buf = need < 32 ? v >> need : 0;
    cnt -= need;
    return v & ((1u << need) - 1u);
  }
};
template <int NSYM> struct Huff { int16_t count[MAXBITS + 1]; int16_t symbol[NSYM]; };

// canonical Huffman code from code lengths; returns how many codes of the longest length are left (0 complete, < 0 over-subscribed)
template <int NSYM>
__device__ int construct(Huff<NSYM>& h, const int16_t* length, int n) {
  for (int l = 0; l <= MAXBITS; ++l) h.count[l] = 0;
  for (int s = 0; s < n; ++s) h.count[length[s]]++;
  if (h.count[0] == n) return 0;
  int left = 1;
  for (int l = 1; l <= MAXBITS; ++l) { left <<= 1; left -= h.count[l]; if (left < 0) return left; }
  int16_t offs[MAXBITS + 1];
  offs[1] = 0;
  for (int l = 1; l < MAXBITS; ++l) offs[l + 1] = offs[l] + h.count[l];
  for (int s = 0; s < n; ++s) if (length[s] != 0) h.symbol[offs[length[s]]++] = (int16_t)s;
  return left;
}
// Primary table of the literal/length code: the next LUT_BITS input bits -> symbol << 4 | code length, 0 when the code is
// longer than LUT_BITS bits (decode() then walks the lengths).  Most symbols of a tile are literals with short codes, so
// the common case is one lookup instead of a bit-by-bit walk.
constexpr int LUT_BITS = 10;
struct FastLit { uint16_t lut[1 << LUT_BITS]; };
__device__ inline void build_fast(FastLit& f, const int16_t* length, int n) {
  uint16_t next_code[MAXBITS + 2];
  int16_t bl_count[MAXBITS + 1];
  for (int l = 0; l <= MAXBITS; ++l) bl_count[l] = 0;
  for (int s = 0; s < n; ++s) bl_count[length[s]]++;
  bl_count[0] = 0;
  uint32_t code = 0;
  for (int l = 1; l <= MAXBITS; ++l) { code = (code + (uint32_t)bl_count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
  for (int i = 0; i < (1 << LUT_BITS); ++i) f.lut[i] = 0;
  for (int s = 0; s < n; ++s) {
    const int l = length[s];
    if (l == 0) continue;
    const uint32_t c = next_code[l]++;
    if (l > LUT_BITS) continue;
    const uint32_t rev = __brev(c) >> (32 - l);   // the stream carries Huffman codes most significant bit first
    for (uint32_t k = rev; k < (1u << LUT_BITS); k += 1u << l) f.lut[k] = (uint16_t)((s << 4) | l);
  }
}

template <int NSYM>
__device__ int decode(BitIn& s, const Huff<NSYM>& h) {
  // top the bit buffer up to a whole code (<= 15 bits) once, then walk the code lengths on a register copy
  while (s.cnt < MAXBITS && s.pos < s.len) { s.buf |= (uint32_t)s.in[s.pos++] << s.cnt; s.cnt += 8; }
  uint32_t bitbuf = s.buf;
  int left = s.cnt, code = 0, first = 0, index = 0;
  for (int l = 1; l <= MAXBITS; ++l) {
    if (left == 0) { s.bad = true; return -1; }
    code |= (int)(bitbuf & 1u); bitbuf >>= 1; --left;
    const int count = h.count[l];
    if (code - count < first) { s.buf = bitbuf; s.cnt = left; return h.symbol[index + (code - first)]; }
    index += count; first += count; first <<= 1; code <<= 1;
  }
  return -1;
}

__device__ __constant__ uint16_t c_lens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ __constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ __constant__ uint16_t c_dists[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ __constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ __constant__ uint8_t c_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

template <int NL, int ND>
__device__ bool codes(BitIn& s, uint8_t* out, uint64_t& pos, uint64_t cap, const Huff<NL>& lencode, const Huff<ND>& distcode, const FastLit& fast) {
  for (;;) {
    while (s.cnt < MAXBITS && s.pos < s.len) { s.buf |= (uint32_t)s.in[s.pos++] << s.cnt; s.cnt += 8; }
    int sym;
    const uint32_t e = fast.lut[s.buf & ((1u << LUT_BITS) - 1u)];
    if (e != 0 && (int)(e & 15u) <= s.cnt) { s.buf >>= (e & 15u); s.cnt -= (int)(e & 15u); sym = (int)(e >> 4); }
    else sym = decode(s, lencode);
    if (sym < 0) return false;
    if (sym < 256) { if (pos >= cap) return false; out[pos++] = (uint8_t)sym; continue; }
    if (sym == 256) return true;
    sym -= 257;
    if (sym >= 29) return false;
    const uint32_t len = c_lens[sym] + s.bits(c_lext[sym]);
    const int ds = decode(s, distcode);
    if (ds < 0 || ds >= 30) return false;
    const uint64_t dist = c_dists[ds] + s.bits(c_dext[ds]);
    if (s.bad || dist > pos || pos + len > cap) return false;
    for (uint32_t i = 0; i < len; ++i, ++pos) out[pos] = out[pos - dist];
  }
}

// one zlib stream; returns INF_*
__device__ inline uint32_t inflate_stream(const uint8_t* in, uint64_t in_len, uint8_t* out, uint64_t out_len) {
  if (in_len < 6) return INF_ERR_HEADER;
  const uint32_t cmf = in[0], flg = in[1];
  if ((cmf & 0x0F) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return INF_ERR_HEADER;
  BitIn s{in, in_len - 4, 2, 0, 0, false};
  FastLit fast;
  uint64_t pos = 0;
  int last;
  do {
    last = (int)s.bits(1);
    const uint32_t type = s.bits(2);
    if (s.bad) return INF_ERR_DATA;
    if (type == 0) {   // stored: back to the byte boundary (decode() may have buffered whole bytes ahead)
      s.pos -= (uint64_t)(s.cnt >> 3); s.buf = 0; s.cnt = 0;
      if (s.pos + 4 > s.len) return INF_ERR_DATA;
      const uint32_t n = in[s.pos] | (in[s.pos + 1] << 8), nn = in[s.pos + 2] | (in[s.pos + 3] << 8);
      s.pos += 4;
      if (n != (~nn & 0xFFFFu) || s.pos + n > s.len || pos + n > out_len) return INF_ERR_DATA;
      for (uint32_t i = 0; i < n; ++i) out[pos++] = in[s.pos++];
    } else if (type == 1) {   // fixed codes
      Huff<FIXLCODES> lencode; Huff<MAXDCODES> distcode;
      int16_t lengths[FIXLCODES];
      int sym = 0;
      for (; sym < 144; ++sym) lengths[sym] = 8;
      for (; sym < 256; ++sym) lengths[sym] = 9;
      for (; sym < 280; ++sym) lengths[sym] = 7;
      for (; sym < FIXLCODES; ++sym) lengths[sym] = 8;
      construct(lencode, lengths, FIXLCODES);
      for (sym = 0; sym < MAXDCODES; ++sym) lengths[sym] = 5;
      construct(distcode, lengths, MAXDCODES);
      for (sym = 0; sym < 144; ++sym) lengths[sym] = 8;   // `lengths` was reused for the distance code
      build_fast(fast, lengths, FIXLCODES);
      if (!codes(s, out, pos, out_len, lencode, distcode, fast)) return INF_ERR_DATA;
    } else if (type == 2) {   // dynamic codes
      const int nlen = (int)s.bits(5) + 257, ndist = (int)s.bits(5) + 1, ncode = (int)s.bits(4) + 4;
      if (s.bad || nlen > MAXLCODES || ndist > MAXDCODES) return INF_ERR_DATA;
      int16_t lengths[MAXLCODES + MAXDCODES];
      Huff<FIXLCODES> lencode; Huff<MAXDCODES> distcode;
      int index = 0;
      for (; index < ncode; ++index) lengths[c_order[index]] = (int16_t)s.bits(3);
      for (; index < 19; ++index) lengths[c_order[index]] = 0;
      if (construct(lencode, lengths, 19) != 0) return INF_ERR_DATA;
      index = 0;
      while (index < nlen + ndist) {
        int sym = decode(s, lencode);
        if (sym < 0) return INF_ERR_DATA;
        if (sym < 16) lengths[index++] = (int16_t)sym;
        else {
          int len = 0, rep;
          if (sym == 16) { if (index == 0) return INF_ERR_DATA; len = lengths[index - 1]; rep = 3 + (int)s.bits(2); }
          else if (sym == 17) rep = 3 + (int)s.bits(3);
          else rep = 11 + (int)s.bits(7);
          if (s.bad || index + rep > nlen + ndist) return INF_ERR_DATA;
          while (rep--) lengths[index++] = (int16_t)len;
        }
      }
      if (lengths[256] == 0) return INF_ERR_DATA;
      int err = construct(lencode, lengths, nlen);
      if (err < 0 || (err > 0 && nlen != lencode.count[0] + lencode.count[1])) return INF_ERR_DATA;   // incomplete only when one 1-bit code
      err = construct(distcode, lengths + nlen, ndist);
      if (err < 0 || (err > 0 && ndist != distcode.count[0] + distcode.count[1])) return INF_ERR_DATA;
      build_fast(fast, lengths, nlen);
      if (!codes(s, out, pos, out_len, lencode, distcode, fast)) return INF_ERR_DATA;
    } else return INF_ERR_DATA;
    if (s.bad) return INF_ERR_DATA;
  } while (!last);
  if (pos != out_len) return INF_ERR_SIZE;
  uint32_t a = 1, b = 0;   // Adler-32 of the output against the big-endian trailer
  for (uint64_t i = 0; i < out_len;) {
    const uint64_t end = i + 5552 < out_len ? i + 5552 : out_len;
    for (; i < end; ++i) { a += out[i]; b += a; }
    a %= 65521u; b %= 65521u;
  }
  const uint8_t* t = in + in_len - 4;
  const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
  return ((b << 16) | a) == want ? INF_OK : INF_ERR_ADLER;
}

}  // namespace inflate_detail

// one thread per tile; err[0] = first INF_ERR_* seen, err[1] = the job it was seen in
__global__ void __launch_bounds__(64) k_inflate(const InflateJob* __restrict__ jobs, uint64_t n_jobs, uint32_t* __restrict__ err) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const InflateJob job = jobs[j];
  const uint32_t rc = job.out_len == 0 && job.in_len == 0 ? (uint32_t)INF_OK : inflate_detail::inflate_stream(job.in, job.in_len, job.out, job.out_len);
  if (rc != INF_OK && atomicCAS(&err[0], 0u, rc) == 0u) err[1] = (uint32_t)j;
}

}  // namespace gdb
