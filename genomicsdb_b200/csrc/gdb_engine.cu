// Host orchestration of the B200 variant-combine path + the extern "C" ABI of include/gdbgpu.h.
// Everything that computes runs in the kernels of gdb_kernels*.cuh; this file stages fragment bytes into HBM, sizes
// buffers, launches, and hands batches back.  No CPU fallback: without a CUDA device every entry point fails.
#include "../../include/gdbgpu.h"

#include <cub/device/device_radix_sort.cuh>

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "gdb_inflate.cuh"
#include "gdb_kernels6.cuh"
#include "gdb_tiledb.h"

using namespace gdb;

// ------------------------------------------------------------------------------------------------ plumbing
struct BatchBuffers;
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct gdbgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;        // kernels and the small control copies
  cudaStream_t copy_stream = nullptr;   // device -> host copies of finished batches (overlap the next batch's kernels)
  std::string err;
  uint64_t launches = 0;
  // optional per-kernel device timing (CUDA events on the launching stream), see gdbgpu_ctx_set_profiling
  bool profiling = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, std::pair<double, uint64_t>> stats;
  std::vector<BatchBuffers*> bb_pool;
  std::multimap<size_t, void*> free_blocks; size_t cached_bytes = 0;   // device blocks waiting for reuse (see dev_alloc)
  // file staging (gdbgpu_array_open*): worker threads pread the fragment files into pinned chunks and copy them on their own streams
  uint8_t* stage_pin = nullptr; std::vector<cudaStream_t> stage_streams; std::vector<cudaEvent_t> stage_events;
  uint32_t mix_attr_set = 0;   // k_remap_mix instantiations whose dynamic shared-memory limit has been raised
  uint32_t tma_attr_set = 0;   // k_remap_tma instantiations whose dynamic shared-memory limit has been raised
  uint32_t* d_cls_tab = nullptr; uint16_t* d_jk_tab = nullptr; int n_sms = 148;   // universal genotype tables (gdb_kernels2.cuh: TileTabs)
  cudaEvent_t get_event() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
  void resolve() {
    if (recs.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& r : recs) { float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b); auto& st = stats[r.name]; st.first += ms; st.second += 1; pool.push_back(r.a); pool.push_back(r.b); }
    recs.clear();
  }
};
struct ProfScope {
  gdbgpu_ctx* c; ProfRec r;
  ProfScope(gdbgpu_ctx* ctx, const char* name) : c(ctx) { if (c->profiling) { r.name = name; r.a = c->get_event(); r.b = c->get_event(); cudaEventRecord(r.a, c->stream); } }
  ~ProfScope() { if (c->profiling) { cudaEventRecord(r.b, c->stream); c->recs.push_back(r); } }
};

namespace {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct LimitError : std::runtime_error { using std::runtime_error::runtime_error; };
struct OverlapError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)

// Device memory goes through a per-context caching allocator: a block given back by an array / scan / batch is kept and
// handed to the next request of (nearly) the same size, so staging an array or growing a batch buffer in steady state
// never reaches cudaMalloc / cudaFree (multi-GB calls cost 100+ ms and synchronise the device).  All work of a context is
// ordered on its compute stream (the copy stream only reads batch buffers that are not recycled before their copies are
// waited for), so immediate reuse is safe.
thread_local gdbgpu_ctx* tl_ctx = nullptr;
inline void* dev_alloc(gdbgpu_ctx* c, size_t bytes, size_t& cap) {
  bytes = (bytes + 511) & ~(size_t)511;
  if (c) {
    auto it = c->free_blocks.lower_bound(bytes);
    if (it != c->free_blocks.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
      void* p = it->second; cap = it->first; c->cached_bytes -= cap; c->free_blocks.erase(it); return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess && c && !c->free_blocks.empty()) {   // give the cache back to the driver and retry
    cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->free_blocks) cudaFree(kv.second);
    c->free_blocks.clear(); c->cached_bytes = 0;
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  cap = bytes;
  return p;
}
inline void dev_free(gdbgpu_ctx* c, void* p, size_t cap) {
  if (!p) return;
  if (c) { c->free_blocks.emplace(cap, p); c->cached_bytes += cap; }
  else cudaFree(p);
}
template <class T>
struct DevBuf {
  T* p = nullptr; size_t n = 0, cap = 0; gdbgpu_ctx* owner = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) { dev_free(owner, p, cap); p = nullptr; } }
  void swap(DevBuf& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); std::swap(owner, o.owner); }
  void alloc(size_t count) {
    release(); n = count; owner = tl_ctx;
    if (count) p = static_cast<T*>(dev_alloc(owner, count * sizeof(T), cap));
  }
  void ensure(size_t count) { if (count > n) alloc(count + count / 4 + 16); }
  size_t bytes() const { return n * sizeof(T); }
  // takes over the block of another buffer (any element type) as `count` elements: no copy, the source is left empty
  template <class U> void adopt(DevBuf<U>& src, size_t count) {
    release();
    if (count * sizeof(T) > src.cap) throw std::runtime_error("internal: adopted block is too small");
    p = reinterpret_cast<T*>(src.p); n = count; cap = src.cap; owner = src.owner;
    src.p = nullptr; src.n = 0; src.cap = 0;
  }
};
template <class T>
struct PinBuf {
  T* p = nullptr; size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete; PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  void ensure(size_t count) { if (count > n) { if (p) cudaFreeHost(p); p = nullptr; n = count + count / 4 + 16; CK(cudaMallocHost((void**)&p, n * sizeof(T))); } }
};

inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)std::max<uint64_t>(1, (n + per - 1) / per); }

#define LAUNCH(ctx, kernel, grid, block, ...) do { { ProfScope ps__(ctx, #kernel); kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); } ++(ctx)->launches; CK(cudaGetLastError()); } while (0)

// exclusive scan; `total` (device) receives the sum.  in/out may alias when TI == TO.
template <class TI, class TO>
void exclusive_scan(gdbgpu_ctx* ctx, const TI* in, TO* out, uint64_t n, TO* total_dev, DevBuf<TO>& tile_sums) {
  uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles == 0) { CK(cudaMemsetAsync(total_dev, 0, sizeof(TO), ctx->stream)); return; }
  tile_sums.ensure(tiles);
  LAUNCH(ctx, (k_scan_tiles<TI, TO>), (unsigned)tiles, SCAN_THREADS, in, out, tile_sums.p, n);
  LAUNCH(ctx, (k_scan_sums<TO>), 1, 1024, tile_sums.p, tiles, total_dev);
  if (tiles > 1) LAUNCH(ctx, (k_scan_add<TO>), (unsigned)tiles, SCAN_THREADS, out, tile_sums.p, n);
}

template <class K, class V>
void radix_sort_pairs(gdbgpu_ctx* ctx, const K* kin, K* kout, const V* vin, V* vout, uint64_t n, int end_bit, DevBuf<uint8_t>& tmp) {
  size_t bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, ctx->stream));
  tmp.ensure(bytes);
  CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, ctx->stream));
  ctx->launches += 4;
}

__global__ void k_narrow_u64(const uint64_t* in, uint64_t n, uint32_t* out) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = (uint32_t)in[i]; }
__global__ void k_iota(uint32_t* p, uint64_t n) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = (uint32_t)i; }
__global__ void k_flag_del(ArrayDev A, uint32_t* flag) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < A.n) flag[i] = (A.desc[i].flags & F_DELETION) && !(A.desc[i].flags & F_ROWOUT) ? 1u : 0u; }
__global__ void k_compact_idx(const uint32_t* flag, const uint32_t* pos, uint64_t n, uint32_t* out) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n && flag[i]) out[pos[i]] = (uint32_t)i; }
__global__ void k_bounds(const int64_t* coords, uint64_t n, int64_t q_lo, int64_t q_hi, uint64_t* out) {
  // out[0] = first cell with col >= q_lo, out[1] = first with col > q_lo, out[2] = first with col > q_hi
  int t = threadIdx.x;
  if (t > 2) return;
  int64_t key = (t == 0) ? q_lo : (t == 1 ? q_lo : q_hi);
  bool strict = (t != 0);
  uint64_t lo = 0, hi = n;
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; int64_t c = coords[2 * mid + 1]; if (strict ? (c <= key) : (c < key)) lo = mid + 1; else hi = mid; }
  out[t] = lo;
}
// rows of the staged array without any cell at or after column `col` (tile-range staging, gdbgpu_array_open_range)
__global__ void k_rows_without_cells(ArrayDev A, int64_t col_lo, uint32_t* __restrict__ count) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.num_rows) return;
  const uint64_t lo = A.row_start[r], hi = A.row_start[r + 1];
  if (lo == hi || A.rs_col[hi - 1] < col_lo) atomicAdd(count, 1u);
}
__global__ void k_add_u64(uint64_t* p, uint64_t n, uint64_t add) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] += add; }
__global__ void k_gather_bytes(const uint8_t* in, const uint32_t* perm, uint64_t n, uint32_t cell_bytes, uint8_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) for (uint32_t b = 0; b < cell_bytes; ++b) out[i * cell_bytes + b] = in[(uint64_t)perm[i] * cell_bytes + b];
}

// Fragment bytes travel disk (page cache) -> pinned chunk -> HBM: STAGE_WORKERS host threads take chunks of the jobs off one
// counter, each with two pinned buffers and its own stream, so reads, copies of different files and PCIe overlap.  (Reading
// whole files into pageable vectors first cost 2 s for a 3.9 GB shard; this path is bound by the page-cache copy.)
struct FileJob { std::string path; uint64_t off, len; uint8_t* dst; };
constexpr int STAGE_WORKERS = 8;
constexpr size_t STAGE_CHUNK = 8u << 20;
void stage_files(gdbgpu_ctx* ctx, const std::vector<FileJob>& jobs) {
  if (jobs.empty()) return;
  if (!ctx->stage_pin) {
    CK(cudaMallocHost((void**)&ctx->stage_pin, (size_t)STAGE_WORKERS * 2 * STAGE_CHUNK));
    for (int w = 0; w < STAGE_WORKERS; ++w) { cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); ctx->stage_streams.push_back(st); }
    for (int w = 0; w < 2 * STAGE_WORKERS; ++w) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->stage_events.push_back(e); }
  }
  struct Chunk { uint32_t job; uint64_t off, len; };
  std::vector<Chunk> chunks;
  for (uint32_t j = 0; j < jobs.size(); ++j) for (uint64_t o = 0; o < jobs[j].len; o += STAGE_CHUNK) chunks.push_back(Chunk{j, o, std::min<uint64_t>(STAGE_CHUNK, jobs[j].len - o)});
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::string first_error;
  std::mutex err_mu;
  auto fail = [&](const std::string& m) { std::lock_guard<std::mutex> g(err_mu); if (!failed.exchange(true)) first_error = m; };
  const int device = ctx->device;
  auto worker = [&](int w) {
    if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice failed in a staging thread"); return; }
    int fd = -1; uint32_t fd_job = 0xFFFFFFFFu;
    bool used[2] = {false, false};
    for (int b = 0;; b ^= 1) {
      const size_t c = next.fetch_add(1);
      if (c >= chunks.size() || failed.load()) break;
      const Chunk& ch = chunks[c]; const FileJob& jb = jobs[ch.job];
      if (fd_job != ch.job) { if (fd >= 0) close(fd); fd = open(jb.path.c_str(), O_RDONLY); fd_job = ch.job; if (fd < 0) { fail("cannot open " + jb.path); break; } }
      uint8_t* pin = ctx->stage_pin + ((size_t)w * 2 + b) * STAGE_CHUNK;
      cudaEvent_t ev = ctx->stage_events[(size_t)w * 2 + b];
      if (used[b] && cudaEventSynchronize(ev) != cudaSuccess) { fail("cudaEventSynchronize failed in a staging thread"); break; }
      uint64_t done = 0;
      while (done < ch.len) {
        ssize_t r = pread(fd, pin + done, (size_t)(ch.len - done), (off_t)(jb.off + ch.off + done));
        if (r <= 0) { fail("short read on " + jb.path); break; }
        done += (uint64_t)r;
      }
      if (done < ch.len) break;
      if (cudaMemcpyAsync(jb.dst + ch.off, pin, ch.len, cudaMemcpyHostToDevice, ctx->stage_streams[w]) != cudaSuccess || cudaEventRecord(ev, ctx->stage_streams[w]) != cudaSuccess) { fail("cudaMemcpyAsync failed in a staging thread"); break; }
      used[b] = true;
    }
    if (fd >= 0) close(fd);
    if (cudaStreamSynchronize(ctx->stage_streams[w]) != cudaSuccess) fail("a staging copy failed");
  };
  std::vector<std::thread> ts;
  const int nw = (int)std::min<size_t>(STAGE_WORKERS, chunks.size());
  for (int w = 0; w < nw; ++w) ts.emplace_back(worker, w);
  for (auto& t : ts) t.join();
  if (failed.load()) throw std::runtime_error(first_error);
}

struct HostVar { const uint64_t* off = nullptr; const void* val = nullptr; uint64_t bytes = 0; };
struct HostCols {
  uint64_t n = 0;
  const int64_t* coords = nullptr; const int64_t* end = nullptr;
  HostVar ref, alt, gt; bool has_gt = false;
  std::vector<HostVar> fields;
  std::vector<HostVar> plain;   // fixed-length plain field: off == nullptr
};
struct PlainMeta { std::string name; int32_t elem, fixed_len; int32_t no_format = 0, info_op = 0; };

struct DevVar { DevBuf<uint64_t> off; DevBuf<uint8_t> val; uint64_t bytes = 0; };
// a column that gdbgpu_array_open* has already placed (and inflated) in HBM: the fixed part (values, or the n + 1 value
// offsets of a variable attribute) and the value stream.  build_array takes the blocks over instead of copying them.
struct DevColumn { DevBuf<uint8_t> fixed, var; uint64_t var_bytes = 0; };

}  // namespace

// ------------------------------------------------------------------------------------------------ array
struct gdbgpu_array {
  gdbgpu_ctx* ctx = nullptr;
  uint64_t n = 0; uint32_t n_fragments = 1;
  int64_t lb_row = 0, num_rows = 0, min_col = 0, max_col = 0, max_end = 0;
  uint64_t n_begin = 0, n_bad = 0, n_del_cells = 0, staged_bytes = 0;
  bool has_gt = false; int gt_phased = 0;
  std::vector<FieldMeta> fm; std::vector<std::string> field_names;
  DevBuf<int64_t> coords, end;
  DevVar ref, alt, gt; std::vector<std::unique_ptr<DevVar>> fields;
  std::vector<std::unique_ptr<DevVar>> plain; std::vector<PlainMeta> pm; PlainSet plain_dev{};   // plain FORMAT fields (K6)
  DevBuf<CellDesc> desc; DevBuf<HotRec> hot; DevBuf<uint32_t> warm, cls; DevBuf<int64_t> effend;
  DevBuf<uint32_t> rs_cell, rs_keys; DevBuf<int64_t> rs_col, rs_effend; DevBuf<uint64_t> row_start; uint64_t n_rs = 0; DevBuf<RsRec> rs_rec;
  DevBuf<uint32_t> bad_cell; DevBuf<int64_t> bad_pred_col;
  ArrayDev dev{};
  uint64_t hbm_bytes() const {
    uint64_t b = coords.bytes() + end.bytes() + ref.off.bytes() + ref.val.bytes() + alt.off.bytes() + alt.val.bytes() + gt.off.bytes() + gt.val.bytes() +
                 desc.bytes() + hot.bytes() + warm.bytes() + cls.bytes() + effend.bytes() + rs_rec.bytes() + rs_cell.bytes() + rs_keys.bytes() + rs_col.bytes() + rs_effend.bytes() + row_start.bytes();
    for (auto& f : fields) b += f->off.bytes() + f->val.bytes();
    for (auto& f : plain) b += f->off.bytes() + f->val.bytes();
    return b;
  }
};

namespace {

constexpr uint32_t BAD_CAP = 4096;

void upload_var(gdbgpu_ctx* ctx, const HostVar& h, uint64_t n, DevVar& d, uint64_t& staged) {
  d.off.alloc(n + 1);
  d.val.alloc(std::max<uint64_t>(h.bytes, 1));
  d.bytes = h.bytes;
  if (n) CK(cudaMemcpyAsync(d.off.p, h.off, n * 8, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(d.off.p + n, &h.bytes, 8, cudaMemcpyDefault, ctx->stream));
  if (h.bytes) CK(cudaMemcpyAsync(d.val.p, h.val, h.bytes, cudaMemcpyDefault, ctx->stream));
  staged += n * 8 + h.bytes;
}

void fill_dev_view(gdbgpu_array* a) {
  ArrayDev& D = a->dev;
  D.n = a->n; D.coords = a->coords.p; D.end = a->end.p;
  D.ref = VarCol{a->ref.off.p, a->ref.val.p}; D.alt = VarCol{a->alt.off.p, a->alt.val.p};
  D.gt = VarCol{a->gt.off.p, a->gt.val.p};
  D.has_gt = a->has_gt; D.gt_phased = a->gt_phased; D.n_fields = (int)a->fields.size();
  for (size_t f = 0; f < a->fields.size(); ++f) { D.f[f] = VarCol{a->fields[f]->off.p, a->fields[f]->val.p}; D.fm[f] = a->fm[f]; }
  D.lb_row = a->lb_row; D.num_rows = a->num_rows;
  D.desc = a->desc.p; D.hot = a->hot.p; D.warm = a->warm.p; D.cls = a->cls.p; D.effend = a->effend.p;
  a->plain_dev.n = (int32_t)a->plain.size();
  for (size_t f = 0; f < a->plain.size(); ++f) a->plain_dev.f[f] = PlainDev{a->pm[f].fixed_len ? nullptr : a->plain[f]->off.p, a->plain[f]->val.p, a->pm[f].elem, a->pm[f].fixed_len, a->pm[f].no_format == 2 ? HC_BLOCK : (a->pm[f].no_format == 3 ? (HC_BLOCK | HC_RAW | HC_SKIP_REDUCED) : (a->pm[f].no_format == 4 ? (HC_BLOCK | HC_SKIP_REDUCED) : 0))};
  D.rs_rec = a->rs_rec.p;
  D.n_rs = a->n_rs; D.rs_cell = a->rs_cell.p; D.rs_col = a->rs_col.p; D.rs_effend = a->rs_effend.p; D.row_start = a->row_start.p;
}

// K1: merge the concatenated cells of several fragments into global (col,row) order, newest fragment wins.
// Returns src[] (device) = for each surviving merged cell, its index in the concatenation.
void fragment_merge(gdbgpu_ctx* ctx, const int64_t* d_coords, uint64_t n_all, const std::vector<uint64_t>& frag_cells, DevBuf<uint32_t>& src, uint64_t& n_out) {
  // runs = the fragments, oldest first; each round merges neighbours (gdb_kernels.cuh: k_merge_path)
  DevBuf<int64_t> col[2], row[2]; DevBuf<uint32_t> cell[2]; DevBuf<uint64_t> d_runs; DevBuf<uint2> d_tiles;
  for (int b = 0; b < 2; ++b) { col[b].alloc(n_all); row[b].alloc(n_all); cell[b].alloc(n_all); }
  unsigned g = blocks_for(n_all, 256);
  LAUNCH(ctx, k_merge_init, g, 256, d_coords, n_all, col[0].p, row[0].p, cell[0].p);
  std::vector<uint64_t> runs(1, 0);
  for (uint64_t c : frag_cells) runs.push_back(runs.back() + c);
  int cur = 0;
  while (runs.size() > 2) {
    const uint32_t n_runs = (uint32_t)runs.size() - 1, n_pairs = (n_runs + 1) / 2;
    std::vector<uint2> tiles; std::vector<uint64_t> next(1, 0);
    for (uint32_t p = 0; p < n_pairs; ++p) {
      const uint64_t lo = runs[2 * p], hi = runs[std::min<uint32_t>(2 * p + 2, n_runs)];
      for (uint64_t t = 0; t * MERGE_TILE < hi - lo; ++t) tiles.push_back(make_uint2(p, (uint32_t)t));
      next.push_back(hi);
    }
    d_runs.alloc(runs.size()); d_tiles.alloc(std::max<size_t>(tiles.size(), 1));
    CK(cudaMemcpyAsync(d_runs.p, runs.data(), runs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
    if (!tiles.empty()) LAUNCH(ctx, k_merge_path, (unsigned)tiles.size(), MERGE_THREADS, col[cur].p, row[cur].p, cell[cur].p, col[cur ^ 1].p, row[cur ^ 1].p, cell[cur ^ 1].p, d_runs.p, n_runs, d_tiles.p);
    CK(cudaStreamSynchronize(ctx->stream));   // the host vectors of this round may go
    runs.swap(next); cur ^= 1;
  }
  DevBuf<uint32_t>& perm_b = cell[cur];
  // perm_b = merged order incl. shadowed duplicates
  DevBuf<uint32_t> keep, pos, total; DevBuf<uint32_t> sums;
  keep.alloc(n_all); pos.alloc(n_all); total.alloc(1);
  LAUNCH(ctx, k_merge_keep, g, 256, d_coords, perm_b.p, n_all, keep.p);
  exclusive_scan<uint32_t, uint32_t>(ctx, keep.p, pos.p, n_all, total.p, sums);
  uint32_t h_total = 0;
  CK(cudaMemcpyAsync(&h_total, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  n_out = h_total;
  src.alloc(std::max<uint64_t>(n_out, 1));
  LAUNCH(ctx, k_merge_compact, g, 256, perm_b.p, keep.p, pos.p, n_all, src.p);
  CK(cudaStreamSynchronize(ctx->stream));
}

void gather_var(gdbgpu_ctx* ctx, const DevVar& in, const uint32_t* src, uint64_t n_out, DevVar& out) {
  DevBuf<uint64_t> len, sums, total;
  len.alloc(n_out + 1); total.alloc(1);
  out.off.alloc(n_out + 1);
  LAUNCH(ctx, k_gather_len, blocks_for(n_out, 256), 256, in.off.p, src, n_out, len.p);
  exclusive_scan<uint64_t, uint64_t>(ctx, len.p, out.off.p, n_out, total.p, sums);
  uint64_t h_total = 0;
  CK(cudaMemcpyAsync(&h_total, total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpyAsync(out.off.p + n_out, &h_total, 8, cudaMemcpyHostToDevice, ctx->stream));
  out.val.alloc(std::max<uint64_t>(h_total, 1)); out.bytes = h_total;
  LAUNCH(ctx, k_gather_var, blocks_for(n_out * 32, 256), 256, in.off.p, in.val.p, src, out.off.p, n_out, out.val.p);
  CK(cudaStreamSynchronize(ctx->stream));
}

// K0 launch: one warp per stream, the longest streams first (a launch lasts as long as its last stream); throws on a corrupt stream
void inflate_jobs(gdbgpu_ctx* ctx, const std::vector<InflateJob>& jobs) {
  if (jobs.empty()) return;
  DevBuf<InflateJob> d_jobs; DevBuf<uint32_t> inf_err, d_order;
  d_jobs.alloc(jobs.size()); inf_err.alloc(2); d_order.alloc(jobs.size());
  CK(cudaMemsetAsync(inf_err.p, 0, 8, ctx->stream));
  CK(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(InflateJob), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<uint32_t> order(jobs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return jobs[x].out_len > jobs[y].out_len; });
  CK(cudaMemcpyAsync(d_order.p, order.data(), order.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaFuncSetAttribute(k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INFLATE_SMEM));
  {
    ProfScope ps__(ctx, "k_inflate");
    k_inflate<<<blocks_for(jobs.size(), inflate_detail::WARPS_PER_BLOCK), INFLATE_THREADS, INFLATE_SMEM, ctx->stream>>>(d_jobs.p, d_order.p, (uint64_t)jobs.size(), inf_err.p);
  }
  ++ctx->launches; CK(cudaGetLastError());
  uint32_t e[2] = {0, 0};
  CK(cudaMemcpyAsync(e, inf_err.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (e[0]) throw std::runtime_error("zlib stream of tile " + std::to_string(e[1]) + " does not inflate (" +
                                     (e[0] == INF_ERR_HEADER ? "bad header" : e[0] == INF_ERR_SIZE ? "size mismatch" : e[0] == INF_ERR_ADLER ? "Adler-32 mismatch" : "corrupt data") + ")");
}

// stage columns (already concatenated over fragments; host memory, or HBM when the tiles were inflated on the device) and
// build the per-array indexes
// `owned` (gdbgpu_array_open*): the same columns, already in HBM, in the order coords, END, REF, ALT, [GT], fields, plain fields
gdbgpu_array* build_array(gdbgpu_ctx* ctx, const HostCols& h, const std::vector<uint64_t>& frag_cells, int64_t lb_row, int64_t num_rows, int gt_phased,
                          const std::vector<FieldMeta>& fm, const std::vector<std::string>& names, const std::vector<PlainMeta>& pm = {},
                          std::vector<DevColumn>* owned = nullptr) {
  static const bool trace = getenv("GDBGPU_TRACE") != nullptr;   // phase times of the staging on stderr
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[gdbgpu] build_array %-10s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  if (h.n >= (1ull << 30)) throw LimitError("more than 2^30 cells in one staged array: shard the column range across GPUs");
  if (h.fields.size() > (size_t)MAX_FIELDS) throw LimitError("more than 15 allele-dependent fields queried");
  if (num_rows < 0 || num_rows >= (1ll << 31)) throw LimitError("row count out of range");
  if (pm.size() > (size_t)MAX_PLAIN || pm.size() != h.plain.size()) throw LimitError("more than 24 plain FORMAT fields");
  std::unique_ptr<gdbgpu_array> a(new gdbgpu_array());
  a->pm = pm;
  a->ctx = ctx; a->n = h.n; a->lb_row = lb_row; a->num_rows = num_rows; a->gt_phased = gt_phased; a->has_gt = h.has_gt; a->fm = fm; a->field_names = names;
  a->n_fragments = (uint32_t)std::max<size_t>(1, frag_cells.size());
  uint64_t n = h.n;
  if (owned) {
    // the columns lie in HBM already: take the blocks over (a variable attribute's fixed block has room for the sentinel offset)
    size_t k = 0;
    auto take_var = [&](DevVar& d) {
      DevColumn& c = (*owned)[k++];
      d.bytes = c.var_bytes;
      d.off.adopt(c.fixed, n + 1); d.val.adopt(c.var, std::max<uint64_t>(c.var_bytes, 1));
      CK(cudaMemcpyAsync(d.off.p + n, &d.bytes, 8, cudaMemcpyHostToDevice, ctx->stream));
    };
    a->coords.adopt((*owned)[k++].fixed, std::max<uint64_t>(2 * n, 2));
    a->end.adopt((*owned)[k++].fixed, std::max<uint64_t>(n, 1));
    take_var(a->ref); take_var(a->alt);
    if (h.has_gt) take_var(a->gt);
    for (size_t f = 0; f < h.fields.size(); ++f) { a->fields.emplace_back(new DevVar()); take_var(*a->fields.back()); }
    for (size_t f = 0; f < h.plain.size(); ++f) {
      a->plain.emplace_back(new DevVar());
      if (!pm[f].fixed_len) { take_var(*a->plain.back()); continue; }
      DevVar& d = *a->plain.back();
      d.bytes = n * (uint64_t)pm[f].fixed_len * (pm[f].elem == ELEM_CHAR ? 1u : 4u);
      d.val.adopt((*owned)[k++].fixed, std::max<uint64_t>(d.bytes, 1));
    }
    CK(cudaStreamSynchronize(ctx->stream));   // (the sentinels came from stack variables)
  } else {
    a->coords.alloc(std::max<uint64_t>(2 * n, 2)); a->end.alloc(std::max<uint64_t>(n, 1));
    if (n) { CK(cudaMemcpyAsync(a->coords.p, h.coords, n * 16, cudaMemcpyDefault, ctx->stream)); CK(cudaMemcpyAsync(a->end.p, h.end, n * 8, cudaMemcpyDefault, ctx->stream)); }
    a->staged_bytes = n * 24;
    upload_var(ctx, h.ref, n, a->ref, a->staged_bytes);
    upload_var(ctx, h.alt, n, a->alt, a->staged_bytes);
    if (h.has_gt) upload_var(ctx, h.gt, n, a->gt, a->staged_bytes);
    for (auto& f : h.fields) { a->fields.emplace_back(new DevVar()); upload_var(ctx, f, n, *a->fields.back(), a->staged_bytes); }
    for (size_t f = 0; f < h.plain.size(); ++f) {
      a->plain.emplace_back(new DevVar());
      if (!pm[f].fixed_len) { upload_var(ctx, h.plain[f], n, *a->plain.back(), a->staged_bytes); continue; }
      DevVar& d = *a->plain.back();
      d.bytes = h.plain[f].bytes;
      if (d.bytes != n * (uint64_t)pm[f].fixed_len * (pm[f].elem == ELEM_CHAR ? 1u : 4u)) throw ArgError("plain field " + pm[f].name + ": value bytes do not match n_cells x fixed_len");
      d.val.alloc(std::max<uint64_t>(d.bytes, 1));
      if (d.bytes) CK(cudaMemcpyAsync(d.val.p, h.plain[f].val, d.bytes, cudaMemcpyDefault, ctx->stream));
      a->staged_bytes += d.bytes;
    }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  lap("upload");

  if (frag_cells.size() > 1 && n) {   // K1
    DevBuf<uint32_t> src; uint64_t n_out = 0;
    fragment_merge(ctx, a->coords.p, n, frag_cells, src, n_out);
    DevBuf<int64_t> c2, e2; c2.alloc(std::max<uint64_t>(2 * n_out, 2)); e2.alloc(std::max<uint64_t>(n_out, 1));
    LAUNCH(ctx, k_gather_fixed, blocks_for(n_out * 2, 256), 256, (const uint8_t*)a->coords.p, src.p, n_out, 16u, (uint8_t*)c2.p);
    LAUNCH(ctx, k_gather_fixed, blocks_for(n_out, 256), 256, (const uint8_t*)a->end.p, src.p, n_out, 8u, (uint8_t*)e2.p);
    CK(cudaStreamSynchronize(ctx->stream));
    a->coords.swap(c2); a->end.swap(e2);
    auto regather = [&](DevVar& v) { DevVar o; gather_var(ctx, v, src.p, n_out, o); v.off.swap(o.off); v.val.swap(o.val); v.bytes = o.bytes; };
    regather(a->ref); regather(a->alt); if (h.has_gt) regather(a->gt);
    for (auto& f : a->fields) regather(*f);
    for (size_t f = 0; f < a->plain.size(); ++f) {
      if (!pm[f].fixed_len) { regather(*a->plain[f]); continue; }
      const uint32_t cb = (uint32_t)pm[f].fixed_len * (pm[f].elem == ELEM_CHAR ? 1u : 4u);
      DevBuf<uint8_t> v2; v2.alloc(std::max<uint64_t>(n_out * cb, 1));
      LAUNCH(ctx, k_gather_bytes, blocks_for(n_out, 256), 256, a->plain[f]->val.p, src.p, n_out, cb, v2.p);
      CK(cudaStreamSynchronize(ctx->stream));
      a->plain[f]->val.swap(v2);
      a->plain[f]->bytes = n_out * cb;
    }
    a->n = n = n_out;
  }
  // K2: classify + row-major index
  a->desc.alloc(std::max<uint64_t>(n, 1)); a->hot.alloc(std::max<uint64_t>(n, 1)); a->effend.alloc(std::max<uint64_t>(n, 1));
  a->rs_cell.alloc(std::max<uint64_t>(n, 1)); a->rs_keys.alloc(std::max<uint64_t>(n, 1)); a->rs_col.alloc(std::max<uint64_t>(n, 1)); a->rs_effend.alloc(std::max<uint64_t>(n, 1));
  a->row_start.alloc((uint64_t)num_rows + 2);
  a->bad_cell.alloc(BAD_CAP); a->bad_pred_col.alloc(BAD_CAP);
  fill_dev_view(a.get());
  if (n) {
    unsigned g = blocks_for(n, 256);
    {   // hot records + warm blobs
      DevBuf<uint32_t> units, offs; DevBuf<uint64_t> total, sums;
      units.alloc(n); offs.alloc(n); total.alloc(1);
      LAUNCH(ctx, k_classify, g, 256, a->dev, a->desc.p, a->hot.p, units.p);
      DevBuf<uint64_t> offs64; offs64.alloc(n);
      exclusive_scan<uint32_t, uint64_t>(ctx, units.p, offs64.p, n, total.p, sums);
      uint64_t h_units = 0;
      CK(cudaMemcpyAsync(&h_units, total.p, 8, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
      if (h_units >= (1ull << 30)) throw LimitError("warm-blob pool above 16 GB: shard the column range across GPUs");
      a->warm.alloc(std::max<uint64_t>(h_units * 4, 4));
      a->dev.warm = a->warm.p;
      LAUNCH(ctx, k_narrow_u64, g, 256, offs64.p, n, offs.p);
      DevBuf<uint32_t> werr; werr.alloc(1);
      CK(cudaMemsetAsync(werr.p, 0, 4, ctx->stream));
      LAUNCH(ctx, k_warm_fill, g, 256, a->dev, a->desc.p, a->hot.p, units.p, offs.p, a->warm.p, werr.p);
      uint32_t h_werr = 0;
      CK(cudaMemcpyAsync(&h_werr, werr.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
      a->cls.alloc(n);
      a->dev.cls = a->cls.p;
      LAUNCH(ctx, k_cls_fill, g, 256, a->hot.p, n, a->cls.p);
      CK(cudaStreamSynchronize(ctx->stream));
      if (h_werr) throw LimitError("a cell's ALT string or GT vector is longer than 2^16 words");
    }
    lap("classify");
    DevBuf<uint32_t> keys, vals; DevBuf<uint8_t> tmp;
    keys.alloc(n); vals.alloc(n);
    LAUNCH(ctx, k_row_keys, g, 256, a->dev, a->desc.p, keys.p, vals.p);
    int bits = 1; while ((1ll << bits) <= num_rows) ++bits;
    radix_sort_pairs(ctx, keys.p, a->rs_keys.p, vals.p, a->rs_cell.p, n, bits, tmp);
    LAUNCH(ctx, k_row_bounds, blocks_for((uint64_t)num_rows + 1, 256), 256, a->rs_keys.p, n, num_rows, a->row_start.p);
    uint64_t n_rs = 0;
    CK(cudaMemcpyAsync(&n_rs, a->row_start.p + num_rows, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    a->n_rs = n_rs;
    DevBuf<unsigned long long> stats; stats.alloc(4);
    long long init[4] = {0, 0, INT64_MIN, 0};
    CK(cudaMemcpyAsync(stats.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_effend_rowout, g, 256, a->dev, a->desc.p, a->effend.p);
    if (n_rs) {
      LAUNCH(ctx, k_rs_cols, blocks_for(n_rs, 256), 256, a->dev, a->rs_cell.p, n_rs, a->rs_col.p);
      LAUNCH(ctx, k_effend, blocks_for(n_rs, 256), 256, a->dev, a->desc.p, a->rs_cell.p, a->rs_col.p, a->rs_keys.p, n_rs, a->rs_effend.p, a->effend.p, stats.p,
             a->bad_cell.p, a->bad_pred_col.p, BAD_CAP);
      LAUNCH(ctx, k_rs_flag_del, blocks_for(n_rs, 256), 256, a->desc.p, a->rs_cell.p, n_rs);
      a->rs_rec.alloc(n_rs);
      LAUNCH(ctx, k_rs_pack, blocks_for(n_rs, 256), 256, a->dev, n_rs, a->rs_rec.p);
    }
    long long hs[4];
    CK(cudaMemcpyAsync(hs, stats.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    int64_t first_last[2];
    CK(cudaMemcpyAsync(&first_last[0], a->coords.p + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&first_last[1], a->coords.p + 2 * (n - 1) + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    a->n_begin = (uint64_t)hs[0]; a->n_bad = (uint64_t)hs[1]; a->max_end = hs[2] == INT64_MIN ? first_last[1] : hs[2]; a->n_del_cells = (uint64_t)hs[3];
    a->min_col = first_last[0]; a->max_col = first_last[1];
    if (a->max_end < a->max_col) a->max_end = a->max_col;
    lap("row index");
  }
  fill_dev_view(a.get());
  return a.release();
}

}  // namespace

// Small host -> device control tables of one batch (strides, bases, site lists, plans) travel in ONE pinned block and one
// copy instead of a dozen pageable cudaMemcpyAsync calls: the device idles while the host prepares a batch's remap launch.
struct CtlArena {
  PinBuf<uint8_t> h; DevBuf<uint8_t> d; size_t used = 0, flushed = 0;
  void begin(size_t cap) { h.ensure(cap); d.ensure(cap); used = flushed = 0; }
  template <class T> T* push(const T* src, size_t n) {
    const size_t off = (used + 15) & ~(size_t)15, bytes = n * sizeof(T);
    if (off + bytes > h.n || off + bytes > d.n) throw CudaError("internal: control arena overflow");
    if (bytes) memcpy(h.p + off, src, bytes);
    used = off + bytes;
    return reinterpret_cast<T*>(d.p + off);
  }
  void flush(cudaStream_t st) {   // copies what was pushed since the last flush
    const size_t from = flushed & ~(size_t)15;
    if (used > from) CK(cudaMemcpyAsync(d.p + from, h.p + from, used - from, cudaMemcpyHostToDevice, st));
    flushed = used;
  }
};

// K3/K4 state of a WINDOW of consecutive sites of a scan: live matrix, class words, site table, nt lists + per-call allele
// LUTs, signature tables, the merged-allele pools.  None of it depends on how the sites are cut into batches, so it is
// computed once for as many sites as a memory budget allows (GDBGPU_FRONT_BYTES, default 6 GiB: 24 bytes per (site, query
// row)) -- k_live_walk / k_site_scan / k_union / k_verify_alleles run with grids of thousands of blocks instead of one
// under-filled wave per batch, no site is scanned twice because a batch kept fewer sites than were guessed, and the one
// host round trip K4 needs (site table -> output layout) happens once per window, not once per batch: the batches that
// follow only plan (host, from the mirrored site table) and launch K5 / K6 back to back.
struct FrontWindow {
  DevBuf<int32_t> live; DevBuf<uint32_t> lcls; DevBuf<SiteRec> sites; DevBuf<uint32_t> agg; DevBuf<uint2> nt, nt_lut; DevBuf<unsigned long long> sig_keys;
  DevBuf<uint32_t> cursors, allele_pos; DevBuf<uint64_t> allele_hash; DevBuf<uint8_t> allele_bytes;
  std::vector<SiteRec> h_sites; uint32_t hc[4] = {0, 0, 0, 0};
  uint64_t first = 0; uint32_t n = 0, sig_slots = 0; bool valid = false;
  // K4d: the permutation rows of the sites' call signatures (k_sig_perm), for the whole window
  DevBuf<uint32_t> sig_perm; CtlArena* ctl = nullptr; uint32_t* sig_off_d = nullptr; uint64_t sig_words = 0;
  ~FrontWindow();
};

FrontWindow::~FrontWindow() { delete ctl; }

// Per-batch device / pinned-host buffers.  They are large (GBs) and expensive to create (cudaMalloc, cudaMallocHost), so a
// context keeps finished scans' buffers and hands them to the next scan.
struct BatchBuffers {
  // per batch (device)
  DevBuf<uint2> blk_map; DevBuf<uint32_t> cursors, lut_first; DevBuf<uint16_t> lut_len; DevBuf<int16_t> lut_in2m;
  const int32_t* live = nullptr;   // the batch's rows of the window's live matrix
  const uint32_t* allele_pos = nullptr; const uint8_t* allele_bytes = nullptr;   // the window's allele pools (SiteRec::allele_first indexes them)
  DevBuf<uint32_t> out_data[MAX_FIELDS + 1]; DevBuf<uint16_t> out_len[MAX_FIELDS + 1];
  CtlArena ctl;   // everything below points into ctl.d
  uint32_t* out_stride[MAX_FIELDS + 1] = {}; uint64_t* out_base[MAX_FIELDS + 1] = {};
  uint8_t* d_site_flags = nullptr; uint32_t *d_n_alleles = nullptr, *d_n_live = nullptr, *d_allele_first = nullptr;
  uint32_t *fast_sites = nullptr, *slow_sites = nullptr, *tile_sites = nullptr, *nt_prefix_d = nullptr;
  uint8_t* fmt_type_d = nullptr; uint32_t* fmt_nps_d = nullptr; uint64_t* fmt_off_d = nullptr;
  // per batch (host mirrors)
  const SiteRec* h_sites = nullptr;   // the batch's part of the window's mirrored site table
  std::vector<uint8_t> h_site_flags; std::vector<uint32_t> h_n_alleles, h_n_live, h_allele_first;
  std::vector<uint64_t> h_base[MAX_FIELDS + 1]; std::vector<uint32_t> h_stride[MAX_FIELDS + 1];
  PinBuf<int64_t> p_site_begin, p_site_end; PinBuf<int32_t> p_live; PinBuf<uint32_t> p_allele_pos, p_lut_first, p_out_data[MAX_FIELDS + 1]; PinBuf<uint8_t> p_allele_bytes;
  PinBuf<uint16_t> p_out_len[MAX_FIELDS + 1], p_lut_len; PinBuf<int16_t> p_lut_in2m;
  // K6 (GDBGPU_SCAN_BCF_FORMAT): per (field, site) plan + the packed FORMAT blocks
  DevBuf<int32_t> fmt_mm; DevBuf<uint8_t> fmt_bytes;
  DevBuf<uint32_t> info_vals, info_cnt, info_value_d; DevBuf<uint8_t> info_valid_d; DevBuf<int32_t> dp_sum_d;
  std::vector<uint32_t> h_info_value; std::vector<uint8_t> h_info_valid; std::vector<int32_t> h_dp_sum;
  DevBuf<uint32_t> esum_values, esum_lens; std::vector<uint32_t> h_vec_len, h_vec_values, h_esum_raw; std::vector<uint64_t> h_vec_off;
  std::vector<int32_t> h_fmt_mm; std::vector<uint8_t> h_fmt_type; std::vector<uint32_t> h_fmt_nps; std::vector<uint64_t> h_fmt_off;
  PinBuf<uint8_t> p_fmt_bytes;
  // state of the batch these buffers hold
  bool valid = false; uint64_t first = 0; uint32_t keep = 0, arg_sites = 0; uint64_t arg_bytes = 0, alg_bytes = 0, fmt_total = 0;
  uint32_t hc[4] = {0, 0, 0, 0};
  DevBuf<uint32_t> site_stats; PinBuf<uint32_t> p_site_stats;   // GDBGPU_SCAN_SITE_STATS: per site, reference-block calls and calls beginning there
  PinBuf<uint32_t> p_tail;            // err[4] + cursors[4] as they stood when the batch's last kernel had run
  cudaEvent_t ev_compute = nullptr, ev_done = nullptr;
  ~BatchBuffers() { if (ev_compute) cudaEventDestroy(ev_compute); if (ev_done) cudaEventDestroy(ev_done); }
};

// ------------------------------------------------------------------------------------------------ scan
struct gdbgpu_scan {
  gdbgpu_array* arr = nullptr; gdbgpu_ctx* ctx = nullptr;
  gdbgpu_scan_params params{}; bool device_output = false;
  ScanDev S{};
  uint64_t cursor = 0;
  std::vector<int64_t> h_rows;
  DevBuf<int32_t> qrow_of_row, init_cell, dlive, ddel, live_excl, del_excl; DevBuf<int64_t> row_of_qrow, init_effend, bp, site_begin, site_end;
  DevBuf<unsigned long long> bitmap; DevBuf<uint32_t> word_rank, bp_cell0, err, del_cells, del_words, site_c0, site_c1; DevBuf<uint64_t> site_start; DevBuf<DelRec> del_recs;
  FrontWindow front;
  std::unique_ptr<BatchBuffers> slot[2];   // borrowed from the context's pool for the lifetime of the scan; [cur] = batch being handed out, the other = prefetched
  int cur = 0; bool bcf_format = false, prefetch = true; int produce_gt = 0, java = 0;
  InfoSet info{};   // INFO scalars to combine + the DP rule's fields (GDBGPU_SCAN_BCF_FORMAT)
  EsumSet esum{};   // vector fields combined with an element-wise sum
  uint32_t cap_allele_recs = 0, cap_allele_bytes = 0, cap_lut = 0;
  int n_out = 0;
  // columnar-iterator view of the query (gdbgpu_scan_cells*): the device list and what was copied out for the caller
  bool site_stats = false, summary_only = false;
  bool cells_only = false, cells_built = false; uint64_t n_cells = 0, n_left = 0;
  uint64_t win_first = 0, win_count = UINT64_MAX;   // the part of the list gdbgpu_scan_cells / _cells_fetch hand out
  DevBuf<uint32_t> cell_list;
  std::vector<uint64_t> h_cell; std::vector<int64_t> h_row, h_begin, h_end;
  std::map<std::pair<int, int>, std::pair<std::vector<uint64_t>, std::vector<uint8_t>>> fetched;
  std::vector<int64_t> ac_column; std::vector<uint64_t> ac_count; std::vector<uint32_t> ac_ref_pos, ac_alt_pos; std::vector<uint8_t> ac_bytes;
};

namespace {

void check_scan_errors(gdbgpu_scan* s, bool& pool_overflow) {
  uint32_t e[4];
  CK(cudaMemcpyAsync(e, s->err.p, sizeof(e), cudaMemcpyDeviceToHost, s->ctx->stream));
  CK(cudaStreamSynchronize(s->ctx->stream));
  if (e[0]) throw OverlapError("Unhandled overlapping variants (a call that is neither a deletion nor a reference block is overlapped by a later call of the same sample)");
  if (e[1] == 1) throw LimitError("a call has more than 1023 ALT alleles");
  if (e[1] == 2) throw LimitError("more than 512 distinct ALT alleles at one site");
  if (e[1] == 3) throw LimitError("ploidy above 8 needs the general-ploidy genotype enumeration, which keeps at most 8 slots");
  if (e[1] == 4) throw LimitError("merged REF assembled from more than 32 pieces");
  if (e[1] == 6) throw LimitError("two different ALT alleles of one site share a 64-bit hash (detected by the byte comparison; never merged silently)");
  pool_overflow = e[2] != 0;
  if (pool_overflow) CK(cudaMemsetAsync(s->err.p + 2, 0, 4, s->ctx->stream));
}

void scan_prepare(gdbgpu_scan* s) {
  gdbgpu_array* a = s->arr; gdbgpu_ctx* ctx = s->ctx; ScanDev& S = s->S;
  const gdbgpu_scan_params& p = s->params;
  S.coords = a->coords.p;
  S.scan_full = (p.flags & GDBGPU_SCAN_FULL) ? 1 : 0;
  S.handle_del = (p.flags & GDBGPU_SCAN_NO_DELETIONS) ? 0 : 1;
  S.min_pl_gt = (p.flags & GDBGPU_SCAN_MIN_PL_GT) ? 1 : 0;
  S.want_lut = (p.flags & GDBGPU_SCAN_WANT_LUT) ? 1 : 0;
  S.max_alt = p.max_diploid_alt_alleles ? p.max_diploid_alt_alleles : 50u;
  s->device_output = (p.flags & GDBGPU_SCAN_DEVICE_OUTPUT) != 0;
  s->bcf_format = (p.flags & GDBGPU_SCAN_BCF_FORMAT) != 0;
  s->produce_gt = (p.flags & GDBGPU_SCAN_PRODUCE_GT) ? 1 : 0;
  s->java = (p.flags & GDBGPU_SCAN_HTSJDK_CONVENTIONS) ? 1 : 0;
  s->prefetch = !(p.flags & GDBGPU_SCAN_NO_PREFETCH);
  s->cells_only = (p.flags & GDBGPU_SCAN_CELLS_ONLY) != 0;
  s->site_stats = (p.flags & GDBGPU_SCAN_SITE_STATS) != 0;
  s->summary_only = (p.flags & GDBGPU_SCAN_SUMMARY_ONLY) != 0;
  if (s->summary_only && ((p.flags & GDBGPU_SCAN_DEVICE_OUTPUT) || (p.flags & GDBGPU_SCAN_BCF_FORMAT) || (p.flags & GDBGPU_SCAN_WANT_LUT)))
    throw ArgError("GDBGPU_SCAN_SUMMARY_ONLY goes with host output of the per-site tables only");
  s->info.n = 0;
  s->esum.n = 0;
  for (size_t f = 0; f < a->fields.size(); ++f) if (p.esum_remapped_mask & (1u << f)) { s->esum.remapped[s->esum.n] = 1; s->esum.index[s->esum.n] = (int32_t)f; ++s->esum.n; }
  for (size_t f = 0; f < a->pm.size(); ++f) {
    if (a->pm[f].info_op == GDBGPU_INFO_ELEMENT_WISE_SUM) { s->esum.remapped[s->esum.n] = 0; s->esum.index[s->esum.n] = (int32_t)f; ++s->esum.n; }
    else if (a->pm[f].info_op) { s->info.field[s->info.n] = (int32_t)f; s->info.op[s->info.n] = a->pm[f].info_op; ++s->info.n; }
  }
  auto dp_index = [&](uint32_t one_based) -> int32_t {
    if (!one_based) return -1;
    if (one_based > a->pm.size() || a->pm[one_based - 1].fixed_len < 1 || a->pm[one_based - 1].elem != ELEM_INT32) throw ArgError("DP rule: not a fixed-length int32 plain field");
    return (int32_t)one_based - 1;
  };
  s->info.dp_info = dp_index(p.dp_info_field); s->info.dp_min_dp = dp_index(p.dp_min_dp_field); s->info.dp_format = dp_index(p.dp_format_field);
  // queried rows: sorted, unique, inside the array (variant_query_config.cc:60-117)
  if (p.query_rows) {
    s->h_rows.assign(p.query_rows, p.query_rows + p.n_query_rows);
    std::sort(s->h_rows.begin(), s->h_rows.end());
    s->h_rows.erase(std::unique(s->h_rows.begin(), s->h_rows.end()), s->h_rows.end());
    s->h_rows.erase(std::remove_if(s->h_rows.begin(), s->h_rows.end(), [&](int64_t r) { return r < a->lb_row || r >= a->lb_row + a->num_rows; }), s->h_rows.end());
  } else { s->h_rows.resize((size_t)a->num_rows); for (int64_t r = 0; r < a->num_rows; ++r) s->h_rows[(size_t)r] = a->lb_row + r; }
  S.nq = s->h_rows.size();
  s->err.alloc(4); CK(cudaMemsetAsync(s->err.p, 0, 16, ctx->stream));
  S.err = s->err.p;
  s->row_of_qrow.alloc(std::max<uint64_t>(S.nq, 1));
  if (S.nq) CK(cudaMemcpyAsync(s->row_of_qrow.p, s->h_rows.data(), S.nq * 8, cudaMemcpyHostToDevice, ctx->stream));
  S.row_of_qrow = s->row_of_qrow.p;
  S.qrow_of_row = nullptr;
  if (p.query_rows) {
    std::vector<int32_t> m((size_t)a->num_rows, -1);
    for (size_t q = 0; q < s->h_rows.size(); ++q) m[(size_t)(s->h_rows[q] - a->lb_row)] = (int32_t)q;
    s->qrow_of_row.alloc(std::max<size_t>(m.size(), 1));
    if (!m.empty()) CK(cudaMemcpyAsync(s->qrow_of_row.p, m.data(), m.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    S.qrow_of_row = s->qrow_of_row.p;
  }
  s->init_cell.alloc(std::max<uint64_t>(S.nq, 1)); s->init_effend.alloc(std::max<uint64_t>(S.nq, 1));
  S.init_cell = s->init_cell.p; S.init_effend = s->init_effend.p;
  S.n_sites = 0; S.m = 0;
  s->n_out = (a->has_gt ? 1 : 0) + (int)a->fields.size();
  if (a->n == 0 || S.nq == 0) return;
  // clamp the interval to where data can be
  int64_t hi_data = std::max(a->max_end, a->max_col);
  if (S.scan_full) { S.q_lo = a->min_col; S.q_hi = hi_data; }
  else { S.q_lo = p.col_lo; S.q_hi = std::min(p.col_hi, hi_data); if (p.col_lo > p.col_hi) throw ArgError("col_lo > col_hi"); }
  if (S.q_lo > S.q_hi) return;
  // bit 0 of the breakpoint bitmap stands for q_lo itself (gdb_kernels.cuh: bp_bit); every other breakpoint is a begin column
  // of a staged cell, or one past a (clipped) end, hence at or after min(first staged column, q_hi + 1)
  S.base = std::max(S.q_lo, std::min(a->min_col - 1, S.q_hi));
  S.span_bits = (uint64_t)(S.q_hi - S.base) + 3ull;
  if (S.span_bits > (1ull << 37)) throw LimitError("query column span above 2^37 positions: split the interval");
  S.n_words = (S.span_bits + 63) / 64;
  // forward range
  DevBuf<uint64_t> d_bounds; d_bounds.alloc(3);
  uint64_t hb[3] = {0, 0, a->n};
  if (!S.scan_full) {
    LAUNCH(ctx, k_bounds, 1, 32, a->coords.p, a->n, S.q_lo, S.q_hi, d_bounds.p);
    CK(cudaMemcpyAsync(hb, d_bounds.p, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  S.iL = hb[0]; S.i0 = hb[1]; S.i1 = hb[2];
  if (s->cells_only) {   // the columnar iterator neither merges calls nor minds overlaps: the left sweep is all it needs
    s->bitmap.alloc(S.n_words);
    CK(cudaMemsetAsync(s->bitmap.p, 0, S.n_words * 8, ctx->stream));
    S.bitmap = s->bitmap.p;
    S.i0 = S.scan_full ? 0 : hb[0];
    LAUNCH(ctx, k_left_sweep, blocks_for(S.nq, 256), 256, a->dev, S);
    return;
  }
  if (a->n_bad) {
    if (a->n_bad > BAD_CAP) throw OverlapError("Unhandled overlapping variants (too many to enumerate)");
    LAUNCH(ctx, k_check_evictions, blocks_for(a->n_bad, 128), 128, a->dev, S, a->bad_cell.p, a->bad_pred_col.p, (uint32_t)a->n_bad);
  }
  // K2b deletion pre-pass
  S.n_del = 0;
  if (S.handle_del && a->n_del_cells) {
    DevBuf<uint32_t> flag, pos, total, sums, words;
    flag.alloc(a->n); pos.alloc(a->n); total.alloc(1);
    LAUNCH(ctx, k_flag_del, blocks_for(a->n, 256), 256, a->dev, flag.p);
    exclusive_scan<uint32_t, uint32_t>(ctx, flag.p, pos.p, a->n, total.p, sums);
    uint32_t nd = 0;
    CK(cudaMemcpyAsync(&nd, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    if (nd) {
      s->del_cells.alloc(nd); s->del_recs.alloc(nd); words.alloc(nd);
      LAUNCH(ctx, k_compact_idx, blocks_for(a->n, 256), 256, flag.p, pos.p, a->n, s->del_cells.p);
      LAUNCH(ctx, k_del_size, blocks_for(nd, 128), 128, a->dev, s->del_cells.p, (uint64_t)nd, words.p);
      exclusive_scan<uint32_t, uint32_t>(ctx, words.p, words.p, nd, total.p, sums);
      uint32_t nw = 0;
      CK(cudaMemcpyAsync(&nw, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
      s->del_words.alloc(std::max<uint32_t>(nw, 1));
      LAUNCH(ctx, k_del_fill, blocks_for(nd, 128), 128, a->dev, S.min_pl_gt, s->del_cells.p, (uint64_t)nd, words.p, s->del_recs.p, s->del_words.p, s->err.p);
      CK(cudaStreamSynchronize(ctx->stream));
      S.n_del = nd; S.del_cells = s->del_cells.p; S.del_recs = s->del_recs.p; S.del_words = s->del_words.p;
    }
  }
  // K3: breakpoints
  s->bitmap.alloc(S.n_words); s->word_rank.alloc(S.n_words + 1);
  CK(cudaMemsetAsync(s->bitmap.p, 0, S.n_words * 8, ctx->stream));
  S.bitmap = s->bitmap.p; S.word_rank = s->word_rank.p;
  LAUNCH(ctx, k_left_sweep, blocks_for(S.nq, 256), 256, a->dev, S);
  if (S.i1 > S.i0) LAUNCH(ctx, k_mark_events, blocks_for(S.i1 - S.i0, 256), 256, a->dev, S);
  DevBuf<uint32_t> cnt, total32, sums32;
  cnt.alloc(S.n_words); total32.alloc(1);
  LAUNCH(ctx, k_popc, blocks_for(S.n_words, 256), 256, S.bitmap, S.n_words, cnt.p);
  exclusive_scan<uint32_t, uint32_t>(ctx, cnt.p, s->word_rank.p, S.n_words, total32.p, sums32);
  uint32_t m = 0;
  CK(cudaMemcpyAsync(&m, total32.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  S.m = m;
  if (m < 2) { S.m = m; return; }
  s->bp.alloc(m); s->bp_cell0.alloc(m); s->dlive.alloc(m); s->ddel.alloc(m); s->live_excl.alloc(m); s->del_excl.alloc(m); s->site_start.alloc((uint64_t)m + 1);
  S.bp = s->bp.p; S.bp_cell0 = s->bp_cell0.p; S.dlive = s->dlive.p; S.ddel = s->ddel.p; S.site_start = s->site_start.p;
  CK(cudaMemsetAsync(s->dlive.p, 0, (size_t)m * 4, ctx->stream)); CK(cudaMemsetAsync(s->ddel.p, 0, (size_t)m * 4, ctx->stream));
  LAUNCH(ctx, k_extract_bp, blocks_for(S.n_words, 256), 256, S);
  LAUNCH(ctx, k_deltas, blocks_for((S.i1 - S.iL) + S.nq, 256), 256, a->dev, S);
  DevBuf<int32_t> total_i, sums_i; total_i.alloc(1);
  exclusive_scan<int32_t, int32_t>(ctx, s->dlive.p, s->live_excl.p, m, total_i.p, sums_i);
  exclusive_scan<int32_t, int32_t>(ctx, s->ddel.p, s->del_excl.p, m, total_i.p, sums_i);
  DevBuf<uint64_t> cnt64, total64, sums64; cnt64.alloc(m); total64.alloc(1);
  LAUNCH(ctx, k_site_counts, blocks_for(m, 256), 256, S, s->live_excl.p, s->del_excl.p, cnt64.p);
  exclusive_scan<uint64_t, uint64_t>(ctx, cnt64.p, s->site_start.p, m, total64.p, sums64);
  uint64_t T = 0;
  CK(cudaMemcpyAsync(&T, total64.p, 8, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpyAsync(s->site_start.p + m, &T, 8, cudaMemcpyHostToDevice, ctx->stream));
  S.n_sites = T;
  if (T) {
    s->site_begin.alloc(T); s->site_end.alloc(T); s->site_c0.alloc(T); s->site_c1.alloc(T);
    S.site_begin = s->site_begin.p; S.site_end = s->site_end.p; S.site_c0 = s->site_c0.p; S.site_c1 = s->site_c1.p;
    LAUNCH(ctx, k_fill_sites, blocks_for(T, 256), 256, S, s->del_excl.p);
  }
  bool ovf = false;
  check_scan_errors(s, ovf);
}

// words a sample's vector needs at this site (the algorithmic figure) ...
uint32_t stride_needed(const gdbgpu_array* a, const SiteRec& r, int o) {
  if (a->has_gt && o == 0) return r.max_gt_len;
  int f = o - (a->has_gt ? 1 : 0);
  if (r.max_in_len[f] == 0) return 0;
  if (!(r.flags & 2u)) return r.max_in_len[f];
  int kind = a->fm[f].kind;
  if (kind == KIND_G && (r.flags & 4u)) return 0;
  return num_elements(kind, r.n_alleles - 1u, r.max_ploidy[f]);
}
// ... and the stride actually used: remapped fields are padded to a multiple of 4 words so that every sample's vector
// is 16-byte aligned (the tile kernel stores 128 bits per lane); the padding holds BCF vector-end words
uint32_t stride_of(const gdbgpu_array* a, const SiteRec& r, int o) {
  uint32_t st = stride_needed(a, r, o);
  if (a->has_gt && o == 0) return st;
  if (r.flags & 2u) st = (st + 3u) & ~3u;
  return st;
}

// Runs every kernel of the batch starting at site `first` into `bb` and queues the device -> host copies of its results
// on the copy stream.  Nothing here waits for those copies: publish_batch does.
// K3/K4 for the window of sites starting at first_site (see FrontWindow)
void compute_front(gdbgpu_scan* s, uint64_t first_site, uint32_t max_sites) {
  gdbgpu_array* a = s->arr; gdbgpu_ctx* ctx = s->ctx; ScanDev& S = s->S; FrontWindow& w = s->front;
  w.valid = false;
  // batches still in flight read the window they were cut from (their device -> host copies run on the copy stream)
  for (auto& sl : s->slot) if (sl && sl->ev_done) CK(cudaStreamWaitEvent(ctx->stream, sl->ev_done, 0));
  const uint64_t budget = [] { const char* e = getenv("GDBGPU_FRONT_BYTES"); const uint64_t v = e ? strtoull(e, nullptr, 10) : 0; return v ? v : (6ull << 30); }();
  const uint64_t remaining = S.n_sites - first_site;
  // a caller paging in small batches gets small windows (every batch copies the window's allele pool)
  const uint32_t nb = (uint32_t)std::min<uint64_t>({remaining, std::max<uint64_t>(1, budget / std::max<uint64_t>(1, S.nq * 24)), 1u << 20, max_sites ? std::max<uint64_t>(64, 16ull * max_sites) : ~0ull});
  for (;;) {
    w.live.ensure((uint64_t)nb * S.nq); w.lcls.ensure((uint64_t)nb * S.nq); w.sites.ensure(nb); w.cursors.ensure(4);
    w.agg.ensure((uint64_t)nb * AGG_WORDS); w.nt.ensure((uint64_t)nb * S.nq); w.nt_lut.ensure((uint64_t)nb * S.nq);
    CK(cudaMemsetAsync(w.agg.p, 0, (size_t)nb * AGG_WORDS * 4, ctx->stream));
    if (!s->cap_allele_recs) { s->cap_allele_recs = 1u << 16; s->cap_allele_bytes = 1u << 20; }
    while (s->cap_allele_recs < 16ull * nb && s->cap_allele_recs < (1u << 28)) { s->cap_allele_recs *= 2; s->cap_allele_bytes *= 2; }
    w.allele_pos.ensure(2ull * s->cap_allele_recs); w.allele_hash.ensure(s->cap_allele_recs); w.allele_bytes.ensure(s->cap_allele_bytes);
    CK(cudaMemsetAsync(w.cursors.p, 0, 16, ctx->stream));
    BatchDev B{};
    B.first_site = first_site; B.n_sites = nb; B.live = w.live.p; B.lcls = w.lcls.p; B.sites = w.sites.p; B.agg = w.agg.p; B.nt = w.nt.p; B.nt_lut = w.nt_lut.p; B.cursors = w.cursors.p;
    // signature tables (gdb_kernels2.cuh: nt_lut): one slot per 64 samples, only for scans wide enough to repeat signatures
    B.sig_slots = 0; B.sig_keys = nullptr;
    if (S.nq >= SIG_MIN_ROWS) { uint32_t sl = 64; while (sl < SIG_MAX_SLOTS && (uint64_t)sl * 64 < S.nq) sl *= 2; B.sig_slots = sl; w.sig_keys.ensure((uint64_t)nb * sl); B.sig_keys = w.sig_keys.p; }
    w.sig_slots = B.sig_slots;
    B.allele_pos = w.allele_pos.p; B.allele_hash = w.allele_hash.p; B.cap_allele_recs = s->cap_allele_recs; B.allele_bytes = w.allele_bytes.p; B.cap_allele_bytes = s->cap_allele_bytes;
    {   // sites walked by one thread after one binary search: long chunks amortise the search, short ones keep the device full
      const uint64_t want_threads = 400000;
      uint32_t chunk = (uint32_t)std::min<uint64_t>(nb, std::max<uint64_t>(32, (uint64_t)nb * S.nq / want_threads));
      LAUNCH(ctx, k_live_walk, dim3(blocks_for(S.nq, 256), (nb + chunk - 1) / chunk), 256, a->dev, S, first_site, nb, chunk, w.live.p, w.lcls.p);
    }
    LAUNCH(ctx, k_site_scan, dim3(nb, blocks_for(S.nq, SSCAN_ROWS)), SSCAN_THREADS, a->dev, S, B);
    LAUNCH(ctx, k_union, nb, UNION_THREADS, a->dev, S, B);
    w.h_sites.resize(nb);
    CK(cudaMemcpyAsync(w.h_sites.data(), w.sites.p, (size_t)nb * sizeof(SiteRec), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(w.hc, w.cursors.p, 16, cudaMemcpyDeviceToHost, ctx->stream));   // allele pools are final after k_union
    bool ovf = false;
    check_scan_errors(s, ovf);
    if (ovf) { s->cap_allele_recs *= 4; s->cap_allele_bytes *= 4; continue; }
    {   // byte comparison behind every hash match (a throughput pass over the window's nt lists; its verdict travels with the batches' tails)
      uint32_t max_nt = 0;
      for (uint32_t i = 0; i < nb; ++i) if (w.h_sites[i].n_alleles >= 2u) max_nt = std::max(max_nt, w.h_sites[i].n_nt);
      if (max_nt && B.sig_keys) CK(cudaMemsetAsync(w.sig_keys.p, 0xFF, (size_t)nb * B.sig_slots * 8, ctx->stream));
      for (uint32_t s0 = 0; s0 < nb && max_nt; s0 += 65535u) {   // grid.y is the site
        BatchDev Bp = B; const uint32_t ns = std::min<uint32_t>(65535u, nb - s0);
        Bp.first_site = first_site + s0; Bp.n_sites = ns; Bp.live = B.live + (uint64_t)s0 * S.nq; Bp.lcls = B.lcls + (uint64_t)s0 * S.nq; Bp.sites = B.sites + s0; Bp.agg = B.agg + (uint64_t)s0 * AGG_WORDS;
        Bp.nt = B.nt + (uint64_t)s0 * S.nq; Bp.nt_lut = B.nt_lut + (uint64_t)s0 * S.nq; if (B.sig_keys) Bp.sig_keys = B.sig_keys + (uint64_t)s0 * B.sig_slots;
        LAUNCH(ctx, k_verify_alleles, dim3(blocks_for(max_nt, VERIFY_THREADS), ns), VERIFY_THREADS, a->dev, S, Bp);
      }
    }
    // K4d, once per window: one permutation row per occupied signature slot of every site the band kernel serves (it needs the
    // output strides, which are a function of the site alone)
    w.sig_words = 0; w.sig_off_d = nullptr;
    {
      const char* k5_env = getenv("GDBGPU_K5");
      const bool k5_mix = !k5_env || !strcmp(k5_env, "mix");
      const int n_out = s->n_out, o_f0 = a->has_gt ? 1 : 0;
      if (k5_mix && !s->summary_only && B.sig_keys && a->fields.size() >= 1 && a->fields.size() <= (size_t)TILE_MAX_FIELDS) {
        std::vector<uint32_t> sig_off(nb, 0xFFFFFFFFu);
        std::vector<std::vector<uint32_t>> stride(n_out, std::vector<uint32_t>(nb, 0));
        uint64_t words = 0;
        for (uint32_t i = 0; i < nb; ++i) {
          const SiteRec& r = w.h_sites[i];
          for (int o = 0; o < n_out; ++o) stride[o][i] = stride_of(a, r, o);
          const bool tile_path = r.n_alleles <= (uint32_t)TILE_MAX_A && r.max_ain <= (uint32_t)TILE_MAX_A && r.max_gt_len <= (uint32_t)TILE_MAX_GT && r.max_ploidy_all <= 2u;
          if (!tile_path || !(r.flags & 2u) || !r.n_nt || r.n_alleles > 0xFDu) continue;
          uint32_t row_words = SIG_HDR_BYTES / 4; bool ok = true;
          for (int o = o_f0; o < n_out; ++o) { row_words += stride[o][i] >> 2; ok = ok && r.max_in_len[o - o_f0] <= PERM_MAX_IN; }
          if (!ok || words + (uint64_t)B.sig_slots * row_words >= (1ull << 29)) continue;   // (2 GiB of rows at most; the sites beyond rebuild their LUTs per call)
          sig_off[i] = (uint32_t)words; words += (uint64_t)B.sig_slots * row_words;
        }
        if (words) {
          if (!w.ctl) w.ctl = new CtlArena();
          w.ctl->begin((size_t)(n_out + 1) * nb * 4 + 256);
          RemapOut Ow{}; Ow.n_out = n_out;
          for (int o = 0; o < n_out; ++o) Ow.f[o].stride = w.ctl->push(stride[o].data(), nb);
          w.sig_off_d = w.ctl->push(sig_off.data(), nb);
          w.ctl->flush(ctx->stream);
          w.sig_perm.ensure(words);
          LAUNCH(ctx, k_sig_perm, nb, SIGPERM_THREADS, a->dev, B, Ow, w.sig_off_d, w.sig_perm.p);
          w.sig_words = words;
        }
      }
    }
    w.first = first_site; w.n = nb; w.valid = true;
    return;
  }
}

void compute_batch(gdbgpu_scan* s, BatchBuffers& bb, uint64_t first_site, uint32_t max_sites, uint64_t max_bytes) {
  gdbgpu_array* a = s->arr; gdbgpu_ctx* ctx = s->ctx; ScanDev& S = s->S;
  bb.valid = false; bb.first = first_site; bb.arg_sites = max_sites; bb.arg_bytes = max_bytes; bb.keep = 0; bb.fmt_total = 0;
  if (!bb.ev_compute) { CK(cudaEventCreateWithFlags(&bb.ev_compute, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&bb.ev_done, cudaEventDisableTiming)); }
  if (max_bytes == 0) max_bytes = 1ull << 30;
  FrontWindow& w = s->front;
  if (!w.valid || first_site < w.first || first_site >= w.first + w.n) compute_front(s, first_site, max_sites);
  const uint32_t woff = (uint32_t)(first_site - w.first);
  const uint32_t nb = (uint32_t)std::min<uint64_t>({(uint64_t)w.n - woff, max_sites ? max_sites : 0xFFFFFFFFull, 1u << 20});
  const int n_out = s->n_out;
  {
    bb.cursors.ensure(4);
    CK(cudaMemsetAsync(bb.cursors.p, 0, 16, ctx->stream));   // [2]: the batch's LUT cursor (the remap kernels advance it)
    bb.h_sites = w.h_sites.data() + woff; bb.live = w.live.p + (uint64_t)woff * S.nq; bb.allele_pos = w.allele_pos.p; bb.allele_bytes = w.allele_bytes.p;
    bb.hc[0] = w.hc[0]; bb.hc[1] = w.hc[1]; bb.hc[2] = bb.hc[3] = 0;
    BatchDev B{};
    B.first_site = first_site; B.n_sites = nb; B.live = w.live.p + (uint64_t)woff * S.nq; B.lcls = w.lcls.p + (uint64_t)woff * S.nq; B.sites = w.sites.p + woff; B.agg = w.agg.p + (uint64_t)woff * AGG_WORDS;
    B.nt = w.nt.p + (uint64_t)woff * S.nq; B.nt_lut = w.nt_lut.p + (uint64_t)woff * S.nq; B.cursors = bb.cursors.p;
    B.sig_slots = w.sig_slots; B.sig_keys = w.sig_slots ? w.sig_keys.p + (uint64_t)woff * w.sig_slots : nullptr;
    B.allele_pos = w.allele_pos.p; B.allele_hash = w.allele_hash.p; B.cap_allele_recs = s->cap_allele_recs; B.allele_bytes = w.allele_bytes.p; B.cap_allele_bytes = s->cap_allele_bytes;
    // strides / bases; keep the prefix of sites that fits
    for (int o = 0; o < n_out; ++o) { bb.h_base[o].assign(1, 0); bb.h_stride[o].clear(); }
    uint64_t bytes = 0, alg_bytes = 0; uint32_t keep = 0; uint64_t lut_words = 0;
    for (uint32_t i = 0; i < nb; ++i) {
      uint64_t site_bytes = 0, site_alg = 0; uint32_t st[MAX_FIELDS + 1];
      for (int o = 0; o < n_out; ++o) {
        st[o] = stride_of(a, bb.h_sites[i], o);
        site_bytes += (((uint64_t)st[o] * S.nq + 3) & ~3ull) * 4 + S.nq * 2;
        site_alg += (uint64_t)stride_needed(a, bb.h_sites[i], o) * S.nq * 4;
      }
      if (i > 0 && bytes + site_bytes > max_bytes) break;
      for (int o = 0; o < n_out; ++o) { bb.h_stride[o].push_back(st[o]); bb.h_base[o].push_back((bb.h_base[o][i] + (uint64_t)st[o] * S.nq + 3) & ~3ull); }   // 16-byte aligned sites
      bytes += site_bytes; alg_bytes += site_alg; lut_words += (uint64_t)bb.h_sites[i].max_ain * S.nq; keep = i + 1;
    }
    // split the batch: the tile kernels take every ordinary site, the generic kernel the pathological ones.  k_remap_tile
    // remaps the non-trivial calls of its sites itself (phase C); k_remap_nt only serves the copy-through sites
    // GDBGPU_K5 selects the remap kernel of the remapped sites: "mix" (default: k_remap_mix, tile blocks and nt blocks in one
    // grid), "tile" (k_remap_tile, the tile's non-trivial calls as a third phase of the block), "tma" (k_remap_tma)
    const char* k5_env = getenv("GDBGPU_K5");
    const bool k5_mix = !k5_env || !strcmp(k5_env, "mix");
    auto nt_blocks = [k5_mix](const SiteRec& r) -> uint32_t { return ((r.flags & 2u) && !k5_mix) ? 0u : (r.n_nt + NT_GROUPS - 1) / NT_GROUPS; };
    std::vector<uint32_t> fast, slow, nt_prefix, order;
    uint32_t jk_n = 4;
    for (uint32_t i = 0; i < keep; ++i) {
      const SiteRec& r = bb.h_sites[i];
      bool ok = a->fields.size() >= 1 && a->fields.size() <= (size_t)TILE_MAX_FIELDS && r.n_alleles <= (uint32_t)TILE_MAX_A && r.max_ain <= (uint32_t)TILE_MAX_A &&
                r.max_gt_len <= (uint32_t)TILE_MAX_GT && r.max_ploidy_all <= 2u;
      if (ok) {
        nt_prefix.push_back(fast.empty() ? 0u : nt_prefix.back() + nt_blocks(bb.h_sites[fast.back()]));
        fast.push_back(i);
        jk_n = std::max(jk_n, r.n_alleles * (r.n_alleles + 1u) / 2u);
      } else slow.push_back(i);
    }
    size_t n_remap = 0, n_tma = 0; uint32_t tma_smem_words = 0;
    if (!fast.empty()) {
      nt_prefix.push_back(nt_prefix.back() + nt_blocks(bb.h_sites[fast.back()]));   // blocks of k_remap_nt, per site
      // remapped sites -> k_remap_tma when every tile of the block fits shared memory (the tiles leave through the bulk-copy
      // engine), else k_remap_tile; copy-through sites -> k_copy_tile.  All three lists live in tile_sites, in that order
      // GDBGPU_K5=tma: tiles built in shared memory and stored by the bulk-copy engine (k_remap_tma).  Measured slower than
      // k_remap_tile's direct STG.128 stream on this workload (profiles/r02_k5_tma_experiments.json), so it is opt-in
      const bool no_tma = !(k5_env && !strcmp(k5_env, "tma"));
      auto tma_words = [&](uint32_t i) -> uint32_t {   // words per sample over the remapped fields, 0 = not eligible
        if (no_tma || !(bb.h_sites[i].flags & 2u)) return 0;
        uint32_t words = 0;
        for (int o = a->has_gt ? 1 : 0; o < n_out; ++o) { if (bb.h_stride[o][i] > 4u * TMA_MAXQ) return 0; words += bb.h_stride[o][i]; }
        return words <= (uint32_t)TMA_SMEM_WORDS ? words : 0;
      };
      order.reserve(fast.size());
      for (uint32_t i : fast) { const uint32_t w = tma_words(i); if (w) { order.push_back(i); tma_smem_words = std::max(tma_smem_words, w); } }
      n_tma = order.size();
      for (uint32_t i : fast) if ((bb.h_sites[i].flags & 2u) && !tma_words(i)) order.push_back(i);
      n_remap = order.size();
      for (uint32_t i : fast) if (!(bb.h_sites[i].flags & 2u)) order.push_back(i);
    }
    bb.h_site_flags.resize(keep); bb.h_n_alleles.resize(keep); bb.h_n_live.resize(keep); bb.h_allele_first.resize(keep);
    for (uint32_t i = 0; i < keep; ++i) { const SiteRec& r = bb.h_sites[i]; bb.h_site_flags[i] = (uint8_t)r.flags; bb.h_n_alleles[i] = r.n_alleles; bb.h_n_live[i] = r.n_live; bb.h_allele_first[i] = r.allele_first; }
    // one pinned block, one copy: bases, strides, site lists, site table (+ the FORMAT plan later)
    bb.ctl.begin((size_t)n_out * ((size_t)keep * 12 + 64) + (size_t)keep * 56 + (size_t)(n_out + MAX_PLAIN) * keep * 13 + 4096);
    RemapOut O{};
    O.n_out = n_out;
    for (int o = 0; o < n_out; ++o) {
      bb.out_data[o].ensure(s->summary_only ? 1 : std::max<uint64_t>(bb.h_base[o][keep], 1)); bb.out_len[o].ensure(s->summary_only ? 1 : (uint64_t)keep * S.nq);
      bb.out_base[o] = bb.ctl.push(bb.h_base[o].data(), (size_t)keep + 1);
      bb.out_stride[o] = bb.ctl.push(bb.h_stride[o].data(), keep);
      O.f[o] = FieldOutDev{bb.out_data[o].p, bb.out_base[o], bb.out_stride[o], bb.out_len[o].p};
    }
    bb.fast_sites = bb.ctl.push(fast.data(), fast.size()); bb.nt_prefix_d = bb.ctl.push(nt_prefix.data(), nt_prefix.size());
    bb.tile_sites = bb.ctl.push(order.data(), order.size()); bb.slow_sites = bb.ctl.push(slow.data(), slow.size());
    if (s->device_output) {
      bb.d_site_flags = bb.ctl.push(bb.h_site_flags.data(), keep); bb.d_n_alleles = bb.ctl.push(bb.h_n_alleles.data(), keep);
      bb.d_n_live = bb.ctl.push(bb.h_n_live.data(), keep); bb.d_allele_first = bb.ctl.push(bb.h_allele_first.data(), keep);
    }
    bb.ctl.flush(ctx->stream);
    if (s->site_stats) {   // the copy rides the compute stream: it is done before ev_done is
      bb.site_stats.ensure((size_t)keep * 2); bb.p_site_stats.ensure((size_t)keep * 2);
      LAUNCH(ctx, k_site_stats, keep, 256, a->dev, S, bb.live, first_site, keep, bb.site_stats.p);
      CK(cudaMemcpyAsync(bb.p_site_stats.p, bb.site_stats.p, (size_t)keep * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (S.want_lut) {
      if (lut_words >= (1ull << 32)) throw LimitError("LUT output of one batch above 2^32 entries: lower max_sites");
      s->cap_lut = (uint32_t)std::max<uint64_t>(lut_words, 1);
      bb.lut_first.ensure((uint64_t)keep * S.nq); bb.lut_len.ensure((uint64_t)keep * S.nq); bb.lut_in2m.ensure(s->cap_lut);
      O.lut_first = bb.lut_first.p; O.lut_len = bb.lut_len.p; O.lut_in2m = bb.lut_in2m.p; O.cap_lut = s->cap_lut;
    }
    B.n_sites = keep;
    if (!fast.empty() && !s->summary_only) {   // per-site summaries need no remapped fields
      const int nf = (int)a->fields.size();
      const TileTabs tabs{ctx->d_cls_tab, ctx->d_jk_tab};
      const uint32_t tiles = (uint32_t)((S.nq + TILE_ROWS - 1) / TILE_ROWS);
      if (tiles > 65535u) throw LimitError("more than 16.7 M query rows");
      const int key = nf * 2 + (a->has_gt ? 1 : 0);
      const uint32_t tma_tiles = (uint32_t)((S.nq + TMA_ROWS - 1) / TMA_ROWS);
      const size_t tma_smem = (size_t)TMA_ROWS * tma_smem_words * 4;
      for (size_t first = 0; first < n_tma; first += 65535) {
        const unsigned ny = (unsigned)std::min<size_t>(65535, n_tma - first);
        auto launch_tma = [&](auto kern) {
          if (!(ctx->tma_attr_set & (1u << key))) {   // tiles + static scratch exceed the default 48 KB per block
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_ROWS * TMA_SMEM_WORDS * 4));
            ctx->tma_attr_set |= 1u << key;
          }
          ProfScope ps__(ctx, "k_remap_tma");
          kern<<<dim3(tma_tiles, ny), TMA_THREADS, tma_smem, ctx->stream>>>(a->dev, S, B, O, bb.tile_sites + first, tabs);
        };
        switch (key) {
          case 2: launch_tma(k_remap_tma<1, false>); break;
          case 3: launch_tma(k_remap_tma<1, true>); break;
          case 4: launch_tma(k_remap_tma<2, false>); break;
          case 5: launch_tma(k_remap_tma<2, true>); break;
          case 6: launch_tma(k_remap_tma<3, false>); break;
          case 7: launch_tma(k_remap_tma<3, true>); break;
          case 8: launch_tma(k_remap_tma<4, false>); break;
          default: launch_tma(k_remap_tma<4, true>); break;
        }
        ++ctx->launches; CK(cudaGetLastError());
      }
      for (size_t first = n_tma; first < n_remap && !k5_mix; first += 65535) {
        const unsigned ny = (unsigned)std::min<size_t>(65535, n_remap - first);
        auto launch_tile = [&](auto kern) {
          ProfScope ps__(ctx, "k_remap_tile");
          kern<<<dim3(tiles, ny), TILE_ROWS, 0, ctx->stream>>>(a->dev, S, B, O, bb.tile_sites + first, tabs);
        };
        switch (key) {
          case 2: launch_tile(k_remap_tile<1, false>); break;
          case 3: launch_tile(k_remap_tile<1, true>); break;
          case 4: launch_tile(k_remap_tile<2, false>); break;
          case 5: launch_tile(k_remap_tile<2, true>); break;
          case 6: launch_tile(k_remap_tile<3, false>); break;
          case 7: launch_tile(k_remap_tile<3, true>); break;
          case 8: launch_tile(k_remap_tile<4, false>); break;
          default: launch_tile(k_remap_tile<4, true>); break;
        }
        ++ctx->launches; CK(cudaGetLastError());
      }
      for (size_t first = n_remap; first < order.size() && !k5_mix; first += 65535) {
        const unsigned ny = (unsigned)std::min<size_t>(65535, order.size() - first);
        auto launch_copy = [&](auto kern) {
          ProfScope ps__(ctx, "k_copy_tile");
          kern<<<dim3(tiles, ny), TILE_ROWS, 0, ctx->stream>>>(a->dev, S, B, O, bb.tile_sites + first);
        };
        switch (key) {
          case 2: launch_copy(k_copy_tile<1, false>); break;
          case 3: launch_copy(k_copy_tile<1, true>); break;
          case 4: launch_copy(k_copy_tile<2, false>); break;
          case 5: launch_copy(k_copy_tile<2, true>); break;
          case 6: launch_copy(k_copy_tile<3, false>); break;
          case 7: launch_copy(k_copy_tile<3, true>); break;
          case 8: launch_copy(k_copy_tile<4, false>); break;
          default: launch_copy(k_copy_tile<4, true>); break;
        }
        ++ctx->launches; CK(cudaGetLastError());
      }
      const uint32_t n_blk_all = nt_prefix.back();
      const uint32_t n_blk = n_blk_all;
      if (n_blk) {
        bb.blk_map.ensure(n_blk);
        LAUNCH(ctx, k_nt_blocks, (unsigned)fast.size(), 256, B, bb.fast_sites, bb.nt_prefix_d, bb.blk_map.p);
      }
      if (k5_mix) {
        // remapped sites' tile blocks and every site's nt blocks, one grid (gdb_kernels6.cuh)
        const uint32_t n_bands = (uint32_t)((fast.size() + MIX_SPB - 1) / MIX_SPB);
        uint64_t n_tb = (uint64_t)n_bands * tiles;
        uint32_t n_blk = n_blk_all;
        if (const char* dbg = getenv("GDBGPU_MIX_ONLY")) { if (!strcmp(dbg, "tile")) n_blk = 0; else if (!strcmp(dbg, "nt")) n_tb = 0; }   // timing experiments only: wrong results
        if (n_tb + n_blk >= (1ull << 31)) throw LimitError("more than 2^31 blocks in one batch: lower max_sites");
        if (n_tb + n_blk) {
          // ceil(2^32 N / G), below 2^32 (a grid of nt blocks only is flagged by tile_groups = 0 and f(b) = b)
          const uint64_t G = n_tb + n_blk;
          const uint32_t nt_ratio = n_tb ? (uint32_t)((((uint64_t)n_blk << 32) + G - 1) / G) : 0xFFFFFFFFu;
          auto launch_mix = [&](auto kern, size_t smem_bytes) {
            if (!(ctx->mix_attr_set & (1u << key))) {   // live entries + hot-record ring + tile buffers exceed the default 48 KB per block
              CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
              ctx->mix_attr_set |= 1u << key;
            }
            ProfScope ps__(ctx, "k_remap_mix");
            kern<<<(unsigned)(n_tb + n_blk), TILE_ROWS, smem_bytes, ctx->stream>>>(a->dev, S, B, O, fast.size() == keep ? nullptr : bb.fast_sites, (uint32_t)fast.size(), tiles, n_tb ? 1u : 0u, bb.blk_map.p, n_blk, nt_ratio, tabs, (w.sig_words && !S.want_lut) ? w.sig_off_d + woff : nullptr, w.sig_perm.p);
          };
          switch (key) {
            case 2: launch_mix(k_remap_mix<1, false>, MixSmem<1>::BYTES); break;
            case 3: launch_mix(k_remap_mix<1, true>, MixSmem<1>::BYTES); break;
            case 4: launch_mix(k_remap_mix<2, false>, MixSmem<2>::BYTES); break;
            case 5: launch_mix(k_remap_mix<2, true>, MixSmem<2>::BYTES); break;
            case 6: launch_mix(k_remap_mix<3, false>, MixSmem<3>::BYTES); break;
            case 7: launch_mix(k_remap_mix<3, true>, MixSmem<3>::BYTES); break;
            case 8: launch_mix(k_remap_mix<4, false>, MixSmem<4>::BYTES); break;
            default: launch_mix(k_remap_mix<4, true>, MixSmem<4>::BYTES); break;
          }
          ++ctx->launches; CK(cudaGetLastError());
        }
      } else if (n_blk) {
        auto launch_nt = [&](auto kern) {
          ProfScope ps__(ctx, "k_remap_nt");
          kern<<<n_blk, NT_THREADS, 0, ctx->stream>>>(a->dev, S, B, O, bb.blk_map.p, jk_n, tabs);
        };
        switch (key) {
          case 2: launch_nt(k_remap_nt<1, false>); break;
          case 3: launch_nt(k_remap_nt<1, true>); break;
          case 4: launch_nt(k_remap_nt<2, false>); break;
          case 5: launch_nt(k_remap_nt<2, true>); break;
          case 6: launch_nt(k_remap_nt<3, false>); break;
          case 7: launch_nt(k_remap_nt<3, true>); break;
          case 8: launch_nt(k_remap_nt<4, false>); break;
          default: launch_nt(k_remap_nt<4, true>); break;
        }
        ++ctx->launches; CK(cudaGetLastError());
      }
    }
    if (!slow.empty() && !s->summary_only) {
      uint32_t tiles = (uint32_t)((S.nq + REMAP_ROWS - 1) / REMAP_ROWS);
      LAUNCH(ctx, k_remap_generic, (unsigned)(slow.size() * tiles), REMAP_THREADS, a->dev, S, B, O, bb.slow_sites);
    }
    // ---- K6: FORMAT collation + BCF typed encoding (remapped fields from K5's output, plain fields through the live matrix)
    if (s->bcf_format) {
      const int n_plain = (int)a->plain.size(), n_fmt = n_out + n_plain;
      const uint64_t n_plan = (uint64_t)n_fmt * keep, n_rp = (uint64_t)n_out * keep, n_pp = (uint64_t)n_plain * keep;
      bb.fmt_mm.ensure(2 * n_rp + PS_WORDS * n_pp + 4);
      bb.h_fmt_mm.resize(2 * n_rp + PS_WORDS * n_pp); bb.h_fmt_type.assign(n_plan, 0); bb.h_fmt_nps.assign(n_plan, 0); bb.h_fmt_off.assign(n_plan, 0);
      for (uint64_t i = 0; i < n_rp; ++i) { bb.h_fmt_mm[2 * i] = INT32_MAX; bb.h_fmt_mm[2 * i + 1] = INT32_MIN + 1; }
      int32_t* h_ps = bb.h_fmt_mm.data() + 2 * n_rp;
      for (uint64_t i = 0; i < n_pp; ++i) { h_ps[PS_WORDS * i + PS_MIN] = INT32_MAX; h_ps[PS_WORDS * i + PS_MAX] = INT32_MIN + 1; h_ps[PS_WORDS * i + PS_MAXLEN] = 0; h_ps[PS_WORDS * i + PS_ANY] = 0; }
      CK(cudaMemcpyAsync(bb.fmt_mm.p, bb.h_fmt_mm.data(), bb.h_fmt_mm.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
      uint64_t max_words = 1;
      for (int o = 0; o < n_out; ++o) for (uint32_t i = 0; i < keep; ++i) max_words = std::max<uint64_t>(max_words, (uint64_t)bb.h_stride[o][i] * S.nq);
      const unsigned chunks = (unsigned)std::min<uint64_t>(64, std::max<uint64_t>(1, max_words / (FMT_THREADS * 16)));
      const unsigned row_blocks = blocks_for(S.nq, FMT_THREADS);
      LAUNCH(ctx, k_fmt_range, dim3(chunks, keep), FMT_THREADS, a->dev, O, keep, (uint32_t)S.nq, s->produce_gt, s->java, bb.fmt_mm.p);
      if (n_plain) LAUNCH(ctx, k_plain_scan, dim3(row_blocks, keep), FMT_THREADS, a->plain_dev, bb.live, keep, (uint32_t)S.nq, bb.fmt_mm.p + 2 * n_rp);
      const bool dp_rule = s->info.dp_info >= 0 || s->info.dp_format >= 0;
      bb.h_info_value.assign((size_t)n_plain * keep, 0); bb.h_info_valid.assign((size_t)n_plain * keep, 0); bb.h_dp_sum.assign(keep, 0);
      const int n_info = s->info.n;
      if (n_info || dp_rule) {   // INFO scalars: ordered compaction of the valid values, then median / sum / mean; DP sum
        bb.info_vals.ensure(std::max<uint64_t>((uint64_t)n_info * keep * S.nq, 1)); bb.info_cnt.ensure(std::max<uint64_t>((uint64_t)n_info * keep, 1));
        bb.info_value_d.ensure(std::max<uint64_t>((uint64_t)n_info * keep, 1)); bb.info_valid_d.ensure(std::max<uint64_t>((uint64_t)n_info * keep, 1)); bb.dp_sum_d.ensure(keep);
        LAUNCH(ctx, k_info_collect, keep, FMT_THREADS, a->plain_dev, s->info, bb.live, keep, (uint32_t)S.nq, bb.info_vals.p, bb.info_cnt.p, bb.dp_sum_d.p);
        if (n_info) LAUNCH(ctx, k_info_reduce, dim3(keep, n_info), FMT_THREADS, a->plain_dev, s->info, keep, (uint32_t)S.nq, bb.info_vals.p, bb.info_cnt.p, bb.info_value_d.p, bb.info_valid_d.p);
        // results land at the plain-field index of every INFO field
        for (int k = 0; k < n_info; ++k) {
          const size_t dst = (size_t)s->info.field[k] * keep;
          CK(cudaMemcpyAsync(bb.h_info_value.data() + dst, bb.info_value_d.p + (size_t)k * keep, (size_t)keep * 4, cudaMemcpyDeviceToHost, ctx->stream));
          CK(cudaMemcpyAsync(bb.h_info_valid.data() + dst, bb.info_valid_d.p + (size_t)k * keep, keep, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaMemcpyAsync(bb.h_dp_sum.data(), bb.dp_sum_d.p, (size_t)keep * 4, cudaMemcpyDeviceToHost, ctx->stream));
      }
      const int n_vec_slots = (int)a->fields.size() + n_plain, n_esum = s->esum.n;
      bb.h_vec_len.assign((size_t)n_vec_slots * keep, 0); bb.h_vec_off.assign((size_t)n_vec_slots * keep, 0); bb.h_vec_values.clear();
      std::vector<uint32_t> h_esum_len;
      if (n_esum) {   // element-wise sums: fixed-capacity slots on the device, compacted on the host below
        bb.esum_values.ensure((uint64_t)n_esum * keep * ESUM_CAP); bb.esum_lens.ensure((uint64_t)n_esum * keep);
        LAUNCH(ctx, k_info_esum, dim3(keep, n_esum), ESUM_THREADS, a->dev, a->plain_dev, O, s->esum, bb.live, reinterpret_cast<const int32_t*>(bb.fmt_mm.p + 2 * n_rp), keep,
               (uint32_t)S.nq, bb.esum_values.p, bb.esum_lens.p, s->err.p);
        h_esum_len.resize((size_t)n_esum * keep); bb.h_esum_raw.resize((size_t)n_esum * keep * ESUM_CAP);
        CK(cudaMemcpyAsync(h_esum_len.data(), bb.esum_lens.p, h_esum_len.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(bb.h_esum_raw.data(), bb.esum_values.p, bb.h_esum_raw.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      }
      CK(cudaMemcpyAsync(bb.h_fmt_mm.data(), bb.fmt_mm.p, bb.h_fmt_mm.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int v = 0; v < n_esum; ++v) {
        const int slot = s->esum.remapped[v] ? s->esum.index[v] : (int)a->fields.size() + s->esum.index[v];
        for (uint32_t i = 0; i < keep; ++i) {
          const uint32_t n = h_esum_len[(size_t)v * keep + i];
          bb.h_vec_len[(size_t)slot * keep + i] = n; bb.h_vec_off[(size_t)slot * keep + i] = bb.h_vec_values.size();
          const uint32_t* src = bb.h_esum_raw.data() + ((size_t)v * keep + i) * ESUM_CAP;
          bb.h_vec_values.insert(bb.h_vec_values.end(), src, src + n);
        }
      }
      // bcf_enc_vint's width rule (vcf.c:1926-1972); a block of ONE value goes through bcf_enc_int1 instead (vcf.c:1931)
      auto int_type = [&](int64_t lo, int64_t hi, uint64_t n_values) -> uint8_t {
        if (n_values == 1) {
          if (lo > hi) return BT_INT8;          // the value is missing / vector_end
          return (lo >= -120 && hi <= 127) ? BT_INT8 : ((lo >= -32760 && hi <= 32767) ? BT_INT16 : BT_INT32);
        }
        return (hi <= 127 && lo > -127) ? BT_INT8 : ((hi <= 32767 && lo > -32767) ? BT_INT16 : BT_INT32);
      };
      uint64_t off = 0;
      for (uint32_t i = 0; i < keep; ++i) for (int o = 0; o < n_fmt; ++o) {   // the blocks of one site are neighbours
        const uint64_t pi = (uint64_t)o * keep + i;
        uint32_t nps; uint8_t type;
        if (o < n_out) {
          const bool is_gt = a->has_gt && o == 0;
          const uint32_t need = stride_needed(a, bb.h_sites[i], o);
          if (need == 0) continue;
          nps = (is_gt && a->gt_phased) ? (need + 1u) / 2u : need;
          if (!is_gt && a->fm[o - (a->has_gt ? 1 : 0)].elem) type = BT_FLOAT;
          else type = int_type(bb.h_fmt_mm[2 * pi], bb.h_fmt_mm[2 * pi + 1], (uint64_t)nps * S.nq);
        } else {
          const int f = o - n_out;
          const int32_t* st = h_ps + ((uint64_t)f * keep + i) * PS_WORDS;
          if (!st[PS_ANY] || a->pm[f].no_format == 1) continue;
          nps = (uint32_t)st[PS_MAXLEN];
          if (nps == 0) continue;
          type = a->pm[f].elem == ELEM_CHAR ? BT_CHAR : (a->pm[f].elem == ELEM_FLOAT32 ? BT_FLOAT : int_type(st[PS_MIN], st[PS_MAX], (uint64_t)nps * S.nq));
        }
        const uint64_t size = (uint64_t)nps * S.nq * ((type == BT_INT8 || type == BT_CHAR) ? 1u : (type == BT_INT16 ? 2u : 4u));
        bb.h_fmt_type[pi] = type; bb.h_fmt_nps[pi] = nps; bb.h_fmt_off[pi] = off;
        off += (size + 15u) & ~15ull;
      }
      bb.fmt_total = off;
      bb.fmt_bytes.ensure(std::max<uint64_t>(off, 16));
      bb.fmt_type_d = bb.ctl.push(bb.h_fmt_type.data(), n_plan); bb.fmt_nps_d = bb.ctl.push(bb.h_fmt_nps.data(), n_plan); bb.fmt_off_d = bb.ctl.push(bb.h_fmt_off.data(), n_plan);
      bb.ctl.flush(ctx->stream);
      const FmtPlanDev P{bb.fmt_type_d, bb.fmt_nps_d, bb.fmt_off_d, bb.fmt_bytes.p};
      LAUNCH(ctx, k_fmt_pack, dim3(chunks, keep), FMT_THREADS, a->dev, O, P, keep, (uint32_t)S.nq, s->produce_gt, s->java);
      LAUNCH(ctx, k_fmt_pack_tile, dim3(blocks_for(S.nq, PACK_ROWS), keep), FMT_THREADS, a->dev, O, P, keep, (uint32_t)S.nq, s->java);
      if (n_plain) LAUNCH(ctx, k_plain_pack, dim3(row_blocks, keep), FMT_THREADS, a->plain_dev, P, n_out, bb.live, keep, (uint32_t)S.nq, s->java);
    }
    // ---- results: site metadata (host mirrors), then the copies of everything the caller asked for
    bb.keep = keep; bb.alg_bytes = alg_bytes;
    bb.p_tail.ensure(8);
    if (S.want_lut) {   // the LUT copy is sized by a cursor only the remap kernels advance
      CK(cudaMemcpyAsync(bb.hc + 2, bb.cursors.p + 2, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMemcpyAsync(bb.p_tail.p, s->err.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(bb.p_tail.p + 4, bb.cursors.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(bb.ev_compute, ctx->stream));
    if (!s->device_output) {
      cudaStream_t cs = ctx->copy_stream;
      CK(cudaStreamWaitEvent(cs, bb.ev_compute, 0));
      bb.p_site_begin.ensure(keep); bb.p_site_end.ensure(keep);
      bb.p_allele_pos.ensure(2ull * std::max<uint32_t>(bb.hc[0], 1)); bb.p_allele_bytes.ensure(std::max<uint32_t>(bb.hc[1], 1));
      CK(cudaMemcpyAsync(bb.p_site_begin.p, s->site_begin.p + first_site, (size_t)keep * 8, cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(bb.p_site_end.p, s->site_end.p + first_site, (size_t)keep * 8, cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(bb.p_allele_pos.p, bb.allele_pos, (size_t)bb.hc[0] * 8, cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(bb.p_allele_bytes.p, bb.allele_bytes, bb.hc[1], cudaMemcpyDeviceToHost, cs));
      if (s->summary_only) {
        // per-site tables only: nothing per call crosses PCIe
      } else if (s->bcf_format) {
        bb.p_fmt_bytes.ensure(std::max<uint64_t>(bb.fmt_total, 16));
        if (bb.fmt_total) CK(cudaMemcpyAsync(bb.p_fmt_bytes.p, bb.fmt_bytes.p, bb.fmt_total, cudaMemcpyDeviceToHost, cs));
      } else {
        bb.p_live.ensure((uint64_t)keep * S.nq);
        CK(cudaMemcpyAsync(bb.p_live.p, bb.live, (size_t)keep * S.nq * 4, cudaMemcpyDeviceToHost, cs));
        for (int o = 0; o < n_out; ++o) {
          bb.p_out_data[o].ensure(std::max<uint64_t>(bb.h_base[o][keep], 1)); bb.p_out_len[o].ensure((uint64_t)keep * S.nq);
          if (bb.h_base[o][keep]) CK(cudaMemcpyAsync(bb.p_out_data[o].p, bb.out_data[o].p, bb.h_base[o][keep] * 4, cudaMemcpyDeviceToHost, cs));
          CK(cudaMemcpyAsync(bb.p_out_len[o].p, bb.out_len[o].p, (size_t)keep * S.nq * 2, cudaMemcpyDeviceToHost, cs));
        }
        if (S.want_lut) {
          bb.p_lut_first.ensure((uint64_t)keep * S.nq); bb.p_lut_len.ensure((uint64_t)keep * S.nq); bb.p_lut_in2m.ensure(std::max<uint32_t>(bb.hc[2], 1));
          CK(cudaMemcpyAsync(bb.p_lut_first.p, bb.lut_first.p, (size_t)keep * S.nq * 4, cudaMemcpyDeviceToHost, cs));
          CK(cudaMemcpyAsync(bb.p_lut_len.p, bb.lut_len.p, (size_t)keep * S.nq * 2, cudaMemcpyDeviceToHost, cs));
          CK(cudaMemcpyAsync(bb.p_lut_in2m.p, bb.lut_in2m.p, (size_t)bb.hc[2] * 2, cudaMemcpyDeviceToHost, cs));
        }
      }
      CK(cudaEventRecord(bb.ev_done, cs));
    } else CK(cudaEventRecord(bb.ev_done, ctx->stream));
    bb.valid = true;
    return;
  }
}

// Waits for the batch in `bb` and fills the caller's view of it.
void publish_batch(gdbgpu_scan* s, BatchBuffers& bb, gdbgpu_batch* out) {
  gdbgpu_array* a = s->arr; ScanDev& S = s->S;
  CK(cudaEventSynchronize(bb.ev_done));
  const uint32_t* e = bb.p_tail.p;
  if (e[0]) throw OverlapError("Unhandled overlapping variants (a call that is neither a deletion nor a reference block is overlapped by a later call of the same sample)");
  if (e[1] == 3) throw LimitError("ploidy above 8 needs the general-ploidy genotype enumeration, which keeps at most 8 slots");
  if (e[1] == 5) throw LimitError("element-wise INFO sum over a vector longer than 1024 values");
  if (e[1] == 6) throw LimitError("two different ALT alleles of one site share a 64-bit hash (detected by the byte comparison; never merged silently)");
  if (e[1]) throw LimitError("engine limit exceeded in the remap kernels");
  if (e[2]) throw CudaError("internal: LUT pool overflow");
  const uint32_t keep = bb.keep; const int n_out = s->n_out;
  out->n_sites = keep; out->out_bytes = bb.alg_bytes;
  out->site_flags = bb.h_site_flags.data(); out->n_alleles = bb.h_n_alleles.data(); out->n_live = bb.h_n_live.data(); out->allele_first = bb.h_allele_first.data();
  if (s->device_output) {
    out->site_begin = s->site_begin.p + bb.first; out->site_end = s->site_end.p + bb.first;
    out->site_flags = bb.d_site_flags; out->n_alleles = bb.d_n_alleles; out->n_live = bb.d_n_live; out->allele_first = bb.d_allele_first;
    out->allele_pos = bb.allele_pos; out->allele_bytes = bb.allele_bytes; out->live = bb.live;
    for (int o = 0; o < n_out; ++o) out->fields[o] = gdbgpu_field_out{bb.out_data[o].p, bb.out_base[o], bb.out_stride[o], bb.out_len[o].p};
    if (S.want_lut) { out->lut_first = bb.lut_first.p; out->lut_len = bb.lut_len.p; out->lut_in2m = bb.lut_in2m.p; }
    if (s->bcf_format) { out->fmt_bytes = bb.fmt_bytes.p; out->fmt_type = bb.fmt_type_d; out->fmt_nps = bb.fmt_nps_d; out->fmt_offset = bb.fmt_off_d; }
  } else {
    out->site_begin = bb.p_site_begin.p; out->site_end = bb.p_site_end.p;
    out->allele_pos = bb.p_allele_pos.p; out->allele_bytes = bb.p_allele_bytes.p;
    if (s->summary_only) {
      out->live = nullptr;
      for (int o = 0; o < n_out; ++o) out->fields[o] = gdbgpu_field_out{nullptr, bb.h_base[o].data(), bb.h_stride[o].data(), nullptr};
    } else if (s->bcf_format) {
      out->fmt_bytes = bb.p_fmt_bytes.p; out->fmt_type = bb.h_fmt_type.data(); out->fmt_nps = bb.h_fmt_nps.data(); out->fmt_offset = bb.h_fmt_off.data();
      for (int o = 0; o < n_out; ++o) out->fields[o] = gdbgpu_field_out{nullptr, bb.h_base[o].data(), bb.h_stride[o].data(), nullptr};
    } else {
      out->live = bb.p_live.p;
      for (int o = 0; o < n_out; ++o) out->fields[o] = gdbgpu_field_out{bb.p_out_data[o].p, bb.h_base[o].data(), bb.h_stride[o].data(), bb.p_out_len[o].p};
      if (S.want_lut) { out->lut_first = bb.p_lut_first.p; out->lut_len = bb.p_lut_len.p; out->lut_in2m = bb.p_lut_in2m.p; }
    }
  }
  out->site_stats = s->site_stats ? bb.p_site_stats.p : nullptr;
  out->fmt_total_bytes = bb.fmt_total;
  out->n_fmt_fields = s->bcf_format ? n_out + (int)a->plain.size() : 0;
  if (s->bcf_format) {
    out->info_valid = bb.h_info_valid.data(); out->info_value = bb.h_info_value.data(); out->info_dp_sum = bb.h_dp_sum.data();
    out->vec_len = bb.h_vec_len.data(); out->vec_offset = bb.h_vec_off.data(); out->vec_values = bb.h_vec_values.data();
  }
  (void)a;
}

BatchBuffers& scan_slot(gdbgpu_scan* s, int i) {
  if (!s->slot[i]) {
    gdbgpu_ctx* ctx = s->ctx;
    if (!ctx->bb_pool.empty()) { s->slot[i].reset(ctx->bb_pool.back()); ctx->bb_pool.pop_back(); } else s->slot[i].reset(new BatchBuffers());
    s->slot[i]->valid = false;
  }
  return *s->slot[i];
}

// Hands out the batch at the cursor and, unless told otherwise, starts the next one before waiting for this one's copies:
// the device -> host transfer of batch n runs while the kernels of batch n+1 do.
int next_batch(gdbgpu_scan* s, uint32_t max_sites, uint64_t max_bytes, gdbgpu_batch* out) {
  ScanDev& S = s->S;
  memset(out, 0, sizeof(*out));
  out->first_site = s->cursor; out->n_query_rows = S.nq; out->n_fields = s->n_out;
  if (s->cursor >= S.n_sites) return GDBGPU_OK;
  auto holds = [&](const BatchBuffers& b, uint64_t first) { return b.valid && b.first == first && b.arg_sites == max_sites && b.arg_bytes == max_bytes; };
  BatchBuffers& c = scan_slot(s, s->cur);
  if (!holds(c, s->cursor)) compute_batch(s, c, s->cursor, max_sites, max_bytes);
  const uint64_t next_first = c.first + c.keep;
  // (device output: the caller still reads this batch's rows of the window's live matrix and allele pools, so the next window waits for the next call)
  if (s->prefetch && next_first < S.n_sites && !(s->device_output && next_first >= s->front.first + s->front.n)) {
    BatchBuffers& n = scan_slot(s, s->cur ^ 1);
    if (!holds(n, next_first)) compute_batch(s, n, next_first, max_sites, max_bytes);
  }
  publish_batch(s, c, out);
  c.valid = false;
  s->cursor = next_first; s->cur ^= 1;
  return GDBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ K7: columnar iterator
// genomicsdb_iterators.cc:181-300,425-510 as a list: [0, n_left) the calls reaching into the interval from the left,
// ordered by (begin, row); then the interval's begin cells of the queried rows in storage order.
void cells_build(gdbgpu_scan* s) {
  if (s->cells_built) return;
  gdbgpu_array* a = s->arr; gdbgpu_ctx* ctx = s->ctx; ScanDev S = s->S;
  s->n_cells = s->n_left = 0; s->cells_built = true;
  if (a->n == 0 || S.nq == 0 || S.q_lo > S.q_hi) return;
  S.i0 = S.iL;   // the traversal starts at the first cell with column >= q_lo
  DevBuf<uint32_t> flag, pos, total, sums, left_cell, left_sorted; DevBuf<uint64_t> key, left_key, left_key_sorted; DevBuf<uint8_t> tmp;
  total.alloc(1);
  uint32_t n_left = 0, n_fwd = 0;
  if (!S.scan_full) {
    flag.alloc(S.nq); pos.alloc(S.nq); key.alloc(S.nq);
    LAUNCH(ctx, k_cells_left_flag, blocks_for(S.nq, 256), 256, a->dev, S, flag.p, key.p);
    exclusive_scan<uint32_t, uint32_t>(ctx, flag.p, pos.p, S.nq, total.p, sums);
    CK(cudaMemcpyAsync(&n_left, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    if (n_left) {
      left_cell.alloc(n_left); left_sorted.alloc(n_left); left_key.alloc(n_left); left_key_sorted.alloc(n_left);
      LAUNCH(ctx, k_cells_left_compact, blocks_for(S.nq, 256), 256, S, flag.p, pos.p, key.p, left_cell.p, left_key.p);
      radix_sort_pairs(ctx, left_key.p, left_key_sorted.p, left_cell.p, left_sorted.p, n_left, 64, tmp);   // stable: rows stay ascending within a column
    }
  }
  const uint64_t span = S.i1 - S.i0;
  if (span) {
    flag.alloc(span); pos.alloc(span);
    LAUNCH(ctx, k_cells_fwd_flag, blocks_for(span, 256), 256, a->dev, S, flag.p);
    exclusive_scan<uint32_t, uint32_t>(ctx, flag.p, pos.p, span, total.p, sums);
    CK(cudaMemcpyAsync(&n_fwd, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  }
  s->n_left = n_left; s->n_cells = (uint64_t)n_left + n_fwd;
  s->cell_list.alloc(std::max<uint64_t>(s->n_cells, 1));
  if (n_left) CK(cudaMemcpyAsync(s->cell_list.p, left_sorted.p, (size_t)n_left * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  if (n_fwd) LAUNCH(ctx, k_cells_fwd_compact, blocks_for(span, 256), 256, S, flag.p, pos.p, s->cell_list.p + n_left);
  CK(cudaStreamSynchronize(ctx->stream));
}

void cells_to_host(gdbgpu_scan* s, gdbgpu_cell_list* out) {
  cells_build(s);
  gdbgpu_ctx* ctx = s->ctx;
  const uint64_t first = std::min(s->win_first, s->n_cells), n = std::min(s->win_count, s->n_cells - first);
  s->h_cell.resize(n); s->h_row.resize(n); s->h_begin.resize(n); s->h_end.resize(n);
  if (n) {
    DevBuf<uint64_t> cell; DevBuf<int64_t> row, begin, end;
    cell.alloc(n); row.alloc(n); begin.alloc(n); end.alloc(n);
    LAUNCH(ctx, k_cells_fill, blocks_for(n, 256), 256, s->arr->dev, s->cell_list.p + first, n, cell.p, row.p, begin.p, end.p);
    CK(cudaMemcpyAsync(s->h_cell.data(), cell.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(s->h_row.data(), row.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(s->h_begin.data(), begin.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(s->h_end.data(), end.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  out->n_cells = n; out->n_left = first < s->n_left ? std::min(s->n_left - first, n) : 0;
  out->n_total = s->n_cells; out->first = first;
  out->cell = s->h_cell.data(); out->row = s->h_row.data(); out->begin = s->h_begin.data(); out->end = s->h_end.data();
}

void cells_histogram(gdbgpu_scan* s, uint64_t h_begin, uint64_t h_end, uint64_t bin_size, uint64_t* counts, uint64_t n_bins) {
  if (!bin_size || h_end < h_begin || n_bins != (h_end - h_begin) / bin_size + 1) throw ArgError("histogram: n_bins must be (end - begin) / bin_size + 1");
  cells_build(s);
  gdbgpu_ctx* ctx = s->ctx;
  memset(counts, 0, n_bins * 8);
  if (!s->n_cells) return;
  // only the bins the query touches are counted and copied (the reference's tool spans 4 x 10^9 columns with 40 M bins)
  DevBuf<unsigned long long> range, d; range.alloc(2);
  unsigned long long h_range[2] = {~0ull, 0ull};
  CK(cudaMemcpyAsync(range.p, h_range, 16, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_cells_bin_range, blocks_for(s->n_cells, 256), 256, s->arr->dev, s->cell_list.p, s->n_cells, h_begin, h_end, bin_size, n_bins, range.p);
  CK(cudaMemcpyAsync(h_range, range.p, 16, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  const uint64_t first = h_range[0], span = h_range[1] - h_range[0] + 1;
  d.alloc(span);
  CK(cudaMemsetAsync(d.p, 0, span * 8, ctx->stream));
  LAUNCH(ctx, k_cells_hist, blocks_for(s->n_cells, 256), 256, s->arr->dev, s->cell_list.p, s->n_cells, h_begin, h_end, bin_size, n_bins, first, d.p);
  CK(cudaMemcpyAsync(counts + first, d.p, span * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
}

// bytes of one column for a device list of cells -> host (offsets[n + 1] + values)
void fetch_column(gdbgpu_scan* s, const DevVar& col, int32_t fixed_bytes, const uint32_t* list, uint64_t n, std::vector<uint64_t>& off, std::vector<uint8_t>& val) {
  gdbgpu_ctx* ctx = s->ctx;
  off.assign(n + 1, 0); val.clear();
  if (!n) return;
  if (fixed_bytes) {
    DevBuf<uint8_t> g; g.alloc(n * (uint64_t)fixed_bytes);
    LAUNCH(ctx, k_gather_bytes, blocks_for(n, 256), 256, col.val.p, list, n, (uint32_t)fixed_bytes, g.p);
    val.resize(n * (uint64_t)fixed_bytes);
    CK(cudaMemcpyAsync(val.data(), g.p, val.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i <= n; ++i) off[i] = i * (uint64_t)fixed_bytes;
    return;
  }
  DevVar g;
  gather_var(ctx, col, list, n, g);
  val.resize(g.bytes);
  CK(cudaMemcpyAsync(off.data(), g.off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (g.bytes) CK(cudaMemcpyAsync(val.data(), g.val.p, g.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
}

void cells_fetch(gdbgpu_scan* s, int32_t kind, int32_t index, const uint64_t** offsets, const uint8_t** bytes) {
  cells_build(s);
  gdbgpu_array* a = s->arr;
  const DevVar* col = nullptr; int32_t fixed_bytes = 0;
  switch (kind) {
    case GDBGPU_CELL_FIELD_REF: col = &a->ref; break;
    case GDBGPU_CELL_FIELD_ALT: col = &a->alt; break;
    case GDBGPU_CELL_FIELD_GT: if (!a->has_gt) throw ArgError("the array has no GT column"); col = &a->gt; break;
    case GDBGPU_CELL_FIELD_REMAPPED: if (index < 0 || (size_t)index >= a->fields.size()) throw ArgError("no such allele-dependent field"); col = a->fields[index].get(); break;
    case GDBGPU_CELL_FIELD_PLAIN:
      if (index < 0 || (size_t)index >= a->plain.size()) throw ArgError("no such plain field");
      col = a->plain[index].get();
      fixed_bytes = a->pm[index].fixed_len * (a->pm[index].elem == GDBGPU_ELEM_CHAR ? 1 : 4);
      break;
    default: throw ArgError("unknown cell field kind");
  }
  auto& slot = s->fetched[std::make_pair((int)kind, (int)index)];
  const uint64_t first = std::min(s->win_first, s->n_cells), n = std::min(s->win_count, s->n_cells - first);
  fetch_column(s, *col, fixed_bytes, s->cell_list.p + first, n, slot.first, slot.second);
  *offsets = slot.first.data(); *bytes = slot.second.data();
}

// AlleleCountOperator (variant_operations.cc:900-1089): counted on the device per (column, normalised REF, normalised ALT)
// hash; only one representative cell per distinct key crosses to the host, where its strings are cut and the records sorted
// the way the reference's nested std::map prints them.
void cells_allele_counts(gdbgpu_scan* s, gdbgpu_allele_counts* out) {
  cells_build(s);
  gdbgpu_array* a = s->arr; gdbgpu_ctx* ctx = s->ctx; const uint64_t n = s->n_cells;
  s->ac_column.clear(); s->ac_count.clear(); s->ac_ref_pos.clear(); s->ac_alt_pos.clear(); s->ac_bytes.clear();
  uint32_t n_ent = 0, n_keys = 0;
  DevBuf<uint32_t> cnt, pos, total, sums, ent_cell, ent_alt, iota, ent_sorted, flag, head_ent, head_start, rep_cell, rep_alt, rep_list;
  DevBuf<uint64_t> key, key_sorted; DevBuf<uint8_t> tmp;
  const uint32_t gt_step = a->gt_phased ? 2u : 1u;
  if (n && a->has_gt) {
    cnt.alloc(n); pos.alloc(n); total.alloc(1);
    LAUNCH(ctx, k_ac_count, blocks_for(n, 256), 256, a->dev, s->cell_list.p, n, gt_step, cnt.p);
    exclusive_scan<uint32_t, uint32_t>(ctx, cnt.p, pos.p, n, total.p, sums);
    CK(cudaMemcpyAsync(&n_ent, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
  }
  std::vector<uint32_t> h_start, h_cell, h_alt; std::vector<int64_t> h_column; std::vector<uint64_t> r_off, a_off; std::vector<uint8_t> r_val, a_val;
  if (n_ent) {
    key.alloc(n_ent); key_sorted.alloc(n_ent); ent_cell.alloc(n_ent); ent_alt.alloc(n_ent); iota.alloc(n_ent); ent_sorted.alloc(n_ent); flag.alloc(n_ent);
    LAUNCH(ctx, k_ac_fill, blocks_for(n, 256), 256, a->dev, s->cell_list.p, n, gt_step, pos.p, key.p, ent_cell.p, ent_alt.p);
    LAUNCH(ctx, k_iota, blocks_for(n_ent, 256), 256, iota.p, (uint64_t)n_ent);
    radix_sort_pairs(ctx, key.p, key_sorted.p, iota.p, ent_sorted.p, n_ent, 64, tmp);
    LAUNCH(ctx, k_ac_heads, blocks_for(n_ent, 256), 256, key_sorted.p, (uint64_t)n_ent, flag.p);
    pos.alloc(n_ent);
    exclusive_scan<uint32_t, uint32_t>(ctx, flag.p, pos.p, n_ent, total.p, sums);
    CK(cudaMemcpyAsync(&n_keys, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    head_ent.alloc(n_keys); head_start.alloc(n_keys); rep_cell.alloc(n_keys); rep_alt.alloc(n_keys); rep_list.alloc(n_keys);
    LAUNCH(ctx, k_ac_compact, blocks_for(n_ent, 256), 256, flag.p, pos.p, ent_sorted.p, (uint64_t)n_ent, head_ent.p, head_start.p);
    LAUNCH(ctx, k_gather_u32, blocks_for(n_keys, 256), 256, ent_cell.p, head_ent.p, (uint64_t)n_keys, rep_cell.p);   // list index of the representative
    LAUNCH(ctx, k_gather_u32, blocks_for(n_keys, 256), 256, ent_alt.p, head_ent.p, (uint64_t)n_keys, rep_alt.p);
    LAUNCH(ctx, k_gather_u32, blocks_for(n_keys, 256), 256, s->cell_list.p, rep_cell.p, (uint64_t)n_keys, rep_list.p);  // its cell
    h_start.resize(n_keys); h_cell.resize(n_keys); h_alt.resize(n_keys);
    CK(cudaMemcpyAsync(h_start.data(), head_start.p, (size_t)n_keys * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_cell.data(), rep_list.p, (size_t)n_keys * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_alt.data(), rep_alt.p, (size_t)n_keys * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    fetch_column(s, a->ref, 0, rep_list.p, n_keys, r_off, r_val);
    fetch_column(s, a->alt, 0, rep_list.p, n_keys, a_off, a_val);
    DevBuf<uint64_t> c64; DevBuf<int64_t> row, begin, end;
    c64.alloc(n_keys); row.alloc(n_keys); begin.alloc(n_keys); end.alloc(n_keys);
    LAUNCH(ctx, k_cells_fill, blocks_for(n_keys, 256), 256, a->dev, rep_list.p, (uint64_t)n_keys, c64.p, row.p, begin.p, end.p);
    h_column.resize(n_keys);
    CK(cudaMemcpyAsync(h_column.data(), begin.p, (size_t)n_keys * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // host: cut the representative strings (normalize_REF_ALT_pair), order as map<column, map<pair<REF, ALT>, count>>
  struct Rec { int64_t column; std::string ref, alt; uint64_t count; };
  std::vector<Rec> recs(n_keys);
  for (uint32_t k = 0; k < n_keys; ++k) {
    Rec& r = recs[k];
    r.count = (k + 1 < n_keys ? h_start[k + 1] : n_ent) - h_start[k];
    r.column = h_column[k];
    std::string ref((const char*)r_val.data() + r_off[k], r_off[k + 1] - r_off[k]);
    const char* ab = (const char*)a_val.data() + a_off[k]; size_t al = a_off[k + 1] - a_off[k];
    size_t t0 = 0, tok = 0;
    while (tok < h_alt[k] && t0 < al) { if (ab[t0] == '|') ++tok; ++t0; }
    size_t t1 = t0; while (t1 < al && ab[t1] != '|') ++t1;
    std::string alt(ab + t0, t1 - t0);
    if (ref.size() > 1 && !alt.empty()) {
      const bool symbolic = (alt.size() == 1 && (alt[0] == '&' || alt[0] == '*')) || (alt.front() == '<' && alt.back() == '>') || alt.find_first_of("[]") != std::string::npos;
      if (symbolic) ref.resize(1);
      else {
        size_t suffix = 0;
        if (alt.size() >= ref.size()) suffix = ref.size() - 1; else if (alt.size() > 1) suffix = alt.size() - 1;
        ref.resize(ref.size() - suffix); alt.resize(alt.size() - suffix);
      }
    }
    r.ref = ref; r.alt = alt;
  }
  std::sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) { return x.column != y.column ? x.column < y.column : (x.ref != y.ref ? x.ref < y.ref : x.alt < y.alt); });
  for (auto& r : recs) {
    s->ac_column.push_back(r.column); s->ac_count.push_back(r.count);
    s->ac_ref_pos.push_back((uint32_t)s->ac_bytes.size()); s->ac_ref_pos.push_back((uint32_t)r.ref.size()); s->ac_bytes.insert(s->ac_bytes.end(), r.ref.begin(), r.ref.end());
    s->ac_alt_pos.push_back((uint32_t)s->ac_bytes.size()); s->ac_alt_pos.push_back((uint32_t)r.alt.size()); s->ac_bytes.insert(s->ac_bytes.end(), r.alt.begin(), r.alt.end());
  }
  out->n = recs.size(); out->column = s->ac_column.data(); out->count = s->ac_count.data();
  out->ref_pos = s->ac_ref_pos.data(); out->alt_pos = s->ac_alt_pos.data(); out->bytes = s->ac_bytes.data();
}

template <class F>
int guarded(gdbgpu_ctx* ctx, F&& f) {
  tl_ctx = ctx;
  try { f(); return GDBGPU_OK; }
  catch (const CudaError& e) { if (ctx) ctx->err = e.what(); return GDBGPU_ERR_CUDA; }
  catch (const LimitError& e) { if (ctx) ctx->err = e.what(); return GDBGPU_ERR_LIMIT; }
  catch (const OverlapError& e) { if (ctx) ctx->err = e.what(); return GDBGPU_ERR_OVERLAP; }
  catch (const ArgError& e) { if (ctx) ctx->err = e.what(); return GDBGPU_ERR_ARG; }
  catch (const std::bad_alloc&) { if (ctx) ctx->err = "host out of memory"; return GDBGPU_ERR_CUDA; }
  catch (const std::exception& e) { if (ctx) ctx->err = e.what(); return GDBGPU_ERR_IO; }
}

}  // namespace

// ================================================================================================ extern "C"
extern "C" {

const char* gdbgpu_version(void) { return "gdbgpu 0.1 (sm_100a)"; }

int gdbgpu_ctx_create(int device, gdbgpu_ctx** out) {
  if (!out) return GDBGPU_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return GDBGPU_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return GDBGPU_ERR_CUDA;
  gdbgpu_ctx* c = new gdbgpu_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return GDBGPU_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return GDBGPU_ERR_CUDA; }
  {   // diploid genotype tables: g -> class of a trivial call's word, g -> (j, k)
    std::vector<uint32_t> cls(256, 0); std::vector<uint16_t> jk(1024, 0);
    uint32_t g = 0;
    for (uint32_t k = 0; g < 1024; ++k) for (uint32_t j = 0; j <= k && g < 1024; ++j, ++g) {
      jk[g] = (uint16_t)((k << 8) | j);
      cls[g >> 2] |= (k == 0 ? 0u : (j == 0 ? 1u : 2u)) << (8 * (g & 3));
    }
    cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaMalloc((void**)&c->d_cls_tab, 1024) != cudaSuccess || cudaMalloc((void**)&c->d_jk_tab, 2048) != cudaSuccess ||
        cudaMemcpy(c->d_cls_tab, cls.data(), 1024, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(c->d_jk_tab, jk.data(), 2048, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaStreamDestroy(c->stream); delete c; return GDBGPU_ERR_CUDA;
    }
  }
  *out = c;
  return GDBGPU_OK;
}
void gdbgpu_ctx_destroy(gdbgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  tl_ctx = ctx;
  ctx->resolve();
  for (auto e : ctx->pool) cudaEventDestroy(e);
  for (auto* b : ctx->bb_pool) delete b;
  cudaStreamSynchronize(ctx->stream);
  tl_ctx = nullptr;
  for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
  ctx->free_blocks.clear();
  if (ctx->stage_pin) cudaFreeHost(ctx->stage_pin);
  for (auto st : ctx->stage_streams) cudaStreamDestroy(st);
  for (auto e : ctx->stage_events) cudaEventDestroy(e);
  if (ctx->d_cls_tab) cudaFree(ctx->d_cls_tab);
  if (ctx->d_jk_tab) cudaFree(ctx->d_jk_tab);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
void gdbgpu_ctx_set_profiling(gdbgpu_ctx* ctx, int enable) { if (ctx) { ctx->resolve(); ctx->profiling = enable != 0; } }
void gdbgpu_ctx_reset_kernel_stats(gdbgpu_ctx* ctx) { if (ctx) { ctx->resolve(); ctx->stats.clear(); } }
int gdbgpu_ctx_kernel_stats(gdbgpu_ctx* ctx, const char* kernel_name, double* total_ms, uint64_t* launches) {
  if (!ctx || !kernel_name) return GDBGPU_ERR_ARG;
  ctx->resolve();
  double ms = 0; uint64_t n = 0;
  for (auto& kv : ctx->stats) if (kv.first.find(kernel_name) != std::string::npos) { ms += kv.second.first; n += kv.second.second; }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return GDBGPU_OK;
}
int gdbgpu_ctx_synchronize(gdbgpu_ctx* ctx) {
  if (!ctx) return GDBGPU_ERR_ARG;
  return (cudaStreamSynchronize(ctx->stream) == cudaSuccess && cudaStreamSynchronize(ctx->copy_stream) == cudaSuccess) ? GDBGPU_OK : GDBGPU_ERR_CUDA;
}
int gdbgpu_ctx_release_cached_memory(gdbgpu_ctx* ctx) {
  if (!ctx) return GDBGPU_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess) return GDBGPU_ERR_CUDA;
  for (auto* b : ctx->bb_pool) { tl_ctx = ctx; delete b; }
  ctx->bb_pool.clear();
  for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
  ctx->free_blocks.clear(); ctx->cached_bytes = 0;
  return GDBGPU_OK;
}
const char* gdbgpu_last_error(const gdbgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }
void* gdbgpu_ctx_stream(gdbgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t gdbgpu_ctx_launch_count(const gdbgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gdbgpu_array_from_columns(gdbgpu_ctx* ctx, const gdbgpu_columns* c, gdbgpu_array** out) {
  if (!ctx || !c || !out) return GDBGPU_ERR_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    CK(cudaSetDevice(ctx->device));
    if (c->n_fields < 0 || c->n_fields > MAX_FIELDS) throw LimitError("more than 15 allele-dependent fields");
    HostCols h; h.n = c->n_cells; h.coords = c->coords; h.end = c->end;
    h.ref = HostVar{c->ref.offsets, c->ref.values, c->ref.values_bytes};
    h.alt = HostVar{c->alt.offsets, c->alt.values, c->alt.values_bytes};
    h.has_gt = c->gt.offsets != nullptr;
    if (h.has_gt) h.gt = HostVar{c->gt.offsets, c->gt.values, c->gt.values_bytes};
    std::vector<FieldMeta> fm; std::vector<std::string> names;
    for (int f = 0; f < c->n_fields; ++f) {
      h.fields.push_back(HostVar{c->fields[f].offsets, c->fields[f].values, c->fields[f].values_bytes});
      std::string nm = c->field_descs[f].name ? c->field_descs[f].name : "";
      fm.push_back(FieldMeta{c->field_descs[f].length_kind, c->field_descs[f].elem_type, nm == "PL" ? 1 : 0});
      names.push_back(nm);
    }
    std::vector<PlainMeta> pm;
    if (c->n_plain_fields < 0 || c->n_plain_fields > MAX_PLAIN) throw LimitError("more than 24 plain FORMAT fields");
    for (int f = 0; f < c->n_plain_fields; ++f) {
      const gdbgpu_plain_field_desc& d = c->plain_descs[f];
      if (d.elem_type < 0 || d.elem_type > GDBGPU_ELEM_CHAR || d.fixed_len < 0) throw ArgError("bad plain field descriptor");
      if (!d.fixed_len && !c->plain_fields[f].offsets) throw ArgError("variable-length plain field without offsets");
      h.plain.push_back(HostVar{d.fixed_len ? nullptr : c->plain_fields[f].offsets, c->plain_fields[f].values, c->plain_fields[f].values_bytes});
      if (d.info_op && d.info_op != GDBGPU_INFO_ELEMENT_WISE_SUM && (d.fixed_len < 1 || d.elem_type == GDBGPU_ELEM_CHAR)) throw ArgError("an INFO combine operation needs a fixed-length 4-byte field");
      if (d.info_op == GDBGPU_INFO_ELEMENT_WISE_SUM && d.elem_type == GDBGPU_ELEM_CHAR) throw ArgError("element-wise sum of a string field");
      pm.push_back(PlainMeta{d.name ? d.name : "", d.elem_type, d.fixed_len, d.no_format, d.info_op});
    }
    *out = build_array(ctx, h, {h.n}, c->lb_row_idx, c->num_rows, c->gt_phased, fm, names, pm);
  });
}

int gdbgpu_array_open(gdbgpu_ctx* ctx, const char* workspace, const char* array_name, const gdbgpu_field_desc* fields, int32_t n_fields, int32_t gt_phased,
                      gdbgpu_array** out) {
  return gdbgpu_array_open_ex(ctx, workspace, array_name, fields, n_fields, nullptr, 0, gt_phased, out);
}
int gdbgpu_array_open_ex(gdbgpu_ctx* ctx, const char* workspace, const char* array_name, const gdbgpu_field_desc* fields, int32_t n_fields,
                         const gdbgpu_plain_field_desc* plain, int32_t n_plain, int32_t gt_phased, gdbgpu_array** out) {
  return gdbgpu_array_open_range(ctx, workspace, array_name, fields, n_fields, plain, n_plain, gt_phased, INT64_MIN, INT64_MAX, out);
}
int gdbgpu_array_open_range(gdbgpu_ctx* ctx, const char* workspace, const char* array_name, const gdbgpu_field_desc* fields, int32_t n_fields,
                            const gdbgpu_plain_field_desc* plain, int32_t n_plain, int32_t gt_phased, int64_t col_lo, int64_t col_hi, gdbgpu_array** out) {
  if (!ctx || !workspace || !array_name || !out || n_fields < 0 || n_plain < 0 || n_plain > MAX_PLAIN || (n_plain && !plain) || col_lo > col_hi) return GDBGPU_ERR_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    CK(cudaSetDevice(ctx->device));
    TdbArray ta(workspace, array_name);
    const TdbSchema& sc = ta.schema();
    const int nattr = (int)sc.attributes.size();
    auto need = [&](const char* nm) { int id = sc.attr_id(nm); if (id < 0) throw std::runtime_error(std::string("array has no attribute ") + nm); return id; };
    int id_end = need("END"), id_ref = need("REF"), id_alt = need("ALT"), id_gt = sc.attr_id("GT");
    std::vector<int> ids = {nattr, id_end, id_ref, id_alt};
    if (id_gt >= 0) ids.push_back(id_gt);
    std::vector<FieldMeta> fm; std::vector<std::string> names;
    for (int f = 0; f < n_fields; ++f) {
      int id = need(fields[f].name);
      if (!sc.is_var(id) || sc.type_size(id) != 4) throw std::runtime_error(std::string("field is not a variable-length 4-byte attribute: ") + fields[f].name);
      ids.push_back(id);
      fm.push_back(FieldMeta{fields[f].length_kind, fields[f].elem_type, std::string(fields[f].name) == "PL" ? 1 : 0});
      names.push_back(fields[f].name);
    }
    std::vector<PlainMeta> pm;
    for (int f = 0; f < n_plain; ++f) {   // the schema decides fixed / variable and the element size; the descriptor must agree
      int id = need(plain[f].name);
      const size_t es = plain[f].elem_type == GDBGPU_ELEM_CHAR ? 1 : 4;
      if (sc.type_size(id) != es) throw std::runtime_error(std::string("plain field element size does not match the array schema: ") + plain[f].name);
      const int fixed = sc.is_var(id) ? 0 : sc.cell_val_num[id];
      if (plain[f].fixed_len != fixed) throw std::runtime_error(std::string("plain field length does not match the array schema: ") + plain[f].name);
      ids.push_back(id);
      if (plain[f].info_op && ((plain[f].info_op != GDBGPU_INFO_ELEMENT_WISE_SUM && fixed < 1) || plain[f].elem_type == GDBGPU_ELEM_CHAR)) throw std::runtime_error(std::string("an INFO combine operation needs a 4-byte field (fixed-length unless element-wise): ") + plain[f].name);
      pm.push_back(PlainMeta{plain[f].name, plain[f].elem_type, fixed, plain[f].no_format, plain[f].info_op});
    }
    int64_t lb = 0, ub = 0;
    const bool have_rows = ta.row_bounds(lb, ub);
    const size_t nfrag = ta.num_fragments();
    const bool whole = col_lo == INT64_MIN && col_hi == INT64_MAX;
    // Tiles to stage, per fragment: those whose bounding coordinates overlap the column range -- the
    // book-keeping's first / last cell of every tile, bisected.  A call that spans the range's left edge is found through its
    // END copy (gt_get_column scans forward from col_lo until every row has shown a cell); when a row shows no cell inside
    // the staged tiles the range grows to the right, tiles doubling, until every row has or the fragments end.
    std::vector<int64_t> t0(nfrag, 0), t1(nfrag, 0);
    // ... starting with a margin of about 16 cells per row past col_hi: in a cohort where every sample has calls all along the
    // genome that already holds every row's next cell, and the first attempt is the last
    int64_t margin_tiles = 0;
    if (!whole && have_rows && sc.capacity > 0) margin_tiles = (16 * (ub - lb + 1) + sc.capacity - 1) / sc.capacity;
    for (size_t fi = 0; fi < nfrag; ++fi) {
      if (whole) { t0[fi] = 0; t1[fi] = ta.fragment(fi).bk.tile_num; }
      else { ta.tile_range(fi, col_lo, col_hi, t0[fi], t1[fi]); t1[fi] = std::min<int64_t>(ta.fragment(fi).bk.tile_num, t1[fi] + margin_tiles); }
    }
    margin_tiles = std::max<int64_t>(margin_tiles, 16);
    for (;;) {
      std::vector<uint64_t> frag_cells;
      uint64_t n_total = 0;
      bool more_to_the_right = false;
      for (size_t fi = 0; fi < nfrag; ++fi) {
        frag_cells.push_back((uint64_t)ta.tile_cells(fi, t0[fi], t1[fi])); n_total += frag_cells.back();
        if (t1[fi] < ta.fragment(fi).bk.tile_num) more_to_the_right = true;
      }
      // Every needed column of every fragment goes to HBM as it lies on disk; compressed tiles are inflated there (k_inflate,
      // one thread per zlib stream), straight into the column's place in the concatenation over fragments.  Var offsets are
      // rebased onto the concatenated value stream.
      uint64_t file_bytes_staged = 0;   // bytes that crossed PCIe: the files' tiles as they lie on disk
      std::vector<DevColumn> cols(ids.size());
      // compressed files of all attributes are collected first: ONE copy of their bytes, ONE k_inflate launch over every tile
      struct Pending { size_t arena_off; const std::vector<TdbTileSpan>* tiles; uint8_t* dst; std::string path; uint64_t f_off, f_len; };
      std::vector<FileJob> file_jobs;
      std::vector<std::vector<TdbRawColumn>> raw(ids.size(), std::vector<TdbRawColumn>(nfrag));
      std::vector<Pending> pending;
      size_t arena_bytes = 0, n_jobs = 0;
      struct Rebase { uint64_t* p; uint64_t n, add; };
      std::vector<Rebase> rebases;
      for (size_t k = 0; k < ids.size(); ++k) {
        uint64_t fixed_total = 0, var_total = 0;
        for (size_t fi = 0; fi < nfrag; ++fi) { ta.read_column_raw(fi, ids[k], raw[k][fi], t0[fi], t1[fi], /*lazy=*/true); fixed_total += raw[k][fi].fixed_bytes; var_total += raw[k][fi].var_bytes; }
        DevColumn& dc = cols[k];
        dc.fixed.alloc(std::max<uint64_t>(fixed_total, 1) + 16); dc.var.alloc(std::max<uint64_t>(var_total, 1)); dc.var_bytes = var_total;
        uint64_t fixed_at = 0, var_at = 0, cells_at = 0;
        for (size_t fi = 0; fi < nfrag; ++fi) {
          const TdbRawColumn& r = raw[k][fi];
          auto place = [&](const std::string& path, uint64_t f_off, uint64_t f_len, const std::vector<TdbTileSpan>& tiles, uint64_t bytes, uint8_t* dst) {
            if (!bytes) return;
            file_bytes_staged += r.compressed ? f_len : bytes;
            if (!r.compressed) { file_jobs.push_back(FileJob{path, f_off, bytes, dst}); return; }
            pending.push_back(Pending{arena_bytes, &tiles, dst, path, f_off, f_len});
            arena_bytes += (f_len + 15) & ~(size_t)15; n_jobs += tiles.size();
          };
          place(r.fixed_path, r.fixed_off, r.fixed_len, r.fixed_tiles, r.fixed_bytes, dc.fixed.p + fixed_at);
          if (r.is_var) {
            place(r.var_path, r.var_off, r.var_len, r.var_tiles, r.var_bytes, dc.var.p + var_at);
            if (var_at != r.var_base && frag_cells[fi]) rebases.push_back(Rebase{reinterpret_cast<uint64_t*>(dc.fixed.p) + cells_at, frag_cells[fi], var_at - r.var_base});
          }
          fixed_at += r.fixed_bytes; var_at += r.var_bytes; cells_at += frag_cells[fi];
        }
      }
      DevBuf<uint8_t> arena;
      if (n_jobs) arena.alloc(arena_bytes);
      for (const Pending& pd : pending) file_jobs.push_back(FileJob{pd.path, pd.f_off, pd.f_len, arena.p + pd.arena_off});
      stage_files(ctx, file_jobs);   // every byte of the range's tiles is in HBM when this returns
      if (n_jobs) {
        std::vector<InflateJob> jobs; jobs.reserve(n_jobs);
        for (const Pending& pd : pending)
          for (const TdbTileSpan& t : *pd.tiles) jobs.push_back(InflateJob{arena.p + pd.arena_off + t.in_off, t.in_len, pd.dst + t.out_off, t.out_len});
        inflate_jobs(ctx, jobs);
      }
      for (const Rebase& rb : rebases) LAUNCH(ctx, k_add_u64, blocks_for(rb.n, 256), 256, rb.p, rb.n, rb.add);
      CK(cudaStreamSynchronize(ctx->stream));   // the host file buffers (pageable) may go once every copy has run
      raw.clear();
      HostCols h; h.n = n_total;
      h.coords = reinterpret_cast<const int64_t*>(cols[0].fixed.p); h.end = reinterpret_cast<const int64_t*>(cols[1].fixed.p);
      auto hv = [&](size_t k) { return HostVar{reinterpret_cast<const uint64_t*>(cols[k].fixed.p), cols[k].var.p, cols[k].var_bytes}; };
      h.ref = hv(2); h.alt = hv(3);
      size_t k = 4;
      h.has_gt = id_gt >= 0;
      if (h.has_gt) h.gt = hv(k++);
      for (int f = 0; f < n_fields; ++f) h.fields.push_back(hv(k++));
      for (int f = 0; f < n_plain; ++f, ++k) h.plain.push_back(pm[f].fixed_len ? HostVar{nullptr, cols[k].fixed.p, (uint64_t)n_total * pm[f].fixed_len * (pm[f].elem == ELEM_CHAR ? 1u : 4u)} : hv(k));
      if (!have_rows) {   // no GenomicsDB metadata: bound by the data
        std::vector<int64_t> hc(2 * n_total);
        if (n_total) CK(cudaMemcpy(hc.data(), h.coords, n_total * 16, cudaMemcpyDeviceToHost));
        ub = -1;
        for (uint64_t i = 0; i < n_total; ++i) ub = std::max(ub, hc[2 * i]);
      }
      std::unique_ptr<gdbgpu_array> arr(build_array(ctx, h, frag_cells, lb, ub - lb + 1, gt_phased, fm, names, pm, &cols));
      arr->staged_bytes = file_bytes_staged;
      if (!whole && more_to_the_right && arr->num_rows > 0) {
        // a query may begin anywhere inside the range: the call live at its left edge is found through the row's first cell
        // at or after that edge (gt_get_column, query_variants.cc:845-941), so every row has to show a cell at or after col_hi
        DevBuf<uint32_t> unseen; unseen.alloc(1);
        CK(cudaMemsetAsync(unseen.p, 0, 4, ctx->stream));
        LAUNCH(ctx, k_rows_without_cells, blocks_for((uint64_t)arr->num_rows, 256), 256, arr->dev, col_hi, unseen.p);
        uint32_t h_unseen = 0;
        CK(cudaMemcpyAsync(&h_unseen, unseen.p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
        if (h_unseen) {
          tl_ctx = ctx;
          arr.reset();
          for (size_t fi = 0; fi < nfrag; ++fi) t1[fi] = std::min<int64_t>(ta.fragment(fi).bk.tile_num, t1[fi] + margin_tiles);
          margin_tiles *= 2;
          continue;
        }
      }
      *out = arr.release();
      return;
    }
  });
}

int gdbgpu_inflate_streams(gdbgpu_ctx* ctx, const void* in, const uint64_t* in_offsets, void* out, const uint64_t* out_offsets, uint64_t n, double* kernel_ms) {
  if (!ctx || !in_offsets || !out_offsets || (n && (!in || !out))) return GDBGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CK(cudaSetDevice(ctx->device));
    if (kernel_ms) *kernel_ms = 0.0;
    if (!n) return;
    for (uint64_t i = 0; i < n; ++i) if (in_offsets[i + 1] < in_offsets[i] || out_offsets[i + 1] < out_offsets[i]) throw ArgError("offsets must not decrease");
    DevBuf<uint8_t> d_in, d_out;
    d_in.alloc(std::max<uint64_t>(in_offsets[n], 16)); d_out.alloc(std::max<uint64_t>(out_offsets[n], 16));
    CK(cudaMemcpyAsync(d_in.p, in, in_offsets[n], cudaMemcpyHostToDevice, ctx->stream));
    std::vector<InflateJob> jobs(n);
    for (uint64_t i = 0; i < n; ++i) jobs[i] = InflateJob{d_in.p + in_offsets[i], in_offsets[i + 1] - in_offsets[i], d_out.p + out_offsets[i], out_offsets[i + 1] - out_offsets[i]};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    try { inflate_jobs(ctx, jobs); } catch (...) { cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaMemcpyAsync(out, d_out.p, out_offsets[n], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;   // the launch + its error word read back
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  });
}

int gdbgpu_array_column_cuts(const char* workspace, const char* array_name, int32_t n_parts, int64_t* cuts) {
  if (!workspace || !array_name || n_parts < 1 || !cuts) return GDBGPU_ERR_ARG;
  try {
    TdbArray ta(workspace, array_name);
    // tiles of all fragments by the column of their first cell; a part closes at a tile boundary once it holds its share
    struct Tile { int64_t first_col, last_col; int64_t cells; };
    std::vector<Tile> tiles;
    for (size_t fi = 0; fi < ta.num_fragments(); ++fi) {
      const TdbBookKeeping& bk = ta.fragment(fi).bk;
      for (int64_t t = 0; t < bk.tile_num; ++t)
        tiles.push_back(Tile{bk.bounding_coords[4 * t + 1], bk.bounding_coords[4 * t + 3], ta.tile_cells(fi, t, t + 1)});
    }
    if (tiles.empty()) { for (int32_t r = 0; r <= n_parts; ++r) cuts[r] = 0; return GDBGPU_OK; }
    std::sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.first_col < b.first_col; });
    int64_t total = 0, hi = tiles[0].last_col;
    for (const Tile& t : tiles) { total += t.cells; hi = std::max(hi, t.last_col); }
    cuts[0] = tiles[0].first_col; cuts[n_parts] = hi + 1;
    int64_t acc = 0; size_t i = 0;
    for (int32_t r = 1; r < n_parts; ++r) {
      const int64_t want = total * r / n_parts;
      while (i < tiles.size() && acc + tiles[i].cells <= want) acc += tiles[i++].cells;
      // the cut sits at the first column of the next tile, never before the previous cut
      int64_t c = i < tiles.size() ? tiles[i].first_col : hi + 1;
      cuts[r] = std::max(c, cuts[r - 1]);
    }
    return GDBGPU_OK;
  } catch (const std::exception&) { return GDBGPU_ERR_IO; }
}

int gdbgpu_array_get_info(const gdbgpu_array* a, gdbgpu_array_info* info) {
  if (!a || !info) return GDBGPU_ERR_ARG;
  info->n_cells = a->n; info->n_begin_cells = a->n_begin; info->n_fragments = a->n_fragments; info->lb_row_idx = a->lb_row; info->num_rows = a->num_rows;
  info->min_col = a->min_col; info->max_col = a->max_col; info->max_end = a->max_end; info->staged_bytes = a->staged_bytes; info->hbm_bytes = a->hbm_bytes();
  info->n_fields = (int)a->fields.size(); info->has_gt = a->has_gt; info->gt_phased = a->gt_phased; info->n_plain_fields = (int)a->plain.size();
  return GDBGPU_OK;
}
int gdbgpu_array_copy_cells(const gdbgpu_array* a, uint64_t first, uint64_t count, int64_t* coords, int64_t* end) {
  if (!a || first + count > a->n) return GDBGPU_ERR_ARG;
  return guarded(a->ctx, [&] {
    CK(cudaSetDevice(a->ctx->device));
    if (coords && count) CK(cudaMemcpy(coords, a->coords.p + 2 * first, count * 16, cudaMemcpyDeviceToHost));
    if (end && count) CK(cudaMemcpy(end, a->end.p + first, count * 8, cudaMemcpyDeviceToHost));
  });
}
void gdbgpu_array_close(gdbgpu_array* a) { if (a) { cudaSetDevice(a->ctx->device); tl_ctx = a->ctx; delete a; } }

int gdbgpu_scan_begin(gdbgpu_array* arr, const gdbgpu_scan_params* params, gdbgpu_scan** out) {
  if (!arr || !params || !out) return GDBGPU_ERR_ARG;
  *out = nullptr;
  std::unique_ptr<gdbgpu_scan> s(new gdbgpu_scan());
  s->arr = arr; s->ctx = arr->ctx; s->params = *params;
  int rc = guarded(arr->ctx, [&] { CK(cudaSetDevice(arr->ctx->device)); scan_prepare(s.get()); });
  if (rc == GDBGPU_OK) *out = s.release();
  return rc;
}
int gdbgpu_scan_get_info(const gdbgpu_scan* s, gdbgpu_scan_info* info) {
  if (!s || !info) return GDBGPU_ERR_ARG;
  info->n_sites = s->S.n_sites; info->n_breakpoints = s->S.m; info->n_query_rows = s->S.nq; info->n_forward_cells = s->S.i1 - s->S.i0;
  return GDBGPU_OK;
}
int gdbgpu_scan_next_batch(gdbgpu_scan* s, uint32_t max_sites, uint64_t max_bytes, gdbgpu_batch* batch) {
  if (!s || !batch) return GDBGPU_ERR_ARG;
  return guarded(s->ctx, [&] {
    if (s->cells_only) throw ArgError("GDBGPU_SCAN_CELLS_ONLY: the scan serves gdbgpu_scan_cells* only");
    CK(cudaSetDevice(s->ctx->device)); next_batch(s, max_sites, max_bytes, batch);
  });
}
int gdbgpu_scan_cells(gdbgpu_scan* s, gdbgpu_cell_list* out) {
  if (!s || !out) return GDBGPU_ERR_ARG;
  return guarded(s->ctx, [&] { CK(cudaSetDevice(s->ctx->device)); cells_to_host(s, out); });
}
int gdbgpu_scan_cells_window(gdbgpu_scan* s, uint64_t first, uint64_t count) {
  if (!s) return GDBGPU_ERR_ARG;
  s->win_first = first; s->win_count = count ? count : UINT64_MAX;
  return GDBGPU_OK;
}
int gdbgpu_scan_cells_histogram(gdbgpu_scan* s, uint64_t begin, uint64_t end, uint64_t bin_size, uint64_t* counts, uint64_t n_bins) {
  if (!s || !counts) return GDBGPU_ERR_ARG;
  return guarded(s->ctx, [&] { CK(cudaSetDevice(s->ctx->device)); cells_histogram(s, begin, end, bin_size, counts, n_bins); });
}
int gdbgpu_scan_cells_allele_counts(gdbgpu_scan* s, gdbgpu_allele_counts* out) {
  if (!s || !out) return GDBGPU_ERR_ARG;
  return guarded(s->ctx, [&] { CK(cudaSetDevice(s->ctx->device)); cells_allele_counts(s, out); });
}
int gdbgpu_scan_cells_fetch(gdbgpu_scan* s, int32_t kind, int32_t index, const uint64_t** offsets, const uint8_t** bytes) {
  if (!s || !offsets || !bytes) return GDBGPU_ERR_ARG;
  return guarded(s->ctx, [&] { CK(cudaSetDevice(s->ctx->device)); cells_fetch(s, kind, index, offsets, bytes); });
}
int gdbgpu_scan_seek(gdbgpu_scan* s, uint64_t site_idx) {
  if (!s || site_idx > s->S.n_sites) return GDBGPU_ERR_ARG;
  s->cursor = site_idx;
  return GDBGPU_OK;
}
void gdbgpu_scan_end(gdbgpu_scan* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  tl_ctx = s->ctx;
  cudaStreamSynchronize(s->ctx->stream);
  cudaStreamSynchronize(s->ctx->copy_stream);
  for (auto& b : s->slot) if (b) { b->valid = false; s->ctx->bb_pool.push_back(b.release()); }
  delete s;
}

}  // extern "C"
