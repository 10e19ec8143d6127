// Write-stream microbenchmark behind the K5 design: what does a pure store stream reach on this
// B200, and does it matter how the bytes leave the SM?
//   stg        : every lane STG.128, a warp covers 512 contiguous bytes (what k_remap_tile's phase B did in round 1)
//   bulk<C>    : the block fills a C-byte tile in shared memory (STS.128), fence.proxy.async, ONE lane issues
//                cp.async.bulk.global.shared::cta (SASS UBLKCP), double-buffered with cp.async.bulk.wait_group.read
//   bulk_rows  : same tile, but every thread issues the bulk copy of its own 224-byte row (rows padded in shared memory)
//   memset     : cudaMemsetAsync
//   copy       : STREAM-style copy, read + write bytes (the definition of MEASURED_PEAKS.json hbm_gbs)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bench store_bench.cu ; prints one JSON object.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(256) k_stg(uint4* __restrict__ out, uint64_t n_quads) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) out[i] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n_quads) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) out[i] = in[i];
}
// persistent blocks; tile t of `chunk` bytes goes to out + t * chunk
__global__ void __launch_bounds__(256) k_bulk(uint8_t* __restrict__ out, uint64_t n_tiles, uint32_t chunk) {
  extern __shared__ __align__(128) uint8_t smem[];
  int buf = 0;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, buf ^= 1) {
    uint4* s = reinterpret_cast<uint4*>(smem + (size_t)buf * chunk);
    if (threadIdx.x == 0) bulk_wait_read<1>();   // the copy that last read this buffer has finished reading it
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < chunk / 16; i += blockDim.x) s[i] = make_uint4((uint32_t)t, i, 2u, 3u);
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0) { bulk_store(out + t * chunk, s, chunk); bulk_commit(); }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();
}
// rows of 224 bytes (a diploid 10-allele PL vector), padded to 240 in shared memory; every thread stores its own row
__global__ void __launch_bounds__(256) k_bulk_rows(uint8_t* __restrict__ out, uint64_t n_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t ROW = 224, PAD = 240;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    uint4* s = reinterpret_cast<uint4*>(smem + (size_t)threadIdx.x * PAD);
    bulk_wait_read<0>();
#pragma unroll
    for (uint32_t i = 0; i < ROW / 16; ++i) s[i] = make_uint4((uint32_t)t, i, 2u, 3u);
    fence_async();
    bulk_store(out + (t * 256 + threadIdx.x) * ROW, s, ROW);
    bulk_commit();
  }
  bulk_wait_read<0>();
}

// row-owner stores: every lane writes its own 224-byte row (a warp instruction touches 32 rows)
__global__ void __launch_bounds__(256) k_rows128(uint8_t* __restrict__ out, uint64_t n_rows) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    uint4* d = reinterpret_cast<uint4*>(out + r * 224);
#pragma unroll
    for (int i = 0; i < 14; ++i) d[i] = make_uint4((uint32_t)r, i, 2u, 3u);
  }
}
__global__ void __launch_bounds__(256) k_rows256(uint8_t* __restrict__ out, uint64_t n_rows) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    uint8_t* d = out + r * 224;
#pragma unroll
    for (int i = 0; i < 7; ++i)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d + 32 * i), "r"((uint32_t)r), "r"(i), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7) : "memory");
  }
}

template <class F>
static double time_ms(F&& f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const uint64_t bytes = 8ull << 30;
  uint8_t *a, *b;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("{\"sms\": %d, \"bytes\": %llu", sms, (unsigned long long)bytes);
  auto gbs = [&](double ms, double factor) { return factor * (double)bytes / 1e6 / ms; };
  printf(", \"copy_gbs\": %.1f", gbs(time_ms([&] { k_copy<<<sms * 16, 256>>>((const uint4*)a, (uint4*)b, bytes / 16); }, 10), 2.0));
  printf(", \"memset_gbs\": %.1f", gbs(time_ms([&] { CK(cudaMemsetAsync(b, 3, bytes)); }, 10), 1.0));
  for (int bps : {4, 8, 16, 32}) printf(", \"stg_gbs_x%d\": %.1f", bps, gbs(time_ms([&] { k_stg<<<sms * bps, 256>>>((uint4*)b, bytes / 16); }, 10), 1.0));
  for (uint32_t chunk : {2048u, 8192u, 16384u, 32768u, 57344u}) for (int bps : {1, 2, 3, 4, 6}) {
    const size_t sm = (size_t)chunk * 2;
    if (sm * bps > 220 * 1024) continue;
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    printf(", \"bulk_%u_x%d_gbs\": %.1f", chunk, bps, gbs(time_ms([&] { k_bulk<<<sms * bps, 256, sm>>>(b, bytes / chunk, chunk); }, 5), 1.0));
  }
  {
    const size_t sm = 256 * 240;
    CK(cudaFuncSetAttribute(k_bulk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const uint64_t n_tiles = bytes / (256 * 224);
    for (int bps : {1, 2, 3}) printf(", \"bulk_rows224_x%d_gbs\": %.1f", bps, (double)n_tiles * 256 * 224 / 1e6 / time_ms([&] { k_bulk_rows<<<sms * bps, 256, sm>>>(b, n_tiles); }, 5));
  }
  {
    const uint64_t n_rows = bytes / 224;
    for (int bps : {4, 8, 16}) {
      printf(", \"rows224_stg128_x%d_gbs\": %.1f", bps, (double)n_rows * 224 / 1e6 / time_ms([&] { k_rows128<<<sms * bps, 256>>>(b, n_rows); }, 5));
      printf(", \"rows224_stg256_x%d_gbs\": %.1f", bps, (double)n_rows * 224 / 1e6 / time_ms([&] { k_rows256<<<sms * bps, 256>>>(b, n_rows); }, 5));
    }
  }
  printf("}\n");
  return 0;
}
