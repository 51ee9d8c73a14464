// Random-access roofline microbenchmark for the pseudo-alignment lookup path.
//
// SURVEY.md §8(d) defines "the random-access lookup roofline" as the measured
// rate at which a B200 serves uniformly random aligned 32-byte sectors from a
// table the size of the k-mer index (>> 126 MB L2).  This program measures it
// for several access shapes so the index layout (slot / bucket size, which load
// instruction) is chosen from data instead of guessed.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gather_bench tools/gather_bench.cu
// Run:    tools/gather_bench [table_GiB=16] [lookups_per_thread=64]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

enum Mode {
  M_NC16 = 0,    // one ld.global.nc.L1::no_allocate.v4.u32 per lookup (16 B of a 32-B sector)
  M_CA16 = 1,    // one plain ld.global.v4.u32 per lookup
  M_NC8 = 2,     // one 8-byte nc load per lookup
  M_NC32 = 3,    // one 256-bit ld.global.nc.v8.u32 per lookup (a whole sector)
  M_NC16x2 = 4,  // two 16-byte nc loads per lookup (both halves of a sector)
  M_NC32x2 = 5,  // two 256-bit loads per lookup (64-B granule)
  M_NC32x4 = 6,  // four 256-bit loads per lookup (128-B line)
  M_PAIR16 = 7,  // lane pairs share a sector: 16 lookups per warp instruction, each lane loads 16 B
};

template <int MODE> struct Traits;
template <> struct Traits<M_NC16>   { static constexpr int G = 32;  static constexpr const char* name = "nc16"; };
template <> struct Traits<M_CA16>   { static constexpr int G = 32;  static constexpr const char* name = "ca16"; };
template <> struct Traits<M_NC8>    { static constexpr int G = 32;  static constexpr const char* name = "nc8"; };
template <> struct Traits<M_NC32>   { static constexpr int G = 32;  static constexpr const char* name = "nc32(256b)"; };
template <> struct Traits<M_NC16x2> { static constexpr int G = 32;  static constexpr const char* name = "nc16x2"; };
template <> struct Traits<M_NC32x2> { static constexpr int G = 64;  static constexpr const char* name = "nc32x2(64B)"; };
template <> struct Traits<M_NC32x4> { static constexpr int G = 128; static constexpr const char* name = "nc32x4(128B)"; };
template <> struct Traits<M_PAIR16> { static constexpr int G = 32;  static constexpr const char* name = "pair16"; };

__device__ __forceinline__ uint32_t ld16nc(const void* p) {
  uint32_t a, b, c, d;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
  return a ^ b ^ c ^ d;
}
__device__ __forceinline__ uint32_t ld16ca(const void* p) {
  uint32_t a, b, c, d;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
  return a ^ b ^ c ^ d;
}
__device__ __forceinline__ uint32_t ld8nc(const void* p) {
  uint32_t a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  return a ^ b;
}
__device__ __forceinline__ uint32_t ld32nc(const void* p) {
  uint32_t a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
  return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ table, uint64_t n_granules,
                                                     int iters, uint64_t seed, unsigned long long* sink) {
  constexpr int G = Traits<MODE>::G;
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t sub = 0;
  if (MODE == M_PAIR16) { sub = (uint32_t)(tid & 1) * 16; tid >>= 1; }
  uint32_t acc = 0;
  for (int it = 0; it < iters; it += UNROLL) {
    uint32_t v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      uint64_t h = mix64(seed + tid * (uint64_t)iters + it + u);
      uint64_t g = (uint64_t)(((unsigned __int128)h * n_granules) >> 64);
      const uint8_t* p = table + g * (uint64_t)G;
      if (MODE == M_NC16) v[u] = ld16nc(p);
      if (MODE == M_CA16) v[u] = ld16ca(p);
      if (MODE == M_NC8) v[u] = ld8nc(p);
      if (MODE == M_NC32) v[u] = ld32nc(p);
      if (MODE == M_NC16x2) v[u] = ld16nc(p) ^ ld16nc(p + 16);
      if (MODE == M_NC32x2) v[u] = ld32nc(p) ^ ld32nc(p + 32);
      if (MODE == M_NC32x4) v[u] = ld32nc(p) ^ ld32nc(p + 32) ^ ld32nc(p + 64) ^ ld32nc(p + 96);
      if (MODE == M_PAIR16) v[u] = ld16nc(p + sub);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u];
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ULL);
}

__global__ void fill_kernel(uint4* p, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x9e3779b9u, 0x7f4a7c15u);
}

__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

template <int MODE, int UNROLL>
static void run(const uint8_t* table, uint64_t bytes, int iters, int ctas_per_sm, int sms, unsigned long long* sink) {
  constexpr int G = Traits<MODE>::G;
  uint64_t n_granules = bytes / G;
  int grid = sms * ctas_per_sm * 4;  // 4 waves of resident CTAs
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    gather_kernel<MODE, UNROLL><<<grid, 256>>>(table, n_granules, iters, 0x1234567ULL * (rep + 1), sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  double lookups = (double)grid * 256.0 * iters / (MODE == M_PAIR16 ? 2 : 1);
  double rate = lookups / (best * 1e-3);
  printf("{\"table_GiB\": %.2f, \"mode\": \"%s\", \"granule\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, "
         "\"lookups_per_s\": %.4e, \"granule_GBps\": %.1f}\n",
         bytes / 1073741824.0, Traits<MODE>::name, G, UNROLL, ctas_per_sm, best, rate, rate * G / 1e9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  double max_gib = argc > 1 ? atof(argv[1]) : 16.0;
  int iters = argc > 2 ? atoi(argv[2]) : 64;
  uint64_t max_bytes = (uint64_t)(max_gib * (1ULL << 30));
  if (argc > 3) {  // L2 fetch granularity hint (bytes): 32, 64 or 128
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[3])));
  }
  size_t gran = 0; CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  printf("{\"l2_fetch_granularity\": %zu}\n", gran);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d}\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize);
  uint8_t* table; CK(cudaMalloc(&table, max_bytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
  int sms = prop.multiProcessorCount;
  fill_kernel<<<sms * 8, 256>>>((uint4*)table, max_bytes / 16);
  CK(cudaDeviceSynchronize());
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint64_t half = max_bytes / 2 / 16;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      copy_kernel<<<sms * 16, 256>>>((const uint4*)table, (uint4*)table + half, half);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::min(best, ms);
    }
    printf("{\"stream_copy_GBps\": %.1f}\n", 2.0 * half * 16 / (best * 1e-3) / 1e9);
  }
  if (argc > 4) {  // single-mode run for ncu: one table size, nc32 only
    uint64_t bytes = max_bytes;
    run<M_NC32, 8>(table, bytes, iters, 8, sms, sink);
    CK(cudaFree(table));
    return 0;
  }
  for (double gib : {0.25, 8.0, 64.0}) {
    if (gib > max_gib) break;
    uint64_t bytes = (uint64_t)(gib * (1ULL << 30));
    for (int cps : {4, 8}) {
      run<M_NC16, 8>(table, bytes, iters, cps, sms, sink);
      run<M_CA16, 8>(table, bytes, iters, cps, sms, sink);
      run<M_NC8, 8>(table, bytes, iters, cps, sms, sink);
      run<M_NC32, 8>(table, bytes, iters, cps, sms, sink);
      run<M_NC16x2, 8>(table, bytes, iters, cps, sms, sink);
      run<M_NC32x2, 4>(table, bytes, iters, cps, sms, sink);
      run<M_NC32x4, 4>(table, bytes, iters, cps, sms, sink);
      run<M_PAIR16, 8>(table, bytes, iters, cps, sms, sink);
    }
  }
  CK(cudaFree(table));
  return 0;
}
