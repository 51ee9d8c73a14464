// Locality microbenchmark for the minimizer-bucketed lookup table.
//
// Question: when GS neighbouring lanes of a warp look up sectors of the SAME random
// 128-byte line in one load instruction (what consecutive k-mer windows sharing a
// minimizer do), how many lane-lookups/s does a B200 serve from a table >> L2?
// GS = 1 is the fully random case of tools/gather_bench.cu (one line per lookup).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/locality_bench tools/locality_bench.cu
// Run:    tools/locality_bench [table_GiB=16] [iters=64]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ uint32_t ld32nc(const void* p) {
  uint32_t a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
  return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// GS lanes share a line per load instruction; SPREAD = 1: each lane takes a pseudo-random sector of the line,
// SPREAD = 0: all lanes of the group take the same sector (pure broadcast).
// REUSE > 1: the same line is used by REUSE successive load instructions of the group (temporal reuse through L2).
template <int GS, int SPREAD, int REUSE, int UNROLL>
__global__ void __launch_bounds__(256) group_kernel(const uint8_t* __restrict__ table, uint64_t n_lines, int iters,
                                                    uint64_t seed, unsigned long long* sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t gid = tid / GS;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it += UNROLL) {
    uint32_t v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      uint64_t h = mix64(seed + gid * (uint64_t)iters + (uint64_t)((it + u) / REUSE));
      uint64_t line = (uint64_t)(((unsigned __int128)h * n_lines) >> 64);
      uint32_t sec = SPREAD ? (uint32_t)(mix64(tid * 1315423911ULL + it + u) >> 62) : 0u;
      v[u] = ld32nc(table + line * 128 + sec * 32);
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

template <int GS, int SPREAD, int REUSE>
static void run(const uint8_t* table, uint64_t bytes, int iters, int sms, unsigned long long* sink) {
  const int grid = sms * 8 * 4;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    group_kernel<GS, SPREAD, REUSE, 8><<<grid, 256>>>(table, bytes / 128, iters, 0x1234567ULL * (rep + 1), sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  double lookups = (double)grid * 256.0 * iters;
  double rate = lookups / (best * 1e-3);
  printf("{\"table_GiB\": %.2f, \"lanes_per_line\": %d, \"spread_sectors\": %d, \"reuse\": %d, \"ms\": %.3f, "
         "\"lane_lookups_per_s\": %.4e, \"distinct_lines_per_s\": %.4e}\n",
         bytes / 1073741824.0, GS, SPREAD, REUSE, best, rate, rate / GS / REUSE);
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
  double gib = argc > 1 ? atof(argv[1]) : 16.0;
  int iters = argc > 2 ? atoi(argv[2]) : 64;
  uint64_t bytes = (uint64_t)(gib * (1ULL << 30));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, prop.multiProcessorCount);
  uint8_t* table; CK(cudaMalloc(&table, bytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
  int sms = prop.multiProcessorCount;
  fill_kernel<<<sms * 8, 256>>>((uint4*)table, bytes / 16);
  CK(cudaDeviceSynchronize());
  run<1, 1, 1>(table, bytes, iters, sms, sink);
  run<2, 1, 1>(table, bytes, iters, sms, sink);
  run<4, 1, 1>(table, bytes, iters, sms, sink);
  run<8, 1, 1>(table, bytes, iters, sms, sink);
  run<16, 1, 1>(table, bytes, iters, sms, sink);
  run<32, 1, 1>(table, bytes, iters, sms, sink);
  run<8, 0, 1>(table, bytes, iters, sms, sink);
  run<1, 1, 2>(table, bytes, iters, sms, sink);
  run<1, 1, 4>(table, bytes, iters, sms, sink);
  run<1, 1, 8>(table, bytes, iters, sms, sink);
  run<8, 1, 2>(table, bytes, iters, sms, sink);
  CK(cudaFree(table));
  return 0;
}
