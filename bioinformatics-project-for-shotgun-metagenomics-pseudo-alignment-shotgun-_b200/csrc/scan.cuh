// scan.cuh -- block- and device-wide exclusive prefix sums used by the CSR passes.
#pragma once
#include "common.cuh"

namespace pa {

// Exclusive scan of one value per thread across a block of THREADS threads.
// `total` receives the block sum in every thread.  Needs THREADS/32 words of shared scratch.
template <int THREADS, typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T& total, T* warp_scratch /* [THREADS/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_scratch[warp] = incl;
  __syncthreads();
  T warp_base = 0, sum = 0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) {
    T t = warp_scratch[w];
    if (w < warp) warp_base += t;
    sum += t;
  }
  __syncthreads();
  total = sum;
  return warp_base + incl - v;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// device-wide exclusive scan: out[i] = sum(in[0..i)), *d_total = sum(in[0..n)).
// d_tile_sums must hold ceil(n / SCAN_TILE) + 1 uint64.
int32_t exclusive_scan_u32(const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_tile_sums, uint64_t* d_total,
                           cudaStream_t s);
inline uint64_t scan_tiles(uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// single-block exclusive scan over n uint64 values in place; total -> *d_total
__global__ void scan_u64_single_block(uint64_t* data, uint64_t n, uint64_t* d_total);

}  // namespace pa
