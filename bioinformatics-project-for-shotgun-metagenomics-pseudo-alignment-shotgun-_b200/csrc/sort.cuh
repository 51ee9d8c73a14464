// sort.cuh -- K2 interface (see sort.cu)
#pragma once
#include "common.cuh"

namespace pa {

// destination of one digit's run in the routed scatter pass (multi-GPU build): device pointers, possibly peer memory
struct PeerRoute {
  uint64_t* keys;   // null: drop the digit
  uint32_t* vals;
};

// Bytes of device scratch radix_sort_pairs needs for n pairs.
size_t radix_sort_temp_bytes(uint64_t n);

// Stable LSD radix sort of n (key, value) pairs on key bits [begin_bit, end_bit).
// (keys_a, vals_a) hold the input; (keys_b, vals_b) are same-sized scratch.
// *result_in_b tells which pair of buffers holds the sorted output.
int32_t radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n,
                         int end_bit, void* d_temp, size_t temp_bytes, cudaStream_t s, int* result_in_b, int begin_bit = 0);

int32_t radix_digit_counts(const uint64_t* keys, uint64_t n, int begin_bit, unsigned long long* h_counts, void* d_temp,
                           size_t temp_bytes, cudaStream_t s);
int32_t radix_scatter_routed(const uint64_t* keys, const uint32_t* vals, uint64_t n, int begin_bit, const PeerRoute* d_route,
                             void* d_temp, size_t temp_bytes, cudaStream_t s);

}  // namespace pa
