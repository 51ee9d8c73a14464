// sort.cuh -- K2 interface (see sort.cu)
#pragma once
#include "common.cuh"

namespace pa {

// Bytes of device scratch radix_sort_pairs needs for n pairs.
size_t radix_sort_temp_bytes(uint64_t n);

// Stable LSD radix sort of n (key, value) pairs on key bits [begin_bit, end_bit).
// (keys_a, vals_a) hold the input; (keys_b, vals_b) are same-sized scratch.
// *result_in_b tells which pair of buffers holds the sorted output.
int32_t radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n,
                         int end_bit, void* d_temp, size_t temp_bytes, cudaStream_t s, int* result_in_b, int begin_bit = 0);

// The same result for keys that are spread evenly over [0, 2^end_bit) (the hashed k-mer keys of a build): digit passes
// over the top `top_bits` only (0 = ceil(log2 n) + 10), then a repair of the rare runs of equal top bits that mix
// different keys; falls back to the remaining passes when the keys turn out not to be spread (correct on any input).
// Synchronises the stream.  *fell_back (optional) reports the fallback.  PA_SORT_FULL=1 forces the plain sort.
int32_t radix_sort_pairs_hashed(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n, int end_bit,
                                void* d_temp, size_t temp_bytes, cudaStream_t s, int* result_in_b, int top_bits = 0,
                                int* fell_back = nullptr);

}  // namespace pa
