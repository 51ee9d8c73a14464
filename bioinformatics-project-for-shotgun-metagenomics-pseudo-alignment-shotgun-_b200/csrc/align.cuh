// align.cuh -- K4 / K8 interface (see align.cu)
#pragma once
#include "index.cuh"
#include <algorithm>

namespace pa {

// thresholds of Read.pseudo_align (kmer.py:482-489), already clamped by the host to ranges in which the
// integer comparisons below cannot overflow (see capi.cu: clamp_params)
struct AlignParams {
  int64_t m;    // unique threshold, >= 0
  int64_t p;    // ambiguous threshold; < 0 disables validate_unique_mappings (kmer.py:469)
  int64_t mrq;  // min read quality  : drop when sum(q) < mrq * len
  int64_t mkq;  // min k-mer quality : skip window when sum(q[window]) < mkq * k
  int64_t mg;   // max genomes       : skip k-mer when n_genomes > mg
  int32_t has_mrq, has_mkq, has_mg;
  int32_t pad;
};

// Result word per read: bits 63:62 type (0 dropped by read quality, 1 unmapped, 2 unique, 3 ambiguous);
// bits 61:40 list length; bits 39:0 the genome index when length == 1, else the offset of the list in d_list.
int32_t align_batch_device(Index& ix, const uint8_t* d_bases, const uint8_t* d_quals, const uint64_t* d_read_off,
                           uint64_t n_reads, uint64_t max_read_len, const AlignParams& prm, uint64_t* d_words,
                           uint32_t* d_list, uint64_t list_cap, unsigned long long* d_cursor,
                           unsigned long long* d_counters, cudaStream_t s, int32_t* launches,
                           const uint32_t* d_planes = nullptr, uint64_t planes_base0 = 0, uint64_t planes_read0 = 0);
// d_planes != nullptr: the reads arrive as host-packed bit planes (see ReadInput in align.cu / hostpack.h) and
// d_bases is not read

int32_t summary_reduce_device(const uint64_t* d_words, const uint32_t* d_list, uint64_t n_reads, uint64_t read_index_base,
                              uint32_t G, unsigned long long* d_stats, unsigned long long* d_unique,
                              unsigned long long* d_ambiguous, unsigned long long* d_first_seen, cudaStream_t s);

}  // namespace pa
