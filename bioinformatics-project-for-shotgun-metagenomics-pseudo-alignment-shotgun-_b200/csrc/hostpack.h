// hostpack.h -- host-side 2-bit packing of read batches for the host-buffer alignment path (pa_align_batch).
//
// PCIe is the bottleneck of the end-to-end path once the reads are aligned at > 5e8 reads/s on the device: a 150-bp
// read is 150 bytes of ASCII but only 2 x 5 32-bit words as bit planes.  The host therefore turns every chunk of
// reads into exactly the planes the kernel's ballots would compute (bit i of low word c = bit 1 of the ASCII code of
// base 32c + i, high word = bit 2; A=0 C=1 T=2 G=3) on all cores while the previous chunk is on the wire.
// Layout per chunk: read i (chunk-local), starting at base offset o, with L bases and nw = ceil(L / 32):
//   words [w0, w0 + nw) low plane, [w0 + nw, w0 + 2 nw) high plane,  w0 = 2 * ((o - base0) / 32 + i)
// (computable on both sides from the offsets alone; at most two spare words per read).
#pragma once
#include <cstdint>
#include <functional>

namespace pa {

// number of 32-bit words a chunk of n_reads reads with n_bases bases needs
inline uint64_t planes_words(uint64_t n_bases, uint64_t n_reads) { return 2 * (n_bases / 32 + n_reads + 1); }

// Packs reads [lo, hi) (offsets read_off[lo .. hi], bases indexed absolutely) into `planes` using up to n_threads
// worker threads.  Returns false when a base outside ACGT was met (the chunk then has to travel as ASCII).
// The offsets must be monotonic (check with scan_offsets first).
bool pack_reads_planes(const uint8_t* bases, const uint64_t* read_off, uint64_t lo, uint64_t hi, uint32_t* planes,
                       int n_threads);

// Parallel check of read_off[lo .. hi]: returns false when it is not monotonic; *max_len / *min_len = longest / shortest read.
bool scan_offsets(const uint64_t* read_off, uint64_t lo, uint64_t hi, uint64_t* max_len, uint64_t* min_len, int n_threads);

// runs fn(0 .. n_tasks-1) on the library's worker pool (blocking)
void host_parallel_for(int n_tasks, const std::function<void(int)>& fn);

int host_pack_threads();   // worker threads available (hardware concurrency, capped; PA_PACK_THREADS overrides)

}  // namespace pa
