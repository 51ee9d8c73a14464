// index.cuh -- the device-resident k-mer index (CSR + lookup table).
#pragma once
#include "common.cuh"
#include <vector>

namespace pa {

// Scratch reuse inside a build.  A build allocates and frees a dozen multi-GB buffers, and cudaMalloc / cudaFree of
// that size map / unmap memory and synchronise the device: measured on the B200 boxes that is tens of ms per call and
// occasionally hundreds.  Inside an AllocScope (one per public build entry point; everything in it runs on the one
// stream of the index) a released DevBuf is parked instead of freed and the next allocation that fits takes it over --
// stream order makes the hand-over safe -- so e.g. the CSR arrays land in the sort's double buffers.  What is still
// parked when the outermost scope ends is freed.  PA_NO_CACHE=1 switches the reuse off.
// (A stream-ordered cudaMallocAsync pool was tried first and measured slower: it re-grows on every build.)
struct AllocScope {
  explicit AllocScope(cudaStream_t s);
  ~AllocScope();
  AllocScope(const AllocScope&) = delete;
  AllocScope& operator=(const AllocScope&) = delete;
  bool outermost;
};
bool scope_take(size_t bytes, void** p, size_t* got);   // a parked block of >= bytes (and not absurdly larger), if any
bool scope_park(void* p, size_t bytes);                 // false outside a scope: the caller frees

// Retired-buffer cache (process-wide, per device): what is still parked when a build ends, and the big buffers of an index
// that is freed, are kept instead of cudaFree'd -- up to PA_CACHE_GB (default: 45 % of the device memory, 0 switches
// it off) -- and the next build takes them over.  cudaMalloc / cudaFree of multi-GB buffers cost milliseconds each and
// synchronise the device; a process that builds more than once (EXTSIM rebuilds, benchmarks, a server) pays them once.
// Every cached buffer is idle: it is retired only after the stream that used it was synchronised.  cudaMalloc failures
// trim the cache and retry; pa_trim_memory() releases everything.
bool cache_take(size_t bytes, void** p, size_t* got);
bool cache_put(void* p, size_t bytes);                  // false: not cached, the caller frees
void cache_trim();
size_t cache_held();                                    // bytes cached for the current device (memory that is free for a build)
void alloc_miss(size_t bytes);                          // PA_TRACE_ALLOC=1: reports every allocation that had to call cudaMalloc

// Owns a device allocation; frees on destruction.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && !scope_park(p, bytes) && !cache_put(p, bytes)) cudaFree(p);
    p = nullptr; bytes = 0;
  }
  int32_t alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    size_t got = 0;
    if (scope_take(n, &p, &got) || cache_take(n, &p, &got)) { bytes = got; return ST_OK; }
    alloc_miss(n);
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaErrorMemoryAllocation) { (void)cudaGetLastError(); cache_trim(); e = cudaMalloc(&p, n); }
    if (e != cudaSuccess) { p = nullptr; (void)cudaGetLastError(); set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e)); return ST_NOMEM; }
    bytes = n;
    return ST_OK;
  }
  void swap(DevBuf& o) { std::swap(p, o.p); std::swap(bytes, o.bytes); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Index = CSR over distinct k-mers (ascending key) + the align lookup structures.
//   ukeys[U]                      distinct keys, ascending
//   run_off[U+1]                  key -> its genome runs
//   run_genome[R]                 genome index of each run (ascending inside a key)
//   pos_off[R+1]                  run -> its positions
//   pos[N]                        positions inside the genome (ascending inside a run)
//   genome_off[G+1]               base offsets of the genomes in the concatenated input
//   slots / stash / mlist         see TableView in common.cuh
struct Index {
  int32_t k = 0;
  int32_t device = 0;
  uint32_t n_genomes = 0;
  uint64_t n_keys = 0, n_runs = 0, n_occ = 0;
  uint64_t total_bases = 0;
  cudaStream_t stream = nullptr;
  DevBuf ukeys, run_off, run_genome, pos_off, pos, genome_off;
  // first_occ[U]: global base position of each k-mer's first occurrence in the ORIGINAL genome list (the dict
  // insertion order of kmer.py:146-147 survives genome removal, kmer.py:237-243).  Materialised lazily.
  DevBuf first_occ;
  bool has_first_occ = false;
  bool align_only = false;   // table-only index (replica of a partitioned build, streamed build): no CSR at all
  bool no_tables = false;    // partition of a partitioned build: CSR only, the lookup table lives in the replica
  float t_dist_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // phases of the partitioned build that produced this replica
  std::vector<uint64_t> h_genome_off;
  // lookup structures
  DevBuf slots, stash, mlist;
  uint32_t bpd = 1, hi_bits = 0, tag_bits = 1, val_bits = 63, gbits = 1, n_inline = 1;   // table geometry, see TableView
  uint32_t min_len = 1;   // minimizer length m = min(k, 16)
  uint64_t stash_cap = 0;
  uint32_t stash_count = 0;
  uint64_t n_msectors = 0;
  MixParams mix{1, 1};
  // per-warp scratch of the align kernel (allocated on first use)
  DevBuf align_scratch;
  DevBuf align_queue;   // {count u64, pad, uint32 read indices}: reads the fast kernel hands to the general kernel
  DevBuf align_qmasks;  // EXTQUALITY: per read a 128-bit window mask, then per read a "dropped" byte (quality_masks_kernel)
  uint64_t align_scratch_warps = 0, align_scratch_stride = 0;
  // host-buffer alignment path (pa_align_batch): chunk slots so that packing / the H2D copy of later chunks overlaps the
  // kernel of chunk i; buffers grow on demand and are kept for the next call
  struct HostSlot {
    DevBuf bases, quals, off, words, planes;
    uint32_t* h_planes = nullptr;      // pinned staging of the host-packed bit planes (hostpack.h)
    uint64_t h_planes_words = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t kernel_done = nullptr, h2d_done = nullptr;
    cudaEvent_t copy_beg = nullptr, copy_end = nullptr;   // bracket a raw ASCII transfer: measures the link rate
    uint64_t copy_bytes = 0;
  } slot[4];
  static constexpr int N_HOST_SLOTS = 4;
  DevBuf host_list, host_state;
  double pack_rate_gbs = 0;   // running estimates (GB/s): host packing rate of ASCII, raw H2D rate -- see pa_align_batch
  double link_rate_gbs = 0;
  // timing of the last build (ms, CUDA events on `stream`)
  float t_encode_ms = 0, t_sort_ms = 0, t_rle_ms = 0, t_table_ms = 0;

  TableView view() const {
    TableView t;
    t.slots = slots.as<uint64_t>();
    t.stash = stash.as<ulonglong2>();
    t.mlist = mlist.as<uint32_t>();
    t.stash_mask = stash_cap ? stash_cap - 1 : 0;
    t.stash_count = stash_count;
    minimizer_params(t, k);
    t.bpd = bpd;
    t.hi_bits = hi_bits;
    t.hmask = hi_bits >= 32 ? 0xFFFFFFFFu : ((1u << hi_bits) - 1);
    t.tag_bits = tag_bits;
    t.val_bits = val_bits;
    t.gbits = gbits;
    t.n_inline = n_inline;
    t.inl_ones = 0;
    for (uint32_t i = 0; i < n_inline; ++i) t.inl_ones |= 1ULL << (i * gbits);
    t.inl_highs = t.inl_ones << (gbits - 1);
    return t;
  }
  uint64_t n_blocks() const { return (uint64_t)bpd << digit_bits_for_k(k); }
  size_t device_bytes() const {
    return ukeys.bytes + run_off.bytes + run_genome.bytes + pos_off.bytes + pos.bytes + genome_off.bytes + first_occ.bytes +
           slots.bytes + stash.bytes + mlist.bytes + align_scratch.bytes + align_queue.bytes + align_qmasks.bytes;
  }
  ~Index() {
    for (DevBuf* b : {&ukeys, &run_off, &run_genome, &pos_off, &pos, &genome_off, &first_occ, &slots, &stash, &mlist,
                      &align_scratch, &align_queue, &align_qmasks, &host_list, &host_state})
      b->release();
    if (stream) cudaStreamSynchronize(stream);
    for (auto& sl : slot) { if (sl.stream) cudaStreamDestroy(sl.stream); if (sl.kernel_done) cudaEventDestroy(sl.kernel_done); if (sl.h2d_done) cudaEventDestroy(sl.h2d_done); if (sl.copy_beg) cudaEventDestroy(sl.copy_beg); if (sl.copy_end) cudaEventDestroy(sl.copy_end); if (sl.h_planes) cudaFreeHost(sl.h_planes); }
    if (stream) cudaStreamDestroy(stream);
  }
};

// K1 options (build.cu: encode_windows)
struct EncodeOpts {
  uint64_t emit_total;        // windows starting at or beyond this base of the slice are left to the next chunk
  uint8_t* owner;             // [total] out: rank that owns the record (255: invalid window / another round); null: not partitioned
  uint32_t n_parts, n_rounds, round;   // part = (digit of the minimizer * n_parts) >> digit_bits; owner = part / n_rounds, in round part % n_rounds
  uint32_t vals_are_genomes;  // vals = genome index instead of global position (table-only builds: no 2^32-base limit)
};

// build.cu
int32_t index_build_from_device_bases(Index& ix, const uint8_t* d_bases);
// K1 over a slice of the concatenated genomes (see encode_windows); *h_n_valid += valid windows emitted
int32_t encode_slice(const uint8_t* d_bases, uint64_t total, uint64_t pos0, const uint64_t* d_genome_off, uint32_t G, int k,
                     uint64_t* d_keys, uint32_t* d_vals, const EncodeOpts& opt, unsigned long long* d_counters /*[2]: valid, bad*/,
                     cudaStream_t s);
// K3: CSR of `ix` from n sorted records (vals = global positions, or genome indices without positions)
int32_t rle_to_csr(Index& ix, const uint64_t* d_keys, const uint32_t* d_vals, uint64_t n, bool vals_are_genomes);
int32_t index_export_order(Index& ix, uint32_t* h_order);
int32_t index_checksum(Index& ix, uint64_t h_out[4]);
int32_t index_ensure_first_occ(Index& ix);
int32_t index_lookup_ranks(Index& ix, const uint8_t* h_kmers, uint64_t n, uint64_t* h_rank);
int32_t index_extsim_stats(Index& ix, const uint32_t* h_group, uint32_t n_groups, uint64_t* h_total, uint64_t* h_unique);
int32_t index_extsim_pairwise(Index& ix, const uint32_t* h_group, uint32_t n_groups, uint64_t* h_inter);
int32_t index_drop_genomes(Index& ix, const uint8_t* h_keep);

}  // namespace pa
