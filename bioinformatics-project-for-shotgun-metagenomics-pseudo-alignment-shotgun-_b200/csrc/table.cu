// table.cu -- construction of the minimizer-bucketed lookup table from a CSR (see table.cuh, TableView in common.cuh).
// Replaces the hashing Python's dict does for KmerReference.kmers (/root/reference/src/kmer.py:130, 146-150, 292-298).
#include "table.cuh"
#include "scan.cuh"
#include "sort.cuh"

#include <algorithm>
#include <cstdlib>

namespace pa {

namespace {

inline unsigned grid_for(uint64_t n, int threads) { return (unsigned)std::max<uint64_t>(1, (n + threads - 1) / threads); }

// Related genomes share long runs of k-mers, so the number of DISTINCT genome sets is tiny next to the number of
// k-mers carrying them (config B: a few hundred sets for 10^7 k-mers).  Storing each set once keeps mlist inside
// L1/L2 and keeps the slot payload (a sector index) short.
//   long_count    number of k-mers with more than n_inline genomes
//   long_collect  (hash of the genome list, u) for each of them, appended in any order
//   [radix sort by hash]
//   set_heads     an entry opens a new set unless its list equals its sorted predecessor's (hash AND content)
//   set_assign    msec_off[u] = first sector of the k-mer's set; heads write their list into mlist
__global__ void __launch_bounds__(256) long_count(const uint64_t* __restrict__ run_off, uint64_t U, uint32_t n_inline,
                                                  unsigned long long* __restrict__ total) {
  const uint64_t stride = (uint64_t)gridDim.x * 256;
  uint32_t mine = 0;
  for (uint64_t u = blockIdx.x * 256ull + threadIdx.x; u < U; u += stride) mine += (run_off[u + 1] - run_off[u]) > n_inline ? 1u : 0u;
  mine = warp_sum(mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, (unsigned long long)mine);
}

// (hash of the genome list, u) for every k-mer with a long list, appended through a cursor that a block advances once per
// tile of 4,096 k-mers (one atomic per warp on a single address was the whole cost of this kernel: 7.9 ms); the order of
// the entries is arbitrary (they are sorted by hash next; which k-mer of a set becomes its head does not matter)
constexpr int LC_ITEMS = 16;
constexpr int LC_TILE = 256 * LC_ITEMS;
__global__ void __launch_bounds__(256) long_collect(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome, uint64_t U,
                                                    uint32_t n_inline, unsigned long long* __restrict__ cursor,
                                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  __shared__ uint32_t scratch[256 / 32];
  __shared__ unsigned long long s_base;
  const uint64_t tiles = (U + LC_TILE - 1) / LC_TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t u0 = tile * LC_TILE + threadIdx.x;
    uint32_t flags = 0;
#pragma unroll
    for (int r = 0; r < LC_ITEMS; ++r) {
      const uint64_t u = u0 + (uint64_t)r * 256;
      if (u < U && run_off[u + 1] - run_off[u] > n_inline) flags |= 1u << r;
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan<256, uint32_t>((uint32_t)__popc(flags), total, scratch);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(cursor, (unsigned long long)total) : 0ULL;
    __syncthreads();
    uint64_t at = s_base + excl;
    while (flags) {
      const int r = __ffs(flags) - 1;
      flags &= flags - 1;
      const uint64_t u = u0 + (uint64_t)r * 256;
      uint64_t h = 0xCBF29CE484222325ULL;
      for (uint64_t q = run_off[u], q1 = run_off[u + 1]; q < q1; ++q) { h ^= run_genome[q]; h *= 0x100000001B3ULL; h ^= h >> 29; }
      keys[at] = h;
      vals[at] = (uint32_t)u;
      ++at;
    }
    __syncthreads();   // s_base is rewritten by the next tile
  }
}

__global__ void set_heads(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n,
                          const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome,
                          uint32_t* __restrict__ head_secs) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t u = vals[i], r0 = run_off[u], c = run_off[u + 1] - r0;
  bool head = true;
  if (i > 0 && keys[i] == keys[i - 1]) {
    const uint64_t v = vals[i - 1], q0 = run_off[v];
    if (run_off[v + 1] - q0 == c) {
      head = false;
      for (uint64_t j = 0; j < c && !head; ++j) head = run_genome[r0 + j] != run_genome[q0 + j];
    }
  }
  head_secs[i] = head ? (uint32_t)((c + MLIST_SECTOR - 1) / MLIST_SECTOR) : 0u;
}

__global__ void set_assign(const uint32_t* __restrict__ vals, uint64_t n, const uint32_t* __restrict__ head_secs,
                           const uint64_t* __restrict__ sec_off, const uint64_t* __restrict__ run_off,
                           const uint32_t* __restrict__ run_genome, uint64_t* __restrict__ msec_off,
                           uint32_t* __restrict__ mlist) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t u = vals[i], r0 = run_off[u], c = run_off[u + 1] - r0;
  const uint64_t secs = (c + MLIST_SECTOR - 1) / MLIST_SECTOR;
  if (head_secs[i]) {
    msec_off[u] = sec_off[i];
    uint32_t* dst = mlist + sec_off[i] * MLIST_SECTOR;
    for (uint64_t j = 0; j < c; ++j) dst[j] = run_genome[r0 + j] | (j + 1 == c ? LIST_END : 0u);
  } else {
    msec_off[u] = sec_off[i] - secs;   // sec_off is an exclusive scan: the set of the nearest head before i ends at sec_off[i]
  }
}

// value field of a distinct k-mer with c genomes rg[0..c) (ascending); see TableView in common.cuh
__device__ __forceinline__ uint64_t entry_value(const TableView& t, uint64_t c, const uint32_t* __restrict__ rg, uint64_t msec) {
  const uint32_t kshift = t.val_bits - 3;
  if (c == 1) return ((uint64_t)KIND_SPECIFIC << kshift) | rg[0];
  if (c <= t.n_inline) {
    uint64_t v = 0;
    for (uint32_t i = 0; i < t.n_inline; ++i) v |= (uint64_t)rg[i < c ? i : c - 1] << (i * t.gbits);
    return ((uint64_t)KIND_INLINE << kshift) | v;
  }
  return ((uint64_t)KIND_MLIST << kshift) | msec;
}

// One thread per distinct k-mer: un-hash the CSR key, find its minimizer, and claim the first free slot of its bucket
// with atomicCAS -- in the home block or, when that bucket is full, in the same bucket of the next blocks of its digit,
// setting CONT on the last slot of every bucket it passes (readers only go on from a full bucket that has CONT).  Slots
// of a bucket fill in order, so "last slot taken" means "bucket full".  K-mers that find no free slot within CHAIN_LEN
// blocks are collected as {raw key, value} pairs for the stash (the last bucket of their chain then has CONT set,
// which sends readers there).
__global__ void table_insert(const uint64_t* __restrict__ ukeys, const uint64_t* __restrict__ run_off,
                             const uint32_t* __restrict__ run_genome, const uint64_t* __restrict__ msec_off, uint64_t msec_base,
                             uint64_t U, TableView t, MixParams mix, unsigned long long* __restrict__ slots,
                             unsigned int* __restrict__ ovf_count, unsigned long long* __restrict__ ovf_pairs, uint32_t ovf_cap) {
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  const uint64_t raw = unmix_key(ukeys[u], mix);
  const uint32_t kmask = (t.k >= 32) ? 0xFFFFFFFFu : ((1u << t.k) - 1);
  const uint32_t lo = (uint32_t)raw & kmask, hi = (uint32_t)(raw >> t.k) & kmask;
  uint32_t mh, p;
  kmer_minimizer(t, lo, hi, &mh, &p);
  const SlotAddr a = slot_addr(t, lo, hi, mh, p);
  const uint64_t r0 = run_off[u], c = run_off[u + 1] - r0;
  const unsigned long long value = entry_value(t, c, run_genome + r0, c > t.n_inline ? msec_off[u] + msec_base : 0);
  const unsigned long long cont_bit = 1ULL << (t.val_bits - 1);
  for (uint32_t d = 0; d < CHAIN_LEN; ++d) {
    unsigned long long* b = slots + (chain_block(t, a.block, d, mh) * BLOCK_BUCKETS + a.bucket) * BUCKET_SLOTS;
    const unsigned long long word = ((a.tag | ((uint64_t)d << t.hi_bits)) << t.val_bits) | value;
    for (uint32_t i = 0; i < BUCKET_SLOTS; ++i)
      if (atomicCAS(b + i, (unsigned long long)EMPTY64, word) == EMPTY64) return;
    atomicOr(b + BUCKET_SLOTS - 1, cont_bit);
  }
  uint32_t at = atomicAdd(ovf_count, 1u);
  if (at < ovf_cap) { ovf_pairs[2 * (size_t)at] = raw; ovf_pairs[2 * (size_t)at + 1] = value; }
}

__global__ void stash_insert(const unsigned long long* __restrict__ pairs, uint32_t n, unsigned long long* __restrict__ stash /* {raw key, value} */,
                             uint64_t stash_mask) {
  uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n) return;
  const unsigned long long raw = pairs[2 * (size_t)i0], value = pairs[2 * (size_t)i0 + 1];
  uint64_t i = stash_slot(raw);
  for (;;) {
    i &= stash_mask;
    if (atomicCAS(&stash[2 * i], (unsigned long long)EMPTY64, raw) == EMPTY64) {
      stash[2 * i + 1] = value;
      return;
    }
    ++i;
  }
}

}  // namespace

// ===========================================================================
// geometry
// ===========================================================================
int32_t table_geometry(int k_in, uint32_t G, uint64_t U, double load, uint32_t min_bpd, TableGeom* g) {
  // tag = 2(k-m) + CHAIN_BITS + hi_bits; the value field (val_bits = 64 - tag_bits) must hold the CONT bit + 2 kind
  // bits + a genome id (with the all-ones id left unused so that no slot equals EMPTY64) + any mlist sector index.
  const int k = k_in < 1 ? 1 : k_in;   // k <= 0: no k-mers at all; the geometry only has to be benign
  const uint32_t m = minimizer_len_for_k(k), tb = digit_bits_for_k(k), dshift = 2 * m - tb;
  const uint32_t gb = std::max(1u, ceil_log2_u64(G));
  const uint32_t need_spec = std::max(1u, ceil_log2_u64((uint64_t)G + 1));
  const int64_t hb_max = 64 - 3 - (int64_t)need_spec - 2 * ((int64_t)k - m) - CHAIN_BITS;
  if (hb_max < 0) { set_error("lookup table: genome ids do not fit (k=%d, G=%u)", k, G); return ST_UNSUPPORTED; }
  if (!(load > 0.01)) load = 0.01;
  if (load > 0.9) load = 0.9;
  const double slots = (double)std::max<uint64_t>(U, 1) / load;
  uint64_t bpd = (uint64_t)((slots / (BLOCK_BUCKETS * BUCKET_SLOTS)) / (double)(1u << tb)) + 1;
  bpd = std::max<uint64_t>(bpd, std::max<uint32_t>(min_bpd, 1));
  for (;;) {
    if (bpd > 0x7FFFFFFFull) { set_error("lookup table: too many blocks"); return ST_UNSUPPORTED; }
    const uint64_t per_block = ((1ULL << dshift) + bpd - 1) / bpd;   // hashes that can share a block: consecutive integers
    const uint32_t hb = ceil_log2_u64(per_block);
    if ((int64_t)hb <= hb_max) {
      g->bpd = (uint32_t)bpd; g->hi_bits = hb;
      g->tag_bits = 2 * ((uint32_t)k - m) + CHAIN_BITS + hb;
      g->val_bits = 64 - g->tag_bits;
      g->payload_bits = g->val_bits - 3;
      g->gbits = gb;
      uint32_t n_in = std::min<uint32_t>(4, g->payload_bits / gb);
      g->n_inline = n_in < 2 ? 1 : n_in;
      return ST_OK;
    }
    bpd *= 2;   // fewer hashes per block -> fewer tag bits
  }
}

double table_load_factor(int k, uint64_t U, size_t free_bytes) {
  // measured on config B (profiles/r02_load_sweep.jsonl): K4 takes 12.3 / 12.8 / 13.3 / 15.7 / 17.0 ms per 10^7 reads at load
  // 0.10 / 0.15 / 0.20 / 0.25 / 0.30 (80 / 53 / 40 / 32 / 26.7 bytes of table per k-mer): sparse while memory is plentiful,
  // dense when it is not
  if (const char* e = getenv("PA_TABLE_LOAD")) { const double v = atof(e); if (v > 0) return v; }
  if (const char* e = getenv("PA_TABLE_DENSE")) {   // tests: n > 0 doubles the load factor 0.2 n times (chains, CONT, stash), n < 0 halves it
    const int n = atoi(e);
    return n >= 0 ? std::min(0.9, 0.2 * (double)(1u << std::min(n, 8))) : 0.2 / (double)(1u << std::min(-n, 8));
  }
  const double bytes_per_slot = 8.0;
  double load = 0.10;
  // up to a quarter of the free device memory for a sparse table, up to half of it before the table gets denser than 1/5
  // (config E: 3.35e9 k-mers end at 0.32 = 84 GB on a 180-GB device; 0.42 is the densest the chains stay short at)
  while (load < 0.199 && (double)U / load * bytes_per_slot > 0.25 * (double)free_bytes) load += 0.02;
  while (load < 0.42 && (double)U / load * bytes_per_slot > 0.5 * (double)free_bytes) load += 0.02;
  (void)k;
  return load;
}

void apply_geometry(Index& ix, const TableGeom& g) {
  ix.min_len = minimizer_len_for_k(ix.k < 1 ? 1 : ix.k);
  ix.bpd = g.bpd; ix.hi_bits = g.hi_bits; ix.tag_bits = g.tag_bits; ix.val_bits = g.val_bits;
  ix.gbits = g.gbits; ix.n_inline = g.n_inline;
}

// ===========================================================================
// genome sets
// ===========================================================================
int32_t build_genome_sets(cudaStream_t s, const CsrView& c, uint32_t n_inline, GenomeSets* out) {
  const uint64_t U = c.U;
  out->n_msec = 0;
  out->msec_off.release(); out->mlist.release();
  DevBuf tile_sums, d_total, set_ka, set_kb, set_va, set_vb, set_tmp, head_secs, sec_off;
  PA_TRY(d_total.alloc(16));
  uint64_t n_msec = 0, n_long = 0;
  uint32_t* set_vals = nullptr;   // k-mers with long lists, sorted by set
  int dev = 0, sms = 148;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned ugrid = (unsigned)std::min<uint64_t>(grid_for(U, 256), (uint64_t)sms * 16);
  if (U) {
    PA_CUDA(cudaMemsetAsync(d_total.p, 0, 16, s));
    long_count<<<ugrid, 256, 0, s>>>(c.run_off, U, n_inline, d_total.as<unsigned long long>());
    PA_CUDA(cudaMemcpyAsync(&n_long, d_total.p, 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  if (n_long) {
    if (n_long >= 0xFFFFFFFFull) { set_error("genome sets: too many k-mers with long lists"); return ST_UNSUPPORTED; }
    PA_TRY(out->msec_off.alloc((U + 1) * 8));
    PA_TRY(set_ka.alloc(n_long * 8)); PA_TRY(set_kb.alloc(n_long * 8)); PA_TRY(set_va.alloc(n_long * 4)); PA_TRY(set_vb.alloc(n_long * 4));
    PA_TRY(set_tmp.alloc(radix_sort_temp_bytes(n_long)));
    PA_TRY(head_secs.alloc(n_long * 4)); PA_TRY(sec_off.alloc(n_long * 8));
    PA_TRY(tile_sums.alloc((scan_tiles(n_long) + 1) * 8));
    long_collect<<<ugrid, 256, 0, s>>>(c.run_off, c.run_genome, U, n_inline, d_total.as<unsigned long long>() + 1,
                                       set_ka.as<uint64_t>(), set_va.as<uint32_t>());
    int in_b = 0;
    PA_TRY(radix_sort_pairs_hashed(set_ka.as<uint64_t>(), set_va.as<uint32_t>(), set_kb.as<uint64_t>(), set_vb.as<uint32_t>(), n_long, 64,
                                   set_tmp.p, set_tmp.bytes, s, &in_b));
    const uint64_t* skeys = in_b ? set_kb.as<uint64_t>() : set_ka.as<uint64_t>();
    set_vals = in_b ? set_vb.as<uint32_t>() : set_va.as<uint32_t>();
    set_heads<<<grid_for(n_long, 256), 256, 0, s>>>(skeys, set_vals, n_long, c.run_off, c.run_genome, head_secs.as<uint32_t>());
    PA_TRY(exclusive_scan_u32(head_secs.as<uint32_t>(), sec_off.as<uint64_t>(), n_long, tile_sums.as<uint64_t>(), d_total.as<uint64_t>(), s));
    PA_CUDA(cudaMemcpyAsync(&n_msec, d_total.p, 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  out->n_msec = n_msec;
  PA_TRY(out->mlist.alloc(std::max<uint64_t>(n_msec, 1) * 32));
  PA_CUDA(cudaMemsetAsync(out->mlist.p, 0xFF, out->mlist.bytes, s));
  if (n_long)
    set_assign<<<grid_for(n_long, 256), 256, 0, s>>>(set_vals, n_long, head_secs.as<uint32_t>(), sec_off.as<uint64_t>(), c.run_off,
                                                     c.run_genome, out->msec_off.as<uint64_t>(), out->mlist.as<uint32_t>());
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));   // the scratch above is released when this returns
  return ST_OK;
}

// ===========================================================================
// slots and stash
// ===========================================================================
int32_t table_insert_csr(Index& ix, const CsrView& c, const GenomeSets& sets, uint64_t msec_base, std::vector<uint64_t>* overflow) {
  if (!c.U) return ST_OK;
  cudaStream_t s = ix.stream;
  DevBuf ovf_count, ovf_pairs;
  PA_TRY(ovf_count.alloc(4));
  const uint32_t ovf_cap = (uint32_t)std::min<uint64_t>(c.U / 64 + 65536, 0x0FFFFFF0ull);
  PA_TRY(ovf_pairs.alloc((size_t)ovf_cap * 16));
  PA_CUDA(cudaMemsetAsync(ovf_count.p, 0, 4, s));
  table_insert<<<grid_for(c.U, 256), 256, 0, s>>>(c.ukeys, c.run_off, c.run_genome, sets.msec_off.as<uint64_t>(), msec_base, c.U,
                                                  ix.view(), ix.mix, ix.slots.as<unsigned long long>(), ovf_count.as<unsigned int>(),
                                                  ovf_pairs.as<unsigned long long>(), ovf_cap);
  PA_CUDA(cudaGetLastError());
  uint32_t n_ovf = 0;
  PA_CUDA(cudaMemcpyAsync(&n_ovf, ovf_count.p, 4, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  if (n_ovf > ovf_cap) { set_error("lookup table: %u k-mers overflowed their chains (table overloaded)", n_ovf); return ST_CAPACITY; }
  if (n_ovf) {
    const size_t at = overflow->size();
    overflow->resize(at + 2 * (size_t)n_ovf);
    PA_CUDA(cudaMemcpyAsync(overflow->data() + at, ovf_pairs.p, (size_t)n_ovf * 16, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  return ST_OK;
}

int32_t table_build_stash(Index& ix, const std::vector<uint64_t>& pairs) {
  cudaStream_t s = ix.stream;
  ix.stash.release();
  ix.stash_cap = 0;
  const uint64_t n = pairs.size() / 2;
  if (n > 0xFFFFFFF0ull) { set_error("stash: too many entries"); return ST_UNSUPPORTED; }
  ix.stash_count = (uint32_t)n;
  if (!n) return ST_OK;
  uint64_t cap = 16;
  while (cap < n * 2) cap <<= 1;
  ix.stash_cap = cap;
  PA_TRY(ix.stash.alloc(cap * 16));
  PA_CUDA(cudaMemsetAsync(ix.stash.p, 0xFF, cap * 16, s));
  DevBuf d_pairs;
  PA_TRY(d_pairs.alloc(n * 16));
  PA_CUDA(cudaMemcpyAsync(d_pairs.p, pairs.data(), n * 16, cudaMemcpyHostToDevice, s));
  stash_insert<<<grid_for(n, 256), 256, 0, s>>>(d_pairs.as<unsigned long long>(), (uint32_t)n, ix.stash.as<unsigned long long>(), cap - 1);
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t index_build_tables(Index& ix) {
  cudaStream_t s = ix.stream;
  const uint64_t U = ix.n_keys;
  ix.mix = mix_params_for_k(ix.k);
  ix.slots.release(); ix.stash.release(); ix.mlist.release();
  ix.stash_cap = 0; ix.stash_count = 0; ix.n_msectors = 0;
  size_t free_b = 0, total_b = 0;
  PA_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += cache_held();   // buffers kept for reuse are memory a build may have
  const double load = table_load_factor(ix.k, U, free_b);
  const CsrView csr{ix.ukeys.as<uint64_t>(), ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), U};
  TableGeom g;
  GenomeSets sets;
  uint32_t min_bpd = 1;
  for (;;) {
    PA_TRY(table_geometry(ix.k, ix.n_genomes, U, load, min_bpd, &g));
    PA_TRY(build_genome_sets(s, csr, g.n_inline, &sets));
    if (ceil_log2_u64(sets.n_msec + 1) <= g.payload_bits) break;
    if (g.bpd >= 0x40000000u) { set_error("lookup table: list references do not fit (k=%d)", ix.k); return ST_UNSUPPORTED; }
    min_bpd = g.bpd * 2;    // a bigger table has shorter tags, i.e. longer payloads
  }
  std::vector<uint64_t> overflow;
  for (;;) {
    apply_geometry(ix, g);
    const uint64_t bytes = ix.n_blocks() * (BLOCK_BUCKETS * BUCKET_SLOTS * 8);
    PA_TRY(ix.slots.alloc(bytes));
    PA_CUDA(cudaMemsetAsync(ix.slots.p, 0xFF, bytes, s));
    overflow.clear();
    const int32_t st = table_insert_csr(ix, csr, sets, 0, &overflow);
    if (st == ST_OK) break;
    if (st != ST_CAPACITY || g.bpd >= 0x40000000u) return st;
    // pathological: grow the table instead of the stash (the genome sets stay valid while n_inline does not change)
    const uint32_t n_in = g.n_inline;
    PA_TRY(table_geometry(ix.k, ix.n_genomes, U, load, g.bpd * 2, &g));
    if (g.n_inline != n_in) PA_TRY(build_genome_sets(s, csr, g.n_inline, &sets));
  }
  ix.n_msectors = sets.n_msec;
  ix.mlist.swap(sets.mlist);
  PA_TRY(table_build_stash(ix, overflow));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

}  // namespace pa
