// build.cu -- K1 (2-bit window encoder), K3 (run-length / CSR pass), lookup-table
// construction, and the CSR consumers (rank lookup, insertion order, EXTSIM
// statistics K5/K6, genome removal K7).
//
// Reference being replaced: KmerReference._build_kmer_mapping
// (/root/reference/src/kmer.py:135-150), get_kmer_references (kmer.py:292-298),
// _compute_genome_stats (kmer.py:152-177), the intersections of
// _apply_greedy_filter (kmer.py:206-207) and _remove_filtered_genomes_from_kmers
// (kmer.py:232-243).  All kernels are integer, HBM-bound streaming passes.
#include "index.cuh"
#include "scan.cuh"
#include "sort.cuh"
#include "table.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace pa {

// ===========================================================================
// generic scans
// ===========================================================================
namespace {

__global__ void __launch_bounds__(SCAN_THREADS) tile_sum_u32(const uint32_t* __restrict__ in, uint64_t n,
                                                             uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t ws[SCAN_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  uint64_t acc = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
    if (i < n) acc += in[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += ws[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_apply_u32(const uint32_t* __restrict__ in,
                                                                    uint64_t* __restrict__ out, uint64_t n,
                                                                    const uint64_t* __restrict__ tile_excl) {
  __shared__ uint64_t ws[SCAN_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint64_t local = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    local += v[j];
  }
  uint64_t total;
  uint64_t excl = block_exclusive_scan<SCAN_THREADS, uint64_t>(local, total, ws) + tile_excl[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = excl;
    excl += v[j];
  }
}

}  // namespace

__global__ void scan_u64_single_block(uint64_t* data, uint64_t n, uint64_t* d_total) {
  __shared__ uint64_t ws[1024 / 32];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n; base += blockDim.x) {
    uint64_t i = base + threadIdx.x;
    uint64_t v = i < n ? data[i] : 0;
    uint64_t total;
    uint64_t excl = block_exclusive_scan<1024, uint64_t>(v, total, ws);
    uint64_t c = carry;
    if (i < n) data[i] = c + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && d_total) *d_total = carry;
}

int32_t exclusive_scan_u32(const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_tile_sums, uint64_t* d_total,
                           cudaStream_t s) {
  if (n == 0) { PA_CUDA(cudaMemsetAsync(d_total, 0, 8, s)); return ST_OK; }
  uint64_t tiles = scan_tiles(n);
  tile_sum_u32<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, n, d_tile_sums);
  scan_u64_single_block<<<1, 1024, 0, s>>>(d_tile_sums, tiles, d_total);
  tile_scan_apply_u32<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, d_out, n, d_tile_sums);
  PA_CUDA(cudaGetLastError());
  return ST_OK;
}

namespace {

// ===========================================================================
// K1: window encoder.  One tile = ENC_TILE window starts.  Phase 1 packs the
// tile's bases into bit planes (low code bit, high code bit, "not ACGT",
// "last base of a genome") with coalesced 16-byte loads; phase 2 extracts each
// window with funnel shifts and writes key / position fully coalesced.
// A window is valid iff it has k ACGT bases inside one genome (kmer.py:93-94
// bounds the window to the genome, kmer.py:145 drops windows containing N).
// ===========================================================================
constexpr int ENC_THREADS = 256;
constexpr int ENC_TILE = 4096;
constexpr int ENC_WORDS = ENC_TILE / 32 + 2;

__global__ void __launch_bounds__(ENC_THREADS)
encode_windows(const uint8_t* __restrict__ bases, uint64_t total, uint64_t pos0, const uint64_t* __restrict__ genome_off, uint32_t G,
               int k, MixParams mix, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* __restrict__ n_valid,
               unsigned int* __restrict__ bad_flag, EncodeOpts opt) {
  // bases[0 .. total) is the slice of the concatenated genomes that starts at global base position pos0 (a multi-GPU
  // build gives every rank a run of genomes, a streamed build walks them chunk by chunk); keys / vals are indexed like
  // bases, vals hold GLOBAL positions (or genome indices: opt.vals_are_genomes).  Only windows starting before
  // opt.emit_total are emitted: the k - 1 bases beyond it are the overlap into the next chunk.
  __shared__ uint32_t s_lo[ENC_WORDS], s_hi[ENC_WORDS], s_inv[ENC_WORDS], s_brk[ENC_WORDS];
  __shared__ uint32_t s_cnt[ENC_THREADS / 32];
  __shared__ uint32_t s_g0;
  __shared__ uint32_t s_ord[ENC_TILE + MINIMIZER_MAX];   // order of the m-mer at every position of the tile (partitioned builds)
  const uint64_t tile_base = (uint64_t)blockIdx.x * ENC_TILE;
  const int tid = threadIdx.x;
  bool bad = false;
  for (int w = tid; w < ENC_WORDS; w += ENC_THREADS) {
    uint64_t b0 = tile_base + (uint64_t)w * 32;
    uint32_t lo = 0, hi = 0, inv = 0;
    uint8_t c[32];
    if (b0 + 32 <= total) {
      const uint4* p = reinterpret_cast<const uint4*>(bases + b0);  // tile_base and the buffer are 32-byte aligned
      *reinterpret_cast<uint4*>(c) = p[0];
      *reinterpret_cast<uint4*>(c + 16) = p[1];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) c[i] = (b0 + i < total) ? bases[b0 + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      uint32_t ch = c[i];
      bool ok = is_acgt(ch);
      uint32_t code = base_code(ch);
      lo |= (code & 1u) << i;
      hi |= (code >> 1) << i;
      inv |= (ok ? 0u : 1u) << i;
      if (!ok && ch != 'N' && b0 + i < total) bad = true;
    }
    s_lo[w] = lo; s_hi[w] = hi; s_inv[w] = inv; s_brk[w] = 0;
  }
  if (bad) atomicOr(bad_flag, 1u);
  __syncthreads();
  // mark the last base of every genome that ends inside this tile's span
  {
    const uint64_t span_beg = pos0 + tile_base;
    const uint64_t span_end = pos0 + min(total, tile_base + (uint64_t)ENC_WORDS * 32);
    if (tile_base < total) {
      uint32_t g0 = genome_of(genome_off, G, span_beg);
      if (tid == 0) s_g0 = g0;
      for (uint32_t g = g0 + tid; g < G; g += ENC_THREADS) {
        uint64_t beg = genome_off[g], end = genome_off[g + 1];
        if (beg >= span_end) break;
        if (end == beg || end > span_end || end - 1 < span_beg) continue;
        uint64_t r = end - 1 - span_beg;
        atomicOr(&s_brk[r >> 5], 1u << (r & 31));
      }
    }
  }
  __syncthreads();
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1);
  TableView mt;                   // only the minimizer fields are used (owner of a record = f(digit of its minimizer))
  minimizer_params(mt, k);
  const uint32_t tb = digit_bits_for_k(k);
  if (opt.owner) {
    // the order of every m-mer once (not once per window it is a candidate of): a window then takes the minimum over its
    // w consecutive entries -- (order << 4 | offset) makes the leftmost candidate win ties, like kmer_minimizer
    for (int q = tid; q < ENC_TILE + (int)mt.w - 1; q += ENC_THREADS) {
      const uint32_t w = q >> 5, b = q & 31;
      const uint32_t xl = __funnelshift_r(s_lo[w], s_lo[w + 1], b) & mt.mmask;
      const uint32_t xh = __funnelshift_r(s_hi[w], s_hi[w + 1], b) & mt.mmask;
      s_ord[q] = mmer_order((xh << mt.m) | xl, mt) << 4;
    }
    __syncthreads();
  }
  uint32_t cnt = 0;
  uint32_t g_walk = (tile_base < total) ? s_g0 : 0;
#pragma unroll 4
  for (int j = 0; j < ENC_TILE / ENC_THREADS; ++j) {
    uint32_t s = j * ENC_THREADS + tid;
    uint64_t gpos = tile_base + s;
    if (gpos >= opt.emit_total) break;
    uint32_t w = s >> 5, b = s & 31;
    uint32_t lo = __funnelshift_r(s_lo[w], s_lo[w + 1], b) & kmask;
    uint32_t hi = __funnelshift_r(s_hi[w], s_hi[w + 1], b) & kmask;
    uint32_t inv = __funnelshift_r(s_inv[w], s_inv[w + 1], b) & kmask;
    uint32_t brk = __funnelshift_r(s_brk[w], s_brk[w + 1], b) & (kmask >> 1);
    bool valid = (inv | brk) == 0 && gpos + (uint64_t)k <= total;
    // the bijectively hashed k-mer is the sort key: equal k-mers still group
    keys[gpos] = valid ? mix_key(((uint64_t)hi << k) | lo, mix) : SENTINEL_KEY;
    if (opt.vals_are_genomes) {
      while (g_walk + 1 < G && genome_off[g_walk + 1] <= pos0 + gpos) ++g_walk;
      vals[gpos] = g_walk;
    } else {
      vals[gpos] = (uint32_t)(pos0 + gpos);
    }
    if (opt.owner) {
      uint32_t own = 255;
      if (valid) {
        uint32_t best = s_ord[s];
        for (uint32_t c = 1; c < mt.w; ++c) best = min(best, s_ord[s + c] + c);
        const uint32_t mp = best & 15u;
        const uint32_t mh = hash_from_order(best >> 4, lo >> mp, mt);
        const uint32_t part = (uint32_t)(((uint64_t)(mh >> mt.dshift) * opt.n_parts) >> tb);
        if (part % opt.n_rounds == opt.round) own = part / opt.n_rounds;
      }
      opt.owner[gpos] = (uint8_t)own;
    }
    cnt += valid;
  }
  cnt = warp_sum(cnt);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int w = 0; w < ENC_THREADS / 32; ++w) t += s_cnt[w];
    if (t) atomicAdd(n_valid, (unsigned long long)t);
  }
}

// ===========================================================================
// K3: run-length pass over the sorted (key, global position) records.
//   key head  : first record of a distinct k-mer
//   run head  : first record of a (k-mer, genome) pair
// rle_count -> per-tile head counts; scan; rle_scatter writes the CSR.
// ===========================================================================
constexpr int RLE_THREADS = 256;
constexpr int RLE_ITEMS = 8;
constexpr int RLE_TILE = RLE_THREADS * RLE_ITEMS;

// genome of a global base position through the coarse map: gmap[pos >> shift] = genome of the first base of that
// block, then a short forward walk (a block rarely holds a genome boundary) instead of a binary search per record
struct GenomeMap {
  const uint32_t* map;
  uint32_t shift;
};
__device__ __forceinline__ uint32_t genome_at(const GenomeMap& gm, const uint64_t* __restrict__ off, uint64_t pos) {
  uint32_t g = gm.map[pos >> gm.shift];
  while (off[g + 1] <= pos) ++g;
  return g;
}

__global__ void genome_map_fill(const uint64_t* __restrict__ off, uint32_t G, uint64_t total, uint32_t shift, uint32_t* __restrict__ map,
                                uint64_t n_entries) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const uint64_t pos = i << shift;
  map[i] = pos < total ? genome_of(off, G, pos) : (G ? G - 1 : 0);
}

// Head flags of one warp's 256 consecutive records, striped: lane l holds records warp_base + 32 j + l (j = 0..7), so
// every load and -- after compaction -- every store of a warp instruction touches consecutive addresses (the first
// version gave each thread 8 consecutive records: its stores were 64 bytes apart and the kernel was bound by L2
// transactions, profiles/r01_rle_scatter_ncu.json).  kh / rh = ballot masks of key heads / (key, genome) heads.
template <bool GEN>   // GEN: vals hold genome indices already (table-only builds), else global positions
__device__ __forceinline__ void rle_flags(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                          const uint64_t* __restrict__ genome_off, GenomeMap G, uint64_t warp_base, uint64_t n,
                                          uint32_t lane, uint64_t (&key)[RLE_ITEMS], uint32_t (&gen)[RLE_ITEMS],
                                          uint32_t (&kh)[RLE_ITEMS], uint32_t (&rh)[RLE_ITEMS]) {
  uint64_t carry_key = 0; uint32_t carry_gen = 0;   // record just before the current row (lane 31 of the previous row)
  bool have_carry = false;
  if (warp_base > 0 && warp_base < n) {
    carry_key = keys[warp_base - 1];
    carry_gen = GEN ? vals[warp_base - 1] : genome_at(G, genome_off, vals[warp_base - 1]);
    have_carry = true;
  }
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; ++j) {
    const uint64_t i = warp_base + (uint64_t)j * 32 + lane;
    const bool ok = i < n;
    key[j] = ok ? keys[i] : 0;
    gen[j] = ok ? (GEN ? vals[i] : genome_at(G, genome_off, vals[i])) : 0;
    uint64_t pk = __shfl_up_sync(0xffffffffu, key[j], 1);
    uint32_t pg = __shfl_up_sync(0xffffffffu, gen[j], 1);
    bool have_prev = true;
    if (lane == 0) { pk = carry_key; pg = carry_gen; have_prev = have_carry; }
    const bool k_head = ok && (!have_prev || key[j] != pk);
    const bool r_head = ok && (k_head || gen[j] != pg);
    kh[j] = __ballot_sync(0xffffffffu, k_head);
    rh[j] = __ballot_sync(0xffffffffu, r_head);
    carry_key = __shfl_sync(0xffffffffu, key[j], 31);
    carry_gen = __shfl_sync(0xffffffffu, gen[j], 31);
    have_carry = true;
  }
}

template <bool GEN>
__global__ void __launch_bounds__(RLE_THREADS)
rle_count(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint64_t* __restrict__ genome_off,
          GenomeMap G, uint64_t n, uint64_t* __restrict__ tile_keys, uint64_t* __restrict__ tile_runs) {
  __shared__ uint32_t ws[2][RLE_THREADS / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_base = (uint64_t)blockIdx.x * RLE_TILE + (uint64_t)warp * (32 * RLE_ITEMS);
  uint64_t key[RLE_ITEMS]; uint32_t gen[RLE_ITEMS], kh[RLE_ITEMS], rh[RLE_ITEMS];
  rle_flags<GEN>(keys, vals, genome_off, G, warp_base, n, lane, key, gen, kh, rh);
  uint32_t a = 0, b = 0;
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; ++j) { a += __popc(kh[j]); b += __popc(rh[j]); }
  if (lane == 0) { ws[0][warp] = a; ws[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t ta = 0, tb = 0;
    for (int w = 0; w < RLE_THREADS / 32; ++w) { ta += ws[0][w]; tb += ws[1][w]; }
    tile_keys[blockIdx.x] = ta; tile_runs[blockIdx.x] = tb;
  }
}

template <bool GEN>
__global__ void __launch_bounds__(RLE_THREADS)
rle_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint64_t* __restrict__ genome_off,
            GenomeMap G, uint64_t n, const uint64_t* __restrict__ tile_keys, const uint64_t* __restrict__ tile_runs,
            uint64_t* __restrict__ ukeys, uint64_t* __restrict__ run_off, uint32_t* __restrict__ run_genome,
            uint64_t* __restrict__ pos_off, uint32_t* __restrict__ pos) {
  __shared__ uint32_t ws[2][RLE_THREADS / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_base = (uint64_t)blockIdx.x * RLE_TILE + (uint64_t)warp * (32 * RLE_ITEMS);
  uint64_t key[RLE_ITEMS]; uint32_t gen[RLE_ITEMS], kh[RLE_ITEMS], rh[RLE_ITEMS];
  rle_flags<GEN>(keys, vals, genome_off, G, warp_base, n, lane, key, gen, kh, rh);
  uint32_t a = 0, b = 0;
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; ++j) { a += __popc(kh[j]); b += __popc(rh[j]); }
  if (lane == 0) { ws[0][warp] = a; ws[1][warp] = b; }
  __syncthreads();
  uint64_t kr = tile_keys[blockIdx.x], rr = tile_runs[blockIdx.x];   // first key / run index of this warp
  for (uint32_t w = 0; w < warp; ++w) { kr += ws[0][w]; rr += ws[1][w]; }
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int j = 0; j < RLE_ITEMS; ++j) {
    const uint64_t i = warp_base + (uint64_t)j * 32 + lane;
    if (i < n) {
      if ((rh[j] >> lane) & 1) {
        const uint64_t r = rr + __popc(rh[j] & lt);
        if ((kh[j] >> lane) & 1) { const uint64_t q = kr + __popc(kh[j] & lt); ukeys[q] = key[j]; run_off[q] = r; }
        run_genome[r] = gen[j];
        if (!GEN) pos_off[r] = i;
      }
      if (!GEN) pos[i] = (uint32_t)((uint64_t)vals[i] - genome_off[gen[j]]);
    }
    kr += __popc(kh[j]);
    rr += __popc(rh[j]);
  }
}

__global__ void iota_u32(uint32_t* v, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

__global__ void set_csr_tails(uint64_t* run_off, uint64_t U, uint64_t R, uint64_t* pos_off, uint64_t N) {
  run_off[U] = R;
  if (pos_off) pos_off[R] = N;
}

// ===========================================================================
// CSR consumers
// ===========================================================================
__global__ void lookup_ranks(const uint64_t* __restrict__ ukeys, uint64_t U, const uint64_t* __restrict__ q, uint64_t n,
                             uint64_t* __restrict__ rank) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = q[i];
  uint64_t lo = 0, hi = U;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
  }
  rank[i] = (key != SENTINEL_KEY && lo < U && ukeys[lo] == key) ? lo : LOOKUP_MISS;
}

// Content checksum of a CSR: four sums modulo 2^64 over {k-mers, (k-mer, genome) pairs, (k-mer, genome, position)
// triples, k-mer count}.  Order-independent and additive over disjoint key sets, so the checksums of the partitions
// of a multi-GPU build add up to the checksum of the single-GPU index exactly when the contents agree.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__global__ void __launch_bounds__(256)
csr_checksum_kernel(const uint64_t* __restrict__ ukeys, const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome,
                    const uint64_t* __restrict__ pos_off, const uint32_t* __restrict__ pos, uint64_t U, int with_pos,
                    unsigned long long* __restrict__ out) {
  unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < U; u += stride) {
    const unsigned long long key = ukeys[u];
    a0 += mix64(key ^ 0x243F6A8885A308D3ULL);
    a3 += 1;
    for (uint64_t r = run_off[u]; r < run_off[u + 1]; ++r) {
      const unsigned long long g = run_genome[r];
      a1 += mix64(key * 3 + g);
      if (with_pos)
        for (uint64_t q = pos_off[r]; q < pos_off[r + 1]; ++q) a2 += mix64(key * 5 + ((g << 32) | pos[q]));
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, a0); atomicAdd(out + 1, a1); atomicAdd(out + 2, a2); atomicAdd(out + 3, a3); }
}

// first occurrence of each distinct k-mer as a global base position: the dict
// insertion order of kmer.py:146-147 is the ascending order of this value.
__global__ void first_occurrence(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome,
                                 const uint64_t* __restrict__ pos_off, const uint32_t* __restrict__ pos,
                                 const uint64_t* __restrict__ genome_off, uint64_t U, uint64_t* __restrict__ keys) {
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  uint64_t r = run_off[u];
  keys[u] = genome_off[run_genome[r]] + pos[pos_off[r]];
}

// K5: per identifier class, number of distinct k-mers holding it / holding it alone.  Classes are few, so every
// block counts in shared memory first (n_groups <= EXT_SMEM_GROUPS) and adds its totals to the global counters once;
// plain global atomics on a thousand addresses serialise in L2 (config D: 43 ms -> a few ms).
constexpr uint32_t EXT_SMEM_GROUPS = 4096;
__global__ void extsim_stats_kernel(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome, uint64_t U,
                                    const uint32_t* __restrict__ group, uint32_t n_groups, int dedupe,
                                    unsigned long long* __restrict__ total, unsigned long long* __restrict__ unique) {
  __shared__ uint32_t s_total[EXT_SMEM_GROUPS], s_unique[EXT_SMEM_GROUPS];
  const bool in_smem = n_groups <= EXT_SMEM_GROUPS;
  if (in_smem) {
    for (uint32_t i = threadIdx.x; i < n_groups; i += blockDim.x) { s_total[i] = 0; s_unique[i] = 0; }
    __syncthreads();
  }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < U; u += stride) {
    uint64_t r0 = run_off[u], c = run_off[u + 1] - r0;
    for (uint64_t j = 0; j < c; ++j) {
      uint32_t gr = group[run_genome[r0 + j]];
      bool first = true;
      if (dedupe)
        for (uint64_t i = 0; i < j && first; ++i) first = group[run_genome[r0 + i]] != gr;
      if (!first) continue;
      if (in_smem) {
        atomicAdd(&s_total[gr], 1u);
        if (c == 1) atomicAdd(&s_unique[gr], 1u);
      } else {
        atomicAdd(&total[gr], 1ULL);
        if (c == 1) atomicAdd(&unique[gr], 1ULL);
      }
    }
  }
  if (in_smem) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_groups; i += blockDim.x) {
      if (s_total[i]) atomicAdd(&total[i], (unsigned long long)s_total[i]);
      if (s_unique[i]) atomicAdd(&unique[i], (unsigned long long)s_unique[i]);
    }
  }
}

// K6: inter[a][b] += 1 for every pair of classes sharing a distinct k-mer (diagonal = total).
//
// Counting pair by pair costs sum(c^2) atomics on a few thousand hot cells (config D: 6x10^9; measured 45-67 ms however
// they are issued: warp or thread per k-mer, 32 or 64 bits, private copies of the matrix).  Near-duplicate genomes make
// the same class list recur millions of times, so the pass counts *lists* first: (a) every k-mer adds one to the
// counter of its list in a hash table keyed by a 64-bit hash of the list (one atomic per k-mer instead of c^2);
// (b) every k-mer compares its list with the table entry's representative - a hash collision between different lists
// sets a flag and the whole pass is redone pair by pair, so the result never depends on the hash; (c) one thread per
// table entry adds the entry's count to the cells of its pairs.  K-mers whose probe window is full (more distinct lists
// than the table holds) and lists longer than EXT_PAIR_SMALL (taken by the whole warp) are counted pair by pair.
// Cells are (min, max) in 32 bits (a count is at most n_keys < 2^32); extsim_pairwise_mirror widens and mirrors.
constexpr uint32_t EXT_PAIR_SMALL = 16;
constexpr uint32_t EXT_PAIR_THREADS = 256;
constexpr uint32_t EXT_SET_PROBES = 16;

struct SetTable {
  unsigned long long* key;   // 0 = empty
  uint32_t* rep;             // smallest k-mer index that holds the list
  uint32_t* count;
  uint32_t mask;             // entries - 1 (0: no table, count pair by pair)
  uint32_t weak_hash;        // tests: every list hashes alike, which must end in the pair-by-pair redo
};

// class list of k-mer [r0, r0+c) into the caller's shared-memory column; returns its length
__device__ __forceinline__ uint32_t class_list(const uint32_t* __restrict__ run_genome, const uint32_t* __restrict__ group,
                                               uint64_t r0, uint32_t c, int dedupe, uint32_t (*col)[EXT_PAIR_THREADS]) {
  uint32_t n = 0;
  for (uint32_t j = 0; j < c; ++j) {
    const uint32_t gr = group[run_genome[r0 + j]];
    bool first = true;
    if (dedupe)
      for (uint32_t i = 0; i < n && first; ++i) first = col[i][threadIdx.x] != gr;
    if (first) col[n++][threadIdx.x] = gr;
  }
  return n;
}

__device__ __forceinline__ unsigned long long class_list_hash(const SetTable& t, uint32_t n, uint32_t (*col)[EXT_PAIR_THREADS]) {
  if (t.weak_hash) return 1;
  unsigned long long h = 0x243F6A8885A308D3ULL + n;
  for (uint32_t i = 0; i < n; ++i) {
    h = (h ^ col[i][threadIdx.x]) * 0x9E3779B97F4A7C15ULL;
    h ^= h >> 29;
  }
  return h ? h : 1;
}

__device__ __forceinline__ void add_pairs(uint32_t n, uint32_t (*col)[EXT_PAIR_THREADS], uint32_t n_groups, uint32_t by,
                                          uint32_t* __restrict__ tri) {
  for (uint32_t a = 0; a < n; ++a) {
    const uint32_t ga = col[a][threadIdx.x];
    for (uint32_t b = a; b < n; ++b) {
      const uint32_t gb = col[b][threadIdx.x];
      atomicAdd(&tri[(size_t)min(ga, gb) * n_groups + max(ga, gb)], by);
    }
  }
}

// slot of `h` in the probe window, claiming an empty one when `claim`; -1 if the window holds other lists only
__device__ __forceinline__ int64_t set_slot(const SetTable& t, unsigned long long h, bool claim) {
  uint32_t at = (uint32_t)(h >> 20) & t.mask;
  for (uint32_t i = 0; i < EXT_SET_PROBES; ++i, at = (at + 1) & t.mask) {
    unsigned long long cur = t.key[at];
    if (cur == 0 && claim) cur = atomicCAS(&t.key[at], 0ULL, h), cur = cur ? cur : h;
    if (cur == h) return at;
    if (cur == 0) return -1;   // (only without claim) never inserted
  }
  return -1;
}

__global__ void __launch_bounds__(EXT_PAIR_THREADS)
extsim_pairwise_kernel(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome, uint64_t U,
                       const uint32_t* __restrict__ group, uint32_t n_groups, int dedupe, SetTable sets,
                       uint32_t* __restrict__ tri) {
  __shared__ uint32_t s_class[EXT_PAIR_SMALL][EXT_PAIR_THREADS];
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < U; base += stride) {   // uniform per block
    const uint64_t u = base + threadIdx.x;
    uint64_t r0 = 0, c = 0;
    if (u < U) { r0 = run_off[u]; c = run_off[u + 1] - r0; }
    if (c <= EXT_PAIR_SMALL) {
      const uint32_t n = class_list(run_genome, group, r0, (uint32_t)c, dedupe, s_class);
      int64_t at = -1;
      if (n >= 2 && sets.mask) at = set_slot(sets, class_list_hash(sets, n, s_class), true);
      if (at >= 0) {
        if (sets.rep[at] > (uint32_t)u) atomicMin(&sets.rep[at], (uint32_t)u);
        atomicAdd(&sets.count[at], 1u);
      } else {
        add_pairs(n, s_class, n_groups, 1u, tri);
      }
    }
    unsigned big = __ballot_sync(0xFFFFFFFFu, c > EXT_PAIR_SMALL);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint64_t wr0 = __shfl_sync(0xFFFFFFFFu, r0, src), wc = __shfl_sync(0xFFFFFFFFu, c, src);
      for (uint64_t t = lane; t < wc * wc; t += 32) {
        uint64_t a = t / wc, b = t % wc;
        if (a > b) continue;
        uint32_t ga = group[run_genome[wr0 + a]], gb = group[run_genome[wr0 + b]];
        if (dedupe) {
          bool fa = true, fb = true;
          for (uint64_t i = 0; i < a && fa; ++i) fa = group[run_genome[wr0 + i]] != ga;
          for (uint64_t i = 0; i < b && fb; ++i) fb = group[run_genome[wr0 + i]] != gb;
          if (!fa || !fb) continue;
        }
        atomicAdd(&tri[(size_t)min(ga, gb) * n_groups + max(ga, gb)], 1u);
      }
    }
  }
}

// K6 (b): a k-mer counted through the table must hold exactly the representative's list.
__global__ void __launch_bounds__(EXT_PAIR_THREADS)
extsim_sets_verify(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome, uint64_t U,
                   const uint32_t* __restrict__ group, int dedupe, SetTable sets, uint32_t* __restrict__ collision) {
  __shared__ uint32_t s_class[EXT_PAIR_SMALL][EXT_PAIR_THREADS];
  const uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  const uint64_t r0 = run_off[u], c = run_off[u + 1] - r0;
  if (c < 2 || c > EXT_PAIR_SMALL) return;
  const uint32_t n = class_list(run_genome, group, r0, (uint32_t)c, dedupe, s_class);
  if (n < 2) return;
  const int64_t at = set_slot(sets, class_list_hash(sets, n, s_class), false);
  if (at < 0) return;                       // was counted pair by pair
  const uint32_t rep = sets.rep[at];
  if (rep == (uint32_t)u) return;
  // walk the representative's list; its distinct classes so far equal mine so far, so mine serve as its seen-set
  const uint64_t q0 = run_off[rep], qc = run_off[rep + 1] - q0;
  uint32_t m = 0;
  bool same = true;
  for (uint64_t j = 0; j < qc && same; ++j) {
    const uint32_t gr = group[run_genome[q0 + j]];
    bool first = true;
    if (dedupe)
      for (uint32_t i = 0; i < m && first; ++i) first = s_class[i][threadIdx.x] != gr;
    if (!first) continue;
    same = m < n && s_class[m][threadIdx.x] == gr;
    ++m;
  }
  if (!same || m != n) atomicOr(collision, 1u);
}

// K6 (c): one thread per table entry adds the entry's count to the cells of its pairs.
__global__ void __launch_bounds__(EXT_PAIR_THREADS)
extsim_sets_flush(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome,
                  const uint32_t* __restrict__ group, uint32_t n_groups, int dedupe, SetTable sets,
                  uint32_t* __restrict__ tri) {
  __shared__ uint32_t s_class[EXT_PAIR_SMALL][EXT_PAIR_THREADS];
  const uint64_t at = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (at > sets.mask || sets.key[at] == 0) return;
  const uint32_t rep = sets.rep[at];
  const uint64_t r0 = run_off[rep], c = run_off[rep + 1] - r0;
  const uint32_t n = class_list(run_genome, group, r0, (uint32_t)c, dedupe, s_class);
  add_pairs(n, s_class, n_groups, sets.count[at], tri);
}

// The matrix is symmetric: K6 counts cell (min, max) in 32 bits (a count is at most n_keys < 2^32); widen and mirror.
__global__ void extsim_pairwise_mirror(const uint32_t* __restrict__ tri, uint32_t n_groups, unsigned long long* __restrict__ inter) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_groups * n_groups) return;
  const uint32_t a = (uint32_t)(i / n_groups), b = (uint32_t)(i % n_groups);
  inter[i] = tri[(size_t)min(a, b) * n_groups + max(a, b)];
}

// K7 (a): per distinct k-mer, how many runs / positions survive the keep mask.
__global__ void drop_count(const uint64_t* __restrict__ run_off, const uint32_t* __restrict__ run_genome,
                           const uint64_t* __restrict__ pos_off, uint64_t U, const uint8_t* __restrict__ keep,
                           uint32_t* __restrict__ key_kept, uint32_t* __restrict__ runs_kept, uint32_t* __restrict__ pos_kept) {
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  uint32_t nr = 0; uint64_t np = 0;
  for (uint64_t r = run_off[u]; r < run_off[u + 1]; ++r)
    if (keep[run_genome[r]]) { ++nr; np += pos_off[r + 1] - pos_off[r]; }
  key_kept[u] = nr > 0; runs_kept[u] = nr; pos_kept[u] = (uint32_t)np;
}

// K7 (b): write the surviving CSR with renumbered genomes.
__global__ void drop_scatter(const uint64_t* __restrict__ ukeys, const uint64_t* __restrict__ run_off,
                             const uint32_t* __restrict__ run_genome, const uint64_t* __restrict__ pos_off,
                             const uint32_t* __restrict__ pos, uint64_t U, const uint8_t* __restrict__ keep,
                             const uint32_t* __restrict__ remap, const uint32_t* __restrict__ key_kept,
                             const uint64_t* __restrict__ key_rank, const uint64_t* __restrict__ run_rank,
                             const uint64_t* __restrict__ pos_rank, const uint64_t* __restrict__ first_occ,
                             uint64_t* __restrict__ n_first_occ, uint64_t* __restrict__ n_ukeys,
                             uint64_t* __restrict__ n_run_off, uint32_t* __restrict__ n_run_genome,
                             uint64_t* __restrict__ n_pos_off, uint32_t* __restrict__ n_pos) {
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= U || !key_kept[u]) return;
  uint64_t ku = key_rank[u], rr = run_rank[u], pp = pos_rank[u];
  n_ukeys[ku] = ukeys[u];
  n_first_occ[ku] = first_occ[u];
  n_run_off[ku] = rr;
  for (uint64_t r = run_off[u]; r < run_off[u + 1]; ++r) {
    uint32_t g = run_genome[r];
    if (!keep[g]) continue;
    n_run_genome[rr] = remap[g];
    n_pos_off[rr] = pp;
    for (uint64_t q = pos_off[r]; q < pos_off[r + 1]; ++q) n_pos[pp++] = pos[q];
    ++rr;
  }
}

inline unsigned grid_for(uint64_t n, int threads) { return (unsigned)std::max<uint64_t>(1, (n + threads - 1) / threads); }

float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

}  // namespace

// ===========================================================================
// host orchestration
// ===========================================================================

// K3 host side: CSR of the index from n_valid sorted (key, global position) records -- or (key, genome index) records,
// which give the keys and genome runs only (no positions: table-only builds)
int32_t rle_to_csr(Index& ix, const uint64_t* d_keys, const uint32_t* d_vals, uint64_t n_valid, bool vals_are_genomes) {
  cudaStream_t s = ix.stream;
  const uint64_t tiles = std::max<uint64_t>(1, (n_valid + RLE_TILE - 1) / RLE_TILE);
  DevBuf tile_keys, tile_runs, totals, gmap;
  PA_TRY(tile_keys.alloc((tiles + 1) * 8)); PA_TRY(tile_runs.alloc((tiles + 1) * 8)); PA_TRY(totals.alloc(16));
  const uint32_t n_genomes = ix.n_genomes;
  GenomeMap G{nullptr, 0};
  if (!vals_are_genomes) {
    const uint32_t gshift = (uint32_t)std::max<int>(0, (int)ceil_log2_u64(ix.total_bases + 1) - 20);
    const uint64_t gentries = (ix.total_bases >> gshift) + 1;
    PA_TRY(gmap.alloc(gentries * 4));
    genome_map_fill<<<grid_for(gentries, 256), 256, 0, s>>>(ix.genome_off.as<uint64_t>(), n_genomes, ix.total_bases, gshift,
                                                            gmap.as<uint32_t>(), gentries);
    G = GenomeMap{gmap.as<uint32_t>(), gshift};
  }
  if (vals_are_genomes)
    rle_count<true><<<(unsigned)tiles, RLE_THREADS, 0, s>>>(d_keys, d_vals, ix.genome_off.as<uint64_t>(), G, n_valid,
                                                            tile_keys.as<uint64_t>(), tile_runs.as<uint64_t>());
  else
    rle_count<false><<<(unsigned)tiles, RLE_THREADS, 0, s>>>(d_keys, d_vals, ix.genome_off.as<uint64_t>(), G, n_valid,
                                                             tile_keys.as<uint64_t>(), tile_runs.as<uint64_t>());
  scan_u64_single_block<<<1, 1024, 0, s>>>(tile_keys.as<uint64_t>(), tiles, totals.as<uint64_t>());
  scan_u64_single_block<<<1, 1024, 0, s>>>(tile_runs.as<uint64_t>(), tiles, totals.as<uint64_t>() + 1);
  uint64_t h_tot[2];
  PA_CUDA(cudaMemcpyAsync(h_tot, totals.p, 16, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  const uint64_t U = h_tot[0], R = h_tot[1];
  if (U >= 0xFFFFFFFFull) { set_error("index build: more than 2^32 - 2 distinct k-mers in one partition"); return ST_UNSUPPORTED; }
  PA_TRY(ix.ukeys.alloc((U + 1) * 8)); PA_TRY(ix.run_off.alloc((U + 1) * 8));
  PA_TRY(ix.run_genome.alloc((R + 1) * 4));
  if (vals_are_genomes) { PA_TRY(ix.pos_off.alloc(8)); PA_TRY(ix.pos.alloc(4)); }
  else { PA_TRY(ix.pos_off.alloc((R + 1) * 8)); PA_TRY(ix.pos.alloc((n_valid + 1) * 4)); }
  if (n_valid) {
    if (vals_are_genomes)
      rle_scatter<true><<<(unsigned)tiles, RLE_THREADS, 0, s>>>(d_keys, d_vals, ix.genome_off.as<uint64_t>(), G, n_valid,
                                                                tile_keys.as<uint64_t>(), tile_runs.as<uint64_t>(), ix.ukeys.as<uint64_t>(),
                                                                ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(),
                                                                ix.pos_off.as<uint64_t>(), ix.pos.as<uint32_t>());
    else
      rle_scatter<false><<<(unsigned)tiles, RLE_THREADS, 0, s>>>(d_keys, d_vals, ix.genome_off.as<uint64_t>(), G, n_valid,
                                                                 tile_keys.as<uint64_t>(), tile_runs.as<uint64_t>(), ix.ukeys.as<uint64_t>(),
                                                                 ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(),
                                                                 ix.pos_off.as<uint64_t>(), ix.pos.as<uint32_t>());
  }
  set_csr_tails<<<1, 1, 0, s>>>(ix.run_off.as<uint64_t>(), U, R, vals_are_genomes ? nullptr : ix.pos_off.as<uint64_t>(), n_valid);
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));
  ix.n_keys = U; ix.n_runs = R; ix.n_occ = n_valid;
  return ST_OK;
}

int32_t encode_slice(const uint8_t* d_bases, uint64_t total, uint64_t pos0, const uint64_t* d_genome_off, uint32_t G, int k,
                     uint64_t* d_keys, uint32_t* d_vals, const EncodeOpts& opt, unsigned long long* d_counters, cudaStream_t s) {
  if (k <= 0 || total == 0 || G == 0 || opt.emit_total == 0) return ST_OK;
  encode_windows<<<grid_for(std::min(total, opt.emit_total), ENC_TILE), ENC_THREADS, 0, s>>>(
      d_bases, total, pos0, d_genome_off, G, k, mix_params_for_k(k), d_keys, d_vals, d_counters,
      reinterpret_cast<unsigned int*>(d_counters + 1), opt);
  PA_CUDA(cudaGetLastError());
  return ST_OK;
}

int32_t index_build_from_device_bases(Index& ix, const uint8_t* d_bases) {
  cudaStream_t s = ix.stream;
  const uint64_t total = ix.total_bases;
  const uint32_t G = ix.n_genomes;
  const int k = ix.k;
  ix.n_keys = ix.n_runs = ix.n_occ = 0;
  cudaEvent_t ev[5];
  for (auto& e : ev) PA_CUDA(cudaEventCreate(&e));
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); } } guard{ev};

  if (k <= 0 || total == 0 || G == 0) {  // kmer.py:91-92: no windows at all
    PA_TRY(ix.ukeys.alloc(8)); PA_TRY(ix.run_off.alloc(8)); PA_TRY(ix.run_genome.alloc(4));
    PA_TRY(ix.pos_off.alloc(8)); PA_TRY(ix.pos.alloc(4));
    PA_CUDA(cudaMemsetAsync(ix.run_off.p, 0, 8, s)); PA_CUDA(cudaMemsetAsync(ix.pos_off.p, 0, 8, s));
    return index_build_tables(ix);
  }
  if (total >= 0xFFFFFFFFull) { set_error("index build: %llu bases exceed the 32-bit position space of this build", (unsigned long long)total); return ST_UNSUPPORTED; }

  // ---- K1 ----
  DevBuf keys_a, keys_b, vals_a, vals_b, counters, sort_tmp;
  PA_TRY(keys_a.alloc(total * 8)); PA_TRY(vals_a.alloc(total * 4));
  PA_TRY(keys_b.alloc(total * 8)); PA_TRY(vals_b.alloc(total * 4));     // allocated up front: the phase events
  PA_TRY(sort_tmp.alloc(radix_sort_temp_bytes(total)));                  // below then bracket kernels only
  PA_TRY(counters.alloc(16));
  PA_CUDA(cudaMemsetAsync(counters.p, 0, 16, s));
  PA_CUDA(cudaEventRecord(ev[0], s));
  PA_TRY(encode_slice(d_bases, total, 0, ix.genome_off.as<uint64_t>(), G, k, keys_a.as<uint64_t>(), vals_a.as<uint32_t>(),
                      EncodeOpts{total, nullptr, 1, 1, 0, 0}, counters.as<unsigned long long>(), s));
  PA_CUDA(cudaEventRecord(ev[1], s));
  unsigned long long h_cnt[2] = {0, 0};
  PA_CUDA(cudaMemcpyAsync(h_cnt, counters.p, 16, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  if (h_cnt[1] & 0xFFFFFFFFull) { set_error("genome sequence contains a character outside ACGTN"); return ST_BAD_BASE; }
  const uint64_t n_valid = h_cnt[0];

  // ---- K2 ----
  int in_b = 0;
  PA_TRY(radix_sort_pairs_hashed(keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), keys_b.as<uint64_t>(), vals_b.as<uint32_t>(), total,
                                 std::min(64, 2 * k + 1), sort_tmp.p, sort_tmp.bytes, s, &in_b));
  PA_CUDA(cudaEventRecord(ev[2], s));
  if (in_b) { keys_a.swap(keys_b); vals_a.swap(vals_b); }
  keys_b.release(); vals_b.release(); sort_tmp.release();

  // ---- K3 ----
  PA_TRY(rle_to_csr(ix, keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), n_valid, false));
  PA_CUDA(cudaEventRecord(ev[3], s));
  keys_a.release(); vals_a.release();

  PA_TRY(index_build_tables(ix));
  PA_CUDA(cudaEventRecord(ev[4], s));
  PA_CUDA(cudaStreamSynchronize(s));
  ix.t_encode_ms = elapsed_ms(ev[0], ev[1]);
  ix.t_sort_ms = elapsed_ms(ev[1], ev[2]);
  ix.t_rle_ms = elapsed_ms(ev[2], ev[3]);
  ix.t_table_ms = elapsed_ms(ev[3], ev[4]);
  return ST_OK;
}

int32_t index_ensure_first_occ(Index& ix) {
  if (ix.has_first_occ) return ST_OK;
  const uint64_t U = ix.n_keys;
  PA_TRY(ix.first_occ.alloc((U + 1) * 8));
  if (U)
    first_occurrence<<<grid_for(U, 256), 256, 0, ix.stream>>>(ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(),
                                                              ix.pos_off.as<uint64_t>(), ix.pos.as<uint32_t>(),
                                                              ix.genome_off.as<uint64_t>(), U, ix.first_occ.as<uint64_t>());
  PA_CUDA(cudaGetLastError());
  ix.has_first_occ = true;
  return ST_OK;
}

int32_t index_export_order(Index& ix, uint32_t* h_order) {
  const uint64_t U = ix.n_keys;
  if (U == 0) return ST_OK;
  cudaStream_t s = ix.stream;
  PA_TRY(index_ensure_first_occ(ix));
  DevBuf ka, kb, va, vb, tmp;
  PA_TRY(ka.alloc(U * 8)); PA_TRY(kb.alloc(U * 8)); PA_TRY(va.alloc(U * 4)); PA_TRY(vb.alloc(U * 4));
  PA_TRY(tmp.alloc(radix_sort_temp_bytes(U)));
  PA_CUDA(cudaMemcpyAsync(ka.p, ix.first_occ.p, U * 8, cudaMemcpyDeviceToDevice, s));
  iota_u32<<<grid_for(U, 256), 256, 0, s>>>(va.as<uint32_t>(), U);
  int in_b = 0;
  PA_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), U, 64, tmp.p,
                          tmp.bytes, s, &in_b));
  PA_CUDA(cudaMemcpyAsync(h_order, in_b ? vb.p : va.p, U * 4, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t index_checksum(Index& ix, uint64_t h_out[4]) {
  cudaStream_t s = ix.stream;
  DevBuf d;
  PA_TRY(d.alloc(32));
  PA_CUDA(cudaMemsetAsync(d.p, 0, 32, s));
  if (ix.n_keys) {
    int dev = 0, sms = 148;
    PA_CUDA(cudaGetDevice(&dev));
    PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)std::min<uint64_t>(grid_for(ix.n_keys, 256), (uint64_t)sms * 16);
    csr_checksum_kernel<<<grid, 256, 0, s>>>(ix.ukeys.as<uint64_t>(), ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(),
                                             ix.pos_off.as<uint64_t>(), ix.pos.as<uint32_t>(), ix.n_keys, 1, d.as<unsigned long long>());
    PA_CUDA(cudaGetLastError());
  }
  PA_CUDA(cudaMemcpyAsync(h_out, d.p, 32, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t index_lookup_ranks(Index& ix, const uint8_t* h_kmers, uint64_t n, uint64_t* h_rank) {
  if (n == 0) return ST_OK;
  if (ix.k <= 0 || ix.n_keys == 0) { for (uint64_t i = 0; i < n; ++i) h_rank[i] = LOOKUP_MISS; return ST_OK; }
  std::vector<uint64_t> q(n);
  for (uint64_t i = 0; i < n; ++i) {
    bool ok;
    uint64_t key = encode_kmer_host(h_kmers + i * (uint64_t)ix.k, ix.k, &ok);
    q[i] = ok ? mix_key(key, ix.mix) : SENTINEL_KEY;
  }
  cudaStream_t s = ix.stream;
  DevBuf dq, dr;
  PA_TRY(dq.alloc(n * 8)); PA_TRY(dr.alloc(n * 8));
  PA_CUDA(cudaMemcpyAsync(dq.p, q.data(), n * 8, cudaMemcpyHostToDevice, s));
  lookup_ranks<<<grid_for(n, 256), 256, 0, s>>>(ix.ukeys.as<uint64_t>(), ix.n_keys, dq.as<uint64_t>(), n, dr.as<uint64_t>());
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaMemcpyAsync(h_rank, dr.p, n * 8, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

static int32_t upload_groups(Index& ix, const uint32_t* h_group, uint32_t n_groups, DevBuf& d_group, int* dedupe) {
  const uint32_t G = ix.n_genomes;
  std::vector<uint8_t> seen(n_groups ? n_groups : 1, 0);
  *dedupe = 0;
  for (uint32_t g = 0; g < G; ++g) {
    if (h_group[g] >= n_groups) { set_error("extsim: group id out of range"); return ST_INVALID_ARG; }
    if (seen[h_group[g]]) *dedupe = 1;
    seen[h_group[g]] = 1;
  }
  PA_TRY(d_group.alloc((size_t)std::max<uint32_t>(G, 1) * 4));
  if (G) PA_CUDA(cudaMemcpyAsync(d_group.p, h_group, (size_t)G * 4, cudaMemcpyHostToDevice, ix.stream));
  return ST_OK;
}

int32_t index_extsim_stats(Index& ix, const uint32_t* h_group, uint32_t n_groups, uint64_t* h_total, uint64_t* h_unique) {
  cudaStream_t s = ix.stream;
  DevBuf d_group, d_out;
  int dedupe = 0;
  PA_TRY(upload_groups(ix, h_group, n_groups, d_group, &dedupe));
  size_t nb = (size_t)std::max<uint32_t>(n_groups, 1) * 8;
  PA_TRY(d_out.alloc(nb * 2));
  PA_CUDA(cudaMemsetAsync(d_out.p, 0, nb * 2, s));
  if (ix.n_keys) {
    int dev = 0, sms = 148;
    PA_CUDA(cudaGetDevice(&dev));
    PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // persistent grid: a block's shared counters (uint32) cannot overflow below 2^32 k-mers per block
    const unsigned grid = (unsigned)std::min<uint64_t>(grid_for(ix.n_keys, 256), (uint64_t)sms * 8);
    extsim_stats_kernel<<<grid, 256, 0, s>>>(
        ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), ix.n_keys, d_group.as<uint32_t>(), n_groups, dedupe,
        d_out.as<unsigned long long>(), d_out.as<unsigned long long>() + std::max<uint32_t>(n_groups, 1));
  }
  PA_CUDA(cudaGetLastError());
  if (n_groups) {
    PA_CUDA(cudaMemcpyAsync(h_total, d_out.p, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaMemcpyAsync(h_unique, d_out.as<char>() + nb, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, s));
  }
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t index_extsim_pairwise(Index& ix, const uint32_t* h_group, uint32_t n_groups, uint64_t* h_inter) {
  cudaStream_t s = ix.stream;
  DevBuf d_group, d_out;
  int dedupe = 0;
  PA_TRY(upload_groups(ix, h_group, n_groups, d_group, &dedupe));
  const size_t cells = std::max<size_t>((size_t)n_groups * n_groups, 1);
  DevBuf d_tri, d_sets;
  PA_TRY(d_out.alloc(cells * 8));
  PA_TRY(d_tri.alloc(cells * 4));
  // table of class lists: 2^20 entries (16 MB: stays in L2); PA_K6_SETS=0 counts pair by pair
  uint32_t entries = 1u << 20;
  if (const char* e = getenv("PA_K6_SETS")) entries = (uint32_t)atoi(e);
  if (entries & (entries - 1)) { set_error("PA_K6_SETS must be a power of two"); return ST_INVALID_ARG; }
  if (ix.n_keys >= 0xFFFFFFFFull) entries = 0;   // representatives are 32-bit k-mer indices
  SetTable sets{nullptr, nullptr, nullptr, 0, 0};
  if (const char* e = getenv("PA_K6_WEAK_HASH")) sets.weak_hash = *e == '1';
  uint32_t* d_collision = nullptr;
  if (entries) {
    PA_TRY(d_sets.alloc((size_t)entries * 16 + 16));
    sets.key = d_sets.as<unsigned long long>();
    sets.rep = (uint32_t*)(sets.key + entries);
    sets.count = sets.rep + entries;
    d_collision = sets.count + entries;
    sets.mask = entries - 1;
  }
  const unsigned grid = (unsigned)grid_for(ix.n_keys, EXT_PAIR_THREADS);
  for (int attempt = 0; attempt < 2; ++attempt) {
    PA_CUDA(cudaMemsetAsync(d_tri.p, 0, cells * 4, s));
    if (sets.mask) {
      PA_CUDA(cudaMemsetAsync(sets.key, 0, (size_t)entries * 8, s));
      PA_CUDA(cudaMemsetAsync(sets.rep, 0xFF, (size_t)entries * 4, s));
      PA_CUDA(cudaMemsetAsync(sets.count, 0, (size_t)entries * 4 + 16, s));
    }
    if (!ix.n_keys) break;
    extsim_pairwise_kernel<<<grid, EXT_PAIR_THREADS, 0, s>>>(
        ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), ix.n_keys, d_group.as<uint32_t>(), n_groups, dedupe, sets,
        d_tri.as<uint32_t>());
    PA_CUDA(cudaGetLastError());
    if (!sets.mask) break;
    extsim_sets_verify<<<grid, EXT_PAIR_THREADS, 0, s>>>(
        ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), ix.n_keys, d_group.as<uint32_t>(), dedupe, sets, d_collision);
    PA_CUDA(cudaGetLastError());
    uint32_t collision = 0;
    PA_CUDA(cudaMemcpyAsync(&collision, d_collision, 4, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
    if (collision) { sets.mask = 0; continue; }   // two lists under one hash: count pair by pair instead
    extsim_sets_flush<<<(unsigned)grid_for(entries, EXT_PAIR_THREADS), EXT_PAIR_THREADS, 0, s>>>(
        ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), d_group.as<uint32_t>(), n_groups, dedupe, sets,
        d_tri.as<uint32_t>());
    PA_CUDA(cudaGetLastError());
    break;
  }
  if (n_groups)
    extsim_pairwise_mirror<<<grid_for(cells, 256), 256, 0, s>>>(d_tri.as<uint32_t>(), n_groups, d_out.as<unsigned long long>());
  PA_CUDA(cudaGetLastError());
  if (n_groups) PA_CUDA(cudaMemcpyAsync(h_inter, d_out.p, (size_t)n_groups * n_groups * 8, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t index_drop_genomes(Index& ix, const uint8_t* h_keep) {
  cudaStream_t s = ix.stream;
  const uint32_t G = ix.n_genomes;
  const uint64_t U = ix.n_keys;
  std::vector<uint32_t> remap(std::max<uint32_t>(G, 1), 0xFFFFFFFFu);
  std::vector<uint64_t> new_off(1, 0);
  uint32_t ng = 0;
  for (uint32_t g = 0; g < G; ++g)
    if (h_keep[g]) {
      remap[g] = ng++;
      new_off.push_back(new_off.back() + (ix.h_genome_off[g + 1] - ix.h_genome_off[g]));
    }
  DevBuf d_keep, d_remap;
  PA_TRY(d_keep.alloc(std::max<uint32_t>(G, 1)));
  PA_TRY(d_remap.alloc((size_t)std::max<uint32_t>(G, 1) * 4));
  if (G) {
    PA_CUDA(cudaMemcpyAsync(d_keep.p, h_keep, G, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaMemcpyAsync(d_remap.p, remap.data(), (size_t)G * 4, cudaMemcpyHostToDevice, s));
  }
  uint64_t nU = 0, nR = 0, nN = 0;
  DevBuf n_ukeys, n_run_off, n_run_genome, n_pos_off, n_pos, n_first;
  PA_TRY(index_ensure_first_occ(ix));  // order keys refer to the genome list before the removal
  if (U) {
    DevBuf key_kept, runs_kept, pos_kept, key_rank, run_rank, pos_rank, tile_sums, totals;
    PA_TRY(key_kept.alloc(U * 4)); PA_TRY(runs_kept.alloc(U * 4)); PA_TRY(pos_kept.alloc(U * 4));
    PA_TRY(key_rank.alloc(U * 8)); PA_TRY(run_rank.alloc(U * 8)); PA_TRY(pos_rank.alloc(U * 8));
    PA_TRY(tile_sums.alloc((scan_tiles(U) + 1) * 8)); PA_TRY(totals.alloc(24));
    drop_count<<<grid_for(U, 256), 256, 0, s>>>(ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(),
                                                 ix.pos_off.as<uint64_t>(), U, d_keep.as<uint8_t>(), key_kept.as<uint32_t>(),
                                                 runs_kept.as<uint32_t>(), pos_kept.as<uint32_t>());
    PA_TRY(exclusive_scan_u32(key_kept.as<uint32_t>(), key_rank.as<uint64_t>(), U, tile_sums.as<uint64_t>(), totals.as<uint64_t>(), s));
    PA_TRY(exclusive_scan_u32(runs_kept.as<uint32_t>(), run_rank.as<uint64_t>(), U, tile_sums.as<uint64_t>(), totals.as<uint64_t>() + 1, s));
    PA_TRY(exclusive_scan_u32(pos_kept.as<uint32_t>(), pos_rank.as<uint64_t>(), U, tile_sums.as<uint64_t>(), totals.as<uint64_t>() + 2, s));
    uint64_t h_tot[3];
    PA_CUDA(cudaMemcpyAsync(h_tot, totals.p, 24, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
    nU = h_tot[0]; nR = h_tot[1]; nN = h_tot[2];
    PA_TRY(n_ukeys.alloc((nU + 1) * 8)); PA_TRY(n_run_off.alloc((nU + 1) * 8)); PA_TRY(n_run_genome.alloc((nR + 1) * 4));
    PA_TRY(n_pos_off.alloc((nR + 1) * 8)); PA_TRY(n_pos.alloc((nN + 1) * 4)); PA_TRY(n_first.alloc((nU + 1) * 8));
    drop_scatter<<<grid_for(U, 256), 256, 0, s>>>(
        ix.ukeys.as<uint64_t>(), ix.run_off.as<uint64_t>(), ix.run_genome.as<uint32_t>(), ix.pos_off.as<uint64_t>(),
        ix.pos.as<uint32_t>(), U, d_keep.as<uint8_t>(), d_remap.as<uint32_t>(), key_kept.as<uint32_t>(),
        key_rank.as<uint64_t>(), run_rank.as<uint64_t>(), pos_rank.as<uint64_t>(), ix.first_occ.as<uint64_t>(),
        n_first.as<uint64_t>(), n_ukeys.as<uint64_t>(),
        n_run_off.as<uint64_t>(), n_run_genome.as<uint32_t>(), n_pos_off.as<uint64_t>(), n_pos.as<uint32_t>());
    set_csr_tails<<<1, 1, 0, s>>>(n_run_off.as<uint64_t>(), nU, nR, n_pos_off.as<uint64_t>(), nN);
    PA_CUDA(cudaGetLastError());
    PA_CUDA(cudaStreamSynchronize(s));
  } else {
    PA_TRY(n_ukeys.alloc(8)); PA_TRY(n_run_off.alloc(8)); PA_TRY(n_run_genome.alloc(4)); PA_TRY(n_pos_off.alloc(8)); PA_TRY(n_pos.alloc(4));
    PA_TRY(n_first.alloc(8));
    PA_CUDA(cudaMemsetAsync(n_run_off.p, 0, 8, s)); PA_CUDA(cudaMemsetAsync(n_pos_off.p, 0, 8, s));
  }
  ix.ukeys.swap(n_ukeys); ix.run_off.swap(n_run_off); ix.run_genome.swap(n_run_genome);
  ix.pos_off.swap(n_pos_off); ix.pos.swap(n_pos); ix.first_occ.swap(n_first);
  ix.n_keys = nU; ix.n_runs = nR; ix.n_occ = nN;
  ix.n_genomes = ng;
  ix.h_genome_off = new_off;
  ix.total_bases = new_off.back();
  PA_TRY(ix.genome_off.alloc(new_off.size() * 8));
  PA_CUDA(cudaMemcpyAsync(ix.genome_off.p, new_off.data(), new_off.size() * 8, cudaMemcpyHostToDevice, s));
  PA_CUDA(cudaStreamSynchronize(s));
  ix.align_scratch.release(); ix.align_scratch_warps = 0;
  return ix.no_tables ? ST_OK : index_build_tables(ix);
}


}  // namespace pa
