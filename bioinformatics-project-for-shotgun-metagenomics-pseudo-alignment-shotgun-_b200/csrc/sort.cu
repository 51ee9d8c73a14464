// sort.cu -- K2: stable LSD radix sort of (uint64 key, uint32 value) pairs.
//
// Replaces the dict / set insertions of KmerReference._build_kmer_mapping
// (/root/reference/src/kmer.py:146-150): equal k-mers become adjacent and, the
// sort being stable and the input being emitted in (genome, position) order,
// the occurrences of one k-mer stay in (genome, position) order.
//
// One-sweep organisation (one read + one write of the data per 8-bit digit):
//   radix_histogram   one pass over the keys, all digit histograms at once
//   radix_scan        exclusive scan of each 256-bin histogram
//   radix_pass        per tile: rank digits in shared memory (match.any per
//                     warp), publish the tile's digit counts, resolve the
//                     tile's global offsets by decoupled look-back over earlier
//                     tiles, reorder through shared memory, write coalesced runs
// HBM-bound: 8 + 2*(8+4) bytes per pair per digit pass are the algorithmic
// bytes; tensor cores are not applicable.
#include "sort.cuh"
#include "scan.cuh"
#include <algorithm>
#include <cstdlib>

namespace pa {

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
#ifndef PA_SORT_ITEMS
#define PA_SORT_ITEMS 16
#endif
#ifndef PA_SORT_MINB
#define PA_SORT_MINB 3
#endif
constexpr int SORT_ITEMS = PA_SORT_ITEMS;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 pairs per tile
constexpr int MAX_PASSES = 8;

constexpr uint64_t LB_FLAG_AGG = 1ULL << 62;
constexpr uint64_t LB_FLAG_INCL = 2ULL << 62;
constexpr uint64_t LB_COUNT_MASK = (1ULL << 62) - 1;

__global__ void __launch_bounds__(SORT_THREADS) radix_histogram(const uint64_t* __restrict__ keys, uint64_t n, int begin_bit,
                                                                int n_passes, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < n_passes * RADIX; i += SORT_THREADS) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * SORT_THREADS;
  for (uint64_t i = blockIdx.x * (uint64_t)SORT_THREADS + threadIdx.x; i < n; i += stride) {
    uint64_t key = keys[i];
    for (int p = 0; p < n_passes; ++p) atomicAdd(&sh[p * RADIX + ((key >> (begin_bit + p * RADIX_BITS)) & (RADIX - 1))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_passes * RADIX; i += SORT_THREADS)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// one block per pass: exclusive scan of 256 bins
__global__ void __launch_bounds__(RADIX) radix_scan(unsigned long long* __restrict__ hist) {
  __shared__ unsigned long long tmp[RADIX];
  unsigned long long* h = hist + (size_t)blockIdx.x * RADIX;
  unsigned long long v = h[threadIdx.x];
  tmp[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < RADIX; o <<= 1) {
    unsigned long long add = threadIdx.x >= o ? tmp[threadIdx.x - o] : 0;
    __syncthreads();
    tmp[threadIdx.x] += add;
    __syncthreads();
  }
  h[threadIdx.x] = tmp[threadIdx.x] - v;
}

struct PassSmem {
  uint64_t keys[SORT_TILE];
  uint32_t vals[SORT_TILE];
  uint32_t warp_hist[SORT_WARPS][RADIX];
  uint32_t tile_base[RADIX];   // exclusive scan of the tile's digit counts
  int64_t adjust[RADIX];       // global offset of digit d minus tile_base[d]
  uint32_t scan_tmp[RADIX];
  uint32_t tile_id;
};

__global__ void __launch_bounds__(SORT_THREADS, PA_SORT_MINB)
radix_pass(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint64_t n, int shift, const unsigned long long* __restrict__ digit_start,
           volatile unsigned long long* __restrict__ lookback, unsigned int* __restrict__ tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // tiles are handed out in launch order so every predecessor of a tile is already running
  if (tid == 0) sm.tile_id = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&sm.warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = sm.tile_id;
  const uint64_t tile_start = (uint64_t)tile * SORT_TILE;
  const uint32_t tile_n = (uint32_t)min((uint64_t)SORT_TILE, n - tile_start);

  uint64_t key[SORT_ITEMS];
  uint32_t val[SORT_ITEMS];
  uint32_t rank[SORT_ITEMS];
  const uint32_t warp_off = warp * (SORT_ITEMS * 32);
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    uint32_t li = warp_off + r * 32 + lane;
    bool ok = li < tile_n;
    key[r] = ok ? keys_in[tile_start + li] : 0;
    val[r] = ok ? vals_in[tile_start + li] : 0;
  }
  // rank each item among equal digits of its warp, in item order (stable).  (Measured on the B200, r02: issuing the
  // match.any of 8 or 16 items back to back before the counter updates, and reading 4 predecessors per look-back round
  // trip, are both slower -- sort 27.2 -> 28.5 / 27.3 / 29.5 ms.)
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    uint32_t li = warp_off + r * 32 + lane;
    bool ok = li < tile_n;
    uint32_t d = ok ? (uint32_t)((key[r] >> shift) & (RADIX - 1)) : (uint32_t)(RADIX + lane);
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = sm.warp_hist[warp][d];
      sm.warp_hist[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();

  // thread d owns digit d: exclusive scan over warps, publish, look back
  {
    const int d = tid;
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t t = sm.warp_hist[w][d];
      sm.warp_hist[w][d] = cnt;
      cnt += t;
    }
    uint64_t excl = 0;
    volatile unsigned long long* mine = lookback + (size_t)tile * RADIX + d;
    if (tile == 0) {
      *mine = LB_FLAG_INCL | cnt;
    } else {
      *mine = LB_FLAG_AGG | cnt;
      int64_t t = (int64_t)tile - 1;
      for (;;) {
        unsigned long long v = lookback[(size_t)t * RADIX + d];
        if ((v >> 62) == 0) continue;  // predecessor has not published yet
        excl += v & LB_COUNT_MASK;
        if (v & LB_FLAG_INCL) break;
        --t;
      }
      *mine = LB_FLAG_INCL | (excl + cnt);
    }
    // exclusive scan of cnt over digits (Hillis-Steele in shared memory)
    sm.scan_tmp[d] = cnt;
    __syncthreads();
    for (int o = 1; o < RADIX; o <<= 1) {
      uint32_t add = d >= o ? sm.scan_tmp[d - o] : 0;
      __syncthreads();
      sm.scan_tmp[d] += add;
      __syncthreads();
    }
    uint32_t base = sm.scan_tmp[d] - cnt;
    sm.tile_base[d] = base;
    sm.adjust[d] = (int64_t)(digit_start[d] + excl) - (int64_t)base;
  }
  __syncthreads();

  // reorder through shared memory so that each digit's items leave as one contiguous run
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    uint32_t li = warp_off + r * 32 + lane;
    if (li < tile_n) {
      uint32_t d = (uint32_t)((key[r] >> shift) & (RADIX - 1));
      uint32_t p = sm.tile_base[d] + sm.warp_hist[warp][d] + rank[r];
      sm.keys[p] = key[r];
      sm.vals[p] = val[r];
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < tile_n; i += SORT_THREADS) {
    uint64_t kk = sm.keys[i];
    uint32_t d = (uint32_t)((kk >> shift) & (RADIX - 1));
    uint64_t dst = (uint64_t)(sm.adjust[d] + (int64_t)i);
    keys_out[dst] = kk;
    vals_out[dst] = sm.vals[i];
  }
}


// ---------------------------------------------------------------------------
// Hashed keys: sort the TOP bits only, then repair the few places where that was not enough.
//
// The k-mer keys of a build are a bijective mix of the k-mer (common.cuh: mix_key), i.e. n distinct keys spread evenly
// over 2^end_bit values.  After a stable sort on the top T bits with 2^T >> n almost every run of equal top bits is a
// run of EQUAL KEYS (the occurrences of one k-mer, already in (genome, position) order); a run that mixes different
// keys ("mixed group") needs a stable sort of its own.  n = 5*10^8, T = 40: ~10^5 mixed groups of two or three records,
// five digit passes instead of eight.
//   fix_detect  one thread per record: a "boundary" is a position whose key differs from its predecessor's inside
//               one run of equal top bits
//   fix_groups  one warp per boundary: walks back to the head of the run and forward to its end; the FIRST boundary
//               of a run (everything before it equals its predecessor) files the run as a group, the others leave
//   fix_small   one warp per group of <= 32 records: stable rank by shuffles, rewritten in place
//   fix_big     one block per longer group (a k-mer with thousands of occurrences next to a stranger): for each
//               distinct key in ascending order a stable compaction into the sort's other buffer, then copied back
// Groups are disjoint and detection finishes (kernel boundary) before anything is rewritten.  Work lists that
// overflow, or a group with too many distinct keys, raise a flag and the caller sorts the remaining low bits with the
// ordinary passes -- correct on any input, fast on hashed keys.
// ---------------------------------------------------------------------------
struct FixCtrl { unsigned int n_bound, n_small, n_big, fallback; };
struct FixGroup { uint64_t head, size; };
constexpr int FIX_THREADS = 256;
constexpr uint32_t FIX_MAX_DISTINCT = 1024;

__global__ void __launch_bounds__(FIX_THREADS) fix_detect(const uint64_t* __restrict__ keys, uint64_t n, int b,
                                                          uint64_t* __restrict__ bound, uint32_t cap, FixCtrl* __restrict__ ctrl) {
  const uint64_t stride = (uint64_t)gridDim.x * FIX_THREADS;
  const int lane = threadIdx.x & 31;
  for (uint64_t i0 = blockIdx.x * (uint64_t)FIX_THREADS + (threadIdx.x & ~31); i0 < n; i0 += stride) {
    const uint64_t i = i0 + lane;
    const uint64_t k1 = i < n ? keys[i] : 0;
    uint64_t k0 = __shfl_up_sync(0xffffffffu, k1, 1);
    if (lane == 0 && i > 0 && i < n) k0 = keys[i - 1];
    if (i > 0 && i < n && k0 != k1 && (k0 >> b) == (k1 >> b)) {
      const unsigned int at = atomicAdd(&ctrl->n_bound, 1u);
      if (at < cap) bound[at] = i; else ctrl->fallback = 1u;
    }
  }
}

__global__ void __launch_bounds__(FIX_THREADS) fix_groups(const uint64_t* __restrict__ keys, uint64_t n, int b,
                                                          const uint64_t* __restrict__ bound, uint32_t cap, FixCtrl* __restrict__ ctrl,
                                                          FixGroup* __restrict__ small, FixGroup* __restrict__ big) {
  const unsigned int nb = min(ctrl->n_bound, cap);
  const int lane = threadIdx.x & 31;
  const unsigned int warps = gridDim.x * (FIX_THREADS / 32);
  for (unsigned int e = blockIdx.x * (FIX_THREADS / 32) + (threadIdx.x >> 5); e < nb; e += warps) {
    const uint64_t i = bound[e];
    const uint64_t top = keys[i] >> b, kref = keys[i - 1];
    // backwards from i-1: the run's head, and whether everything in [head, i) equals keys[i-1]
    uint64_t head = 0;
    bool leader = true;
    for (uint64_t base = i;; base -= 32) {            // lane l looks at position base-1-l
      const bool valid = base >= 1 + (uint64_t)lane;
      const uint64_t kq = valid ? keys[base - 1 - lane] : 0;
      const bool in = valid && (kq >> b) == top;
      const unsigned int out_mask = __ballot_sync(0xffffffffu, !in);
      const unsigned int diff_mask = __ballot_sync(0xffffffffu, in && kq != kref);
      const unsigned int first_out = out_mask ? (unsigned int)__ffs(out_mask) - 1 : 32u;
      const unsigned int before = first_out >= 32 ? 0xffffffffu : ((1u << first_out) - 1);
      if (diff_mask & before) { leader = false; break; }
      if (first_out < 32) { head = base - first_out; break; }
    }
    if (!leader) continue;
    uint64_t end = n;
    for (uint64_t base = i;; base += 32) {            // lane l looks at position base+l
      const uint64_t q = base + lane;
      const bool in = q < n && (keys[q] >> b) == top;
      const unsigned int out_mask = __ballot_sync(0xffffffffu, !in);
      if (out_mask) { end = base + (unsigned int)__ffs(out_mask) - 1; break; }
    }
    if (lane == 0) {
      const FixGroup g{head, end - head};
      if (g.size <= 32) small[atomicAdd(&ctrl->n_small, 1u)] = g;
      else big[atomicAdd(&ctrl->n_big, 1u)] = g;
    }
  }
}

__global__ void __launch_bounds__(FIX_THREADS) fix_small(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                         const FixGroup* __restrict__ small, const FixCtrl* __restrict__ ctrl) {
  const unsigned int ng = ctrl->n_small;
  const int lane = threadIdx.x & 31;
  const unsigned int warps = gridDim.x * (FIX_THREADS / 32);
  for (unsigned int e = blockIdx.x * (FIX_THREADS / 32) + (threadIdx.x >> 5); e < ng; e += warps) {
    const uint64_t h = small[e].head;
    const int S = (int)small[e].size;
    const bool act = lane < S;
    const uint64_t key = act ? keys[h + lane] : ~0ULL;
    const uint32_t val = act ? vals[h + lane] : 0u;
    unsigned int r = 0;
    for (int j = 0; j < S; ++j) {
      const uint64_t kj = __shfl_sync(0xffffffffu, key, j);
      r += (kj < key || (kj == key && j < lane)) ? 1u : 0u;
    }
    __syncwarp();
    if (act) { keys[h + r] = key; vals[h + r] = val; }
  }
}

__global__ void __launch_bounds__(FIX_THREADS) fix_big(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                       uint64_t* __restrict__ keys_s, uint32_t* __restrict__ vals_s,
                                                       const FixGroup* __restrict__ big, FixCtrl* __restrict__ ctrl) {
  __shared__ uint64_t red[FIX_THREADS / 32];
  __shared__ uint32_t scan_scratch[FIX_THREADS / 32];
  const unsigned int ng = ctrl->n_big;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (unsigned int e = blockIdx.x; e < ng; e += gridDim.x) {
    const uint64_t h = big[e].head, S = big[e].size;
    uint64_t last = 0, out = 0;
    uint32_t distinct = 0;
    bool gave_up = false;
    while (out < S) {
      if (distinct >= FIX_MAX_DISTINCT || (uint64_t)(distinct + 1) * S > (1ULL << 32)) { gave_up = true; break; }
      // smallest key above `last` (the first round takes the smallest key of all)
      uint64_t mn = ~0ULL;
      for (uint64_t j = tid; j < S; j += FIX_THREADS) {
        const uint64_t kk = keys[h + j];
        if ((distinct == 0 || kk > last) && kk < mn) mn = kk;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const uint64_t t = __shfl_xor_sync(0xffffffffu, mn, o); mn = t < mn ? t : mn; }
      if (lane == 0) red[warp] = mn;
      __syncthreads();
      uint64_t cur = red[0];
#pragma unroll
      for (int w = 1; w < FIX_THREADS / 32; ++w) cur = red[w] < cur ? red[w] : cur;
      __syncthreads();
      // its records, in their present order, go next
      for (uint64_t c0 = 0; c0 < S; c0 += FIX_THREADS) {
        const uint64_t j = c0 + tid;
        const bool f = j < S && keys[h + j] == cur;
        uint32_t total;
        const uint32_t pos = block_exclusive_scan<FIX_THREADS, uint32_t>(f ? 1u : 0u, total, scan_scratch);
        if (f) { keys_s[h + out + pos] = cur; vals_s[h + out + pos] = vals[h + j]; }
        out += total;
      }
      last = cur;
      ++distinct;
    }
    if (gave_up) { if (tid == 0) ctrl->fallback = 1u; continue; }   // nothing of this group was rewritten
    __syncthreads();
    for (uint64_t j = tid; j < S; j += FIX_THREADS) { keys[h + j] = keys_s[h + j]; vals[h + j] = vals_s[h + j]; }
    __syncthreads();
  }
}

}  // namespace

size_t radix_sort_temp_bytes(uint64_t n) {
  uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
  // [hist: MAX_PASSES*256 u64][tile counter (padded to 256 B)][lookback: tiles*256 u64]
  return (size_t)MAX_PASSES * RADIX * 8 + 256 + (size_t)(tiles ? tiles : 1) * RADIX * 8;
}

int32_t radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n,
                         int end_bit, void* d_temp, size_t temp_bytes, cudaStream_t s, int* result_in_b, int begin_bit) {
  *result_in_b = 0;
  if (end_bit > 64) end_bit = 64;
  if (begin_bit < 0) begin_bit = 0;
  if (n == 0 || end_bit <= begin_bit) return ST_OK;
  if (temp_bytes < radix_sort_temp_bytes(n)) { set_error("radix sort: temp buffer too small"); return ST_INVALID_ARG; }
  const int n_passes = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  const uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
  if (tiles > 0xFFFFFFFFull) { set_error("radix sort: too many tiles"); return ST_UNSUPPORTED; }
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(d_temp);
  unsigned int* tile_counter = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(d_temp) + (size_t)MAX_PASSES * RADIX * 8);
  unsigned long long* lookback = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(tile_counter) + 256);

  // the opt-in to more than 48 KB of dynamic shared memory is per device: set it on every call (a process may build on
  // several devices)
  PA_CUDA(cudaFuncSetAttribute(radix_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)));
  int dev = 0, sms = 148;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  PA_CUDA(cudaMemsetAsync(hist, 0, (size_t)MAX_PASSES * RADIX * 8, s));
  int hgrid = (int)std::min<uint64_t>((n + SORT_THREADS - 1) / SORT_THREADS, (uint64_t)sms * 8);
  radix_histogram<<<hgrid, SORT_THREADS, 0, s>>>(keys_a, n, begin_bit, n_passes, hist);
  radix_scan<<<n_passes, RADIX, 0, s>>>(hist);
  PA_CUDA(cudaGetLastError());

  uint64_t* kin = keys_a; uint32_t* vin = vals_a; uint64_t* kout = keys_b; uint32_t* vout = vals_b;
  for (int p = 0; p < n_passes; ++p) {
    PA_CUDA(cudaMemsetAsync(tile_counter, 0, 256 + (size_t)tiles * RADIX * 8, s));
    radix_pass<<<(unsigned)tiles, SORT_THREADS, sizeof(PassSmem), s>>>(kin, vin, kout, vout, n, begin_bit + p * RADIX_BITS,
                                                                       hist + (size_t)p * RADIX, lookback, tile_counter);
    PA_CUDA(cudaGetLastError());
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *result_in_b = (n_passes & 1);
  return ST_OK;
}

int32_t radix_sort_pairs_hashed(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n, int end_bit,
                                void* d_temp, size_t temp_bytes, cudaStream_t s, int* result_in_b, int top_bits, int* fell_back) {
  if (fell_back) *fell_back = 0;
  if (end_bit > 64) end_bit = 64;
  if (top_bits <= 0) top_bits = (int)ceil_log2_u64(n ? n : 1) + 10;   // expected mixed groups: n / 2^11 at most
  const int passes_full = (end_bit + RADIX_BITS - 1) / RADIX_BITS;
  const int passes_top = (top_bits + RADIX_BITS - 1) / RADIX_BITS;
  const char* full = getenv("PA_SORT_FULL");
  if ((full && *full == '1') || passes_top >= passes_full || n < 2)
    return radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, n, end_bit, d_temp, temp_bytes, s, result_in_b, 0);
  const int begin_bit = end_bit - passes_top * RADIX_BITS;
  PA_TRY(radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, n, end_bit, d_temp, temp_bytes, s, result_in_b, begin_bit));
  uint64_t* k_cur = *result_in_b ? keys_b : keys_a; uint32_t* v_cur = *result_in_b ? vals_b : vals_a;
  uint64_t* k_oth = *result_in_b ? keys_a : keys_b; uint32_t* v_oth = *result_in_b ? vals_a : vals_b;

  // the histogram and look-back areas of the passes are free again: control block + three work lists
  const uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
  FixCtrl* ctrl = reinterpret_cast<FixCtrl*>(d_temp);
  char* lists = reinterpret_cast<char*>(d_temp) + (size_t)MAX_PASSES * RADIX * 8 + 256;
  const uint32_t cap = (uint32_t)std::min<uint64_t>(tiles * RADIX * 8 / 40, 0x7FFFFFFFull);
  uint64_t* bound = reinterpret_cast<uint64_t*>(lists);
  FixGroup* small = reinterpret_cast<FixGroup*>(lists + (size_t)cap * 8);
  FixGroup* big = small + cap;
  int dev = 0, sms = 148;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PA_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(FixCtrl), s));
  const unsigned grid = (unsigned)std::min<uint64_t>((n + FIX_THREADS - 1) / FIX_THREADS, (uint64_t)sms * 8);
  fix_detect<<<grid, FIX_THREADS, 0, s>>>(k_cur, n, begin_bit, bound, cap, ctrl);
  fix_groups<<<sms * 2, FIX_THREADS, 0, s>>>(k_cur, n, begin_bit, bound, cap, ctrl, small, big);
  fix_small<<<sms * 2, FIX_THREADS, 0, s>>>(k_cur, v_cur, small, ctrl);
  fix_big<<<sms, FIX_THREADS, 0, s>>>(k_cur, v_cur, k_oth, v_oth, big, ctrl);
  PA_CUDA(cudaGetLastError());
  FixCtrl h{};
  PA_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(FixCtrl), cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  if (h.fallback) {
    // stable passes over the low bits and the top bits again: equal keys keep their present (= original) order
    if (fell_back) *fell_back = 1;
    int in_oth = 0;
    PA_TRY(radix_sort_pairs(k_cur, v_cur, k_oth, v_oth, n, end_bit, d_temp, temp_bytes, s, &in_oth, 0));
    if (in_oth) *result_in_b ^= 1;
  }
  return ST_OK;
}

}  // namespace pa
