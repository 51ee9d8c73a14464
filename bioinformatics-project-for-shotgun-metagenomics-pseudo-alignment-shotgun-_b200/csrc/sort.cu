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
#include <algorithm>

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
  // rank each item among equal digits of its warp, in item order (stable)
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

}  // namespace pa
