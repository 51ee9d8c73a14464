// dist.cu -- partitioned / streamed index build (SURVEY.md 8(e) "Build: one exchange step") and the C ABI of the
// communicator.  Replaces KmerReference._build_kmer_mapping (/root/reference/src/kmer.py:135-150) across ranks and,
// for indexes larger than one device pass, across rounds.
//
//   owner of a record   = f(top bits of the minimizer hash of its k-mer)          (K1, build.cu: encode_windows)
//   exchange            = one stable scatter pass that stores every record straight into the owner's receive buffer:
//                         peer memory mapped through CUDA IPC, one long coalesced run per (tile, owner); the stores
//                         travel over NVLink and overlap the pass tile by tile (owner_scatter below)
//   per rank and round  K2 sort + K3 CSR of the received key range, then its k-mers go into THIS rank's slice of the
//                       full-size lookup table (table.cu)
//   replica             the slices are all-gathered in place (Comm::allgatherv_device_inplace); genome sets and stash
//                       entries are small and travel through the host
#include "comm.h"
#include "index.cuh"
#include "scan.cuh"
#include "sort.cuh"
#include "table.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <new>
#include <vector>

namespace pa {

namespace {

// ===========================================================================
// exchange kernels
// ===========================================================================
constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;   // 4096 records per tile
constexpr int SC_MAX_OWNERS = 256;               // owner ids are bytes; 255 = "not emitted"
constexpr uint32_t NO_OWNER = 255;

struct Route {            // where the records of one owner go: the owner's receive buffer (possibly peer memory) at
  uint64_t* keys;         // this sender's segment
  uint32_t* vals;
};

// records of every owner in every tile
__global__ void __launch_bounds__(SC_THREADS)
owner_count(const uint8_t* __restrict__ owner, uint64_t n, uint32_t W, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t hist[SC_MAX_OWNERS];
  for (int i = threadIdx.x; i < SC_MAX_OWNERS; i += SC_THREADS) hist[i] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < SC_ITEMS; ++r) {
    const uint64_t i = base + (uint64_t)r * SC_THREADS + threadIdx.x;
    const uint32_t o = i < n ? owner[i] : NO_OWNER;
    const uint32_t peers = __match_any_sync(0xffffffffu, o);
    if (o != NO_OWNER && lane == (uint32_t)__ffs(peers) - 1) atomicAdd(&hist[o], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (uint32_t o = threadIdx.x; o < W; o += SC_THREADS) tile_counts[(size_t)blockIdx.x * W + o] = hist[o];
}

// block o: exclusive scan over the tiles of owner o's counts; totals[o] = records of owner o
__global__ void __launch_bounds__(1024)
owner_scan(const uint32_t* __restrict__ tile_counts, uint64_t tiles, uint32_t W, uint64_t* __restrict__ tile_off,
           unsigned long long* __restrict__ totals) {
  __shared__ uint64_t ws[1024 / 32];
  __shared__ uint64_t carry;
  const uint32_t o = blockIdx.x;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < tiles; base += 1024) {
    const uint64_t t = base + threadIdx.x;
    const uint64_t v = t < tiles ? tile_counts[t * W + o] : 0;
    uint64_t total;
    const uint64_t excl = block_exclusive_scan<1024, uint64_t>(v, total, ws);
    const uint64_t c = carry;
    if (t < tiles) tile_off[t * W + o] = c + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[o] = carry;
}

struct ScatterSmem {
  uint64_t keys[SC_TILE];
  uint32_t vals[SC_TILE];
  uint8_t own[SC_TILE];
  uint32_t warp_hist[SC_WARPS][SC_MAX_OWNERS];
  uint32_t tile_base[SC_MAX_OWNERS + 1];
};

// Stable split of one tile by owner, through shared memory, then one contiguous run per owner to route[owner] at
// tile_off[tile][owner]: the partition pass and the all-to-all are this one kernel.  A warp stores 256 contiguous
// bytes of keys (128 of values) per instruction; with 8 owners a run is ~512 records = 4 KB.
__global__ void __launch_bounds__(SC_THREADS)
owner_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint8_t* __restrict__ owner, uint64_t n,
              uint32_t W, const uint64_t* __restrict__ tile_off, const Route* __restrict__ route) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < SC_WARPS * SC_MAX_OWNERS; i += SC_THREADS) (&sm.warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint64_t tile_start = (uint64_t)blockIdx.x * SC_TILE;
  const uint32_t tile_n = (uint32_t)min((uint64_t)SC_TILE, n - tile_start);
  uint64_t key[SC_ITEMS];
  uint32_t val[SC_ITEMS], own[SC_ITEMS], rank[SC_ITEMS];
  const uint32_t warp_off = warp * (SC_ITEMS * 32);
#pragma unroll
  for (int r = 0; r < SC_ITEMS; ++r) {
    const uint32_t li = warp_off + r * 32 + lane;
    const bool ok = li < tile_n;
    own[r] = ok ? owner[tile_start + li] : NO_OWNER;
    key[r] = ok ? keys[tile_start + li] : 0;
    val[r] = ok ? vals[tile_start + li] : 0;
  }
  // rank of every record among the records of its owner inside its warp, in input order (stable)
#pragma unroll
  for (int r = 0; r < SC_ITEMS; ++r) {
    const uint32_t o = own[r];
    const uint32_t peers = __match_any_sync(0xffffffffu, o);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (o != NO_OWNER && lane == leader) {
      old = sm.warp_hist[warp][o];
      sm.warp_hist[warp][o] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();
  if ((uint32_t)tid < W) {   // exclusive scan over the warps of owner `tid`
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < SC_WARPS; ++w) { const uint32_t t = sm.warp_hist[w][tid]; sm.warp_hist[w][tid] = cnt; cnt += t; }
    sm.tile_base[tid + 1] = cnt;
  }
  __syncthreads();
  if (tid == 0) {   // counts -> starts: tile_base[o] = first slot of owner o, tile_base[W] = records kept
    uint32_t acc = 0;
    for (uint32_t o = 0; o < W; ++o) { const uint32_t c = sm.tile_base[o + 1]; sm.tile_base[o] = acc; acc += c; }
    sm.tile_base[W] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SC_ITEMS; ++r) {
    const uint32_t o = own[r];
    if (o != NO_OWNER) {
      const uint32_t p = sm.tile_base[o] + sm.warp_hist[warp][o] + rank[r];
      sm.keys[p] = key[r];
      sm.vals[p] = val[r];
      sm.own[p] = (uint8_t)o;
    }
  }
  __syncthreads();
  const uint32_t kept = sm.tile_base[W];
  for (uint32_t i = tid; i < kept; i += SC_THREADS) {
    const uint32_t o = sm.own[i];
    const Route rt = route[o];
    const uint64_t j = tile_off[(size_t)blockIdx.x * W + o] + (i - sm.tile_base[o]);
    rt.keys[j] = sm.keys[i];
    rt.vals[j] = sm.vals[i];
  }
}

// ===========================================================================
// Table slices on the wire.  A slice is mostly empty slots (load factor 0.2: four of five are all ones), so a rank does
// not ship its slice but, per 512-byte block, a 64-bit occupancy bitmap plus the occupied slot words -- 110 instead of
// 512 bytes per block at load 0.2 -- and every rank expands what it receives into its table: HBM bandwidth (cheap)
// traded for NVLink bytes (the bound of the gather).  One warp per block, two slots per lane.
// ===========================================================================
constexpr int CMP_THREADS = 256;
constexpr uint32_t BLOCK_SLOTS = BLOCK_BUCKETS * BUCKET_SLOTS;   // 64

// A warp takes CMP_BLOCKS consecutive blocks and has all their loads in flight before it looks at the first.
constexpr int CMP_BLOCKS = 8;
__global__ void __launch_bounds__(CMP_THREADS)
slice_bitmaps(const uint64_t* __restrict__ slots, uint64_t b_lo, uint64_t b_hi, unsigned long long* __restrict__ bitmap,
              uint32_t* __restrict__ cnt) {
  const uint64_t b0 = b_lo + ((blockIdx.x * (uint64_t)CMP_THREADS + threadIdx.x) / 32) * CMP_BLOCKS;
  const uint32_t lane = threadIdx.x & 31;
  if (b0 >= b_hi) return;
  uint64_t s0[CMP_BLOCKS], s1[CMP_BLOCKS];
#pragma unroll
  for (int j = 0; j < CMP_BLOCKS; ++j) {
    const bool ok = b0 + j < b_hi;
    const uint64_t* p = slots + (b0 + j) * BLOCK_SLOTS;
    s0[j] = ok ? p[lane] : EMPTY64;
    s1[j] = ok ? p[32 + lane] : EMPTY64;
  }
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < CMP_BLOCKS; ++j) {
    const uint32_t m0 = __ballot_sync(0xffffffffu, s0[j] != EMPTY64), m1 = __ballot_sync(0xffffffffu, s1[j] != EMPTY64);
    if (lane == (uint32_t)j) mine = ((unsigned long long)m1 << 32) | m0;
  }
  if (lane < CMP_BLOCKS && b0 + lane < b_hi) { bitmap[b0 + lane] = mine; cnt[b0 + lane - b_lo] = __popcll(mine); }
}

__global__ void __launch_bounds__(CMP_THREADS)
slice_words(const uint64_t* __restrict__ slots, uint64_t b_lo, uint64_t b_hi, const uint64_t* __restrict__ off /* per block of the slice */,
            uint64_t* __restrict__ words) {
  const uint64_t b0 = b_lo + ((blockIdx.x * (uint64_t)CMP_THREADS + threadIdx.x) / 32) * CMP_BLOCKS;
  const uint32_t lane = threadIdx.x & 31;
  if (b0 >= b_hi) return;
  uint64_t s0[CMP_BLOCKS], s1[CMP_BLOCKS];
#pragma unroll
  for (int j = 0; j < CMP_BLOCKS; ++j) {
    const bool ok = b0 + j < b_hi;
    const uint64_t* p = slots + (b0 + j) * BLOCK_SLOTS;
    s0[j] = ok ? p[lane] : EMPTY64;
    s1[j] = ok ? p[32 + lane] : EMPTY64;
  }
  const uint64_t my_off = (lane < CMP_BLOCKS && b0 + lane < b_hi) ? off[b0 + lane - b_lo] : 0;
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int j = 0; j < CMP_BLOCKS; ++j) {
    const uint32_t m0 = __ballot_sync(0xffffffffu, s0[j] != EMPTY64), m1 = __ballot_sync(0xffffffffu, s1[j] != EMPTY64);
    uint64_t* dst = words + __shfl_sync(0xffffffffu, my_off, j);
    if (s0[j] != EMPTY64) dst[__popc(m0 & lt)] = s0[j];
    if (s1[j] != EMPTY64) dst[__popc(m0) + __popc(m1 & lt)] = s1[j];
  }
}

__global__ void __launch_bounds__(CMP_THREADS)
bitmap_counts(const unsigned long long* __restrict__ bitmap, uint64_t n_blocks, uint32_t* __restrict__ cnt) {
  const uint64_t b = blockIdx.x * (uint64_t)CMP_THREADS + threadIdx.x;
  if (b < n_blocks) cnt[b] = __popcll(bitmap[b]);
}

// the blocks [b_lo, b_hi) of one source rank: bitmap + words -> 64 slots each.  A warp expands EXP_BLOCKS consecutive
// blocks and has all their word loads in flight before the first store (one block per warp left the kernel at 2.3 TB/s:
// four dependent loads in front of every 512 bytes written).  off = offsets of the blocks' words in the dense global
// order, delta = where this rank's words really start minus where they would start in that order.
constexpr int EXP_BLOCKS = 8;
__global__ void __launch_bounds__(CMP_THREADS)
slice_expand(const unsigned long long* __restrict__ bitmap, const uint64_t* __restrict__ off, const uint64_t* __restrict__ words,
             uint64_t b_lo, uint64_t b_hi, long long delta, uint64_t* __restrict__ slots) {
  const uint64_t b0 = b_lo + ((blockIdx.x * (uint64_t)CMP_THREADS + threadIdx.x) / 32) * EXP_BLOCKS;
  const uint32_t lane = threadIdx.x & 31;
  if (b0 >= b_hi) return;
  unsigned long long my_bm = 0, my_off = 0;
  if (lane < EXP_BLOCKS && b0 + lane < b_hi) { my_bm = bitmap[b0 + lane]; my_off = (unsigned long long)((long long)off[b0 + lane] + delta); }
  const uint32_t lt = (1u << lane) - 1;
  uint64_t v0[EXP_BLOCKS], v1[EXP_BLOCKS];
#pragma unroll
  for (int j = 0; j < EXP_BLOCKS; ++j) {
    const unsigned long long bm = __shfl_sync(0xffffffffu, my_bm, j), of = __shfl_sync(0xffffffffu, my_off, j);
    const uint32_t m0 = (uint32_t)bm, m1 = (uint32_t)(bm >> 32);
    const uint64_t* src = words + of;
    v0[j] = ((m0 >> lane) & 1u) ? __ldcs(src + __popc(m0 & lt)) : EMPTY64;
    v1[j] = ((m1 >> lane) & 1u) ? __ldcs(src + __popc(m0) + __popc(m1 & lt)) : EMPTY64;
  }
#pragma unroll
  for (int j = 0; j < EXP_BLOCKS; ++j)
    if (b0 + j < b_hi) {
      uint64_t* p = slots + (b0 + j) * BLOCK_SLOTS;
      __stcs(p + lane, v0[j]);
      __stcs(p + 32 + lane, v1[j]);
    }
}

double ms_since(std::chrono::steady_clock::time_point a) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
}

// one chunk's encoded records + their per-tile owner offsets
struct ChunkRecords {
  DevBuf keys, vals, owner, tile_counts, tile_off, totals;
  uint64_t n = 0, tiles = 0;
};

struct Piece { uint64_t base; std::vector<uint32_t> ids; };   // genome sets of one (round, rank): sectors [base, base + ids/8)

}  // namespace

// contiguous runs of whole genomes per rank, balanced by bases (FASTA order is kept, so genome indices ascend with the
// rank and equal k-mers arrive at their owner in genome order)
void genome_shard(const uint64_t* off, uint32_t G, int n_ranks, int rank, uint32_t* g_lo, uint32_t* g_hi) {
  const uint64_t total = G ? off[G] - off[0] : 0;
  std::vector<uint32_t> bounds(1, 0);
  uint32_t g = 0;
  uint64_t acc = 0;
  for (int r = 1; r < n_ranks; ++r) {
    const uint64_t target = total / (uint64_t)n_ranks * (uint64_t)r + (total % (uint64_t)n_ranks) * (uint64_t)r / (uint64_t)n_ranks;
    while (g < G && acc + (off[g + 1] - off[g]) / 2 < target) { acc += off[g + 1] - off[g]; ++g; }
    bounds.push_back(g);
  }
  bounds.push_back(G);
  *g_lo = bounds[rank]; *g_hi = bounds[rank + 1];
}

struct BuildTimings { float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };

// ===========================================================================
// the build
// ===========================================================================
struct DistBuild {
  Comm* comm;                 // never null here (a single-rank communicator stands in for "no communicator")
  Index* rep;                 // the replica being built (owns the stream everything runs on)
  Index* part;                // CSR of the current round (kept as the partition when rounds == 1 and !table_only)
  const uint8_t* bases;       // my genomes [g_lo, g_hi) concatenated (device or host)
  bool host_bases, table_only;
  uint32_t G, g_lo, g_hi;
  int k;
  uint32_t W, R;              // ranks, rounds
  uint64_t my_base0, my_bases;   // global position of my first base, number of my bases
  uint64_t chunk;             // bases per K1 launch
  BuildTimings tm;
  // state of the table being filled
  std::vector<uint64_t> overflow;          // {raw key, value} pairs of my k-mers that left their chains
  std::vector<Piece> pieces;               // my genome sets, one piece per round
  uint64_t msec_cursor = 0;                // global sector count so far (all ranks, all rounds)
  uint64_t total_keys = 0, total_runs = 0, total_occ = 0, valid_windows = 0;
  TableGeom geom;
  bool have_table = false;

  uint32_t first_digit(uint32_t part_id) const {   // smallest digit d with (d * n_parts) >> tb >= part_id
    const uint32_t tb = digit_bits_for_k(k), n_parts = W * R;
    return (uint32_t)((((uint64_t)part_id << tb) + n_parts - 1) / n_parts);
  }

  // K1 over chunk c of my genomes into `out` (+ per-tile owner counts / offsets)
  int32_t encode_chunk(uint64_t c0, uint64_t c1, uint32_t round, DevBuf& staging, DevBuf& counters, ChunkRecords& out) {
    cudaStream_t s = rep->stream;
    const uint64_t visible = std::min(my_bases, c1 + (uint64_t)std::max(k - 1, 0)) - c0, emit = c1 - c0;
    const uint8_t* d_src;
    if (host_bases) {
      PA_TRY(staging.alloc(visible + 64));
      PA_CUDA(cudaMemcpyAsync(staging.p, bases + c0, visible, cudaMemcpyHostToDevice, s));
      d_src = staging.as<uint8_t>();
    } else {
      d_src = bases + c0;
    }
    out.n = emit;
    out.tiles = (emit + SC_TILE - 1) / SC_TILE;
    if (out.keys.bytes < emit * 8) PA_TRY(out.keys.alloc(emit * 8));
    if (out.vals.bytes < emit * 4) PA_TRY(out.vals.alloc(emit * 4));
    if (out.owner.bytes < emit) PA_TRY(out.owner.alloc(emit));
    if (out.tile_counts.bytes < out.tiles * W * 4) PA_TRY(out.tile_counts.alloc(out.tiles * W * 4));
    if (out.tile_off.bytes < out.tiles * W * 8) PA_TRY(out.tile_off.alloc(out.tiles * W * 8));
    if (out.totals.bytes < (size_t)W * 8) PA_TRY(out.totals.alloc((size_t)W * 8));
    EncodeOpts opt{emit, out.owner.as<uint8_t>(), W * R, R, round, table_only ? 1u : 0u};
    PA_TRY(encode_slice(d_src, visible, my_base0 + c0, rep->genome_off.as<uint64_t>(), G, k, out.keys.as<uint64_t>(),
                        out.vals.as<uint32_t>(), opt, counters.as<unsigned long long>(), s));
    owner_count<<<(unsigned)out.tiles, SC_THREADS, 0, s>>>(out.owner.as<uint8_t>(), emit, W, out.tile_counts.as<uint32_t>());
    owner_scan<<<W, 1024, 0, s>>>(out.tile_counts.as<uint32_t>(), out.tiles, W, out.tile_off.as<uint64_t>(),
                                  out.totals.as<unsigned long long>());
    PA_CUDA(cudaGetLastError());
    return ST_OK;
  }

  int32_t run(Index** partition_out);
  int32_t table_round(uint32_t round);
  int32_t finalize();
  int32_t gather_slices_compressed(const std::vector<uint64_t>& blk_off);
};

int32_t DistBuild::run(Index** partition_out) {
  cudaStream_t s = rep->stream;
  const auto t_all = std::chrono::steady_clock::now();
  const uint32_t me = (uint32_t)comm->rank;
  PA_CUDA(cudaFuncSetAttribute(owner_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));

  std::vector<std::pair<uint64_t, uint64_t>> chunks;   // [c0, c1) in my-slice coordinates; starts stay 32-byte aligned
  for (uint64_t c0 = 0; c0 < my_bases; c0 += chunk) chunks.push_back({c0, std::min(my_bases, c0 + chunk)});
  const bool keep_chunk = chunks.size() <= 1;          // one chunk: the records of the counting sweep are scattered as they are

  DevBuf staging, counters, d_route;
  PA_TRY(counters.alloc(16));
  PA_TRY(d_route.alloc(sizeof(Route) * SC_MAX_OWNERS));
  ChunkRecords rec;
  unsigned long long h_cnt[2] = {0, 0};

  for (uint32_t round = 0; round < R; ++round) {
    // ---- sweep 1: encode + count records per owner ----
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint64_t> my_counts(W, 0);
    std::vector<std::vector<uint64_t>> chunk_totals(chunks.size(), std::vector<uint64_t>(W, 0));
    PA_CUDA(cudaMemsetAsync(counters.p, 0, 16, s));
    for (size_t c = 0; c < chunks.size(); ++c) {
      PA_TRY(encode_chunk(chunks[c].first, chunks[c].second, round, staging, counters, rec));
      PA_CUDA(cudaMemcpyAsync(chunk_totals[c].data(), rec.totals.p, (size_t)W * 8, cudaMemcpyDeviceToHost, s));
      PA_CUDA(cudaStreamSynchronize(s));
      for (uint32_t o = 0; o < W; ++o) my_counts[o] += chunk_totals[c][o];
    }
    PA_CUDA(cudaMemcpyAsync(h_cnt, counters.p, 16, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
    uint64_t bad = h_cnt[1] & 0xFFFFFFFFull;
    PA_TRY(comm->allreduce_u64_host(&bad, 1, false));
    if (bad) { set_error("genome sequence contains a character outside ACGTN"); return ST_BAD_BASE; }
    std::vector<uint64_t> matrix((size_t)W * W, 0);   // [sender][owner]
    PA_TRY(comm->allgather_host(my_counts.data(), matrix.data(), (size_t)W * 8));
    std::vector<uint64_t> need(W, 0);
    for (uint32_t snd = 0; snd < W; ++snd)
      for (uint32_t o = 0; o < W; ++o) need[o] += matrix[(size_t)snd * W + o];
    const uint64_t n_recv = need[me];
    tm.ms[0] += (float)ms_since(t0); t0 = std::chrono::steady_clock::now();

    // ---- sweep 2: the exchange ----
    PA_TRY(comm->ensure_exchange(need.data()));   // includes the barrier "every receive buffer has been consumed"
    uint64_t* recv_k = static_cast<uint64_t*>(comm->ex.my_k);
    uint32_t* recv_v = static_cast<uint32_t*>(comm->ex.my_v);
    DevBuf send_k, send_v;                        // only without peer mapping: local partition, then NCCL send / recv
    const bool direct = comm->single() || comm->ex.ipc;
    std::vector<uint64_t> seg_base(W, 0);         // where my segment starts in every owner's receive buffer
    for (uint32_t o = 0; o < W; ++o)
      for (uint32_t snd = 0; snd < me; ++snd) seg_base[o] += matrix[(size_t)snd * W + o];
    std::vector<uint64_t> send_off(W + 1, 0);
    for (uint32_t o = 0; o < W; ++o) send_off[o + 1] = send_off[o] + my_counts[o];
    if (!direct) { PA_TRY(send_k.alloc(std::max<uint64_t>(send_off[W], 1) * 8)); PA_TRY(send_v.alloc(std::max<uint64_t>(send_off[W], 1) * 4)); }
    std::vector<uint64_t> running(W, 0);
    for (size_t c = 0; c < chunks.size(); ++c) {
      if (!keep_chunk) PA_TRY(encode_chunk(chunks[c].first, chunks[c].second, round, staging, counters, rec));
      Route h_route[SC_MAX_OWNERS];
      for (uint32_t o = 0; o < W; ++o) {
        if (direct) {
          h_route[o].keys = static_cast<uint64_t*>(comm->ex.peer_k[o]) + seg_base[o] + running[o];
          h_route[o].vals = static_cast<uint32_t*>(comm->ex.peer_v[o]) + seg_base[o] + running[o];
        } else {
          h_route[o].keys = send_k.as<uint64_t>() + send_off[o] + running[o];
          h_route[o].vals = send_v.as<uint32_t>() + send_off[o] + running[o];
        }
        running[o] += chunk_totals[c][o];
      }
      PA_CUDA(cudaMemcpyAsync(d_route.p, h_route, sizeof(Route) * W, cudaMemcpyHostToDevice, s));
      if (rec.n)
        owner_scatter<<<(unsigned)rec.tiles, SC_THREADS, sizeof(ScatterSmem), s>>>(
            rec.keys.as<uint64_t>(), rec.vals.as<uint32_t>(), rec.owner.as<uint8_t>(), rec.n, W, rec.tile_off.as<uint64_t>(),
            d_route.as<Route>());
      PA_CUDA(cudaGetLastError());
      PA_CUDA(cudaStreamSynchronize(s));   // h_route is reused; the stores into peer memory are complete
    }
    if (!direct) {
      std::vector<uint64_t> recv_off(W + 1, 0);
      for (uint32_t snd = 0; snd < W; ++snd) recv_off[snd + 1] = recv_off[snd] + matrix[(size_t)snd * W + me];
      PA_TRY(comm->alltoallv_records(send_k.as<uint64_t>(), send_v.as<uint32_t>(), send_off.data(), recv_k, recv_v, recv_off.data(), s));
      PA_CUDA(cudaStreamSynchronize(s));
      send_k.release(); send_v.release();
    }
    PA_TRY(comm->barrier());               // every rank's stores have landed: the receive buffers are final
    tm.ms[1] += (float)ms_since(t0); t0 = std::chrono::steady_clock::now();

    // ---- K2 + K3 over what this rank owns in this round ----
    part->n_keys = part->n_runs = part->n_occ = 0;
    {
      DevBuf keys_b, vals_b, sort_tmp;
      const uint64_t* sk = recv_k; const uint32_t* sv = recv_v;
      if (n_recv) {
        PA_TRY(keys_b.alloc(n_recv * 8)); PA_TRY(vals_b.alloc(n_recv * 4));
        PA_TRY(sort_tmp.alloc(radix_sort_temp_bytes(n_recv)));
        int in_b = 0;
        PA_TRY(radix_sort_pairs_hashed(recv_k, recv_v, keys_b.as<uint64_t>(), vals_b.as<uint32_t>(), n_recv, std::min(64, 2 * k),
                                       sort_tmp.p, sort_tmp.bytes, s, &in_b));
        if (in_b) { sk = keys_b.as<uint64_t>(); sv = vals_b.as<uint32_t>(); }
        PA_CUDA(cudaStreamSynchronize(s));
      }
      tm.ms[2] += (float)ms_since(t0); t0 = std::chrono::steady_clock::now();
      PA_TRY(rle_to_csr(*part, sk, sv, n_recv, table_only));
      tm.ms[3] += (float)ms_since(t0); t0 = std::chrono::steady_clock::now();
    }
    valid_windows = (uint64_t)h_cnt[0];
    PA_TRY(table_round(round));
    tm.ms[4] += (float)ms_since(t0);
    if (table_only || R > 1) {   // the CSR of the round is done with
      part->ukeys.release(); part->run_off.release(); part->run_genome.release(); part->pos_off.release(); part->pos.release();
    }
  }
  rec.keys.release(); rec.vals.release(); rec.owner.release(); rec.tile_counts.release(); rec.tile_off.release();
  staging.release();
  PA_TRY(finalize());
  if (partition_out) *partition_out = (table_only || R > 1) ? nullptr : part;
  tm.ms[6] = (float)ms_since(t_all);
  return ST_OK;
}

// geometry (first call), genome sets and table insertion of part's current CSR into my slice of the replica's table
int32_t DistBuild::table_round(uint32_t round) {
  cudaStream_t s = rep->stream;
  const uint32_t me = (uint32_t)comm->rank;
  const uint64_t block_bytes = BLOCK_BUCKETS * BUCKET_SLOTS * 8;
  const CsrView csr{part->ukeys.as<uint64_t>(), part->run_off.as<uint64_t>(), part->run_genome.as<uint32_t>(), part->n_keys};
  GenomeSets sets;
  if (!have_table) {
    uint64_t sums[3] = {part->n_keys, part->n_occ, valid_windows};   // distinct + records of this round, valid windows of my genomes
    PA_TRY(comm->allreduce_u64_host(sums, 3, false));
    // distinct k-mers of the whole index: exact with one round; else round 0's distinct / record ratio (the owner is a
    // hash of the k-mer, so a round is a uniform sample of the key space) times all valid windows
    uint64_t U_est = sums[0];
    if (R > 1) U_est = sums[1] ? (uint64_t)((double)sums[0] / (double)sums[1] * (double)sums[2] * 1.01) + 1024 : sums[2];
    size_t free_b = 0, total_b = 0;
    PA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    uint64_t free_min = free_b + cache_held();   // buffers kept for reuse are memory a build may have
    PA_TRY(comm->allreduce_u64_host(&free_min, 1, true));
    const double load = table_load_factor(k, U_est, (size_t)free_min);
    uint32_t min_bpd = 1;
    for (;;) {
      PA_TRY(table_geometry(k, rep->n_genomes, U_est, load, min_bpd, &geom));
      PA_TRY(build_genome_sets(s, csr, geom.n_inline, &sets));
      uint64_t msec_all = sets.n_msec;
      PA_TRY(comm->allreduce_u64_host(&msec_all, 1, false));
      // later rounds hold about as many sets as this one; leave a factor of two
      if (ceil_log2_u64(msec_all * (uint64_t)R * (R > 1 ? 2 : 1) + 1) <= geom.payload_bits) break;
      if (geom.bpd >= 0x40000000u) { set_error("lookup table: list references do not fit (k=%d)", k); return ST_UNSUPPORTED; }
      min_bpd = geom.bpd * 2;
    }
    apply_geometry(*rep, geom);
    PA_TRY(rep->slots.alloc(rep->n_blocks() * block_bytes));
    // my slice: the digits of my parts (one per round), contiguous
    const uint64_t lo = (uint64_t)first_digit(me * R) * geom.bpd, hi = (uint64_t)first_digit((me + 1) * R) * geom.bpd;
    if (hi > lo) PA_CUDA(cudaMemsetAsync(rep->slots.as<char>() + lo * block_bytes, 0xFF, (hi - lo) * block_bytes, s));
    have_table = true;
  } else {
    PA_TRY(build_genome_sets(s, csr, geom.n_inline, &sets));
  }
  // global sector indices: pieces are laid out round-major, rank-minor
  std::vector<uint64_t> msecs(W, 0);
  const uint64_t my_msec = sets.n_msec;
  PA_TRY(comm->allgather_host(&my_msec, msecs.data(), 8));
  uint64_t my_sec_base = msec_cursor;
  for (uint32_t r = 0; r < W; ++r) { if (r < me) my_sec_base += msecs[r]; msec_cursor += msecs[r]; }
  if (ceil_log2_u64(msec_cursor + 1) > geom.payload_bits) { set_error("lookup table: list references do not fit (k=%d, round %u)", k, round); return ST_UNSUPPORTED; }
  const int32_t st = table_insert_csr(*rep, csr, sets, my_sec_base, &overflow);
  if (st == ST_CAPACITY) { set_error("lookup table: overloaded slice (lower PA_TABLE_LOAD)"); return ST_UNSUPPORTED; }
  PA_TRY(st);
  Piece pc; pc.base = my_sec_base; pc.ids.resize(sets.n_msec * MLIST_SECTOR);
  if (sets.n_msec) {
    PA_CUDA(cudaMemcpyAsync(pc.ids.data(), sets.mlist.p, sets.n_msec * 32, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  pieces.push_back(std::move(pc));
  total_keys += part->n_keys; total_runs += part->n_runs; total_occ += part->n_occ;
  return ST_OK;
}

// All-gather of the table slices in compressed form (see slice_bitmaps): blk_off[r] .. blk_off[r + 1] = the blocks of rank r
int32_t DistBuild::gather_slices_compressed(const std::vector<uint64_t>& blk_off) {
  cudaStream_t s = rep->stream;
  const uint32_t me = (uint32_t)comm->rank;
  const uint64_t n_blocks = blk_off[W], b_lo = blk_off[me], b_hi = blk_off[me + 1], nb_me = b_hi - b_lo;
  if (n_blocks >= 0xFFFFFFFFull / 2) { set_error("table gather: too many blocks"); return ST_UNSUPPORTED; }
  DevBuf bitmap, cnt, off_all, tile_sums, d_total;
  PA_TRY(bitmap.alloc((n_blocks + 1) * 8));
  PA_TRY(cnt.alloc((n_blocks + 1) * 4));
  PA_TRY(off_all.alloc((n_blocks + 1) * 8));
  PA_TRY(tile_sums.alloc((scan_tiles(n_blocks) + 1) * 8));
  PA_TRY(d_total.alloc(8));
  const uint64_t* slots = rep->slots.as<uint64_t>();
  const bool trace = getenv("PA_TRACE") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  double t_compress = 0, t_bitmaps = 0, t_words = 0, t_expand = 0;
  auto warps_grid = [](uint64_t blocks) { return (unsigned)std::max<uint64_t>(1, (blocks * 32 + CMP_THREADS - 1) / CMP_THREADS); };
  // ---- my slice -> bitmaps + word counts -> offsets -> words ----
  uint64_t my_words = 0;
  if (nb_me) {
    slice_bitmaps<<<warps_grid((nb_me + CMP_BLOCKS - 1) / CMP_BLOCKS), CMP_THREADS, 0, s>>>(slots, b_lo, b_hi, bitmap.as<unsigned long long>(), cnt.as<uint32_t>());
    PA_TRY(exclusive_scan_u32(cnt.as<uint32_t>(), off_all.as<uint64_t>(), nb_me, tile_sums.as<uint64_t>(), d_total.as<uint64_t>(), s));
    PA_CUDA(cudaMemcpyAsync(&my_words, d_total.p, 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<uint64_t> nwords(W, 0), w_off(W + 1, 0);
  PA_TRY(comm->allgather_host(&my_words, nwords.data(), 8));
  for (uint32_t r = 0; r < W; ++r) w_off[r + 1] = w_off[r] + nwords[r];
  // every rank's words at r * stride: equal segments, so that the words travel in one ncclAllGather
  uint64_t stride = 0;
  for (uint32_t r = 0; r < W; ++r) stride = std::max(stride, nwords[r]);
  stride = (stride + 15) & ~15ull;
  DevBuf words;
  PA_TRY(words.alloc((stride * W + 1) * 8));
  if (nb_me)
    slice_words<<<warps_grid((nb_me + CMP_BLOCKS - 1) / CMP_BLOCKS), CMP_THREADS, 0, s>>>(slots, b_lo, b_hi, off_all.as<uint64_t>(), words.as<uint64_t>() + (size_t)me * stride);
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));
  t_compress = ms_since(tp); tp = std::chrono::steady_clock::now();
  // ---- all ranks' bitmaps (block order) and words: two in-place all-gathers ----
  std::vector<uint64_t> seg(W + 1, 0);
  for (uint32_t r = 0; r <= W; ++r) seg[r] = blk_off[r] * 8;
  PA_TRY(comm->allgatherv_device_inplace(bitmap.p, seg.data(), s));
  if (trace) { PA_CUDA(cudaStreamSynchronize(s)); t_bitmaps = ms_since(tp); tp = std::chrono::steady_clock::now(); }
  PA_TRY(comm->allgather_device_strided(words.p, stride * 8, s));
  if (trace) { PA_CUDA(cudaStreamSynchronize(s)); t_words = ms_since(tp); tp = std::chrono::steady_clock::now(); }
  // ---- expand every other rank's blocks into my table ----
  bitmap_counts<<<(unsigned)std::max<uint64_t>(1, (n_blocks + CMP_THREADS - 1) / CMP_THREADS), CMP_THREADS, 0, s>>>(
      bitmap.as<unsigned long long>(), n_blocks, cnt.as<uint32_t>());
  PA_TRY(exclusive_scan_u32(cnt.as<uint32_t>(), off_all.as<uint64_t>(), n_blocks, tile_sums.as<uint64_t>(), d_total.as<uint64_t>(), s));
  for (uint32_t r = 0; r < W; ++r) {
    if (r == me || blk_off[r + 1] == blk_off[r]) continue;
    const uint64_t nb = blk_off[r + 1] - blk_off[r];
    slice_expand<<<warps_grid((nb + EXP_BLOCKS - 1) / EXP_BLOCKS), CMP_THREADS, 0, s>>>(
        bitmap.as<unsigned long long>(), off_all.as<uint64_t>(), words.as<uint64_t>(), blk_off[r], blk_off[r + 1],
        (long long)((uint64_t)r * stride) - (long long)w_off[r], rep->slots.as<uint64_t>());
  }
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));
  t_expand = ms_since(tp);
  if (trace)
    fprintf(stderr, "[pa gather] rank %u: compress %.2f ms (%.2f GB of words), bitmaps %.2f ms, words %.2f ms, expand %.2f ms\n", me,
            t_compress, (double)my_words * 8 / 1e9, t_bitmaps, t_words, t_expand);
  return ST_OK;
}

// replicate: table slices in place, genome sets and stash entries through the host
int32_t DistBuild::finalize() {
  cudaStream_t s = rep->stream;
  const uint64_t block_bytes = BLOCK_BUCKETS * BUCKET_SLOTS * 8;

  auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<uint64_t> blk(W + 1, 0), off(W + 1, 0);
    for (uint32_t r = 0; r <= W; ++r) { blk[r] = (uint64_t)first_digit(r * R) * geom.bpd; off[r] = blk[r] * block_bytes; }
    const char* how = getenv("PA_TABLE_GATHER");   // "raw" / "ipc" / "nccl": ship the slices as they are
    if (W > 1 && !how) PA_TRY(gather_slices_compressed(blk));
    else PA_TRY(comm->allgatherv_device_inplace(rep->slots.p, off.data(), s));
  }
  {
    // my pieces as one blob: [n_pieces][base, n_ids][ids ...]
    std::vector<uint8_t> blob;
    auto put = [&](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); blob.insert(blob.end(), b, b + n); };
    const uint64_t np = pieces.size();
    put(&np, 8);
    for (const Piece& pc : pieces) { const uint64_t hdr[2] = {pc.base, (uint64_t)pc.ids.size()}; put(hdr, 16); if (!pc.ids.empty()) put(pc.ids.data(), pc.ids.size() * 4); }
    std::vector<uint64_t> sizes; std::vector<uint8_t> all;
    PA_TRY(comm->allgatherv_host(blob.data(), blob.size(), &sizes, &all));
    std::vector<uint32_t> mlist(std::max<uint64_t>(msec_cursor, 1) * MLIST_SECTOR, 0xFFFFFFFFu);
    size_t at = 0;
    for (uint32_t r = 0; r < W; ++r) {
      const uint8_t* p = all.data() + at;
      uint64_t n; memcpy(&n, p, 8); p += 8;
      for (uint64_t i = 0; i < n; ++i) {
        uint64_t hdr[2]; memcpy(hdr, p, 16); p += 16;
        if (hdr[1]) memcpy(mlist.data() + hdr[0] * MLIST_SECTOR, p, hdr[1] * 4);
        p += hdr[1] * 4;
      }
      at += sizes[r];
    }
    rep->n_msectors = msec_cursor;
    PA_TRY(rep->mlist.alloc(mlist.size() * 4));
    PA_CUDA(cudaMemcpyAsync(rep->mlist.p, mlist.data(), mlist.size() * 4, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> osizes; std::vector<uint8_t> oall;
    PA_TRY(comm->allgatherv_host(overflow.data(), overflow.size() * 8, &osizes, &oall));
    std::vector<uint64_t> pairs(oall.size() / 8);
    if (!pairs.empty()) memcpy(pairs.data(), oall.data(), oall.size());
    PA_TRY(table_build_stash(*rep, pairs));
  }
  uint64_t totals[3] = {total_keys, total_runs, total_occ};
  PA_TRY(comm->allreduce_u64_host(totals, 3, false));
  rep->n_keys = totals[0]; rep->n_runs = totals[1]; rep->n_occ = totals[2];
  rep->align_only = true;
  tm.ms[5] = (float)ms_since(t0);
  return ST_OK;
}


// a fresh Index shell on `device` (genome offsets rebased to 0, own stream)
static int32_t make_index(int32_t k, uint32_t G, const uint64_t* genome_off, int32_t device, Index** out) {
  *out = nullptr;
  if (k > 31) { set_error("k = %d is outside the built scope (k <= 31: one k-mer per 64-bit word)", k); return ST_UNSUPPORTED; }
  if (G && !genome_off) { set_error("genome_off is null"); return ST_INVALID_ARG; }
  for (uint32_t g = 0; g < G; ++g)
    if (genome_off[g + 1] < genome_off[g]) { set_error("genome_off is not monotonic"); return ST_INVALID_ARG; }
  PA_CUDA(cudaSetDevice(device));
  Index* ix = new (std::nothrow) Index();
  if (!ix) { set_error("out of host memory"); return ST_NOMEM; }
  ix->k = k; ix->device = device; ix->n_genomes = G;
  ix->mix = mix_params_for_k(k);
  ix->h_genome_off.assign(G + 1, 0);
  for (uint32_t g = 0; g <= G && G; ++g) ix->h_genome_off[g] = genome_off[g] - genome_off[0];
  ix->total_bases = G ? ix->h_genome_off[G] : 0;
  cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete ix; return ST_CUDA; }
  int32_t st = ix->genome_off.alloc((size_t)(G + 1) * 8);
  if (st == ST_OK) {
    e = cudaMemcpyAsync(ix->genome_off.p, ix->h_genome_off.data(), (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) { set_error("genome_off upload failed: %s", cudaGetErrorString(e)); st = ST_CUDA; }
  }
  if (st != ST_OK) { delete ix; return st; }
  *out = ix;
  return ST_OK;
}

int32_t build_partitioned(Comm* comm_in, const uint8_t* bases, const uint64_t* genome_off, uint32_t G, uint32_t g_lo, uint32_t g_hi,
                          int32_t k, int32_t device, uint32_t flags, uint32_t n_rounds, Index** partition, Index** replica) {
  if (!replica) { set_error("null output handle"); return ST_INVALID_ARG; }
  *replica = nullptr;
  if (partition) *partition = nullptr;
  if (g_lo > g_hi || g_hi > G) { set_error("genome range out of bounds"); return ST_INVALID_ARG; }
  const bool table_only = (flags & PA_BUILD_TABLE_ONLY) != 0;
  if (!table_only && !partition) { set_error("a partition handle is needed unless PA_BUILD_TABLE_ONLY is set"); return ST_INVALID_ARG; }
  Comm* local = nullptr;
  Comm* comm = comm_in;
  if (!comm) { PA_TRY(comm_create_callbacks(1, 0, device, nullptr, &local)); comm = local; }
  struct Guard { Comm* c; ~Guard() { delete c; } } guard{local};
  if (comm->device != device) { set_error("the communicator lives on device %d, the build was asked for device %d", comm->device, device); return ST_INVALID_ARG; }
  Index *rep = nullptr, *part = nullptr;
  PA_TRY(make_index(k, G, genome_off, device, &rep));
  int32_t st = make_index(k, G, genome_off, device, &part);
  if (st != ST_OK) { delete rep; return st; }
  part->no_tables = true;
  const uint64_t total = rep->total_bases;
  DistBuild b;
  b.comm = comm; b.rep = rep; b.part = part; b.bases = bases;
  b.host_bases = (flags & PA_BUILD_HOST_BASES) != 0; b.table_only = table_only;
  b.G = G; b.g_lo = g_lo; b.g_hi = g_hi; b.k = k;
  b.W = (uint32_t)comm->n_ranks;
  b.my_base0 = G ? rep->h_genome_off[g_lo] : 0;
  b.my_bases = (k >= 1 && G) ? rep->h_genome_off[g_hi] - rep->h_genome_off[g_lo] : 0;
  b.chunk = 1ull << 28;
  if (const char* e = getenv("PA_BUILD_CHUNK")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 4096) b.chunk = v & ~4095ull; }
  // rounds: a rank sorts about total / (W * R) records at a time; 2^29 of them (6 GB + 6 GB + CSR) is a comfortable pass
  uint32_t R = n_rounds;
  if (R == 0) R = table_only ? (uint32_t)std::max<uint64_t>(1, (total + ((uint64_t)b.W << 29) - 1) / ((uint64_t)b.W << 29)) : 1;
  if (!table_only && R != 1) { set_error("several rounds need PA_BUILD_TABLE_ONLY (a kept partition is the CSR of one round)"); st = ST_INVALID_ARG; }
  if (st == ST_OK && (uint64_t)b.W * R > (1ull << digit_bits_for_k(k < 1 ? 1 : k))) { set_error("%u ranks x %u rounds do not fit the %u-bit digit space of k = %d", b.W, R, digit_bits_for_k(k < 1 ? 1 : k), k); st = ST_INVALID_ARG; }
  if (st == ST_OK && !table_only && total >= 0xFFFFFFFFull) { set_error("index build: %llu bases exceed the 32-bit position space of a build that keeps positions (use PA_BUILD_TABLE_ONLY)", (unsigned long long)total); st = ST_UNSUPPORTED; }
  if (st == ST_OK && b.my_bases && !bases) { set_error("bases is null"); st = ST_INVALID_ARG; }
  if (st == ST_OK && !b.host_bases && (reinterpret_cast<uintptr_t>(bases) & 15) != 0) { set_error("device bases must be 16-byte aligned"); st = ST_INVALID_ARG; }
  b.R = R;
  Index* kept = nullptr;
  if (st == ST_OK) { AllocScope pool(rep->stream); st = b.run(&kept); }
  if (st != ST_OK) { delete rep; delete part; return st; }
  for (int i = 0; i < 8; ++i) rep->t_dist_ms[i] = b.tm.ms[i];
  if (kept) { if (partition) *partition = kept; else delete kept; }
  else delete part;
  *replica = rep;
  return ST_OK;
}

int32_t rebuild_replica(Comm* comm_in, Index* part, Index** replica) {
  *replica = nullptr;
  Comm* local = nullptr;
  Comm* comm = comm_in;
  if (!comm) { PA_TRY(comm_create_callbacks(1, 0, part->device, nullptr, &local)); comm = local; }
  struct Guard { Comm* c; ~Guard() { delete c; } } guard{local};
  Index* rep = nullptr;
  PA_TRY(make_index(part->k, part->n_genomes, part->h_genome_off.data(), part->device, &rep));
  DistBuild b;
  b.comm = comm; b.rep = rep; b.part = part; b.bases = nullptr; b.host_bases = false; b.table_only = false;
  b.G = part->n_genomes; b.g_lo = b.g_hi = 0; b.k = part->k; b.W = (uint32_t)comm->n_ranks; b.R = 1;
  b.my_base0 = b.my_bases = 0; b.chunk = 1ull << 28;
  int32_t st;
  { AllocScope pool(rep->stream); st = b.table_round(0); if (st == ST_OK) st = b.finalize(); }
  if (st != ST_OK) { delete rep; return st; }
  *replica = rep;
  return ST_OK;
}

}  // namespace pa

using namespace pa;
#define IDX(h) (reinterpret_cast<pa::Index*>(h))
#define COMM(h) (reinterpret_cast<pa::Comm*>(h))
#define NEED(cond, msg) do { if (!(cond)) { pa::set_error(msg); return PA_ERR_INVALID_ARG; } } while (0)

extern "C" {

int32_t pa_comm_unique_id(uint8_t id[128]) {
  NEED(id, "null argument");
  return comm_unique_id(id);
}

int32_t pa_comm_init(int32_t n_ranks, int32_t rank, const uint8_t id[128], int32_t device, pa_comm** out) {
  NEED(id && out, "null argument");
  Comm* c = nullptr;
  PA_TRY(comm_create_nccl(n_ranks, rank, id, device, &c));
  *out = reinterpret_cast<pa_comm*>(c);
  return PA_OK;
}

int32_t pa_comm_init_callbacks(int32_t n_ranks, int32_t rank, int32_t device, const pa_comm_callbacks* cb, pa_comm** out) {
  NEED(out, "null argument");
  Comm* c = nullptr;
  PA_TRY(comm_create_callbacks(n_ranks, rank, device, cb, &c));
  *out = reinterpret_cast<pa_comm*>(c);
  return PA_OK;
}

int32_t pa_comm_free(pa_comm* comm) {
  delete COMM(comm);
  return PA_OK;
}

int32_t pa_comm_info(pa_comm* comm, int32_t* n_ranks, int32_t* rank, int32_t* device, int32_t* has_nccl) {
  NEED(comm, "null communicator");
  if (n_ranks) *n_ranks = COMM(comm)->n_ranks;
  if (rank) *rank = COMM(comm)->rank;
  if (device) *device = COMM(comm)->device;
  if (has_nccl) *has_nccl = COMM(comm)->has_nccl() ? 1 : 0;
  return PA_OK;
}

int32_t pa_comm_allreduce_summary(pa_comm* comm, uint64_t* d_sum, uint64_t n_sum, uint64_t* d_min, uint64_t n_min, void* stream) {
  NEED(comm, "null communicator");
  NEED((n_sum == 0 || d_sum) && (n_min == 0 || d_min), "null device buffer");
  Comm& c = *COMM(comm);
  NEED(c.device >= 0, "a host-only communicator cannot reduce device memory");
  PA_CUDA(cudaSetDevice(c.device));
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
  PA_TRY(c.allreduce_u64_device(d_sum, n_sum, false, s));
  PA_TRY(c.allreduce_u64_device(d_min, n_min, true, s));
  if (!stream) PA_CUDA(cudaStreamSynchronize(s));
  return PA_OK;
}

int32_t pa_comm_allreduce_host(pa_comm* comm, uint64_t* values, uint64_t n, int32_t op) {
  NEED(comm, "null communicator");
  NEED(n == 0 || values, "null argument");
  NEED(op == 0 || op == 1, "op must be 0 (SUM) or 1 (MIN)");
  return COMM(comm)->allreduce_u64_host(values, n, op == 1);
}

int32_t pa_comm_allgather_host(pa_comm* comm, const void* in, void* out, uint64_t bytes_per_rank) {
  NEED(comm, "null communicator");
  NEED(bytes_per_rank == 0 || (in && out), "null argument");
  return COMM(comm)->allgather_host(in, out, (size_t)bytes_per_rank);
}

int32_t pa_comm_barrier(pa_comm* comm) {
  NEED(comm, "null communicator");
  return COMM(comm)->barrier();
}

int32_t pa_genome_shard(const uint64_t* genome_off, uint32_t n_genomes, int32_t n_ranks, int32_t rank, uint32_t* g_lo, uint32_t* g_hi) {
  NEED(g_lo && g_hi && (n_genomes == 0 || genome_off), "null argument");
  NEED(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank");
  genome_shard(genome_off, n_genomes, n_ranks, rank, g_lo, g_hi);
  return PA_OK;
}

int32_t pa_index_build_partitioned(pa_comm* comm, const uint8_t* bases, const uint64_t* genome_off, uint32_t n_genomes,
                                   uint32_t g_lo, uint32_t g_hi, int32_t k, int32_t device, uint32_t flags, uint32_t n_rounds,
                                   pa_index** partition, pa_index** replica) {
  Index *part = nullptr, *rep = nullptr;
  PA_TRY(build_partitioned(COMM(comm), bases, genome_off, n_genomes, g_lo, g_hi, k, device, flags, n_rounds,
                           partition ? &part : nullptr, &rep));
  if (partition) *partition = reinterpret_cast<pa_index*>(part);
  *replica = reinterpret_cast<pa_index*>(rep);
  return PA_OK;
}

int32_t pa_index_rebuild_replica(pa_comm* comm, pa_index* partition, pa_index** replica) {
  NEED(partition && replica, "null argument");
  Index* rep = nullptr;
  PA_CUDA(cudaSetDevice(IDX(partition)->device));
  PA_TRY(rebuild_replica(COMM(comm), IDX(partition), &rep));
  *replica = reinterpret_cast<pa_index*>(rep);
  return PA_OK;
}

int32_t pa_build_timings(pa_index* replica, float ms[8]) {
  NEED(replica && ms, "null argument");
  for (int i = 0; i < 8; ++i) ms[i] = IDX(replica)->t_dist_ms[i];
  return PA_OK;
}

int32_t pa_partition_of_kmer(int32_t k, const uint8_t* kmer_ascii, uint32_t n_parts, uint32_t* part) {
  NEED(part && kmer_ascii && k >= 1 && k <= 31 && n_parts >= 1, "bad argument");
  NEED(n_parts <= (1u << digit_bits_for_k(k)), "too many parts for this k");
  bool ok;
  const uint64_t raw = encode_kmer_host(kmer_ascii, k, &ok);
  NEED(ok, "k-mer with a base outside ACGT");
  TableView t{};
  minimizer_params(t, k);
  const uint32_t kmask = (1u << k) - 1;
  uint32_t mh, p;
  kmer_minimizer(t, (uint32_t)raw & kmask, (uint32_t)(raw >> k) & kmask, &mh, &p);
  *part = (uint32_t)(((uint64_t)(mh >> t.dshift) * n_parts) >> digit_bits_for_k(k));
  return PA_OK;
}

int32_t pa_build_exchange(pa_comm* comm_h, const uint64_t* d_keys, const uint32_t* d_vals, const uint8_t* d_owner, uint64_t n,
                          uint64_t** recv_keys, uint32_t** recv_vals, uint64_t* n_recv, void* stream) {
  NEED(comm_h && recv_keys && recv_vals && n_recv, "null argument");
  NEED(n == 0 || (d_keys && d_vals && d_owner), "null device buffer");
  Comm& c = *COMM(comm_h);
  PA_CUDA(cudaSetDevice(c.device));
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
  const uint32_t W = (uint32_t)c.n_ranks, me = (uint32_t)c.rank;
  const uint64_t tiles = std::max<uint64_t>(1, (n + SC_TILE - 1) / SC_TILE);
  DevBuf tile_counts, tile_off, totals, d_route;
  PA_TRY(tile_counts.alloc(tiles * W * 4)); PA_TRY(tile_off.alloc(tiles * W * 8)); PA_TRY(totals.alloc((size_t)W * 8));
  PA_TRY(d_route.alloc(sizeof(Route) * SC_MAX_OWNERS));
  PA_CUDA(cudaFuncSetAttribute(owner_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
  owner_count<<<(unsigned)tiles, SC_THREADS, 0, s>>>(d_owner, n, W, tile_counts.as<uint32_t>());
  owner_scan<<<W, 1024, 0, s>>>(tile_counts.as<uint32_t>(), tiles, W, tile_off.as<uint64_t>(), totals.as<unsigned long long>());
  PA_CUDA(cudaGetLastError());
  std::vector<uint64_t> mine(W, 0), matrix((size_t)W * W, 0), need(W, 0);
  PA_CUDA(cudaMemcpyAsync(mine.data(), totals.p, (size_t)W * 8, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  PA_TRY(c.allgather_host(mine.data(), matrix.data(), (size_t)W * 8));
  for (uint32_t snd = 0; snd < W; ++snd)
    for (uint32_t o = 0; o < W; ++o) need[o] += matrix[(size_t)snd * W + o];
  PA_TRY(c.ensure_exchange(need.data()));
  const bool direct = c.single() || c.ex.ipc;
  DevBuf send_k, send_v;
  std::vector<uint64_t> send_off(W + 1, 0);
  for (uint32_t o = 0; o < W; ++o) send_off[o + 1] = send_off[o] + mine[o];
  if (!direct) { PA_TRY(send_k.alloc(std::max<uint64_t>(send_off[W], 1) * 8)); PA_TRY(send_v.alloc(std::max<uint64_t>(send_off[W], 1) * 4)); }
  Route h_route[SC_MAX_OWNERS];
  for (uint32_t o = 0; o < W; ++o) {
    uint64_t seg = 0;
    for (uint32_t snd = 0; snd < me; ++snd) seg += matrix[(size_t)snd * W + o];
    h_route[o].keys = direct ? static_cast<uint64_t*>(c.ex.peer_k[o]) + seg : send_k.as<uint64_t>() + send_off[o];
    h_route[o].vals = direct ? static_cast<uint32_t*>(c.ex.peer_v[o]) + seg : send_v.as<uint32_t>() + send_off[o];
  }
  PA_CUDA(cudaMemcpyAsync(d_route.p, h_route, sizeof(Route) * W, cudaMemcpyHostToDevice, s));
  if (n)
    owner_scatter<<<(unsigned)tiles, SC_THREADS, sizeof(ScatterSmem), s>>>(d_keys, d_vals, d_owner, n, W, tile_off.as<uint64_t>(),
                                                                           d_route.as<Route>());
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(s));
  if (!direct) {
    std::vector<uint64_t> recv_off(W + 1, 0);
    for (uint32_t snd = 0; snd < W; ++snd) recv_off[snd + 1] = recv_off[snd] + matrix[(size_t)snd * W + me];
    PA_TRY(c.alltoallv_records(send_k.as<uint64_t>(), send_v.as<uint32_t>(), send_off.data(), static_cast<uint64_t*>(c.ex.my_k),
                               static_cast<uint32_t*>(c.ex.my_v), recv_off.data(), s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  PA_TRY(c.barrier());
  *recv_keys = static_cast<uint64_t*>(c.ex.my_k);
  *recv_vals = static_cast<uint32_t*>(c.ex.my_v);
  *n_recv = need[me];
  return PA_OK;
}

}  // extern "C"
