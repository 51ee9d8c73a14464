// capi.cu -- the extern "C" boundary declared in include/pa_b200.h.
// Plain pointers and sizes only; no torch types, no C++ exceptions across the ABI.
#include "../../include/pa_b200.h"
#include "align.cuh"
#include "hostpack.h"
#include "index.cuh"
#include "sort.cuh"
#include "table.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

namespace pa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---- scratch reuse inside a build (see AllocScope in index.cuh) ----
static thread_local int tl_scope_depth = 0;
static thread_local std::vector<std::pair<void*, size_t>>* tl_parked = nullptr;

AllocScope::AllocScope(cudaStream_t) : outermost(tl_scope_depth == 0) {
  static const bool disabled = getenv("PA_NO_CACHE") && *getenv("PA_NO_CACHE") == '1';
  if (outermost && !disabled) tl_parked = new std::vector<std::pair<void*, size_t>>();
  ++tl_scope_depth;
}
AllocScope::~AllocScope() {
  --tl_scope_depth;
  if (outermost && tl_parked) {
    std::vector<std::pair<void*, size_t>>* parked = tl_parked;
    tl_parked = nullptr;
    // every public entry point synchronises its stream before it returns, so what is parked here is idle
    for (auto& b : *parked)
      if (!cache_put(b.first, b.second)) cudaFree(b.first);
    delete parked;
  }
}

// ---- retired-buffer cache (see index.cuh) ----
namespace {
struct CacheEntry { void* p; size_t bytes; int device; };
std::mutex g_cache_m;
std::vector<CacheEntry> g_cache;
constexpr size_t CACHE_MIN = 1u << 20;   // megabyte-sized scratch (tile counters, genome map) is worth keeping too: a cudaMalloc costs
                                         // milliseconds on a process that holds tens of GB, whatever its size

size_t cache_limit(int device) {
  static double gb = getenv("PA_CACHE_GB") ? atof(getenv("PA_CACHE_GB")) : -1.0;
  if (gb >= 0) return (size_t)(gb * 1e9);
  // default: 45 % of the device memory -- the buffers of one whole configs[1] build (index with the load-0.1 table 49 GB +
  // the sort's double buffers) fit, so a rebuild finds every buffer it needs (a quarter held the round-1 index only)
  static size_t share[64] = {0};
  if (device < 0 || device >= 64) return 0;
  if (!share[device]) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    share[device] = total_b / 20 * 9;
  }
  return share[device];
}
}  // namespace

bool cache_take(size_t bytes, void** p, size_t* got) {
  if (bytes < CACHE_MIN) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  std::lock_guard<std::mutex> g(g_cache_m);
  size_t best = SIZE_MAX; int at = -1;
  for (int i = 0; i < (int)g_cache.size(); ++i) {
    const CacheEntry& e = g_cache[i];
    if (e.device == dev && e.bytes >= bytes && e.bytes <= bytes + bytes / 8 + std::min<size_t>(64u << 20, bytes) && e.bytes < best) { best = e.bytes; at = i; }
  }
  if (at < 0) return false;
  *p = g_cache[at].p; *got = best;
  g_cache.erase(g_cache.begin() + at);
  return true;
}

bool cache_put(void* p, size_t bytes) {
  if (bytes < CACHE_MIN) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  const size_t limit = cache_limit(dev);
  if (bytes > limit) return false;
  std::vector<void*> evict;
  {
    std::lock_guard<std::mutex> g(g_cache_m);
    size_t held = 0;
    for (const CacheEntry& e : g_cache) if (e.device == dev) held += e.bytes;
    for (size_t i = 0; i < g_cache.size() && held + bytes > limit;) {   // oldest first
      if (g_cache[i].device == dev) { held -= g_cache[i].bytes; evict.push_back(g_cache[i].p); g_cache.erase(g_cache.begin() + i); }
      else ++i;
    }
    g_cache.push_back({p, bytes, dev});
  }
  for (void* q : evict) cudaFree(q);
  return true;
}

void alloc_miss(size_t bytes) {
  static const bool on = getenv("PA_TRACE_ALLOC") != nullptr;
  if (on && bytes >= (1u << 20)) fprintf(stderr, "[pa alloc] cudaMalloc of %.1f MB (no parked or cached buffer fits)\n", (double)bytes / 1e6);
}

size_t cache_held() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  std::lock_guard<std::mutex> g(g_cache_m);
  size_t held = 0;
  for (const CacheEntry& e : g_cache) if (e.device == dev) held += e.bytes;
  return held;
}

void cache_trim() {
  std::vector<CacheEntry> all;
  { std::lock_guard<std::mutex> g(g_cache_m); all.swap(g_cache); }
  int cur = 0;
  (void)cudaGetDevice(&cur);
  for (const CacheEntry& e : all) { cudaSetDevice(e.device); cudaFree(e.p); }
  cudaSetDevice(cur);
}
bool scope_take(size_t bytes, void** p, size_t* got) {
  if (!tl_parked || bytes < (1u << 20)) return false;       // small buffers are not worth the bookkeeping
  size_t best = SIZE_MAX; int at = -1;
  for (int i = 0; i < (int)tl_parked->size(); ++i) {
    const size_t sz = (*tl_parked)[i].second;
    if (sz >= bytes && sz <= bytes + bytes / 2 + (64u << 20) && sz < best) { best = sz; at = i; }
  }
  if (at < 0) return false;
  *p = (*tl_parked)[at].first; *got = best;
  tl_parked->erase(tl_parked->begin() + at);
  return true;
}
bool scope_park(void* p, size_t bytes) {
  if (!tl_parked || bytes < (1u << 20)) return false;
  tl_parked->push_back({p, bytes});
  return true;
}

namespace {

__global__ void fill_offsets_kernel(uint64_t* __restrict__ off, uint64_t n, uint64_t first, uint64_t len) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) off[i] = first + i * len;
}

__global__ void debug_lookup_kernel(TableView t, MixParams mix, const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ n_genomes,
                                    uint32_t* __restrict__ first_genome) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = keys[i];
  uint32_t c = 0, g0 = 0xFFFFFFFFu;
  if (key != SENTINEL_KEY) {
    uint64_t v = table_lookup(t, unmix_key(key, mix));  // `key` is the hashed k-mer (pa_encode_kmers)
    if (v != LOOKUP_MISS) {
      const uint32_t kind = value_kind(t, v);
      const uint64_t payload = value_payload(t, v);
      if (kind == KIND_SPECIFIC) { c = 1; g0 = (uint32_t)payload; }
      else if (kind == KIND_INLINE) { c = inline_count(t, payload); g0 = (uint32_t)payload & ((1u << t.gbits) - 1); }
      else {
        uint64_t sector = payload;
        g0 = t.mlist[sector * MLIST_SECTOR] & ~LIST_END;
        for (bool more = true; more; ++sector)
          for (int j = 0; j < 8 && more; ++j) { ++c; if (t.mlist[sector * MLIST_SECTOR + j] & LIST_END) more = false; }
      }
    }
  }
  n_genomes[i] = c; first_genome[i] = g0;
}

int32_t new_index(int32_t k, uint32_t G, const uint64_t* genome_off, int32_t device, Index** out) {
  if (!out) { set_error("null output handle"); return ST_INVALID_ARG; }
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
    if (e != cudaSuccess) { set_error("genome_off upload failed: %s", cudaGetErrorString(e)); st = ST_CUDA; }
  }
  if (st != ST_OK) { delete ix; return st; }
  *out = ix;
  return ST_OK;
}

AlignParams clamp_params(const pa_align_params* p) {
  // Clamping keeps every comparison's outcome and makes the int64 products overflow-free:
  //   quality bytes are 0..255, so a threshold <= 0 never filters and one >= 256 always does;
  //   genome counts are 1..2^32, so max_genomes < 0 behaves like -1 and > 2^32 like 2^32;
  //   specific/total counts are < 2^32, so m or p beyond 2^33 behave like 2^33.
  AlignParams a{};
  auto cl = [](int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
  a.m = cl(p->m, 0, 1LL << 33);
  a.p = cl(p->p, -1, 1LL << 33);
  a.has_mrq = p->has_min_read_quality != 0; a.has_mkq = p->has_min_kmer_quality != 0; a.has_mg = p->has_max_genomes != 0;
  a.mrq = cl(p->min_read_quality, 0, 256);
  a.mkq = cl(p->min_kmer_quality, 0, 256);
  a.mg = cl(p->max_genomes, -1, 1LL << 32);
  return a;
}

}  // namespace
}  // namespace pa

using namespace pa;

#define IDX(h) (reinterpret_cast<pa::Index*>(h))
#define NEED(cond, msg) do { if (!(cond)) { pa::set_error(msg); return PA_ERR_INVALID_ARG; } } while (0)

extern "C" {

int32_t pa_abi_version(void) { return PA_ABI_VERSION; }

int32_t pa_last_error(char* buf, size_t n) {
  const char* e = get_error();
  size_t len = strlen(e);
  if (buf && n) { size_t c = std::min(len, n - 1); memcpy(buf, e, c); buf[c] = 0; }
  return (int32_t)len;
}

int32_t pa_trim_memory(void) {
  cache_trim();
  return PA_OK;
}

int32_t pa_device_count(int32_t* n) {
  NEED(n, "null argument");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e)); *n = 0; return PA_ERR_CUDA; }
  *n = c;
  return PA_OK;
}

int32_t pa_index_build_device(const uint8_t* d_bases, const uint64_t* genome_off, uint32_t n_genomes, int32_t k,
                              int32_t device, pa_index** out) {
  Index* ix = nullptr;
  PA_TRY(new_index(k, n_genomes, genome_off, device, &ix));
  if (ix->total_bases && !d_bases) { delete ix; set_error("bases is null"); return PA_ERR_INVALID_ARG; }
  if ((reinterpret_cast<uintptr_t>(d_bases) & 15) != 0) { delete ix; set_error("device bases must be 16-byte aligned"); return PA_ERR_INVALID_ARG; }
  int32_t st;
  { AllocScope pool(ix->stream); st = index_build_from_device_bases(*ix, d_bases + (n_genomes ? genome_off[0] : 0)); }
  if (st != ST_OK) { delete ix; return st; }
  *out = reinterpret_cast<pa_index*>(ix);
  return PA_OK;
}

int32_t pa_index_build(const uint8_t* bases, const uint64_t* genome_off, uint32_t n_genomes, int32_t k, int32_t device,
                       pa_index** out) {
  Index* ix = nullptr;
  PA_TRY(new_index(k, n_genomes, genome_off, device, &ix));
  if (ix->total_bases && !bases) { delete ix; set_error("bases is null"); return PA_ERR_INVALID_ARG; }
  int32_t st = ST_OK;
  {
    AllocScope pool(ix->stream);
    DevBuf d_bases;
    st = d_bases.alloc(ix->total_bases + 64);
    if (st == ST_OK && ix->total_bases) {
      cudaError_t e = cudaMemcpyAsync(d_bases.p, bases + genome_off[0], ix->total_bases, cudaMemcpyHostToDevice, ix->stream);
      if (e != cudaSuccess) { set_error("bases upload failed: %s", cudaGetErrorString(e)); st = ST_CUDA; }
    }
    if (st == ST_OK) st = index_build_from_device_bases(*ix, d_bases.as<uint8_t>());
    cudaStreamSynchronize(ix->stream);
  }
  if (st != ST_OK) { delete ix; return st; }
  *out = reinterpret_cast<pa_index*>(ix);
  return PA_OK;
}

static int32_t import_fill(Index* ix, uint64_t n_keys, uint64_t n_runs, uint64_t n_occ, const uint64_t* keys,
                           const uint64_t* run_off, const uint32_t* run_genome, const uint64_t* pos_off, const uint32_t* pos,
                           const uint64_t* first_occ) {
  cudaStream_t s = ix->stream;
  PA_TRY(ix->ukeys.alloc((n_keys + 1) * 8)); PA_TRY(ix->run_off.alloc((n_keys + 1) * 8));
  PA_TRY(ix->run_genome.alloc((n_runs + 1) * 4)); PA_TRY(ix->pos_off.alloc((n_runs + 1) * 8));
  PA_TRY(ix->pos.alloc((n_occ + 1) * 4));
  const uint64_t zero = 0;
  if (n_keys) {
    PA_CUDA(cudaMemcpyAsync(ix->ukeys.p, keys, n_keys * 8, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaMemcpyAsync(ix->run_off.p, run_off, (n_keys + 1) * 8, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaMemcpyAsync(ix->run_genome.p, run_genome, n_runs * 4, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaMemcpyAsync(ix->pos_off.p, pos_off, (n_runs + 1) * 8, cudaMemcpyHostToDevice, s));
    if (n_occ) PA_CUDA(cudaMemcpyAsync(ix->pos.p, pos, n_occ * 4, cudaMemcpyHostToDevice, s));
  } else {
    PA_CUDA(cudaMemcpyAsync(ix->run_off.p, &zero, 8, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaMemcpyAsync(ix->pos_off.p, &zero, 8, cudaMemcpyHostToDevice, s));
  }
  PA_CUDA(cudaStreamSynchronize(s));
  ix->n_keys = n_keys; ix->n_runs = n_runs; ix->n_occ = n_occ;
  if (first_occ && n_keys) {
    PA_TRY(ix->first_occ.alloc((n_keys + 1) * 8));
    PA_CUDA(cudaMemcpyAsync(ix->first_occ.p, first_occ, n_keys * 8, cudaMemcpyHostToDevice, s));
    PA_CUDA(cudaStreamSynchronize(s));
    ix->has_first_occ = true;
  }
  return index_build_tables(*ix);
}

int32_t pa_index_import(int32_t k, uint32_t n_genomes, const uint64_t* genome_off, uint64_t n_keys, uint64_t n_runs,
                        uint64_t n_occ, const uint64_t* keys, const uint64_t* run_off, const uint32_t* run_genome,
                        const uint64_t* pos_off, const uint32_t* pos, const uint64_t* first_occ, int32_t device,
                        pa_index** out) {
  Index* ix = nullptr;
  PA_TRY(new_index(k, n_genomes, genome_off, device, &ix));
  if (n_keys && (!keys || !run_off || !run_genome || !pos_off || !pos)) { delete ix; set_error("null CSR array"); return PA_ERR_INVALID_ARG; }
  if (n_keys >= 0xFFFFFFFFull) { delete ix; set_error("too many distinct k-mers"); return PA_ERR_UNSUPPORTED; }
  int32_t st;
  { AllocScope pool(ix->stream); st = import_fill(ix, n_keys, n_runs, n_occ, keys, run_off, run_genome, pos_off, pos, first_occ); }
  if (st != ST_OK) { delete ix; return st; }
  *out = reinterpret_cast<pa_index*>(ix);
  return PA_OK;
}

int32_t pa_index_free(pa_index* idx) {
  if (idx) {
    cudaSetDevice(IDX(idx)->device);
    cudaStreamSynchronize(IDX(idx)->stream);
    delete IDX(idx);
  }
  return PA_OK;
}

int32_t pa_index_info_get(pa_index* idx, pa_index_info* info) {
  NEED(idx && info, "null argument");
  Index& ix = *IDX(idx);
  memset(info, 0, sizeof(*info));
  info->k = ix.k; info->device = ix.device; info->n_genomes = ix.n_genomes;
  info->blocks_per_digit = ix.bpd; info->digit_bits = digit_bits_for_k(ix.k); info->n_blocks = ix.n_blocks();
  info->table_bytes = ix.slots.bytes + ix.stash.bytes + ix.mlist.bytes; info->align_only = ix.align_only ? 1 : 0;
  info->minimizer_len = ix.min_len; info->tag_bits = ix.tag_bits; info->stash_count = ix.stash_count;
  info->n_keys = ix.n_keys; info->n_runs = ix.n_runs; info->n_occ = ix.n_occ; info->total_bases = ix.total_bases;
  info->n_list_sectors = ix.n_msectors; info->device_bytes = ix.device_bytes();
  info->build_encode_ms = ix.t_encode_ms; info->build_sort_ms = ix.t_sort_ms;
  info->build_rle_ms = ix.t_rle_ms; info->build_table_ms = ix.t_table_ms;
  return PA_OK;
}

int32_t pa_index_export(pa_index* idx, uint64_t* keys, uint64_t* run_off, uint32_t* run_genome, uint64_t* pos_off,
                        uint32_t* pos, uint32_t* order, uint64_t* first_occ) {
  NEED(idx, "null index");
  Index& ix = *IDX(idx);
  NEED(!ix.align_only, "a table-only index holds no CSR (export the partitions instead)");
  PA_CUDA(cudaSetDevice(ix.device));
  cudaStream_t s = ix.stream;
  if (keys && ix.n_keys) PA_CUDA(cudaMemcpyAsync(keys, ix.ukeys.p, ix.n_keys * 8, cudaMemcpyDeviceToHost, s));
  if (run_off) PA_CUDA(cudaMemcpyAsync(run_off, ix.run_off.p, (ix.n_keys + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (run_genome && ix.n_runs) PA_CUDA(cudaMemcpyAsync(run_genome, ix.run_genome.p, ix.n_runs * 4, cudaMemcpyDeviceToHost, s));
  if (pos_off) PA_CUDA(cudaMemcpyAsync(pos_off, ix.pos_off.p, (ix.n_runs + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (pos && ix.n_occ) PA_CUDA(cudaMemcpyAsync(pos, ix.pos.p, ix.n_occ * 4, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  AllocScope pool(ix.stream);
  if (order) PA_TRY(index_export_order(ix, order));
  if (first_occ && ix.n_keys) {
    PA_TRY(index_ensure_first_occ(ix));
    PA_CUDA(cudaMemcpyAsync(first_occ, ix.first_occ.p, ix.n_keys * 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaStreamSynchronize(s));
  }
  return PA_OK;
}

int32_t pa_index_csr_device(pa_index* idx, uint64_t** d_keys, uint64_t** d_run_off, uint32_t** d_run_genome) {
  NEED(idx, "null index");
  Index& ix = *IDX(idx);
  NEED(!ix.align_only, "a table-only index holds no CSR");
  if (d_keys) *d_keys = ix.ukeys.as<uint64_t>();
  if (d_run_off) *d_run_off = ix.run_off.as<uint64_t>();
  if (d_run_genome) *d_run_genome = ix.run_genome.as<uint32_t>();
  return PA_OK;
}

int32_t pa_index_checksum(pa_index* idx, uint64_t sums[4]) {
  NEED(idx && sums, "null argument");
  NEED(!IDX(idx)->align_only, "a table-only index holds no CSR");
  PA_CUDA(cudaSetDevice(IDX(idx)->device));
  return index_checksum(*IDX(idx), sums);
}

int32_t pa_decode_kmers(int32_t k, const uint64_t* keys, uint64_t n, uint8_t* ascii) {
  NEED(k >= 0 && k <= 31, "k out of range");
  NEED(n == 0 || (keys && ascii), "null argument");
  static const char dec[4] = {'A', 'C', 'T', 'G'};
  const MixParams mix = mix_params_for_k(k);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t key = unmix_key(keys[i], mix);  // exported keys are the hashed k-mers the index is sorted by
    for (int j = 0; j < k; ++j) {
      uint32_t lo = (key >> j) & 1, hi = (key >> (k + j)) & 1;
      ascii[i * (uint64_t)k + j] = dec[(hi << 1) | lo];
    }
  }
  return PA_OK;
}

int32_t pa_encode_kmers(int32_t k, const uint8_t* ascii, uint64_t n, uint64_t* keys) {
  NEED(k >= 1 && k <= 31, "k out of range");
  NEED(n == 0 || (keys && ascii), "null argument");
  for (uint64_t i = 0; i < n; ++i) {
    bool ok;
    uint64_t key = encode_kmer_host(ascii + i * (uint64_t)k, k, &ok);
    keys[i] = ok ? mix_key(key, mix_params_for_k(k)) : SENTINEL_KEY;
  }
  return PA_OK;
}

int32_t pa_index_lookup(pa_index* idx, const uint8_t* kmers_ascii, uint64_t n, uint64_t* rank) {
  NEED(idx, "null index");
  NEED(n == 0 || (kmers_ascii && rank), "null argument");
  NEED(!IDX(idx)->align_only, "a table-only index holds no CSR (look the k-mer up in the partitions instead)");
  PA_CUDA(cudaSetDevice(IDX(idx)->device));
  return index_lookup_ranks(*IDX(idx), kmers_ascii, n, rank);
}

int32_t pa_index_entries(pa_index* idx, const uint64_t* ranks, uint64_t n, uint64_t* run_off, uint32_t* run_genome,
                         uint64_t run_cap, uint64_t* pos_off, uint32_t* pos, uint64_t pos_cap, uint64_t* run_total,
                         uint64_t* pos_total) {
  NEED(idx && run_total && pos_total, "null argument");
  NEED(n == 0 || (ranks && run_off), "null argument");
  Index& ix = *IDX(idx);
  NEED(!ix.align_only, "a table-only index holds no CSR (fetch the entries from the partitions instead)");
  PA_CUDA(cudaSetDevice(ix.device));
  cudaStream_t s = ix.stream;
  // a handful of k-mers at a time (one read's windows at most): plain small copies, no kernel
  std::vector<uint64_t> r01(2 * n), p_lo(n, 0);
  uint64_t runs = 0, npos = 0;
  for (uint64_t i = 0; i < n; ++i) {
    NEED(ranks[i] < ix.n_keys, "rank out of range");
    PA_CUDA(cudaMemcpyAsync(&r01[2 * i], ix.run_off.as<uint64_t>() + ranks[i], 16, cudaMemcpyDeviceToHost, s));
  }
  PA_CUDA(cudaStreamSynchronize(s));
  std::vector<std::vector<uint64_t>> poffs(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t c = r01[2 * i + 1] - r01[2 * i];
    poffs[i].resize(c + 1);
    PA_CUDA(cudaMemcpyAsync(poffs[i].data(), ix.pos_off.as<uint64_t>() + r01[2 * i], (c + 1) * 8, cudaMemcpyDeviceToHost, s));
    run_off[i] = runs;
    runs += c;
  }
  if (n) run_off[n] = runs;
  PA_CUDA(cudaStreamSynchronize(s));
  for (uint64_t i = 0; i < n; ++i) npos += poffs[i].back() - poffs[i].front();
  *run_total = runs; *pos_total = npos;
  if (runs > run_cap || npos > pos_cap || !run_genome || !pos_off || !pos) {
    if (runs == 0 && npos == 0) return PA_OK;
    set_error("entries: %llu runs / %llu positions needed", (unsigned long long)runs, (unsigned long long)npos);
    return PA_ERR_CAPACITY;
  }
  uint64_t at_r = 0, at_p = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t c = r01[2 * i + 1] - r01[2 * i], np_i = poffs[i].back() - poffs[i].front();
    if (c) PA_CUDA(cudaMemcpyAsync(run_genome + at_r, ix.run_genome.as<uint32_t>() + r01[2 * i], c * 4, cudaMemcpyDeviceToHost, s));
    if (np_i) PA_CUDA(cudaMemcpyAsync(pos + at_p, ix.pos.as<uint32_t>() + poffs[i].front(), np_i * 4, cudaMemcpyDeviceToHost, s));
    for (uint64_t j = 0; j < c; ++j) pos_off[at_r + j] = at_p + (poffs[i][j] - poffs[i].front());
    at_r += c; at_p += np_i;
  }
  pos_off[at_r] = at_p;
  PA_CUDA(cudaStreamSynchronize(s));
  return PA_OK;
}

int32_t pa_extsim_stats(pa_index* idx, const uint32_t* group, uint32_t n_groups, uint64_t* total, uint64_t* unique) {
  NEED(idx, "null index");
  NEED(IDX(idx)->n_genomes == 0 || group, "null group map");
  NEED(n_groups == 0 || (total && unique), "null output");
  NEED(!IDX(idx)->align_only, "a table-only index holds no CSR (EXTSIM runs on the partitions)");
  PA_CUDA(cudaSetDevice(IDX(idx)->device));
  return index_extsim_stats(*IDX(idx), group, n_groups, total, unique);
}

int32_t pa_extsim_pairwise(pa_index* idx, const uint32_t* group, uint32_t n_groups, uint64_t* inter) {
  NEED(idx, "null index");
  NEED(IDX(idx)->n_genomes == 0 || group, "null group map");
  NEED(n_groups == 0 || inter, "null output");
  NEED(!IDX(idx)->align_only, "a table-only index holds no CSR (EXTSIM runs on the partitions)");
  PA_CUDA(cudaSetDevice(IDX(idx)->device));
  return index_extsim_pairwise(*IDX(idx), group, n_groups, inter);
}

int32_t pa_index_drop_genomes(pa_index* idx, const uint8_t* keep) {
  NEED(idx, "null index");
  NEED(IDX(idx)->n_genomes == 0 || keep, "null keep mask");
  NEED(!IDX(idx)->align_only, "a table-only index holds no CSR (drop the genomes from the partitions, then pa_index_rebuild_replica)");
  PA_CUDA(cudaSetDevice(IDX(idx)->device));
  AllocScope pool(IDX(idx)->stream);
  return index_drop_genomes(*IDX(idx), keep);
}


int32_t pa_align_batch_device(pa_index* idx, const uint8_t* d_bases, const uint8_t* d_quals, const uint64_t* d_read_off,
                              uint64_t n_reads, uint64_t max_read_len, const pa_align_params* params,
                              uint64_t* d_words, uint32_t* d_list, uint64_t list_cap, uint64_t* d_state, void* stream,
                              int32_t* n_launches) {
  NEED(idx && params, "null argument");
  NEED(n_reads == 0 || (d_bases && d_read_off && d_words && d_state), "null device buffer");
  NEED(params->m >= 0, "m must be bigger than or equal to 0");
  Index& ix = *IDX(idx);
  NEED(!ix.no_tables, "a partition holds no lookup table: align against the replica");
  PA_CUDA(cudaSetDevice(ix.device));
  AlignParams prm = clamp_params(params);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : ix.stream;
  return align_batch_device(ix, d_bases, d_quals, d_read_off, n_reads, max_read_len, prm, d_words, d_list, list_cap,
                            reinterpret_cast<unsigned long long*>(d_state), reinterpret_cast<unsigned long long*>(d_state) + 2,
                            s, n_launches);
}

static int32_t ensure(DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return ST_OK;
  return b.alloc(bytes + bytes / 8 + 256);
}

int32_t pa_align_batch(pa_index* idx, const uint8_t* bases, const uint8_t* quals, const uint64_t* read_off,
                       uint64_t n_reads, const pa_align_params* params, uint64_t* out_words, uint32_t* out_list,
                       uint64_t list_cap, uint64_t* list_len, uint64_t counters[3]) {
  NEED(idx && params, "null argument");
  NEED(params->m >= 0, "m must be bigger than or equal to 0");
  if (list_len) *list_len = 0;
  if (n_reads == 0) return PA_OK;
  NEED(bases && read_off && out_words && counters, "null host buffer");
  Index& ix = *IDX(idx);
  NEED(!ix.no_tables, "a partition holds no lookup table: align against the replica");
  PA_CUDA(cudaSetDevice(ix.device));
  const bool need_q = params->has_min_read_quality || params->has_min_kmer_quality;
  NEED(!need_q || quals, "quality filters requested without quality data");
  const AlignParams prm = clamp_params(params);
  const bool trace = getenv("PA_TRACE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();

  // Chunked pipeline over N_SLOTS slots (each its own stream): chunk c uses slot c % N_SLOTS:
  //   [2-bit packing on the host cores] -> H2D(planes | bases, quals, offsets) -> K4 -> D2H(words).
  // The copy engines move later chunks in while the SMs align earlier ones; kernels are chained with an event
  // because they share the queue and the per-warp scratch.  The list cursor and the filter counters live in one
  // device state block shared by all chunks, so list offsets in the result words are global to the call.
  constexpr int N_SLOTS = Index::N_HOST_SLOTS;
  for (auto& sl : ix.slot) {
    if (!sl.stream) PA_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.kernel_done) PA_CUDA(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
    if (!sl.h2d_done) PA_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    if (!sl.copy_beg) PA_CUDA(cudaEventCreate(&sl.copy_beg));
    if (!sl.copy_end) PA_CUDA(cudaEventCreate(&sl.copy_end));
    sl.copy_bytes = 0;
  }
  PA_TRY(ensure(ix.host_list, std::max<uint64_t>(list_cap, 1) * 4));
  PA_TRY(ensure(ix.host_state, 64));
  PA_CUDA(cudaMemsetAsync(ix.host_state.p, 0, 40, ix.slot[0].stream));
  PA_CUDA(cudaEventRecord(ix.slot[N_SLOTS - 1].kernel_done, ix.slot[0].stream));  // "previous kernel" of chunk 0 = the memset
  uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>(n_reads / 20, 1u << 16), 1u << 20);
  if (const char* e = getenv("PA_CHUNK_READS")) { uint64_t v = strtoull(e, nullptr, 10); if (v) chunk = v; }
  // PA_HOST_PACK: 0 = never pack, 1 = always pack, unset / 2 = the two-resource rule below
  const int pack_mode = getenv("PA_HOST_PACK") ? atoi(getenv("PA_HOST_PACK")) : 2;
  const bool use_pack = pack_mode != 0;
  if (ix.pack_rate_gbs <= 0) ix.pack_rate_gbs = 5.6 * host_pack_threads();
  if (ix.link_rate_gbs <= 0) ix.link_rate_gbs = getenv("PA_LINK_GBS") ? atof(getenv("PA_LINK_GBS")) : 45.0;
  double link_busy_until = 0;   // host-clock estimate (ms since t0) of when the queued H2D transfers end

  uint64_t c = 0, n_packed = 0;
  double t_wait = 0, t_pack = 0, t_enq = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
  const bool fixed_chunk = getenv("PA_CHUNK_READS") != nullptr;
  for (uint64_t lo = 0, step = 0; lo < n_reads; lo += step, ++c) {
    // the last chunks are halved: what is left in flight when the host has enqueued everything is the tail of the call
    step = (!fixed_chunk && n_reads - lo <= 2 * chunk && chunk >= (1u << 17)) ? chunk / 2 : chunk;
    const uint64_t hi = std::min(n_reads, lo + step), n = hi - lo;
    Index::HostSlot& sl = ix.slot[c % N_SLOTS];
    Index::HostSlot& prev = ix.slot[(c + N_SLOTS - 1) % N_SLOTS];
    const uint64_t b0 = read_off[lo], nb = read_off[hi] - b0;
    uint64_t max_len = 0, min_len = 0;
    auto tv0 = now();
    NEED(scan_offsets(read_off, lo, hi, &max_len, &min_len, host_pack_threads()), "read_off is not monotonic");
    t_pack += ms_since(tv0);
    // Which way does this chunk travel?  Packing to 2-bit planes (hostpack.h) costs host time (all cores) and leaves
    // a quarter of the bytes for the link; raw ASCII costs no host time.  Two resources work in parallel -- the host
    // cores and the copy engine -- so the greedy rule is: pack while the link still has about one packing time of
    // transfers queued (the cores would idle otherwise), else send the chunk raw at once (the link would idle).
    // Both rates are measured as the call goes.  With 16 cores against PCIe 5 x16 this settles at ~2 packed : 1 raw;
    // a node whose ranks share few cores settles at mostly raw.  A chunk with a base outside ACGT always travels raw.
    auto tp = now();
    bool packed = false;
    const uint64_t n_words = planes_words(nb, n);
    const double t_now = ms_since(t0);
    const double pack_est_ms = (double)nb / (ix.pack_rate_gbs * 1e6);
    const bool link_backed_up = (link_busy_until - t_now) >= 0.8 * pack_est_ms;
    if (use_pack && (pack_mode == 1 || link_backed_up) && ix.k >= 1 && nb) {
      PA_CUDA(cudaEventSynchronize(sl.h2d_done));   // the staging buffer's previous transfer has left the host
      if (sl.h_planes_words < n_words) {
        if (sl.h_planes) { cudaFreeHost(sl.h_planes); sl.h_planes = nullptr; sl.h_planes_words = 0; }
        const uint64_t want = n_words + n_words / 8 + 1024;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&sl.h_planes), want * 4, cudaHostAllocDefault);
        if (e != cudaSuccess) { (void)cudaGetLastError(); sl.h_planes = nullptr; }
        else sl.h_planes_words = want;
      }
      if (sl.h_planes) {
        auto tk = now();
        packed = pack_reads_planes(bases, read_off, lo, hi, sl.h_planes, host_pack_threads());
        const double ms = ms_since(tk);
        if (ms > 0) ix.pack_rate_gbs = 0.5 * ix.pack_rate_gbs + 0.5 * ((double)nb / ms / 1e6);
      }
    }
    t_pack += ms_since(tp); tp = now();
    // slot buffers are free once the slot's previous chunk (c - N_SLOTS) has finished: its stream is in order
    PA_CUDA(cudaStreamSynchronize(sl.stream));
    t_wait += ms_since(tp); tp = now();
    if (sl.copy_bytes) {   // the slot's previous raw transfer has been timed: refresh the link rate
      float ms = 0;
      if (cudaEventElapsedTime(&ms, sl.copy_beg, sl.copy_end) == cudaSuccess && ms > 0)
        ix.link_rate_gbs = 0.5 * ix.link_rate_gbs + 0.5 * ((double)sl.copy_bytes / ms / 1e6);
      (void)cudaGetLastError();
      sl.copy_bytes = 0;
    }
    if (need_q) PA_TRY(ensure(sl.quals, nb + 64));
    PA_TRY(ensure(sl.off, (n + 1) * 8));
    PA_TRY(ensure(sl.words, n * 8));
    if (packed) {
      PA_TRY(ensure(sl.planes, n_words * 4));
      PA_CUDA(cudaMemcpyAsync(sl.planes.p, sl.h_planes, n_words * 4, cudaMemcpyHostToDevice, sl.stream));
      PA_CUDA(cudaEventRecord(sl.h2d_done, sl.stream));
    } else {
      PA_TRY(ensure(sl.bases, nb + 64));
      if (nb) {
        PA_CUDA(cudaEventRecord(sl.copy_beg, sl.stream));
        PA_CUDA(cudaMemcpyAsync(sl.bases.p, bases + b0, nb, cudaMemcpyHostToDevice, sl.stream));
        PA_CUDA(cudaEventRecord(sl.copy_end, sl.stream));
        sl.copy_bytes = nb;
      }
    }
    if (need_q && nb) PA_CUDA(cudaMemcpyAsync(sl.quals.p, quals + b0, nb, cudaMemcpyHostToDevice, sl.stream));
    if (min_len == max_len) {   // fixed-length reads (the usual FASTQ): the offsets are an arithmetic sequence, no copy
      fill_offsets_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, sl.stream>>>(sl.off.as<uint64_t>(), n + 1, b0, max_len);
      PA_CUDA(cudaGetLastError());
    } else {
      PA_CUDA(cudaMemcpyAsync(sl.off.p, read_off + lo, (n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
    }
    {
      const double bytes = (packed ? (double)n_words * 4 : (double)nb) + (need_q ? (double)nb : 0.0) +
                           (min_len == max_len ? 0.0 : (double)(n + 1) * 8);
      link_busy_until = std::max(link_busy_until, ms_since(t0)) + bytes / (ix.link_rate_gbs * 1e6);
    }
    PA_CUDA(cudaStreamWaitEvent(sl.stream, prev.kernel_done, 0));
    // the kernel indexes bases with the absolute offsets: rebase the pointers instead of rewriting the offsets
    PA_TRY(align_batch_device(ix, packed ? nullptr : sl.bases.as<uint8_t>() - b0, need_q ? sl.quals.as<uint8_t>() - b0 : nullptr,
                              sl.off.as<uint64_t>(), n, max_len, prm, sl.words.as<uint64_t>(), ix.host_list.as<uint32_t>(),
                              list_cap, ix.host_state.as<unsigned long long>(), ix.host_state.as<unsigned long long>() + 2,
                              sl.stream, nullptr, packed ? sl.planes.as<uint32_t>() : nullptr, b0));
    n_packed += packed;
    t_enq += ms_since(tp);
    PA_CUDA(cudaEventRecord(sl.kernel_done, sl.stream));
    PA_CUDA(cudaMemcpyAsync(out_words + lo, sl.words.p, n * 8, cudaMemcpyDeviceToHost, sl.stream));
  }
  auto tq = now();
  for (auto& sl : ix.slot) PA_CUDA(cudaStreamSynchronize(sl.stream));
  const double t_drain = ms_since(tq);
  uint64_t h_state[5];
  PA_CUDA(cudaMemcpy(h_state, ix.host_state.p, 40, cudaMemcpyDeviceToHost));
  if (list_len) *list_len = h_state[0];
  if (h_state[1] == 2) { set_error("align: a read is longer than the length the batch was sized for"); return PA_ERR_INVALID_ARG; }
  if (h_state[0] > list_cap) {
    set_error("out_list too small: %llu entries needed", (unsigned long long)h_state[0]);
    return PA_ERR_CAPACITY;
  }
  if (h_state[0]) {
    NEED(out_list, "null out_list");
    PA_CUDA(cudaMemcpy(out_list, ix.host_list.p, h_state[0] * 4, cudaMemcpyDeviceToHost));
  }
  counters[0] += h_state[2]; counters[1] += h_state[3]; counters[2] += h_state[4];
  if (trace)
    fprintf(stderr, "[pa_align_batch] %llu reads in %llu chunks (%llu packed, %d pack threads, pack %.0f GB/s, link %.0f GB/s; host: wait %.2f pack %.2f enqueue %.2f drain %.2f ms): %.3f ms\n", (unsigned long long)n_reads,
            (unsigned long long)c, (unsigned long long)n_packed, host_pack_threads(), ix.pack_rate_gbs, ix.link_rate_gbs, t_wait, t_pack, t_enq, t_drain, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return PA_OK;
}

int32_t pa_pack_reads(const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t* planes, uint64_t planes_cap,
                      int32_t* all_acgt) {
  NEED(read_off && planes && all_acgt, "null argument");
  NEED(n_reads == 0 || bases, "null argument");
  uint64_t mx = 0, mn = 0;
  NEED(scan_offsets(read_off, 0, n_reads, &mx, &mn, host_pack_threads()), "read_off is not monotonic");
  NEED(planes_cap >= planes_words(read_off[n_reads] - read_off[0], n_reads), "planes buffer too small");
  *all_acgt = pack_reads_planes(bases, read_off, 0, n_reads, planes, host_pack_threads()) ? 1 : 0;
  return PA_OK;
}

int32_t pa_align_batch_packed(pa_index* idx, const uint32_t* planes, const uint8_t* quals, const uint64_t* read_off,
                              uint64_t n_reads, const pa_align_params* params, uint64_t* out_words, uint32_t* out_list,
                              uint64_t list_cap, uint64_t* list_len, uint64_t counters[3]) {
  NEED(idx && params, "null argument");
  NEED(params->m >= 0, "m must be bigger than or equal to 0");
  if (list_len) *list_len = 0;
  if (n_reads == 0) return PA_OK;
  NEED(planes && read_off && out_words && counters, "null host buffer");
  Index& ix = *IDX(idx);
  NEED(!ix.no_tables, "a partition holds no lookup table: align against the replica");
  NEED(n_reads <= (1ull << 31), "packed input is limited to 2^31 reads per call");
  PA_CUDA(cudaSetDevice(ix.device));
  const bool need_q = params->has_min_read_quality || params->has_min_kmer_quality;
  NEED(!need_q || quals, "quality filters requested without quality data");
  const AlignParams prm = clamp_params(params);
  // the pipeline of pa_align_batch without its packing stage: chunk c -> slot c % N_SLOTS: H2D(plane words, [quals], offsets) -> K4 -> D2H(words)
  constexpr int N_SLOTS = Index::N_HOST_SLOTS;
  for (auto& sl : ix.slot) {
    if (!sl.stream) PA_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.kernel_done) PA_CUDA(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
    if (!sl.h2d_done) PA_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    if (!sl.copy_beg) PA_CUDA(cudaEventCreate(&sl.copy_beg));
    if (!sl.copy_end) PA_CUDA(cudaEventCreate(&sl.copy_end));
  }
  PA_TRY(ensure(ix.host_list, std::max<uint64_t>(list_cap, 1) * 4));
  PA_TRY(ensure(ix.host_state, 64));
  PA_CUDA(cudaMemsetAsync(ix.host_state.p, 0, 40, ix.slot[0].stream));
  PA_CUDA(cudaEventRecord(ix.slot[N_SLOTS - 1].kernel_done, ix.slot[0].stream));
  uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>(n_reads / 20, 1u << 16), 1u << 20);
  if (const char* e = getenv("PA_CHUNK_READS")) { uint64_t v = strtoull(e, nullptr, 10); if (v) chunk = v; }
  const uint64_t B0 = read_off[0];
  uint64_t c = 0;
  for (uint64_t lo = 0; lo < n_reads; lo += chunk, ++c) {
    const uint64_t hi = std::min(n_reads, lo + chunk), n = hi - lo;
    Index::HostSlot& sl = ix.slot[c % N_SLOTS];
    Index::HostSlot& prev = ix.slot[(c + N_SLOTS - 1) % N_SLOTS];
    const uint64_t b0 = read_off[lo], nb = read_off[hi] - b0;
    uint64_t max_len = 0, min_len = 0;
    NEED(scan_offsets(read_off, lo, hi, &max_len, &min_len, host_pack_threads()), "read_off is not monotonic");
    // the words of reads [lo, hi) inside the batch-wide layout (hostpack.h): from w_lo up to the start of read hi
    const uint64_t w_lo = 2 * ((b0 - B0) / 32 + lo), w_hi = 2 * ((read_off[hi] - B0) / 32 + hi);
    PA_CUDA(cudaStreamSynchronize(sl.stream));
    if (need_q) PA_TRY(ensure(sl.quals, nb + 64));
    PA_TRY(ensure(sl.off, (n + 1) * 8));
    PA_TRY(ensure(sl.words, n * 8));
    PA_TRY(ensure(sl.planes, (w_hi - w_lo + 2) * 4));
    PA_CUDA(cudaMemcpyAsync(sl.planes.p, planes + w_lo, (w_hi - w_lo) * 4, cudaMemcpyHostToDevice, sl.stream));
    if (need_q && nb) PA_CUDA(cudaMemcpyAsync(sl.quals.p, quals + b0, nb, cudaMemcpyHostToDevice, sl.stream));
    if (min_len == max_len) {
      fill_offsets_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, sl.stream>>>(sl.off.as<uint64_t>(), n + 1, b0, max_len);
      PA_CUDA(cudaGetLastError());
    } else {
      PA_CUDA(cudaMemcpyAsync(sl.off.p, read_off + lo, (n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
    }
    PA_CUDA(cudaStreamWaitEvent(sl.stream, prev.kernel_done, 0));
    // the kernel finds the words of its read i at 2 ((o - B0) / 32 + lo + i) - w_lo of the slot buffer: rebase the pointer
    PA_TRY(align_batch_device(ix, nullptr, need_q ? sl.quals.as<uint8_t>() - b0 : nullptr, sl.off.as<uint64_t>(), n, max_len, prm,
                              sl.words.as<uint64_t>(), ix.host_list.as<uint32_t>(), list_cap, ix.host_state.as<unsigned long long>(),
                              ix.host_state.as<unsigned long long>() + 2, sl.stream, nullptr, sl.planes.as<uint32_t>() - w_lo, B0, lo));
    PA_CUDA(cudaEventRecord(sl.kernel_done, sl.stream));
    PA_CUDA(cudaMemcpyAsync(out_words + lo, sl.words.p, n * 8, cudaMemcpyDeviceToHost, sl.stream));
  }
  for (auto& sl : ix.slot) PA_CUDA(cudaStreamSynchronize(sl.stream));
  uint64_t h_state[5];
  PA_CUDA(cudaMemcpy(h_state, ix.host_state.p, 40, cudaMemcpyDeviceToHost));
  if (list_len) *list_len = h_state[0];
  if (h_state[1] == 2) { set_error("align: a read is longer than the length the batch was sized for"); return PA_ERR_INVALID_ARG; }
  if (h_state[0] > list_cap) { set_error("out_list too small: %llu entries needed", (unsigned long long)h_state[0]); return PA_ERR_CAPACITY; }
  if (h_state[0]) {
    NEED(out_list, "null out_list");
    PA_CUDA(cudaMemcpy(out_list, ix.host_list.p, h_state[0] * 4, cudaMemcpyDeviceToHost));
  }
  counters[0] += h_state[2]; counters[1] += h_state[3]; counters[2] += h_state[4];
  return PA_OK;
}

int32_t pa_summary_reduce_device(const uint64_t* d_words, const uint32_t* d_list, uint64_t n_reads,
                                 uint64_t read_index_base, uint32_t n_genomes, uint64_t* d_stats, uint64_t* d_unique_reads,
                                 uint64_t* d_ambiguous_reads, uint64_t* d_first_seen, void* stream) {
  NEED(n_reads == 0 || (d_words && d_stats && d_unique_reads && d_ambiguous_reads && d_first_seen), "null device buffer");
  return summary_reduce_device(d_words, d_list, n_reads, read_index_base, n_genomes,
                               reinterpret_cast<unsigned long long*>(d_stats), reinterpret_cast<unsigned long long*>(d_unique_reads),
                               reinterpret_cast<unsigned long long*>(d_ambiguous_reads),
                               reinterpret_cast<unsigned long long*>(d_first_seen), reinterpret_cast<cudaStream_t>(stream));
}

int32_t pa_summary_reduce(pa_index* idx, const uint64_t* words, const uint32_t* list, uint64_t n_reads, uint64_t list_len,
                          uint64_t read_index_base, uint64_t stats[4], uint64_t* unique_reads, uint64_t* ambiguous_reads,
                          uint64_t* first_seen) {
  NEED(idx && stats, "null argument");
  Index& ix = *IDX(idx);
  const uint32_t G = ix.n_genomes;
  NEED(G == 0 || (unique_reads && ambiguous_reads && first_seen), "null output");
  PA_CUDA(cudaSetDevice(ix.device));
  cudaStream_t s = ix.stream;
  const size_t gb = (size_t)std::max<uint32_t>(G, 1) * 8;
  DevBuf d_words, d_list, d_acc;
  PA_TRY(d_words.alloc(std::max<uint64_t>(n_reads, 1) * 8));
  PA_TRY(d_list.alloc(std::max<uint64_t>(list_len, 1) * 4));
  PA_TRY(d_acc.alloc(32 + 3 * gb));
  unsigned long long* acc = d_acc.as<unsigned long long>();
  PA_CUDA(cudaMemsetAsync(acc, 0, 32 + 2 * gb, s));
  PA_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(acc) + 32 + 2 * gb, 0xFF, gb, s));
  if (n_reads) PA_CUDA(cudaMemcpyAsync(d_words.p, words, n_reads * 8, cudaMemcpyHostToDevice, s));
  if (list_len) PA_CUDA(cudaMemcpyAsync(d_list.p, list, list_len * 4, cudaMemcpyHostToDevice, s));
  unsigned long long* d_unique = acc + 4;
  unsigned long long* d_amb = d_unique + std::max<uint32_t>(G, 1);
  unsigned long long* d_first = d_amb + std::max<uint32_t>(G, 1);
  PA_TRY(summary_reduce_device(d_words.as<uint64_t>(), d_list.as<uint32_t>(), n_reads, read_index_base, G, acc, d_unique,
                               d_amb, d_first, s));
  PA_CUDA(cudaMemcpyAsync(stats, acc, 32, cudaMemcpyDeviceToHost, s));
  if (G) {
    PA_CUDA(cudaMemcpyAsync(unique_reads, d_unique, (size_t)G * 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaMemcpyAsync(ambiguous_reads, d_amb, (size_t)G * 8, cudaMemcpyDeviceToHost, s));
    PA_CUDA(cudaMemcpyAsync(first_seen, d_first, (size_t)G * 8, cudaMemcpyDeviceToHost, s));
  }
  PA_CUDA(cudaStreamSynchronize(s));
  return PA_OK;
}

int32_t pa_debug_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t n, int32_t end_bit, int32_t device) {
  NEED(n == 0 || (keys && vals), "null argument");
  if (n == 0) return PA_OK;
  PA_CUDA(cudaSetDevice(device));
  DevBuf ka, kb, va, vb, tmp;
  PA_TRY(ka.alloc(n * 8)); PA_TRY(kb.alloc(n * 8)); PA_TRY(va.alloc(n * 4)); PA_TRY(vb.alloc(n * 4));
  PA_TRY(tmp.alloc(radix_sort_temp_bytes(n)));
  PA_CUDA(cudaMemcpy(ka.p, keys, n * 8, cudaMemcpyHostToDevice));
  PA_CUDA(cudaMemcpy(va.p, vals, n * 4, cudaMemcpyHostToDevice));
  int in_b = 0;
  PA_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, end_bit, tmp.p,
                          tmp.bytes, 0, &in_b));
  PA_CUDA(cudaDeviceSynchronize());
  PA_CUDA(cudaMemcpy(keys, in_b ? kb.p : ka.p, n * 8, cudaMemcpyDeviceToHost));
  PA_CUDA(cudaMemcpy(vals, in_b ? vb.p : va.p, n * 4, cudaMemcpyDeviceToHost));
  return PA_OK;
}

int32_t pa_debug_sort_pairs_hashed(uint64_t* keys, uint32_t* vals, uint64_t n, int32_t end_bit, int32_t top_bits, int32_t device,
                                   int32_t* fell_back) {
  NEED(n == 0 || (keys && vals), "null argument");
  if (fell_back) *fell_back = 0;
  if (n == 0) return PA_OK;
  PA_CUDA(cudaSetDevice(device));
  DevBuf ka, kb, va, vb, tmp;
  PA_TRY(ka.alloc(n * 8)); PA_TRY(kb.alloc(n * 8)); PA_TRY(va.alloc(n * 4)); PA_TRY(vb.alloc(n * 4));
  PA_TRY(tmp.alloc(radix_sort_temp_bytes(n)));
  PA_CUDA(cudaMemcpy(ka.p, keys, n * 8, cudaMemcpyHostToDevice));
  PA_CUDA(cudaMemcpy(va.p, vals, n * 4, cudaMemcpyHostToDevice));
  int in_b = 0, fb = 0;
  PA_TRY(radix_sort_pairs_hashed(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, end_bit, tmp.p,
                                 tmp.bytes, 0, &in_b, top_bits, &fb));
  PA_CUDA(cudaDeviceSynchronize());
  if (fell_back) *fell_back = fb;
  PA_CUDA(cudaMemcpy(keys, in_b ? kb.p : ka.p, n * 8, cudaMemcpyDeviceToHost));
  PA_CUDA(cudaMemcpy(vals, in_b ? vb.p : va.p, n * 4, cudaMemcpyDeviceToHost));
  return PA_OK;
}

int32_t pa_debug_pack_reads(const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t* planes, uint64_t planes_cap,
                            int32_t n_threads, int32_t* all_acgt) {
  NEED(read_off && planes && all_acgt, "null argument");
  NEED(planes_cap >= planes_words(read_off[n_reads] - read_off[0], n_reads), "planes buffer too small");
  *all_acgt = pack_reads_planes(bases, read_off, 0, n_reads, planes, n_threads > 0 ? n_threads : host_pack_threads()) ? 1 : 0;
  return PA_OK;
}

int32_t pa_debug_minimizer(int32_t k, const uint8_t* kmers_ascii, uint64_t n, uint32_t* mhash, uint32_t* offset) {
  NEED(k >= 1 && k <= 31, "k out of range");
  NEED(n == 0 || (kmers_ascii && mhash && offset), "null argument");
  TableView t{};
  minimizer_params(t, k);
  const uint32_t kmask = (1u << k) - 1;
  for (uint64_t i = 0; i < n; ++i) {
    bool ok;
    const uint64_t raw = encode_kmer_host(kmers_ascii + i * (uint64_t)k, k, &ok);
    NEED(ok, "k-mer with a base outside ACGT");
    kmer_minimizer(t, (uint32_t)raw & kmask, (uint32_t)(raw >> k) & kmask, &mhash[i], &offset[i]);
  }
  return PA_OK;
}

int32_t pa_debug_table_lookup(pa_index* idx, const uint8_t* kmers_ascii, uint64_t n, uint32_t* n_genomes,
                              uint32_t* first_genome) {
  NEED(idx, "null index");
  NEED(n == 0 || (kmers_ascii && n_genomes && first_genome), "null argument");
  if (n == 0) return PA_OK;
  Index& ix = *IDX(idx);
  NEED(!ix.no_tables, "a partition holds no lookup table: query the replica");
  PA_CUDA(cudaSetDevice(ix.device));
  if (ix.k < 1 || ix.n_keys == 0) { for (uint64_t i = 0; i < n; ++i) { n_genomes[i] = 0; first_genome[i] = 0xFFFFFFFFu; } return PA_OK; }
  std::vector<uint64_t> q(n);
  PA_TRY(pa_encode_kmers(ix.k, kmers_ascii, n, q.data()));
  cudaStream_t s = ix.stream;
  DevBuf dq, dn, dg;
  PA_TRY(dq.alloc(n * 8)); PA_TRY(dn.alloc(n * 4)); PA_TRY(dg.alloc(n * 4));
  PA_CUDA(cudaMemcpyAsync(dq.p, q.data(), n * 8, cudaMemcpyHostToDevice, s));
  debug_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ix.view(), ix.mix, dq.as<uint64_t>(), n, dn.as<uint32_t>(), dg.as<uint32_t>());
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaMemcpyAsync(n_genomes, dn.p, n * 4, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaMemcpyAsync(first_genome, dg.p, n * 4, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return PA_OK;
}

}  // extern "C"
