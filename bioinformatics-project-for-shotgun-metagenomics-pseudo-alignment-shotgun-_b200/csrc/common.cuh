// common.cuh -- shared device/host helpers for the sm_100a k-mer index and
// pseudo-alignment kernels.  See DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>

namespace pa {

// ---------------------------------------------------------------------------
// status codes of the C ABI (include/pa_b200.h)
// ---------------------------------------------------------------------------
enum : int32_t {
  ST_OK = 0,
  ST_INVALID_ARG = -1,
  ST_BAD_BASE = -2,
  ST_CUDA = -3,
  ST_NOMEM = -4,
  ST_CAPACITY = -5,
  ST_UNSUPPORTED = -6,
};

void set_error(const char* fmt, ...);
const char* get_error();

#define PA_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      pa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);  \
      return (e__ == cudaErrorMemoryAllocation) ? pa::ST_NOMEM : pa::ST_CUDA;                      \
    }                                                                                              \
  } while (0)

#define PA_TRY(expr)                  \
  do {                                \
    int32_t s__ = (expr);             \
    if (s__ != pa::ST_OK) return s__; \
  } while (0)

constexpr uint64_t EMPTY64 = 0xFFFFFFFFFFFFFFFFULL;
constexpr uint64_t SENTINEL_KEY = 0xFFFFFFFFFFFFFFFFULL;  // window without a valid k-mer; sorts last
constexpr uint32_t LIST_END = 0x80000000u;                 // bit 31 marks the last genome id of an mlist entry
constexpr uint32_t MLIST_SECTOR = 8;                       // genome ids per 32-byte sector

// ---------------------------------------------------------------------------
// Base encoding.  code = (ascii >> 1) & 3  (A=0, C=1, T=2, G=3).  A k-mer key
// holds two k-bit planes: bits [0,k) = low code bits of bases 0..k-1, bits
// [k,2k) = high code bits.  Plane form is what a warp ballot produces, so the
// align kernel encodes 32 windows with three ballots and three funnel shifts.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool is_acgt(uint32_t c) {
  return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t c) { return (c >> 1) & 3u; }

inline uint64_t encode_kmer_host(const uint8_t* s, int k, bool* ok) {
  uint64_t lo = 0, hi = 0;
  *ok = true;
  for (int i = 0; i < k; ++i) {
    if (!is_acgt(s[i])) { *ok = false; return 0; }
    uint32_t c = base_code(s[i]);
    lo |= (uint64_t)(c & 1) << i;
    hi |= (uint64_t)(c >> 1) << i;
  }
  return (hi << k) | lo;
}

// ---------------------------------------------------------------------------
// Bijective mix on n = 2k bits: xorshift and odd multiplication mod 2^n are
// both invertible, so distinct k-mers keep distinct hashes and only the low
// tag_bits of the hash need to be stored next to the bucket index.
// ---------------------------------------------------------------------------
struct MixParams {
  uint64_t mask;   // 2^n - 1
  uint32_t shift;  // xorshift distance
};
__host__ __device__ __forceinline__ uint64_t mix_key(uint64_t x, const MixParams& p) {
  x ^= x >> p.shift;
  x = (x * 0x9E3779B97F4A7C15ULL) & p.mask;
  x ^= x >> p.shift;
  x = (x * 0xD6E8FEB86659FD93ULL) & p.mask;
  x ^= x >> p.shift;
  return x;
}

// ---------------------------------------------------------------------------
// Lookup table view (device pointers).  One bucket = one 32-byte sector = four
// 8-byte slots.  slot = (tag << val_bits) | value;  value = kind (2 bits) | payload:
//   KIND_SPECIFIC  k-mer of exactly one genome, payload = genome id
//   KIND_INLINE    2..n_inline genomes packed in the payload, gbits each, ascending;
//                  a field equal to its predecessor means "no more genomes"
//   KIND_MLIST     longer lists: payload = first 32-byte sector of the list in mlist
// so that the common cases resolve inside the one sector a lookup has to fetch anyway
// (a B200 moves a whole 128-byte line from HBM per missing sector; see
// profiles/r01_gather_ncu_dram_per_request.csv).
// Keys that do not fit their bucket live in the stash (full hashed key, linear
// probing); a bucket is only ever followed into the stash when it is full.
// ---------------------------------------------------------------------------
struct TableView {
  const uint64_t* buckets;
  const ulonglong2* stash;    // {hashed key (EMPTY64 = free), value} pairs, linear probing
  const uint32_t* mlist;
  uint64_t stash_mask;        // capacity - 1 (capacity is a power of two), 0 when there is no stash
  uint32_t stash_count;
  uint32_t tag_bits;
  uint32_t val_bits;
  uint32_t k;
  uint32_t gbits;     // bits per genome id inside an inline list
  uint32_t n_inline;  // longest inline list (1 = inline lists unused)
  MixParams mix;
};

enum : uint32_t { KIND_MLIST = 0, KIND_INLINE = 2, KIND_SPECIFIC = 3 };

// inverse of mix_key (host side: decoding exported keys back into k-mer strings)
inline uint64_t unmix_key(uint64_t x, const MixParams& p) {
  auto unxorshift = [&](uint64_t v) { uint64_t r = v; for (uint32_t s = p.shift; s < 64; s += p.shift) r = v ^ (r >> p.shift); return r & p.mask; };
  auto inv_odd = [](uint64_t a) { uint64_t inv = a; for (int i = 0; i < 6; ++i) inv *= 2 - a * inv; return inv; };  // Newton, mod 2^64
  x = unxorshift(x);
  x = (x * inv_odd(0xD6E8FEB86659FD93ULL)) & p.mask;
  x = unxorshift(x);
  x = (x * inv_odd(0x9E3779B97F4A7C15ULL)) & p.mask;
  x = unxorshift(x);
  return x;
}
inline MixParams mix_params_for_k(int k) {
  MixParams m;
  m.mask = (k < 1) ? 1 : ((2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1));
  m.shift = (uint32_t)(k < 1 ? 1 : k);
  return m;
}

constexpr uint64_t LOOKUP_MISS = 0xFFFFFFFFFFFFFFFFULL;

__device__ __forceinline__ void ld_sector_nc(const void* p, uint64_t (&s)[4]) {
  // one 256-bit request = one 32-byte sector (see profiles/r01_gather_roofline.jsonl)
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
}
__device__ __forceinline__ void ld_sector_u32_nc(const void* p, uint32_t (&s)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]) : "l"(p));
}

__host__ __device__ __forceinline__ uint64_t stash_slot(uint64_t h) { return (h * 0xA24BAED4963EE407ULL) >> 20; }

__device__ __forceinline__ ulonglong2 ld_stash_nc(const ulonglong2* p) {
  ulonglong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

// probe the stash starting at slot i (masked by the caller or not)
__device__ __forceinline__ uint64_t stash_lookup_from(const TableView& t, uint64_t h, uint64_t i) {
  for (;;) {
    i &= t.stash_mask;
    ulonglong2 e = ld_stash_nc(t.stash + i);
    if (e.x == h) return e.y;
    if (e.x == EMPTY64) return LOOKUP_MISS;
    ++i;
  }
}
__device__ __forceinline__ uint64_t stash_lookup(const TableView& t, uint64_t h) { return stash_lookup_from(t, h, stash_slot(h)); }

// Resolve a bucket that has already been loaded, without following it into the stash.  Returns the value field or
// LOOKUP_MISS; *overflow tells whether the stash has to be consulted (the bucket is full and held no match).
__device__ __forceinline__ uint64_t bucket_resolve_local(const TableView& t, const uint64_t (&s)[4], uint64_t h, bool* overflow) {
  const uint64_t tag = h & ((1ULL << t.tag_bits) - 1);
  const uint64_t vmask = (1ULL << t.val_bits) - 1;
  *overflow = false;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if ((s[i] >> t.val_bits) == tag && s[i] != EMPTY64) return s[i] & vmask;
  *overflow = s[3] != EMPTY64 && t.stash_count != 0;
  return LOOKUP_MISS;
}

__device__ __forceinline__ uint64_t bucket_resolve(const TableView& t, const uint64_t (&s)[4], uint64_t h) {
  bool overflow;
  uint64_t v = bucket_resolve_local(t, s, h, &overflow);
  return overflow ? stash_lookup(t, h) : v;
}

__device__ __forceinline__ uint64_t table_lookup(const TableView& t, uint64_t key) {
  uint64_t h = mix_key(key, t.mix);
  uint64_t s[4];
  ld_sector_nc(t.buckets + (h >> t.tag_bits) * 4, s);
  return bucket_resolve(t, s, h);
}

__host__ __device__ __forceinline__ uint32_t value_kind(const TableView& t, uint64_t v) { return (uint32_t)(v >> (t.val_bits - 2)) & 3u; }
__host__ __device__ __forceinline__ uint64_t value_payload(const TableView& t, uint64_t v) {
  return v & ((1ULL << (t.val_bits - 2)) - 1);
}
// number of genomes of an inline list
__device__ __forceinline__ uint32_t inline_count(const TableView& t, uint64_t payload) {
  const uint32_t gm = (1u << t.gbits) - 1;
  uint32_t c = 1, prev = (uint32_t)payload & gm;
  for (uint32_t i = 1; i < t.n_inline; ++i) {
    uint32_t g = (uint32_t)(payload >> (i * t.gbits)) & gm;
    if (g == prev) break;
    prev = g; ++c;
  }
  return c;
}

// ---------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// genome index of a global base position: largest g with off[g] <= pos (off has G+1 entries)
__device__ __forceinline__ uint32_t genome_of(const uint64_t* __restrict__ off, uint32_t G, uint64_t pos) {
  uint32_t lo = 0, hi = G;  // invariant: off[lo] <= pos < off[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

inline uint32_t ceil_log2_u64(uint64_t x) {
  uint32_t b = 0;
  while ((1ULL << b) < x && b < 63) ++b;
  return b;
}

}  // namespace pa
