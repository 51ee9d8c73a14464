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
  // 'A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54: the bytes 0x40 + i with bit i of 0x0010008A set
  const uint32_t i = c - 0x40u;
  return i < 32u && ((0x0010008Au >> i) & 1u);
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
// Lookup table view (device pointers) -- minimizer-bucketed.
//
// A B200 serves random table reads at a fixed rate of distinct 128-byte LINES per second; the lanes of one load
// instruction that fall into the same line are served together (profiles/r01_locality_roofline.jsonl: 8 lanes per
// line = 7x the lane-lookups/s of one lane per line), while the same line requested by separate in-flight
// instructions costs full price each time.  The k-mer windows of a read are looked up 32 consecutive windows per
// instruction, so the table puts the k-mers that are consecutive in a sequence next to each other:
//
//   minimizer of a k-mer = its m-mer (m = min(k, 16), w = k-m+1 <= 16 candidates) with the smallest bijective
//   2m-bit hash, leftmost on ties; consecutive windows share their minimizer (8.5 windows on average at w = 16,
//   27 % of the groups have all 16), and inside a group the minimizer offset p takes consecutive values.
//   digit  = top digit_bits (<= 8) of the minimizer hash: the unit of ownership of a multi-GPU build (rank r owns a
//            contiguous range of digits, i.e. a contiguous slice of the table) and the range a chain stays inside
//   block  = (hash * blocks_per_digit) >> (2m - digit_bits)   -- any number of blocks per digit, so the load factor
//            is what was asked for and not a power of two away from it (one block = 16 buckets = 512 bytes = 4 lines)
//   bucket = (p + hash) & 15                        (one bucket = one 32-byte sector = 4 slots of 8 bytes)
//            -- a group of consecutive windows reads consecutive sectors: 32 windows touch ~12 lines instead of 32
//   tag    = [k-mer without the minimizer's bases : 2(k-m)] [chain distance d : 2] [low hi_bits of the hash]
//            -- the hashes that share a block are consecutive integers, fewer than 2^hi_bits of them, so block,
//            bucket and tag identify the k-mer exactly (no false positives)
//   word   = tag | CONT | kind (2 bits) | payload
//     KIND_SPECIFIC  k-mer of exactly one genome, payload = genome id
//     KIND_INLINE    2..n_inline genomes packed in the payload, gbits each, ascending;
//                    a field equal to its predecessor means "no more genomes"
//     KIND_MLIST     longer lists: payload = first 32-byte sector of the (de-duplicated) genome set in mlist
// The four slots of a bucket absorb the k-mers of other minimizers that hash to the same block and the strain
// variants that share minimizer and offset; a k-mer whose bucket is full moves to the same bucket of the next
// block of its digit (d = 1..3, wrapping inside the digit's blocks, then the stash) and sets CONT on the last slot of
// every bucket it passed.  A lookup reads ONE sector and goes on only when that bucket is full, has no match AND has
// CONT set, so a miss costs one sector too.  The stash holds {raw k-mer key, value} pairs with linear probing.
// ---------------------------------------------------------------------------
constexpr uint32_t BLOCK_BUCKETS = 16;
constexpr uint32_t BUCKET_SLOTS = 4;
constexpr uint32_t CHAIN_BITS = 2;
constexpr uint32_t CHAIN_LEN = 1u << CHAIN_BITS;

struct TableView {
  const uint64_t* slots;      // 2^block_bits blocks of BLOCK_BUCKETS buckets of BUCKET_SLOTS slots
  const ulonglong2* stash;    // {raw k-mer key (EMPTY64 = free), value} pairs, linear probing
  const uint32_t* mlist;
  uint64_t stash_mask;        // capacity - 1 (capacity is a power of two), 0 when there is no stash
  uint32_t stash_count;
  uint32_t k;
  uint32_t m;          // minimizer length
  uint32_t w;          // minimizer candidates per k-mer = k - m + 1 (1..16)
  uint32_t bpd;        // blocks per digit (>= 1); the table has bpd << digit_bits blocks
  uint32_t dshift;     // 2m - digit_bits: hash bits below the digit
  uint32_t hi_bits;    // low bits of the minimizer hash kept in the tag: 2^hi_bits >= hashes per block
  uint32_t hmask;      // 2^hi_bits - 1
  uint32_t tag_bits;   // 2(k-m) + CHAIN_BITS + hi_bits
  uint32_t val_bits;   // 64 - tag_bits = CONT bit + 2 kind bits + payload
  uint32_t gbits;      // bits per genome id inside an inline list
  uint32_t n_inline;   // longest inline list (1 = inline lists unused)
  uint64_t inl_ones;   // bit 0 of every inline field: sum of 1 << (i * gbits), i < n_inline
  uint64_t inl_highs;  // the top bit of every inline field
  uint32_t mmask;      // 2^m - 1
  uint32_t hdrop;      // low bits of an m-mer that stay raw in its hash: max(0, 2m - 28)
  uint32_t ymask;      // 2^(2m - hdrop) - 1: the mixed part of the hash, also the minimizer order
  uint32_t yshift;     // xorshift distance of the mix: half of its width
};

enum : uint32_t { KIND_MLIST = 0, KIND_INLINE = 2, KIND_SPECIFIC = 3 };

constexpr uint32_t MINIMIZER_MAX = 16;
constexpr uint32_t DIGIT_BITS_MAX = 8;
__host__ __device__ __forceinline__ uint32_t minimizer_len_for_k(int k) { return k < 1 ? 1u : (k > (int)MINIMIZER_MAX ? MINIMIZER_MAX : (uint32_t)k); }
// digit of a minimizer hash = its top digit_bits: ownership unit of the multi-GPU build, range of a chain
__host__ __device__ __forceinline__ uint32_t digit_bits_for_k(int k) { const uint32_t b = 2 * minimizer_len_for_k(k); return b < DIGIT_BITS_MAX ? b : DIGIT_BITS_MAX; }

// the minimizer fields of a TableView (k, m, w, mmask and the split of the m-mer hash) for a k-mer length
__host__ __device__ __forceinline__ void minimizer_params(TableView& t, int k) {
  t.k = (uint32_t)k;                       // k <= 0: no k-mer exists, nothing is ever looked up
  t.m = minimizer_len_for_k(k);
  t.w = k >= 1 ? (uint32_t)k - t.m + 1 : 1;
  t.mmask = (1u << t.m) - 1;
  t.hdrop = 2 * t.m > 28 ? 2 * t.m - 28 : 0;
  t.ymask = (1u << (2 * t.m - t.hdrop)) - 1;
  t.yshift = (2 * t.m - t.hdrop + 1) / 2;
  t.dshift = 2 * t.m - digit_bits_for_k(k);
}

// Two bijective functions of an m-mer x (2m <= 32 bits):
//   mmer_order  the upper 2m - hdrop bits of x through an invertible xorshift / odd-multiplication mix.  Minimizers are
//               ordered by it alone: it has at most 28 bits, so the align kernel slides (order << 4 | offset) through one
//               32-bit shuffle per step.
//   mmer_hash   addresses the table: [low hdrop bits of x, raw][order * odd constant mod 2^(2m - hdrop)].
//               The minimizer of a k-mer is the SMALLEST order among its candidates, so the orders of minimizers crowd
//               near zero and their top bits are nearly constant; the (Fibonacci) multiplication spreads exactly such a
//               range evenly over its top bits.  The raw bits sit on top, where the table block and the owner rank are
//               selected: m-mers that differ only in those bits share their order, so a substitution there (strain
//               variants) keeps the minimizer in place -- with the raw bits below, all variants of a minimizer would
//               crowd into one block (measured: 48x the stash entries, K4 21 % slower).
constexpr uint32_t MMER_SPREAD = 0x9E3779B1u;
__host__ __device__ __forceinline__ uint32_t mmer_order(uint32_t x, const TableView& t) {
  uint32_t y = x >> t.hdrop;
  y ^= y >> t.yshift;
  y = (y * 0x7FEB352DU) & t.ymask;
  y ^= y >> t.yshift;
  return y;
}
// hash from the order and the raw m-mer bits (the align kernel holds the order of the winning candidate)
__host__ __device__ __forceinline__ uint32_t hash_from_order(uint32_t order, uint32_t x_low, const TableView& t) {
  return ((x_low & ((1u << t.hdrop) - 1)) << (2 * t.m - t.hdrop)) | ((order * MMER_SPREAD) & t.ymask);
}
__host__ __device__ __forceinline__ uint32_t mmer_hash(uint32_t x, const TableView& t) {
  return hash_from_order(mmer_order(x, t), x, t);
}

struct SlotAddr {
  uint64_t block;   // home block
  uint64_t tag;     // tag at chain distance 0; at distance d add (d << hi_bits)
  uint32_t bucket;  // bucket inside the block, 0..15 (the same in every block of the chain)
};

// home block / bucket / tag of a k-mer given as planes (lo, hi: exactly k bits each) and its minimizer (hash, offset p)
__host__ __device__ __forceinline__ SlotAddr slot_addr(const TableView& t, uint32_t lo, uint32_t hi, uint32_t mhash, uint32_t p) {
  SlotAddr a;
  const uint32_t km = t.k - t.m;                     // bases outside the minimizer
  const uint32_t below = (1u << p) - 1;              // p <= 15
  // drop the m bits at [p, p+m): bits below p stay, bit j >= p comes from bit j+m -- a bit select between x and x >> m
  const uint32_t rl = (lo & below) | ((lo >> t.m) & ~below);   // p + m <= k <= 31
  const uint32_t rh = (hi & below) | ((hi >> t.m) & ~below);
  const uint32_t rest = (rh << km) | rl;             // 2(k-m) <= 30 bits
  a.block = ((uint64_t)mhash * t.bpd) >> t.dshift;
  a.tag = ((uint64_t)rest << (CHAIN_BITS + t.hi_bits)) | (mhash & t.hmask);
  a.bucket = (p + mhash) & (BLOCK_BUCKETS - 1);
  return a;
}

// minimizer of a whole k-mer (sequential; build side and debug lookups -- the align kernel computes it with a
// sliding window over the read instead).  Leftmost candidate wins ties.
__host__ __device__ __forceinline__ void kmer_minimizer(const TableView& t, uint32_t lo, uint32_t hi, uint32_t* mhash, uint32_t* p) {
  uint32_t best = 0xFFFFFFFFu, bp = 0;
  for (uint32_t j = 0; j < t.w; ++j) {
    const uint32_t y = mmer_order((((hi >> j) & t.mmask) << t.m) | ((lo >> j) & t.mmask), t);
    if (j == 0 || y < best) { best = y; bp = j; }
  }
  *mhash = hash_from_order(best, lo >> bp, t);
  *p = bp;
}

// inverse of mix_key (decoding exported keys back into k-mer strings; the table build un-hashes the CSR keys)
__host__ __device__ inline uint64_t unmix_key(uint64_t x, const MixParams& p) {
  // inverses of the two odd multipliers modulo 2^64 (valid modulo every 2^n)
  const uint64_t inv2 = 0xCFEE444D8B59A89BULL;   // 0xD6E8FEB86659FD93 ^ -1
  const uint64_t inv1 = 0xF1DE83E19937733DULL;   // 0x9E3779B97F4A7C15 ^ -1
  auto unxorshift = [&](uint64_t v) { uint64_t r = v; for (uint32_t s = p.shift; s < 64; s += p.shift) r = v ^ (r >> p.shift); return r & p.mask; };
  x = unxorshift(x);
  x = (x * inv2) & p.mask;
  x = unxorshift(x);
  x = (x * inv1) & p.mask;
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

__host__ __device__ __forceinline__ uint64_t stash_slot(uint64_t key) { return (key * 0xA24BAED4963EE407ULL) >> 20; }

__device__ __forceinline__ ulonglong2 ld_stash_nc(const ulonglong2* p) {
  ulonglong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

// probe the stash for the raw k-mer key
__device__ __forceinline__ uint64_t stash_lookup(const TableView& t, uint64_t key) {
  uint64_t i = stash_slot(key);
  for (;;) {
    i &= t.stash_mask;
    ulonglong2 e = ld_stash_nc(t.stash + i);
    if (e.x == key) return e.y;
    if (e.x == EMPTY64) return LOOKUP_MISS;
    ++i;
  }
}

__device__ __forceinline__ const uint64_t* bucket_ptr(const TableView& t, uint64_t block, uint32_t bucket) {
  return t.slots + (block * BLOCK_BUCKETS + bucket) * BUCKET_SLOTS;
}

// Look for `tag` in one loaded bucket.  Returns the value field (kind | payload) or LOOKUP_MISS; *cont tells whether
// the chain goes on (the bucket is full, holds no match, and some k-mer moved past it).
__device__ __forceinline__ uint64_t bucket_resolve(const TableView& t, const uint64_t (&s)[4], uint64_t tag, bool* cont) {
  const uint64_t vmask = (1ULL << (t.val_bits - 1)) - 1;
  *cont = false;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if ((s[i] >> t.val_bits) == tag && s[i] != EMPTY64) return s[i] & vmask;
  *cont = s[3] != EMPTY64 && ((s[3] >> (t.val_bits - 1)) & 1);
  return LOOKUP_MISS;
}

// The same for a k-mer's home bucket (chain distance 0), branch-free.  There the chain field of `tag` is 00, so an empty
// slot (all ones) can never equal it and needs no test of its own; at most one slot matches.
__device__ __forceinline__ uint64_t bucket_resolve_home(const TableView& t, const uint64_t (&s)[4], uint64_t tag, bool* cont) {
  uint64_t hit = 0;
  bool found = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool m = (s[i] >> t.val_bits) == tag;
    hit = m ? s[i] : hit;
    found |= m;
  }
  *cont = !found && s[3] != EMPTY64 && ((s[3] >> (t.val_bits - 1)) & 1);
  return found ? (hit & (((uint64_t)1 << (t.val_bits - 1)) - 1)) : LOOKUP_MISS;
}

// Continue a lookup past its home bucket: the same bucket of the next blocks, then the stash.
// block at chain distance d (< CHAIN_LEN) of a home block: the next blocks of the same digit, wrapping inside the
// digit; mhash = the minimizer hash the home block came from (its top bits are the digit: no division here)
__host__ __device__ __forceinline__ uint64_t chain_block(const TableView& t, uint64_t home, uint32_t d, uint32_t mhash) {
  const uint64_t first = (uint64_t)(mhash >> t.dshift) * t.bpd;
  uint64_t local = home - first + d;
  while (local >= t.bpd) local -= t.bpd;
  return first + local;
}

static __device__ __noinline__ uint64_t lookup_chain(const TableView& t, const SlotAddr& a, uint32_t mhash, uint64_t raw_key) {
  for (uint32_t d = 1; d < CHAIN_LEN; ++d) {
    uint64_t s[4];
    ld_sector_nc(bucket_ptr(t, chain_block(t, a.block, d, mhash), a.bucket), s);
    bool cont;
    uint64_t v = bucket_resolve(t, s, a.tag | ((uint64_t)d << t.hi_bits), &cont);
    if (!cont) return v;
  }
  return t.stash_count ? stash_lookup(t, raw_key) : LOOKUP_MISS;
}

// full lookup of a raw k-mer key (planes: bits [0,k) low code bits, [k,2k) high code bits)
__device__ __forceinline__ uint64_t table_lookup(const TableView& t, uint64_t raw_key) {
  const uint32_t kmask = (t.k >= 32) ? 0xFFFFFFFFu : ((1u << t.k) - 1);
  const uint32_t lo = (uint32_t)raw_key & kmask, hi = (uint32_t)(raw_key >> t.k) & kmask;
  uint32_t mh, p;
  kmer_minimizer(t, lo, hi, &mh, &p);
  SlotAddr a = slot_addr(t, lo, hi, mh, p);
  uint64_t s[4];
  ld_sector_nc(bucket_ptr(t, a.block, a.bucket), s);
  bool cont;
  uint64_t v = bucket_resolve(t, s, a.tag, &cont);
  return cont ? lookup_chain(t, a, mh, raw_key) : v;
}

__host__ __device__ __forceinline__ uint32_t value_kind(const TableView& t, uint64_t v) { return (uint32_t)(v >> (t.val_bits - 3)) & 3u; }
__host__ __device__ __forceinline__ uint64_t value_payload(const TableView& t, uint64_t v) {
  return v & ((1ULL << (t.val_bits - 3)) - 1);
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
