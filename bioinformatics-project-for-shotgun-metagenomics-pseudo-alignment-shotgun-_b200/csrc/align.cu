// align.cu -- K4 (warp-per-read pseudo-alignment), quality_masks_kernel (the
// EXTQUALITY filters of a batch, evaluated ahead of K4: one bit per window, one
// byte per read) and K8 (summary reduction).
//
// Reference being replaced, per read: Read.mean_quality / kmer_quality /
// extract_kmer_references / generate_genome_counts / try_to_align_specific /
// validate_unique_mappings / pseudo_align (/root/reference/src/kmer.py:394-526)
// and the per-read part of PseudoAlignment.add_read_from_read_record
// (kmer.py:586-598); K8 replaces PseudoAlignment.get_summary (kmer.py:622-657).
//
// Bound: random 32-byte sector requests into the bucket table (one per k-mer
// window; measured roofline in profiles/r01_gather_roofline.jsonl).  All
// arithmetic is integer; the reference's float means are compared as
// sum < threshold * length, which is exact.
#include "align.cuh"
#include <cstdlib>

namespace pa {

namespace {

constexpr int AL_THREADS = 256;
constexpr int AL_WARPS = AL_THREADS / 32;
constexpr int AL_ROUNDS = 4;                 // windows per lane per super-round
constexpr int AL_SUPER = 32 * AL_ROUNDS;     // 128 windows per super-round
constexpr uint32_t NOPOS = 0xFFFFFFFFu;

struct WarpScratch {
  uint32_t* gtab;        // [G][4] = {S count, T count, first specific pos, first pos}
  uint32_t* touched;     // [G]
  unsigned long long* kset_key;  // [kset_cap]
  uint32_t* kset_pos;            // [kset_cap]
  uint32_t kset_mask;
};

__device__ __forceinline__ uint32_t vld(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void vst(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
__device__ __forceinline__ void vst64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
__device__ __forceinline__ void reset_entry(uint32_t* e) { vst(e, 0u); vst(e + 1, 0u); vst(e + 2, 0xFFFFFFFFu); vst(e + 3, 0xFFFFFFFFu); }

// walk an mlist entry: count the ids (needed by the max-genomes filter, kmer.py:425)
__device__ __forceinline__ uint32_t mlist_count(const uint32_t* __restrict__ mlist, uint64_t sector) {
  uint32_t c = 0;
  for (;;) {
    uint32_t ids[8];
    ld_sector_u32_nc(mlist + sector * MLIST_SECTOR, ids);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ++c;
      if (ids[i] & LIST_END) return c;
    }
    ++sector;
  }
}

// generate_genome_counts (kmer.py:431-442), warp-aggregated.  Every lane may hold one kept, de-duplicated k-mer of
// round `r` (window position pos = round_base + lane).  The lanes walk their genome lists in lock step; lanes naming
// the same genome are grouped with match.any and only the lowest lane of a group (which also holds the smallest
// position) updates that genome's table entry, so the 32 lanes of a read that all hit the same few genomes cost a
// handful of plain updates instead of 32-way serialised atomics.
__device__ __forceinline__ void count_round(const TableView& t, const WarpScratch& ws, uint32_t& nT, bool active,
                                            uint64_t val, uint32_t round_base, uint32_t lane) {
  const uint32_t kind = value_kind(t, val);
  const uint64_t payload = value_payload(t, val);
  const uint32_t gm = (1u << t.gbits) - 1;
  bool more = active;
  uint32_t i = 0;
  while (__any_sync(0xffffffffu, more)) {
    uint32_t g = 0x80000000u | lane;  // matches nobody
    bool have = more;
    if (more) {
      if (kind == KIND_SPECIFIC) {
        g = (uint32_t)payload; more = false;
      } else if (kind == KIND_INLINE) {
        g = (uint32_t)(payload >> (i * t.gbits)) & gm;
        more = (i + 1 < t.n_inline) && (((uint32_t)(payload >> ((i + 1) * t.gbits)) & gm) != g);
      } else {
        uint32_t id = __ldg(t.mlist + payload * MLIST_SECTOR + i);
        g = id & ~LIST_END; more = !(id & LIST_END);
      }
      ++i;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, g);
    const uint32_t spec_lanes = __ballot_sync(0xffffffffu, have && kind == KIND_SPECIFIC);
    const bool leader = have && (lane == (uint32_t)__ffs(peers) - 1);
    uint32_t oldT = 1;
    if (leader) {
      uint32_t* e = ws.gtab + (size_t)g * 4;
      oldT = vld(e + 1);
      vst(e + 1, oldT + __popc(peers));
      const uint32_t pos = round_base + lane;
      if (pos < vld(e + 3)) vst(e + 3, pos);
      const uint32_t sp = peers & spec_lanes;
      if (sp) {
        vst(e, vld(e) + __popc(sp));
        const uint32_t spos = round_base + (uint32_t)__ffs(sp) - 1;
        if (spos < vld(e + 2)) vst(e + 2, spos);
      }
    }
    const uint32_t first = __ballot_sync(0xffffffffu, leader && oldT == 0);
    if (leader && oldT == 0) vst(ws.touched + nT + __popc(first & ((1u << lane) - 1)), g);
    nT += __popc(first);
    __syncwarp();
  }
}

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { uint64_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { uint32_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}


// One doubling step of the sliding minimum over the per-position minimizer keys.  Entry (c, lane) stands for position
// 32c + lane and holds (order << 4 | offset of the best candidate seen so far, relative to this position); it takes the
// smaller of itself and the entry d positions to its right, whose offset grows by d (offsets stay below w <= 16, so the
// addition never reaches the order bits).  Equal orders compare by offset: the leftmost candidate wins, as in
// kmer_minimizer.
__device__ __forceinline__ void window_min_step(uint32_t (&key)[AL_ROUNDS + 1], uint32_t d, uint32_t lane) {
  const uint32_t src = (lane + d) & 31;
  const bool wrap = lane + d >= 32;
  // every chunk rotated by d lanes once; the neighbour d positions to the right is in the own chunk's rotation or,
  // for the last d lanes, in the next chunk's
  uint32_t rk[AL_ROUNDS + 1];
#pragma unroll
  for (int c = 0; c <= AL_ROUNDS; ++c) rk[c] = __shfl_sync(0xffffffffu, key[c], src);
#pragma unroll
  for (int c = 0; c <= AL_ROUNDS; ++c) {
    if (c < AL_ROUNDS) key[c] = min(key[c], (wrap ? rk[c + 1] : rk[c]) + d);
    else if (!wrap) key[c] = min(key[c], rk[c] + d);
  }
}

// minimizer (full hash, offset) of the window whose planes are wl / wh, from its slid key
__device__ __forceinline__ void window_minimizer(const TableView& t, uint32_t key, uint32_t wl, uint32_t* mhash, uint32_t* p) {
  *p = key & 15u;
  *mhash = hash_from_order(key >> 4, wl >> *p, t);
}

// rare continuation of a lookup whose home bucket was full, without a match, and has CONT set (kept out of line:
// the fast path only carries the tag)
__device__ __noinline__ uint64_t lookup_chain_window(const TableView& t, uint64_t raw, uint32_t mhash, uint32_t p) {
  const uint32_t kmask = (t.k >= 32) ? 0xFFFFFFFFu : ((1u << t.k) - 1);
  const SlotAddr a = slot_addr(t, (uint32_t)raw & kmask, (uint32_t)(raw >> t.k) & kmask, mhash, p);
  return lookup_chain(t, a, mhash, raw);
}


// ---------------------------------------------------------------------------
// Read input.  ASCII: `bases` indexed with the absolute read offsets.  PACKED (host path): the host already turned
// every read into the two bit planes the ballots below would produce (hostpack.cpp); read i of the chunk, starting
// at base offset o, owns the words planes[w0 .. w0 + 2 nw) with w0 = 2 ((o - base0) / 32 + i), nw = ceil(L / 32):
// nw low-plane words, then nw high-plane words.  Packed reads contain only ACGT (the host falls back to ASCII for a
// chunk that does not), so their "invalid" plane is zero.
// ---------------------------------------------------------------------------
struct ReadInput {
  const uint8_t* bases;
  const uint32_t* planes;
  uint64_t base0;
  uint64_t read0;   // index of the kernel's read 0 inside the packed buffer (a chunk of a batch that was packed as a whole)
};

template <bool PACKED> struct Prefetch { uint32_t v[PACKED ? 1 : AL_ROUNDS + 1]; };

// lane j < 5 loads low-plane word cb + j, lane 5 + j loads high-plane word cb + j (zero beyond the read)
__device__ __forceinline__ uint32_t load_plane_words(const uint32_t* __restrict__ planes, uint64_t w0, uint32_t nw, uint32_t cb,
                                                     uint32_t lane) {
  const bool high = lane >= AL_ROUNDS + 1;
  const uint32_t c = cb + (high ? lane - (AL_ROUNDS + 1) : lane);
  return (lane < 2 * (AL_ROUNDS + 1) && c < nw) ? __ldg(planes + w0 + (high ? nw : 0u) + c) : 0u;
}

template <bool PACKED>
__device__ __forceinline__ void prefetch_read(const ReadInput& in, bool valid, uint64_t read, uint64_t beg, uint64_t L,
                                              uint32_t lane, Prefetch<PACKED>& p) {
  if constexpr (PACKED) {
    p.v[0] = valid ? load_plane_words(in.planes, 2 * ((beg - in.base0) / 32 + in.read0 + read), (uint32_t)((L + 31) / 32), 0, lane) : 0u;
  } else {
    const uint8_t* src = in.bases + beg + lane;
    const uint32_t L32 = !valid ? 0u : (L > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)L);
#pragma unroll
    for (int c = 0; c <= AL_ROUNDS; ++c) p.v[c] = (32u * c + lane < L32) ? src[32 * c] : 0;
  }
}

// bit planes of the AL_ROUNDS + 1 chunks of 32 bases starting at base wbase of the read
template <bool PACKED>
__device__ __forceinline__ void encode_planes(const ReadInput& in, const Prefetch<PACKED>& p, uint64_t read, uint64_t beg,
                                              uint64_t L, uint64_t wbase, uint32_t lane, uint32_t (&lo)[AL_ROUNDS + 1],
                                              uint32_t (&hi)[AL_ROUNDS + 1], uint32_t (&inv)[AL_ROUNDS + 1]) {
  if constexpr (PACKED) {
    uint32_t w = p.v[0];
    if (wbase != 0)
      w = load_plane_words(in.planes, 2 * ((beg - in.base0) / 32 + in.read0 + read), (uint32_t)((L + 31) / 32), (uint32_t)(wbase / 32), lane);
#pragma unroll
    for (int c = 0; c <= AL_ROUNDS; ++c) {
      lo[c] = __shfl_sync(0xffffffffu, w, c);
      hi[c] = __shfl_sync(0xffffffffu, w, AL_ROUNDS + 1 + c);
      inv[c] = 0;
    }
  } else {
#pragma unroll
    for (int c = 0; c <= AL_ROUNDS; ++c) {
      const uint64_t bi = wbase + 32 * c + lane;
      const uint32_t ch = wbase == 0 ? p.v[c] : (bi < L ? in.bases[beg + bi] : 0);
      const uint32_t code = base_code(ch);
      lo[c] = __ballot_sync(0xffffffffu, code & 1u);
      hi[c] = __ballot_sync(0xffffffffu, code >> 1);
      inv[c] = __ballot_sync(0xffffffffu, !is_acgt(ch));
    }
  }
}

struct Emit {
  uint64_t* out_word;
  uint32_t* out_list;
  uint64_t out_cap;
  unsigned long long* cursor;  // [0] list cursor, [1] overflow flag
};

__device__ __forceinline__ uint64_t make_word(uint32_t type, uint64_t len, uint64_t payload) {
  return ((uint64_t)type << 62) | (len << 40) | payload;
}

// try_to_align_specific (kmer.py:444-462) + validate_unique_mappings (kmer.py:464-480) on the
// per-genome table of one read; writes the result word and, for lists longer than one, the list.
__device__ void decide_and_emit(const WarpScratch& ws, uint32_t nT, const AlignParams& prm, const Emit& em, uint64_t read,
                                uint32_t lane) {
  // pass 1: nS, top specific count (ties -> earliest first specific position), max total
  uint32_t nS = 0, maxT = 0;
  uint64_t best = 0;
  for (uint32_t i = lane; i < nT; i += 32) {
    uint32_t g = vld(ws.touched + i);
    const uint32_t* e = ws.gtab + (size_t)g * 4;
    uint32_t S = vld(e), T = vld(e + 1), fS = vld(e + 2);
    if (S) { ++nS; uint64_t key = ((uint64_t)S << 32) | (uint32_t)(NOPOS - fS); best = key > best ? key : best; }
    maxT = T > maxT ? T : maxT;
  }
  nS = warp_sum(nS);
  maxT = warp_max_u32(maxT);
  best = warp_max_u64(best);
  if (nS == 0) {  // kept k-mers but none specific: ambiguous with an empty list (kmer.py:461)
    if (lane == 0) em.out_word[read] = make_word(3, 0, 0);
    return;
  }
  const uint32_t topS = (uint32_t)(best >> 32), top_fS = NOPOS - (uint32_t)best;
  // pass 2: identify the top genome and the runner-up count
  uint32_t top_g = 0, second = 0;
  for (uint32_t i = lane; i < nT; i += 32) {
    uint32_t g = vld(ws.touched + i);
    const uint32_t* e = ws.gtab + (size_t)g * 4;
    uint32_t S = vld(e), fS = vld(e + 2);
    if (S == topS && fS == top_fS) top_g = g + 1;
    else if (S > second) second = S;
  }
  top_g = warp_max_u32(top_g) - 1;
  second = warp_max_u32(second);
  const bool unique = (nS == 1) || ((int64_t)topS >= (int64_t)second + prm.m);
  uint32_t Tm = 0;
  bool flip = false;
  if (unique && prm.p >= 0) {
    Tm = vld(ws.gtab + (size_t)top_g * 4 + 1);
    flip = (int64_t)maxT - (int64_t)Tm > prm.p;
  }
  if (unique && !flip) {
    if (lane == 0) em.out_word[read] = make_word(2, 1, top_g);
    return;
  }
  // ambiguous: ordered list.  not unique -> genomes with S>0 by first specific position (dict order of
  // kmer.py:461); flipped -> [mapped] + genomes with T >= T[mapped] by (first position, genome) (kmer.py:476-479).
  uint32_t n = 0;
  for (uint32_t i = lane; i < nT; i += 32) {
    uint32_t g = vld(ws.touched + i);
    const uint32_t* e = ws.gtab + (size_t)g * 4;
    n += flip ? (vld(e + 1) >= Tm) : (vld(e) > 0);
  }
  n = warp_sum(n);
  const uint32_t len = n + (flip ? 1u : 0u);
  unsigned long long off = 0;
  if (lane == 0) off = atomicAdd(em.cursor, (unsigned long long)len);
  off = __shfl_sync(0xffffffffu, off, 0);
  if (lane == 0) em.out_word[read] = make_word(3, len, off);
  if (off + len > em.out_cap) {
    if (lane == 0) atomicExch(em.cursor + 1, 1ULL);
    return;
  }
  uint32_t* dst = em.out_list + off;
  if (flip && lane == 0) dst[0] = top_g;
  if (flip) ++dst;
  for (uint32_t i = lane; i < nT; i += 32) {
    uint32_t g = vld(ws.touched + i);
    const uint32_t* e = ws.gtab + (size_t)g * 4;
    bool in = flip ? (vld(e + 1) >= Tm) : (vld(e) > 0);
    if (!in) continue;
    uint64_t key = flip ? (((uint64_t)vld(e + 3) << 32) | g) : (uint64_t)vld(e + 2);
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nT; ++j) {
      uint32_t gj = vld(ws.touched + j);
      const uint32_t* ej = ws.gtab + (size_t)gj * 4;
      bool inj = flip ? (vld(ej + 1) >= Tm) : (vld(ej) > 0);
      uint64_t kj = flip ? (((uint64_t)vld(ej + 3) << 32) | gj) : (uint64_t)vld(ej + 2);
      rank += (inj && kj < key);
    }
    dst[rank] = g;
  }
}

template <bool QUAL, bool PACKED, int MIN_BLOCKS>
__global__ void __launch_bounds__(AL_THREADS, MIN_BLOCKS)
align_kernel(TableView t, ReadInput in, const uint8_t* __restrict__ quals,
             const uint64_t* __restrict__ read_off, uint64_t n_reads, AlignParams prm, Emit em,
             unsigned long long* __restrict__ counters, unsigned char* __restrict__ scratch, uint64_t scratch_stride,
             uint32_t G, uint32_t kset_cap, int gtab_in_smem, int kset_in_smem,
             const uint32_t* __restrict__ queue, const unsigned long long* __restrict__ queue_count) {
  // queue != nullptr: process the reads queue[0 .. *queue_count) the fast kernel could not finish (general path);
  // queue == nullptr: process every read 0 .. n_reads
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint64_t n_items = queue ? (uint64_t)*queue_count : n_reads;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_global = (uint64_t)blockIdx.x * AL_WARPS + warp;
  const uint64_t n_warps = (uint64_t)gridDim.x * AL_WARPS;
  const int k = (int)t.k;
  const uint32_t kmask = (k >= 1 && k < 32) ? ((1u << k) - 1) : 0u;

  // carve the per-warp scratch (shared memory when it fits, else this warp's slice of the global scratch)
  WarpScratch ws;
  {
    unsigned char* gp = scratch + warp_global * scratch_stride;
    size_t kset_bytes = (size_t)kset_cap * 12, gtab_bytes = (size_t)G * 20;
    size_t per_warp_smem = (kset_in_smem ? kset_bytes : 0) + (gtab_in_smem ? gtab_bytes : 0);
    per_warp_smem = (per_warp_smem + 15) & ~(size_t)15;
    unsigned char* sp = dyn_smem + warp * per_warp_smem;
    if (kset_in_smem) { ws.kset_key = (unsigned long long*)sp; ws.kset_pos = (uint32_t*)(sp + (size_t)kset_cap * 8); sp += kset_bytes; }
    else { ws.kset_key = (unsigned long long*)gp; ws.kset_pos = (uint32_t*)(gp + (size_t)kset_cap * 8); gp += (kset_bytes + 15) & ~(size_t)15; }
    if (gtab_in_smem) { ws.gtab = (uint32_t*)sp; ws.touched = (uint32_t*)(sp + (size_t)G * 16); }
    else { ws.gtab = (uint32_t*)gp; ws.touched = (uint32_t*)(gp + (size_t)G * 16); }
    ws.kset_mask = kset_cap - 1;
    if (kset_in_smem) for (uint32_t i = lane; i < kset_cap; i += 32) { ws.kset_key[i] = EMPTY64; ws.kset_pos[i] = NOPOS; }
    if (gtab_in_smem) for (uint32_t i = lane; i < G; i += 32) { ws.gtab[i * 4] = 0; ws.gtab[i * 4 + 1] = 0; ws.gtab[i * 4 + 2] = NOPOS; ws.gtab[i * 4 + 3] = NOPOS; }
    __syncwarp();
  }

  unsigned long long c_drop = 0, c_nq = 0, c_nr = 0;  // per-lane partial counters

  // Software pipeline over this warp's reads: the offsets of read i+2 and the first 160 bases of read i+1 are
  // requested while read i is processed, so a read never starts by waiting on its own (sequential) input.
  uint64_t nx_beg = 0, nx_end = 0;      // offsets of the next read
  Prefetch<PACKED> nx_ch, cur_ch;       // its first AL_ROUNDS+1 chunks of bases (one byte per lane / one plane word per lane)
  uint64_t cur_beg = 0, cur_end = 0;
  uint64_t cur_read = 0, nx_read = 0;
  {
    uint64_t r0 = warp_global, r1 = warp_global + n_warps;
    if (r0 < n_items) { cur_read = queue ? queue[r0] : r0; cur_beg = read_off[cur_read]; cur_end = read_off[cur_read + 1]; }
    if (r1 < n_items) { nx_read = queue ? queue[r1] : r1; nx_beg = read_off[nx_read]; nx_end = read_off[nx_read + 1]; }
    prefetch_read<PACKED>(in, r0 < n_items, cur_read, cur_beg, cur_end - cur_beg, lane, cur_ch);
  }

  for (uint64_t item = warp_global; item < n_items; item += n_warps) {
    const uint64_t read = cur_read;
    const uint64_t beg = cur_beg;
    const uint64_t L = cur_end - cur_beg;
    const uint8_t* rq = QUAL ? quals + beg : nullptr;
    // issue the loads of the following reads (consumed at the bottom of the loop)
    uint64_t n2_beg = 0, n2_end = 0, n2_read = 0;
    {
      const uint64_t r1 = item + n_warps, r2 = item + 2 * n_warps;
      prefetch_read<PACKED>(in, r1 < n_items, nx_read, nx_beg, nx_end - nx_beg, lane, nx_ch);
      if (r2 < n_items) { n2_read = queue ? queue[r2] : r2; n2_beg = read_off[n2_read]; n2_end = read_off[n2_read + 1]; }
    }

    bool dropped = false;
    if (QUAL && prm.has_mrq) {  // Read.mean_quality() < min_read_quality  (kmer.py:587)
      uint64_t s = 0;
      for (uint64_t i = lane; i < L; i += 32) s += rq[i];
      s = warp_sum(s);
      if ((int64_t)s < prm.mrq * (int64_t)L) {
        if (lane == 0) { em.out_word[read] = 0; ++c_drop; }
        dropped = true;
      }
    }
    const uint64_t W = (!dropped && k >= 1 && L >= (uint64_t)k) ? L - k + 1 : 0;  // kmer.py:91-92
    const bool single = W <= AL_SUPER;
    bool slow_used = false;
    uint32_t nT = 0;  // genomes touched by this read (warp-uniform)
    uint32_t read_nq = 0, read_nr = 0;
    bool done = dropped;

    for (uint64_t wbase = 0; wbase < W; wbase += AL_SUPER) {
      // ---- bit planes of AL_ROUNDS+1 chunks of 32 bases (ballots over the ASCII bases, or the host-packed words) ----
      uint32_t lo[AL_ROUNDS + 1], hi[AL_ROUNDS + 1], inv[AL_ROUNDS + 1];
      uint32_t qex[AL_ROUNDS + 1];  // exclusive quality prefix at this lane's base, relative to wbase
      encode_planes<PACKED>(in, cur_ch, read, beg, L, wbase, lane, lo, hi, inv);
      if (QUAL && prm.has_mkq) {
        uint32_t carry = 0;
#pragma unroll
        for (int c = 0; c <= AL_ROUNDS; ++c) {
          const uint64_t bi = wbase + 32 * c + lane;
          uint32_t q = bi < L ? rq[bi] : 0, incl = q;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
          qex[c] = carry + incl - q;
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      // ---- minimizers: hash of the m-mer at every base position, sliding minimum over the w candidates ----
      uint32_t mkey[AL_ROUNDS + 1], mh[AL_ROUNDS], mpos[AL_ROUNDS];
#pragma unroll
      for (int c = 0; c <= AL_ROUNDS; ++c) {
        const uint32_t nl = c < AL_ROUNDS ? lo[c + 1] : 0u, nh = c < AL_ROUNDS ? hi[c + 1] : 0u;
        const uint32_t xl = __funnelshift_r(lo[c], nl, lane) & t.mmask;
        const uint32_t xh = __funnelshift_r(hi[c], nh, lane) & t.mmask;
        mkey[c] = mmer_order((xh << t.m) | xl, t) << 4;
      }
      {
        uint32_t span = 1;
        for (; 2 * span <= t.w; span <<= 1) window_min_step(mkey, span, lane);
        if (t.w > span) window_min_step(mkey, t.w - span, lane);   // two overlapping spans cover the w candidates
      }
      // ---- per window: quality filter, block / bucket / tag, one sector load ----
      uint64_t raw[AL_ROUNDS], tag[AL_ROUNDS];
      uint64_t sector[AL_ROUNDS][4];
      bool look[AL_ROUNDS];
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        const uint64_t s = wbase + 32 * r + lane;
        bool exists = s < W;
        bool qf = false;
        if (QUAL && prm.has_mkq) {  // kmer_quality(start, k) < min_kmer_quality, before the lookup (kmer.py:420-422)
          uint32_t tl = lane + k;   // prefix at base s + k lives in chunk r (tl < 32) or r + 1
          uint32_t p_a = __shfl_sync(0xffffffffu, qex[r], tl & 31);
          uint32_t p_b = __shfl_sync(0xffffffffu, qex[r + 1], tl & 31);
          uint32_t end = tl < 32 ? p_a : p_b;
          qf = exists && ((int64_t)(end - qex[r]) < prm.mkq * (int64_t)k);
          read_nq += qf;
        }
        uint32_t wl = __funnelshift_r(lo[r], lo[r + 1], lane) & kmask;
        uint32_t wh = __funnelshift_r(hi[r], hi[r + 1], lane) & kmask;
        uint32_t wi = __funnelshift_r(inv[r], inv[r + 1], lane) & kmask;
        look[r] = exists && !qf && wi == 0;
        raw[r] = ((uint64_t)wh << k) | wl;
        window_minimizer(t, mkey[r], wl, &mh[r], &mpos[r]);   // full hash and offset inside the window, 0 .. w-1
        const SlotAddr a = slot_addr(t, wl, wh, mh[r], mpos[r]);
        tag[r] = a.tag;
        if (look[r]) ld_sector_nc(bucket_ptr(t, a.block, a.bucket), sector[r]);
      }
      // ---- resolve; classify what can be decided without touching mlist ----
      uint64_t val[AL_ROUNDS];
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        val[r] = LOOKUP_MISS;
        if (look[r]) {
          bool cont;
          val[r] = bucket_resolve_home(t, sector[r], tag[r], &cont);
          if (cont) val[r] = lookup_chain_window(t, raw[r], mh[r], mpos[r]);
        }
      }
      uint32_t cnt[AL_ROUNDS];   // genomes of the k-mer; 0 = unknown yet (mlist kind)
      bool l_unknown = false, l_kept_spec = false, l_kept_multi = false;
      uint32_t l_filtered = 0;
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        cnt[r] = 0;
        if (val[r] == LOOKUP_MISS) continue;
        const uint32_t kind = value_kind(t, val[r]);
        if (kind == KIND_SPECIFIC) cnt[r] = 1;
        else if (kind == KIND_INLINE) cnt[r] = inline_count(t, value_payload(t, val[r]));
        if (cnt[r] == 0) {
          if (prm.has_mg) l_unknown = true; else l_kept_multi = true;
        } else if (prm.has_mg && (int64_t)cnt[r] > prm.mg) {  // kmer.py:425-427
          ++l_filtered;
        } else if (cnt[r] == 1) {
          l_kept_spec = true;
        } else {
          l_kept_multi = true;
        }
      }
      // ---- fast paths (reads of one super-round whose outcome needs no per-genome table) ----
      if (single && !__any_sync(0xffffffffu, l_unknown)) {
        const bool w_spec = __any_sync(0xffffffffu, l_kept_spec);
        const bool w_multi = __any_sync(0xffffffffu, l_kept_multi);
        if (!w_spec) {
          // no specific k-mer kept: unmapped when nothing was kept (kmer.py:516-517), else the specific-count dict
          // is empty and the read is ambiguous with an empty list (kmer.py:461)
          read_nr += l_filtered;
          if (lane == 0) em.out_word[read] = make_word(w_multi ? 3 : 1, 0, 0);
          done = true; break;
        }
        if (!w_multi) {
          // only specific k-mers kept.  If they all name one genome, S == T == {g: n}: unique whatever m and p
          // are, and duplicated k-mers cannot change that.
          uint32_t mine = 0xFFFFFFFFu;
          bool same = true;
#pragma unroll
          for (int r = 0; r < AL_ROUNDS; ++r)
            if (cnt[r] == 1 && !(prm.has_mg && prm.mg < 1)) {
              uint32_t g = (uint32_t)value_payload(t, val[r]);
              if (mine == 0xFFFFFFFFu) mine = g; else same &= (g == mine);
            }
          uint32_t have = __ballot_sync(0xffffffffu, mine != 0xFFFFFFFFu);
          uint32_t g0 = __shfl_sync(0xffffffffu, mine, __ffs(have) - 1);
          same &= (mine == 0xFFFFFFFFu) || (mine == g0);
          if (__all_sync(0xffffffffu, same)) {
            read_nr += l_filtered;
            if (lane == 0) em.out_word[read] = make_word(2, 1, g0);
            done = true; break;
          }
        }
      }

      // ---- general path: max-genomes filter, in-read de-duplication, per-genome counting ----
      slow_used = true;
      uint32_t slot[AL_ROUNDS];
      bool kept[AL_ROUNDS];
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        kept[r] = false;
        if (val[r] == LOOKUP_MISS) continue;
        if (prm.has_mg) {
          uint32_t c = cnt[r] ? cnt[r] : mlist_count(t.mlist, value_payload(t, val[r]));
          if ((int64_t)c > prm.mg) { ++read_nr; continue; }
        }
        kept[r] = true;
        // self.kmers[kmer] = ...: a repeated k-mer keeps the position of its first kept occurrence (kmer.py:429)
        const uint32_t pos = (uint32_t)(wbase + 32 * r + lane);
        uint32_t sl = ((((uint32_t)raw[r] ^ (uint32_t)(raw[r] >> 31)) * 0x9E3779B1u) >> 7) & ws.kset_mask;
        uint32_t probes = 0;
        for (;; ++probes) {
          if (probes > ws.kset_mask) break;   // set full: a read longer than the declared max_read_len (flagged below)
          unsigned long long old = atomicCAS(ws.kset_key + sl, (unsigned long long)EMPTY64, (unsigned long long)raw[r]);
          if (old == EMPTY64 || old == raw[r]) break;
          sl = (sl + 1) & ws.kset_mask;
        }
        if (probes > ws.kset_mask) { atomicExch(em.cursor + 1, 2ULL); kept[r] = false; continue; }
        atomicMin(ws.kset_pos + sl, pos);
        slot[r] = sl;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        const uint32_t pos = (uint32_t)(wbase + 32 * r + lane);
        // a repeated k-mer is represented by its first kept occurrence only
        const bool rep = kept[r] && vld(ws.kset_pos + slot[r]) == pos;
        count_round(t, ws, nT, rep, val[r], (uint32_t)(wbase + 32 * r), lane);
      }
      if (single) {
        // one super-round: decide now, while slot[] is still in scope for a targeted cleanup
        if (nT == 0) { if (lane == 0) em.out_word[read] = make_word(1, 0, 0); }
        else decide_and_emit(ws, nT, prm, em, read, lane);
        __syncwarp();
        for (uint32_t i = lane; i < nT; i += 32) reset_entry(ws.gtab + (size_t)vld(ws.touched + i) * 4);
#pragma unroll
        for (int r = 0; r < AL_ROUNDS; ++r)
          if (kept[r]) { vst64(ws.kset_key + slot[r], EMPTY64); vst(ws.kset_pos + slot[r], NOPOS); }
        __threadfence_block();
        __syncwarp();
        done = true;
      }
    }

    if (!done) {
      // reads longer than one super-round (or without any window)
      if (nT == 0) { if (lane == 0) em.out_word[read] = make_word(1, 0, 0); }
      else decide_and_emit(ws, nT, prm, em, read, lane);
      __syncwarp();
      if (slow_used) {
        for (uint32_t i = lane; i < nT; i += 32) reset_entry(ws.gtab + (size_t)vld(ws.touched + i) * 4);
        for (uint32_t i = lane; i < kset_cap; i += 32) { vst64(ws.kset_key + i, EMPTY64); vst(ws.kset_pos + i, NOPOS); }
        __threadfence_block();
        __syncwarp();
      }
    }
    c_nq += read_nq;
    c_nr += read_nr;
    // rotate the pipeline registers
    cur_beg = nx_beg; cur_end = nx_end; cur_read = nx_read;
    nx_beg = n2_beg; nx_end = n2_end; nx_read = n2_read;
    cur_ch = nx_ch;
  }

  c_drop = warp_sum(c_drop); c_nq = warp_sum(c_nq); c_nr = warp_sum(c_nr);
  if (lane == 0) {
    if (c_drop) atomicAdd(counters + 0, c_drop);
    if (c_nq) atomicAdd(counters + 1, c_nq);
    if (c_nr) atomicAdd(counters + 2, c_nr);
  }
}


// ===========================================================================
// K4 fast kernel.  Almost every read is decided by rules that need no per-genome table (SURVEY.md 8(a) rows
// 10-14; the cases are proved next to the code), so they run in a small kernel -- no shared memory, few registers,
// many resident warps -- and only the remaining reads are queued for the general kernel above:
//   * no kept k-mer                                    -> UNMAPPED            (kmer.py:516-517)
//   * kept k-mers, none specific                       -> AMBIGUOUS []        (kmer.py:461: the specific-count dict is empty)
//   * every kept specific k-mer names one genome g, and (p < 0, or no multi-genome k-mer kept, or every kept
//     multi-genome k-mer's set contains g)             -> UNIQUE [g]
//     proof: S = {g: n} has one entry, so try_to_align_specific maps uniquely whatever m is (kmer.py:452-454);
//     validate_unique_mappings (kmer.py:464-480) compares max(T) with T[g]: every kept distinct k-mer contains g, so
//     T[g] is the number of kept distinct k-mers and no genome can exceed it -- max(T) - T[g] = 0 <= p.  Repeated
//     k-mers inside the read change n and T[g] but not the argument.
// A read sequenced from genome g only carries k-mers of g (its private ones and the ones it shares), so the last
// rule covers all reads whose erroneous windows miss the index.
// ===========================================================================
#ifndef PA_FAST_MINB
#define PA_FAST_MINB 2
#endif
#ifndef PA_FAST_THREADS
#define PA_FAST_THREADS 256
#endif
constexpr int FA_THREADS = PA_FAST_THREADS;
constexpr int FA_WARPS = FA_THREADS / 32;
constexpr uint32_t NO_GENOME = 0xFFFFFFFFu;

// 8 genome ids of a set sector, cached in L1 (the sets are de-duplicated, so the hot ones stay resident)
__device__ __forceinline__ void ld_set_sector(const uint32_t* p, uint32_t (&ids)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  ids[0] = a.x; ids[1] = a.y; ids[2] = a.z; ids[3] = a.w; ids[4] = b.x; ids[5] = b.y; ids[6] = b.z; ids[7] = b.w;
}

// ---------------------------------------------------------------------------
// Stage C of the fast kernels: resolve the loaded sectors and apply the rules that need no per-genome table.
// VAL32: a slot word splits at bit 32 -- the high word is exactly [k-mer without its minimizer : 2(k-m) = 30 bits][chain
// distance : 2] and the low word [minimizer hash bits kept in the tag][CONT, kind, payload] (k = 31, m = 16: the case
// this path is tuned for).  Values, kinds and payloads are then 32-bit quantities, the tag word is built without 64-bit
// shifts, and a tag matches when the high words are equal and the low words agree above val_bits: three instructions per
// slot instead of two 64-bit shifts and a compare.  Every other geometry runs the same code on 64-bit words.
// ---------------------------------------------------------------------------
// block / bucket of a window and the tag as it sits in a slot word, as two 32-bit halves
struct SlotWord { uint32_t block, bucket, tw_lo, tw_hi; };
template <bool VAL32>
__device__ __forceinline__ SlotWord slot_word(const TableView& t, uint32_t lo, uint32_t hi, uint32_t mhash, uint32_t p) {
  SlotWord a;
  if constexpr (VAL32) {
    const uint32_t km = t.k - t.m;
    const uint32_t below = (1u << p) - 1;
    const uint32_t rl = (lo & below) | ((lo >> t.m) & ~below);
    const uint32_t rh = (hi & below) | ((hi >> t.m) & ~below);
    a.tw_hi = ((rh << km) | rl) << CHAIN_BITS;                 // 2(k-m) + CHAIN_BITS = 32
    a.tw_lo = (mhash & t.hmask) << t.val_bits;
    a.block = (uint32_t)(((uint64_t)mhash * t.bpd) >> t.dshift);
    a.bucket = (p + mhash) & (BLOCK_BUCKETS - 1);
  } else {
    const SlotAddr g = slot_addr(t, lo, hi, mhash, p);
    const uint64_t tagword = g.tag << t.val_bits;
    a.tw_lo = (uint32_t)tagword; a.tw_hi = (uint32_t)(tagword >> 32);
    a.block = (uint32_t)g.block; a.bucket = g.bucket;
  }
  return a;
}

template <bool VAL32> struct ValT { using type = uint64_t; };
template <> struct ValT<true> { using type = uint32_t; };

template <bool VAL32>
__device__ __forceinline__ typename ValT<VAL32>::type resolve_home(const TableView& t, const uint64_t (&s)[4], uint32_t tw_lo, uint32_t tw_hi, bool* cont) {
  if constexpr (VAL32) {
    const uint32_t lomask = t.val_bits >= 32 ? 0u : (0xFFFFFFFFu << t.val_bits);
    uint32_t hit = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = (uint32_t)s[i], hi = (uint32_t)(s[i] >> 32);
      const bool m = ((((lo ^ tw_lo) & lomask) | (hi ^ tw_hi)) == 0);   // an empty slot (all ones) never equals a distance-0 tag
      hit = m ? lo : hit;
      found |= m;
    }
    const uint32_t last = (uint32_t)s[3];
    *cont = !found && s[3] != EMPTY64 && ((last >> (t.val_bits - 1)) & 1u);
    return found ? (hit & ((1u << (t.val_bits - 1)) - 1u)) : 0xFFFFFFFFu;
  } else {
    const uint64_t tagword = ((uint64_t)tw_hi << 32) | tw_lo;
    const uint64_t tmask = ~0ULL << t.val_bits;
    uint64_t hit = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool m = ((s[i] ^ tagword) & tmask) == 0;
      hit = m ? s[i] : hit;
      found |= m;
    }
    *cont = !found && s[3] != EMPTY64 && ((s[3] >> (t.val_bits - 1)) & 1);
    return found ? (hit & (((uint64_t)1 << (t.val_bits - 1)) - 1)) : LOOKUP_MISS;
  }
}

template <typename V> __device__ __forceinline__ bool is_miss(V v) { return v == (V)~(V)0; }
template <typename V> __device__ __forceinline__ uint32_t vkind(const TableView& t, V v) { return (uint32_t)(v >> (t.val_bits - 3)) & 3u; }
template <typename V> __device__ __forceinline__ V vpayload(const TableView& t, V v) { return v & ((((V)1) << (t.val_bits - 3)) - 1); }
template <typename V>
__device__ __forceinline__ uint32_t vinline_count(const TableView& t, V payload) {
  const uint32_t gm = (1u << t.gbits) - 1;
  uint32_t c = 1, prev = (uint32_t)payload & gm;
  for (uint32_t i = 1; i < t.n_inline; ++i) {
    const uint32_t g = (uint32_t)(payload >> (i * t.gbits)) & gm;
    if (g == prev) break;
    prev = g; ++c;
  }
  return c;
}
// does the genome set of a multi-genome k-mer contain genome g?
template <typename V>
__device__ __forceinline__ bool vset_contains(const TableView& t, V val, uint32_t g) {
  const V payload = vpayload(t, val);
  if (vkind(t, val) == KIND_INLINE) {
    // is one of the n_inline fields equal to g?  All fields at once: xor with g replicated into every field, then the
    // classic "has a zero field" test (exact for existence); padding fields repeat a member, so they cannot add a hit
    const V x = payload ^ ((V)g * (V)t.inl_ones);
    return ((x - (V)t.inl_ones) & ~x & (V)t.inl_highs) != 0;
  }
  for (uint64_t sector = payload;; ++sector) {
    uint32_t ids[8];
    ld_set_sector(t.mlist + sector * MLIST_SECTOR, ids);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if ((ids[i] & ~LIST_END) == g) return true;
      if (ids[i] & LIST_END) return false;
    }
  }
}

// lo / hi / mkey: the read's planes and slid minimizer keys (to redo the address of a window whose chain goes on)
template <bool VAL32>
__device__ __forceinline__ void fast_stage_c(const TableView& t, const AlignParams& prm, uint32_t look,
                                             const uint64_t (&sector)[AL_ROUNDS][4], const uint32_t (&tw_lo)[AL_ROUNDS],
                                             const uint32_t (&tw_hi)[AL_ROUNDS],
                                             const uint32_t (&lo)[AL_ROUNDS + 1], const uint32_t (&hi)[AL_ROUNDS + 1],
                                             const uint32_t (&mkey)[AL_ROUNDS + 1], uint32_t lane, int k, uint32_t kmask,
                                             uint64_t& res, bool& defer, uint32_t& read_nr) {
  using V = typename ValT<VAL32>::type;
  V val[AL_ROUNDS];
  uint32_t chain = 0;   // bit r: the chain of window r goes on (rare)
#pragma unroll
  for (int r = 0; r < AL_ROUNDS; ++r) {
    bool cont;
    const V v = resolve_home<VAL32>(t, sector[r], tw_lo[r], tw_hi[r], &cont);
    const bool on = (look >> r) & 1;
    val[r] = on ? v : (V)~(V)0;
    chain |= (on && cont) ? (1u << r) : 0u;
  }
  if (chain) {
    // (requesting the distance-1 sectors of all such windows first -- L2 prefetches, or loads held in registers -- was
    // measured slower: 13.28 -> 13.46 ms, the latter spills)
#pragma unroll
    for (int r = 0; r < AL_ROUNDS; ++r)
      if ((chain >> r) & 1) {
        const uint32_t wl = __funnelshift_r(lo[r], lo[r + 1], lane) & kmask;
        const uint32_t wh = __funnelshift_r(hi[r], hi[r + 1], lane) & kmask;
        uint32_t mh, mp;
        window_minimizer(t, mkey[r], wl, &mh, &mp);
        val[r] = (V)lookup_chain_window(t, ((uint64_t)wh << k) | wl, mh, mp);   // LOOKUP_MISS truncates to the 32-bit miss
      }
  }
  uint32_t mine = NO_GENOME, l_filtered = 0;
  bool same = true, l_multi = false;
#pragma unroll
  for (int r = 0; r < AL_ROUNDS; ++r) {
    const V v = val[r];
    const bool hitr = !is_miss(v);
    const uint32_t kind = vkind(t, v);
    bool keep = hitr;
    if (prm.has_mg) {  // max-genomes filter, per occurrence (kmer.py:425-427)
      uint32_t c = 1;
      // an inline list holds at most n_inline genomes: it is only counted when max_genomes is smaller than that
      if (hitr && kind == KIND_INLINE) c = (prm.mg >= (int64_t)t.n_inline) ? 2u : vinline_count(t, vpayload(t, v));
      else if (hitr && kind == KIND_MLIST) c = (prm.mg <= (int64_t)t.n_inline) ? t.n_inline + 1 : mlist_count(t.mlist, vpayload(t, v));
      const bool f = hitr && (int64_t)c > prm.mg;
      l_filtered += f;
      keep = hitr && !f;
    }
    const bool spec = keep && kind == KIND_SPECIFIC;
    const uint32_t g = (uint32_t)vpayload(t, v);
    same &= !spec || mine == NO_GENOME || g == mine;
    mine = (spec && mine == NO_GENOME) ? g : mine;
    const bool multi = keep && !spec;
    l_multi |= multi;
    val[r] = multi ? v : (V)~(V)0;   // from here on: the kept multi-genome values of this lane
  }
  read_nr += l_filtered;
  const uint32_t have = __ballot_sync(0xffffffffu, mine != NO_GENOME);
  const bool w_multi = __any_sync(0xffffffffu, l_multi);
  if (have == 0) {
    res = make_word(w_multi ? 3 : 1, 0, 0);
  } else {
    const uint32_t g0 = __shfl_sync(0xffffffffu, mine, __ffs(have) - 1);
    bool ok = same && (mine == NO_GENOME || mine == g0);
    if (w_multi && prm.p >= 0) {
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r)
        if (!is_miss(val[r])) ok = ok && vset_contains(t, val[r], g0);
    }
    if (__all_sync(0xffffffffu, ok)) res = make_word(2, 1, g0); else defer = true;
  }
}

template <bool QUAL, bool PACKED, bool VAL32>
__global__ void __launch_bounds__(FA_THREADS, PA_FAST_MINB)
align_fast_kernel(TableView t, ReadInput in, const uint8_t* __restrict__ quals,
                  const uint64_t* __restrict__ read_off, uint64_t n_reads, AlignParams prm, uint64_t* __restrict__ out_word,
                  unsigned long long* __restrict__ counters, uint32_t* __restrict__ queue,
                  unsigned long long* __restrict__ queue_count) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp_global = (uint64_t)blockIdx.x * FA_WARPS + (threadIdx.x >> 5);
  const uint64_t n_warps = (uint64_t)gridDim.x * FA_WARPS;
  const int k = (int)t.k;
  const uint32_t kmask = (k >= 1 && k < 32) ? ((1u << k) - 1) : 0u;
  unsigned long long c_drop = 0, c_nq = 0, c_nr = 0;  // per-lane partial counters

  // software pipeline: offsets of read i+2 and the first 160 bases of read i+1 are requested while read i is processed
  uint64_t nx_beg = 0, nx_end = 0, cur_beg = 0, cur_end = 0;
  Prefetch<PACKED> nx_ch, cur_ch;
  Prefetch<false> nx_q, cur_q;   // QUAL: the quality bytes travel through the same pipeline (one byte per lane per chunk)
  const ReadInput qin{quals, nullptr, 0, 0};
  {
    uint64_t r0 = warp_global, r1 = warp_global + n_warps;
    if (r0 < n_reads) { cur_beg = read_off[r0]; cur_end = read_off[r0 + 1]; }
    if (r1 < n_reads) { nx_beg = read_off[r1]; nx_end = read_off[r1 + 1]; }
    prefetch_read<PACKED>(in, r0 < n_reads, r0, cur_beg, cur_end - cur_beg, lane, cur_ch);
    if (QUAL) prefetch_read<false>(qin, r0 < n_reads, r0, cur_beg, cur_end - cur_beg, lane, cur_q);
  }

  for (uint64_t read = warp_global; read < n_reads; read += n_warps) {
    const uint64_t L = cur_end - cur_beg;
    const uint8_t* rq = QUAL ? quals + cur_beg : nullptr;
    uint64_t n2_beg = 0, n2_end = 0;
    {
      const uint64_t r1 = read + n_warps, r2 = read + 2 * n_warps;
      prefetch_read<PACKED>(in, r1 < n_reads, r1, nx_beg, nx_end - nx_beg, lane, nx_ch);
      if (QUAL) prefetch_read<false>(qin, r1 < n_reads, r1, nx_beg, nx_end - nx_beg, lane, nx_q);
      if (r2 < n_reads) { n2_beg = read_off[r2]; n2_end = read_off[r2 + 1]; }
    }

    uint64_t res = make_word(1, 0, 0);   // UNMAPPED unless decided otherwise
    bool defer = false;
    uint32_t read_nq = 0, read_nr = 0;
    bool dropped = false;
    if (QUAL && prm.has_mrq) {  // Read.mean_quality() < min_read_quality  (kmer.py:587)
      uint64_t s = 0;
      if (L <= 32 * (AL_ROUNDS + 1)) {   // the prefetched bytes cover the read (bytes beyond L were loaded as 0)
#pragma unroll
        for (int c = 0; c <= AL_ROUNDS; ++c) s += cur_q.v[c];
      } else {
        for (uint64_t i = lane; i < L; i += 32) s += rq[i];
      }
      s = warp_sum(s);
      if ((int64_t)s < prm.mrq * (int64_t)L) { dropped = true; res = 0; if (lane == 0) ++c_drop; }
    }
    const uint64_t W = (!dropped && k >= 1 && L >= (uint64_t)k) ? L - k + 1 : 0;  // kmer.py:91-92
    if (W > AL_SUPER) {
      defer = true;   // longer than one super-round: the general kernel loops over super-rounds
    } else if (W > 0) {
      // ---- bit planes of the read (ballots over the ASCII bases, or the host-packed words) ----
      uint32_t lo[AL_ROUNDS + 1], hi[AL_ROUNDS + 1], inv[AL_ROUNDS + 1];
      uint32_t qex[AL_ROUNDS + 1];  // exclusive quality prefix at this lane's base
      encode_planes<PACKED>(in, cur_ch, read, cur_beg, L, 0, lane, lo, hi, inv);
      if (QUAL && prm.has_mkq) {
        // warp prefix scans of two chunks per register: a 16-bit field holds the sum of 160 bytes
        uint32_t carry = 0;
#pragma unroll
        for (int c = 0; c <= AL_ROUNDS; c += 2) {
          const uint32_t q0 = cur_q.v[c], q1 = c + 1 <= AL_ROUNDS ? cur_q.v[c + 1 <= AL_ROUNDS ? c + 1 : c] : 0u;
          uint32_t incl = q0 | (q1 << 16);
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
          const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
          qex[c] = carry + (incl & 0xFFFFu) - q0;
          carry += tot & 0xFFFFu;
          if (c + 1 <= AL_ROUNDS) { qex[c + 1 <= AL_ROUNDS ? c + 1 : c] = carry + (incl >> 16) - q1; carry += tot >> 16; }
        }
      }
      // ---- minimizers ----
      uint32_t mkey[AL_ROUNDS + 1];
#pragma unroll
      for (int c = 0; c <= AL_ROUNDS; ++c) {
        const uint32_t nl = c < AL_ROUNDS ? lo[c + 1] : 0u, nh = c < AL_ROUNDS ? hi[c + 1] : 0u;
        const uint32_t xl = __funnelshift_r(lo[c], nl, lane) & t.mmask;
        const uint32_t xh = __funnelshift_r(hi[c], nh, lane) & t.mmask;
        mkey[c] = mmer_order((xh << t.m) | xl, t) << 4;
      }
      {
        uint32_t span = 1;
        for (; 2 * span <= t.w; span <<= 1) window_min_step(mkey, span, lane);
        if (t.w > span) window_min_step(mkey, t.w - span, lane);
      }
      // ---- per window: quality filter, block / bucket / tag, one sector load ----
      uint32_t tw_lo[AL_ROUNDS], tw_hi[AL_ROUNDS];
      uint64_t sector[AL_ROUNDS][4];
      uint32_t look = 0;   // bit r: window r of this lane is looked up
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        const uint32_t s = 32 * r + lane;
        const bool exists = s < W;
        bool qf = false;
        if (QUAL && prm.has_mkq) {  // kmer_quality(start, k) < min_kmer_quality, before the lookup (kmer.py:420-422)
          const uint32_t tl = lane + k;
          const uint32_t p_a = __shfl_sync(0xffffffffu, qex[r], tl & 31);
          const uint32_t p_b = __shfl_sync(0xffffffffu, qex[r + 1], tl & 31);
          const uint32_t end = tl < 32 ? p_a : p_b;
          qf = exists && ((int64_t)(end - qex[r]) < prm.mkq * (int64_t)k);
          read_nq += qf;
        }
        const uint32_t wl = __funnelshift_r(lo[r], lo[r + 1], lane) & kmask;
        const uint32_t wh = __funnelshift_r(hi[r], hi[r + 1], lane) & kmask;
        const uint32_t wi = __funnelshift_r(inv[r], inv[r + 1], lane) & kmask;
        uint32_t mh, mpos;
        window_minimizer(t, mkey[r], wl, &mh, &mpos);   // full hash and offset inside the window, 0 .. w-1
        const SlotWord a = slot_word<VAL32>(t, wl, wh, mh, mpos);
        tw_lo[r] = a.tw_lo; tw_hi[r] = a.tw_hi;         // the tag as it sits in a slot word
        if (exists && !qf && wi == 0) { look |= 1u << r; ld_sector_nc(bucket_ptr(t, a.block, a.bucket), sector[r]); }
      }
      // ---- resolve and classify ----
      fast_stage_c<VAL32>(t, prm, look, sector, tw_lo, tw_hi, lo, hi, mkey, lane, k, kmask, res, defer, read_nr);
    }
    if (defer) {
      if (lane == 0) queue[atomicAdd(queue_count, 1ULL)] = (uint32_t)read;
    } else {
      if (lane == 0) out_word[read] = res;
      c_nq += read_nq;
      c_nr += read_nr;
    }
    cur_beg = nx_beg; cur_end = nx_end;
    nx_beg = n2_beg; nx_end = n2_end;
    cur_ch = nx_ch;
    if (QUAL) cur_q = nx_q;
  }
  c_drop = warp_sum(c_drop); c_nq = warp_sum(c_nq); c_nr = warp_sum(c_nr);
  if (lane == 0) {
    if (c_drop) atomicAdd(counters + 0, c_drop);
    if (c_nq) atomicAdd(counters + 1, c_nq);
    if (c_nr) atomicAdd(counters + 2, c_nr);
  }
}


// ---------------------------------------------------------------------------
// K4 fast kernel, staggered (PA_FAST_SPLIT=1).  ncu attributes a fifth of the samples of the kernel above to the tag
// compare waiting for its sector: a warp issues its 128 loads and needs them a few instructions later.  Here the body
// is cut into three stages -- A: planes, minimizer keys (no table access); B: addresses + sector loads; C: resolve +
// rules -- and run as B(i), A(i+1), C(i): the ~500 instructions of the next read's stage A execute while the sectors of
// read i are in flight, with only A's result (planes + five keys per lane) carried from one iteration to the next.
// ---------------------------------------------------------------------------
#ifndef PA_FAST_SPLIT
#define PA_FAST_SPLIT 1
#endif

// QUAL: 0 = no quality filters, 2 = the filters were evaluated by quality_masks_kernel beforehand (one bit per window +
// one "dropped" byte per read): the quality state carried from stage A to stage B is one register.  (1 = quality bytes
// scanned in the kernel exists in the plain stage order only, align_fast_kernel: its five prefix sums per lane made the
// staggered order spill 108 bytes -- 26.6 ms against 17.5.)
template <int QUAL>
struct FrontState {
  uint32_t lo[AL_ROUNDS + 1], hi[AL_ROUNDS + 1];                       // bit planes of the read (warp-uniform)
  uint32_t ok;                                                         // bit r: window r of this lane holds ACGT only
  uint32_t mkey[AL_ROUNDS + 1];                                        // slid minimizer keys
  uint32_t qf;                                                         // QUAL 2: bit r = window r of this lane fails min-kmer-quality
  uint32_t W;                                                          // windows to look up (0: none)
  bool dropped, defer;
};

// Quality filters of a batch, ahead of K4 (kmer.py:394-408, 420-422, 587): per read a 128-bit mask -- bit 32 r + l = window
// 32 r + l has sum(q[w..w+k)) < mkq * k -- and a byte: 1 = sum(q) < mrq * len (the read is dropped).  Integer tests equal to
// the reference's float comparisons.  Reads with more than 128 windows get no mask: the general kernel
// scans their qualities itself.  Streaming: 1 byte in per base, 17 bytes out per read.
// One THREAD per read: a warp stages the quality bytes of its 32 reads (one contiguous span of the input) in shared memory
// with coalesced 16-byte loads, then every thread slides the k-byte window sum along its own read -- ~50 warp instructions
// per read where the warp-per-read scan of K4 (prefix sums by shuffles) takes ~250.
constexpr int QM_WARP_BYTES = 32 * 160 + 32;
__global__ void __launch_bounds__(256) quality_masks_kernel(const uint8_t* __restrict__ quals, const uint64_t* __restrict__ read_off,
                                                            uint64_t n_reads, AlignParams prm, int k, uint4* __restrict__ masks,
                                                            uint8_t* __restrict__ drop) {
  __shared__ __align__(16) uint8_t sh[8][QM_WARP_BYTES];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* my = sh[warp];
  const uint64_t n_groups = (n_reads + 31) / 32;
  const int64_t thr = prm.mkq * (int64_t)k;                       // window filter: sum < thr, sum is a uint32
  const bool thr_all = thr > (int64_t)0xFFFFFFFFLL;
  const uint32_t thr32 = thr <= 0 ? 0u : (thr_all ? 0xFFFFFFFFu : (uint32_t)thr);
  for (uint64_t grp = (uint64_t)blockIdx.x * 8 + warp; grp < n_groups; grp += (uint64_t)gridDim.x * 8) {
    const uint64_t read = grp * 32 + lane;
    const bool valid = read < n_reads;
    const uint64_t beg = read_off[valid ? read : n_reads];
    const uint64_t last = read_off[min(grp * 32 + 32, n_reads)];
    uint64_t end = __shfl_down_sync(0xffffffffu, beg, 1);
    if (lane == 31) end = last;
    const uint64_t L = valid ? end - beg : 0;
    const uint64_t span_beg = __shfl_sync(0xffffffffu, beg, 0), span = last - span_beg;
    const uint8_t* src = quals + span_beg;
    const uint32_t pre = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);   // the staged copy starts at the 16-byte line of the span
    const bool staged = span + pre <= (uint64_t)QM_WARP_BYTES;
    if (staged) {
      const uint8_t* a0 = src - pre;
      const uint32_t nbytes = (uint32_t)span + pre;
      for (uint32_t i = lane * 16; i < nbytes; i += 512) {
        if (i + 16 <= nbytes) *reinterpret_cast<uint4*>(my + i) = __ldcs(reinterpret_cast<const uint4*>(a0 + i));
        else for (uint32_t j = i; j < nbytes; ++j) my[j] = a0[j];
      }
    }
    __syncwarp();
    const uint8_t* qp = staged ? my + pre + (beg - span_beg) : quals + beg;
    uint32_t m[AL_ROUNDS] = {0, 0, 0, 0};
    uint64_t total = 0;
    const uint64_t W = (k >= 1 && L >= (uint64_t)k) ? L - k + 1 : 0;
    if (prm.has_mkq && W > 0 && W <= AL_SUPER) {
      uint32_t s = 0;
      for (int i = 0; i < k; ++i) s += qp[i];
      uint32_t tot32 = s;
      const uint32_t Wn = (uint32_t)W;
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        uint32_t word = 0;
        const uint32_t w_end = min(Wn, 32u * (r + 1));
        for (uint32_t w = 32u * r; w < w_end; ++w) {
          word |= ((thr_all || s < thr32) ? 1u : 0u) << (w & 31);
          if (w + 1 < Wn) { const uint32_t in = qp[w + k]; s += in - qp[w]; tot32 += in; }
        }
        m[r] = word;
      }
      total = tot32;
    } else if (prm.has_mrq) {
      for (uint64_t i = 0; i < L; ++i) total += qp[i];
    }
    __syncwarp();   // the staging buffer is rewritten by the next group
    if (valid) {
      masks[read] = make_uint4(m[0], m[1], m[2], m[3]);
      drop[read] = (prm.has_mrq && (int64_t)total < prm.mrq * (int64_t)L) ? 1 : 0;
    }
  }
}

// the mask record of a read, through the quality slot of the input pipeline (every lane loads the same 17 bytes)
__device__ __forceinline__ void prefetch_masks(const uint8_t* __restrict__ qm, uint64_t n_reads, bool valid, uint64_t read,
                                               Prefetch<false>& p) {
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  uint32_t d = 0;
  if (valid) { m = __ldg(reinterpret_cast<const uint4*>(qm) + read); d = __ldg(qm + 16 * n_reads + read); }
  p.v[0] = m.x; p.v[1] = m.y; p.v[2] = m.z; p.v[3] = m.w; p.v[AL_ROUNDS] = d;
}

template <int QUAL, bool PACKED>
__device__ __forceinline__ void fast_stage_a(const TableView& t, const AlignParams& prm, const ReadInput& in,
                                             const uint8_t* __restrict__ quals, const Prefetch<PACKED>& ch,
                                             const Prefetch<false>& q, uint64_t read, uint64_t beg, uint64_t L, uint32_t lane,
                                             FrontState<QUAL>& f, unsigned long long& c_drop) {
  const int k = (int)t.k;
  f.dropped = false; f.defer = false; f.W = 0; f.qf = 0;
  if (QUAL == 2) {
    if (q.v[AL_ROUNDS]) { f.dropped = true; if (lane == 0) ++c_drop; }
#pragma unroll
    for (int r = 0; r < AL_ROUNDS; ++r) f.qf |= ((q.v[r] >> lane) & 1u) << r;
  }
  const uint64_t W = (!f.dropped && k >= 1 && L >= (uint64_t)k) ? L - k + 1 : 0;  // kmer.py:91-92
  if (W > AL_SUPER) { f.defer = true; return; }   // longer than one super-round: the general kernel loops over super-rounds
  if (W == 0) return;
  f.W = (uint32_t)W;
  {
    // the "not ACGT" plane is needed for one test per window only: reduce it to four bits here instead of carrying five
    // words through the pipeline (the carried state is what limits the registers of the staggered kernel)
    uint32_t inv[AL_ROUNDS + 1];
    encode_planes<PACKED>(in, ch, read, beg, L, 0, lane, f.lo, f.hi, inv);
    const uint32_t kmask = (k >= 1 && k < 32) ? ((1u << k) - 1) : 0u;
    f.ok = 0;
#pragma unroll
    for (int r = 0; r < AL_ROUNDS; ++r)
      f.ok |= ((__funnelshift_r(inv[r], inv[r + 1], lane) & kmask) == 0 ? 1u : 0u) << r;
  }
#pragma unroll
  for (int c = 0; c <= AL_ROUNDS; ++c) {
    const uint32_t nl = c < AL_ROUNDS ? f.lo[c + 1] : 0u, nh = c < AL_ROUNDS ? f.hi[c + 1] : 0u;
    const uint32_t xl = __funnelshift_r(f.lo[c], nl, lane) & t.mmask;
    const uint32_t xh = __funnelshift_r(f.hi[c], nh, lane) & t.mmask;
    f.mkey[c] = mmer_order((xh << t.m) | xl, t) << 4;
  }
  uint32_t span = 1;
  for (; 2 * span <= t.w; span <<= 1) window_min_step(f.mkey, span, lane);
  if (t.w > span) window_min_step(f.mkey, t.w - span, lane);
}

template <int QUAL, bool PACKED, bool VAL32>
__global__ void __launch_bounds__(FA_THREADS, PA_FAST_MINB)
align_fast_split_kernel(TableView t, ReadInput in, const uint8_t* __restrict__ quals,
                        const uint64_t* __restrict__ read_off, uint64_t n_reads, AlignParams prm, uint64_t* __restrict__ out_word,
                        unsigned long long* __restrict__ counters, uint32_t* __restrict__ queue,
                        unsigned long long* __restrict__ queue_count) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp_global = (uint64_t)blockIdx.x * FA_WARPS + (threadIdx.x >> 5);
  const uint64_t n_warps = (uint64_t)gridDim.x * FA_WARPS;
  const int k = (int)t.k;
  const uint32_t kmask = (k >= 1 && k < 32) ? ((1u << k) - 1) : 0u;
  unsigned long long c_drop = 0, c_nq = 0, c_nr = 0;  // per-lane partial counters

  // input pipeline: while read i is resolved, stage A runs on read i+1 (bases requested one iteration earlier), the
  // bases of read i+2 and the offsets of read i+3 are requested
  uint64_t nx_beg = 0, nx_end = 0, n2_beg = 0, n2_end = 0;
  Prefetch<PACKED> nx_ch;
  Prefetch<false> nx_q;
  FrontState<QUAL> cur;
  {
    const uint64_t r0 = warp_global, r1 = r0 + n_warps, r2 = r0 + 2 * n_warps;
    uint64_t beg0 = 0, end0 = 0;
    if (r0 < n_reads) { beg0 = read_off[r0]; end0 = read_off[r0 + 1]; }
    if (r1 < n_reads) { nx_beg = read_off[r1]; nx_end = read_off[r1 + 1]; }
    if (r2 < n_reads) { n2_beg = read_off[r2]; n2_end = read_off[r2 + 1]; }
    Prefetch<PACKED> ch0;
    Prefetch<false> q0;
    prefetch_read<PACKED>(in, r0 < n_reads, r0, beg0, end0 - beg0, lane, ch0);
    if (QUAL == 2) prefetch_masks(quals, n_reads, r0 < n_reads, r0, q0);
    prefetch_read<PACKED>(in, r1 < n_reads, r1, nx_beg, nx_end - nx_beg, lane, nx_ch);
    if (QUAL == 2) prefetch_masks(quals, n_reads, r1 < n_reads, r1, nx_q);
    if (r0 < n_reads) fast_stage_a<QUAL, PACKED>(t, prm, in, quals, ch0, q0, r0, beg0, end0 - beg0, lane, cur, c_drop);
  }

  for (uint64_t read = warp_global; read < n_reads; read += n_warps) {
    // ---- stage B: per window quality filter, block / bucket / tag, one sector load ----
    uint32_t tw_lo[AL_ROUNDS], tw_hi[AL_ROUNDS];
    uint64_t sector[AL_ROUNDS][4];
    uint32_t look = 0;   // bit r: window r of this lane is looked up
    uint32_t read_nq = 0, read_nr = 0;
    if (cur.W) {
#pragma unroll
      for (int r = 0; r < AL_ROUNDS; ++r) {
        const uint32_t s = 32 * r + lane;
        const bool exists = s < cur.W;
        bool qf = false;
        // kmer_quality(start, k) < min_kmer_quality, before the lookup (kmer.py:420-422): bit r of the read's mask
        if (QUAL == 2) { qf = exists && ((cur.qf >> r) & 1u); read_nq += qf; }
        const uint32_t wl = __funnelshift_r(cur.lo[r], cur.lo[r + 1], lane) & kmask;
        const uint32_t wh = __funnelshift_r(cur.hi[r], cur.hi[r + 1], lane) & kmask;
        const uint32_t wi = ((cur.ok >> r) & 1u) ^ 1u;
        uint32_t mh, mp;
        window_minimizer(t, cur.mkey[r], wl, &mh, &mp);
        const SlotWord a = slot_word<VAL32>(t, wl, wh, mh, mp);
        tw_lo[r] = a.tw_lo; tw_hi[r] = a.tw_hi;         // the tag as it sits in a slot word
        if (exists && !qf && wi == 0) { look |= 1u << r; ld_sector_nc(bucket_ptr(t, a.block, a.bucket), sector[r]); }
      }
    }

    // ---- stage A of the next read, inputs of the reads after it ----
    FrontState<QUAL> nx;
    nx.W = 0; nx.dropped = false; nx.defer = false;
    {
      const uint64_t r1 = read + n_warps, r2 = read + 2 * n_warps, r3 = read + 3 * n_warps;
      Prefetch<PACKED> n2_ch;
      Prefetch<false> n2_q;
      prefetch_read<PACKED>(in, r2 < n_reads, r2, n2_beg, n2_end - n2_beg, lane, n2_ch);
      if (QUAL == 2) prefetch_masks(quals, n_reads, r2 < n_reads, r2, n2_q);
      uint64_t n3_beg = 0, n3_end = 0;
      if (r3 < n_reads) { n3_beg = read_off[r3]; n3_end = read_off[r3 + 1]; }
      if (r1 < n_reads) fast_stage_a<QUAL, PACKED>(t, prm, in, quals, nx_ch, nx_q, r1, nx_beg, nx_end - nx_beg, lane, nx, c_drop);
      nx_ch = n2_ch;
      if (QUAL != 0) nx_q = n2_q;
      nx_beg = n2_beg; nx_end = n2_end;
      n2_beg = n3_beg; n2_end = n3_end;
    }

    // ---- stage C: resolve and classify ----
    uint64_t res = cur.dropped ? 0 : make_word(1, 0, 0);   // UNMAPPED unless decided otherwise
    bool defer = cur.defer;
    if (cur.W) fast_stage_c<VAL32>(t, prm, look, sector, tw_lo, tw_hi, cur.lo, cur.hi, cur.mkey, lane, k, kmask, res, defer, read_nr);
    if (defer) {
      if (lane == 0) queue[atomicAdd(queue_count, 1ULL)] = (uint32_t)read;
    } else {
      if (lane == 0) out_word[read] = res;
      c_nq += read_nq;
      c_nr += read_nr;
    }
    cur = nx;
  }
  c_drop = warp_sum(c_drop); c_nq = warp_sum(c_nq); c_nr = warp_sum(c_nr);
  if (lane == 0) {
    if (c_drop) atomicAdd(counters + 0, c_drop);
    if (c_nq) atomicAdd(counters + 1, c_nq);
    if (c_nr) atomicAdd(counters + 2, c_nr);
  }
}

__global__ void scratch_init(unsigned char* scratch, uint64_t n_warps, uint64_t stride, uint32_t G, uint32_t kset_cap,
                             int gtab_in_smem, int kset_in_smem) {
  uint64_t w = blockIdx.x;
  if (w >= n_warps) return;
  unsigned char* gp = scratch + w * stride;
  if (!kset_in_smem) {
    unsigned long long* kk = (unsigned long long*)gp;
    uint32_t* kp = (uint32_t*)(gp + (size_t)kset_cap * 8);
    for (uint32_t i = threadIdx.x; i < kset_cap; i += blockDim.x) { kk[i] = EMPTY64; kp[i] = NOPOS; }
    gp += ((size_t)kset_cap * 12 + 15) & ~(size_t)15;
  }
  if (!gtab_in_smem) {
    uint32_t* gt = (uint32_t*)gp;
    for (uint32_t i = threadIdx.x; i < G; i += blockDim.x) { gt[i * 4] = 0; gt[i * 4 + 1] = 0; gt[i * 4 + 2] = NOPOS; gt[i * 4 + 3] = NOPOS; }
  }
}

// ===========================================================================
// K8: PseudoAlignment.get_summary (kmer.py:622-657) over the result words.
// stats = {unique, ambiguous, unmapped, dropped}; per genome one count per LIST
// ELEMENT; first_seen[g] = min over (global read index << 22 | list position),
// which orders the "Summary" keys by first appearance.
// ===========================================================================
constexpr int SUM_THREADS = 256;
constexpr uint32_t SUM_SMEM_GENOMES = 4096;

// A genome occurs at most once in a read's list, so a per-block counter is bounded by the reads the block sees (32 bits
// are enough) and the per-genome counts are taken in shared memory, added to the global ones once per block: 10^7
// global atomics on a hundred addresses serialise in L2 (1.3 ms per 10^7 reads; 0.1 ms this way).  first_seen only moves
// down, so a (possibly stale) plain load filters almost every atomicMin.
template <bool SMEM>
__global__ void __launch_bounds__(SUM_THREADS)
summary_kernel(const uint64_t* __restrict__ words, const uint32_t* __restrict__ list, uint64_t n_reads, uint64_t read_index_base,
               uint32_t G, unsigned long long* __restrict__ stats, unsigned long long* __restrict__ unique_reads,
               unsigned long long* __restrict__ ambiguous_reads, unsigned long long* first_seen) {
  __shared__ uint32_t s_cnt[SMEM ? 2 * SUM_SMEM_GENOMES : 1];
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < 2 * G; i += SUM_THREADS) s_cnt[i] = 0;
    __syncthreads();
  }
  unsigned long long st[4] = {0, 0, 0, 0};
  const uint64_t stride = (uint64_t)gridDim.x * SUM_THREADS;
  for (uint64_t i = blockIdx.x * (uint64_t)SUM_THREADS + threadIdx.x; i < n_reads; i += stride) {
    uint64_t w = words[i];
    uint32_t type = (uint32_t)(w >> 62);
    uint64_t len = (w >> 40) & 0x3FFFFF, payload = w & 0xFFFFFFFFFFULL;
    if (type == 0) { ++st[3]; continue; }
    if (type == 1) { ++st[2]; continue; }
    ++st[type == 2 ? 0 : 1];
    unsigned long long* cnt = type == 2 ? unique_reads : ambiguous_reads;
    const uint32_t s_base = type == 2 ? 0 : G;
    const uint64_t order_base = (read_index_base + i) << 22;
    for (uint64_t j = 0; j < len; ++j) {
      const uint32_t g = len == 1 ? (uint32_t)payload : list[payload + j];
      if (SMEM) atomicAdd(&s_cnt[s_base + g], 1u); else atomicAdd(&cnt[g], 1ULL);
      const unsigned long long order = order_base | j;
      if (order < *(volatile unsigned long long*)&first_seen[g]) atomicMin(&first_seen[g], order);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    unsigned long long v = warp_sum(st[c]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats[c], v);
  }
  if (SMEM) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 2 * G; i += SUM_THREADS)
      if (s_cnt[i]) atomicAdd(i < G ? &unique_reads[i] : &ambiguous_reads[i - G], (unsigned long long)s_cnt[i]);
  }
}

}  // namespace

int32_t align_batch_device(Index& ix, const uint8_t* d_bases, const uint8_t* d_quals, const uint64_t* d_read_off,
                           uint64_t n_reads, uint64_t max_read_len, const AlignParams& prm_in, uint64_t* d_words,
                           uint32_t* d_list, uint64_t list_cap, unsigned long long* d_cursor /*[2]*/,
                           unsigned long long* d_counters /*[3]*/, cudaStream_t s, int32_t* launches,
                           const uint32_t* d_planes, uint64_t planes_base0, uint64_t planes_read0) {
  if (launches) *launches = 0;
  if (n_reads == 0) return ST_OK;
  constexpr uint64_t MAX_BATCH = 1ULL << 31;   // read indices travel through the queue as uint32
  if (n_reads > MAX_BATCH) {
    if (d_planes) { set_error("align: packed input is limited to 2^31 reads per call"); return ST_UNSUPPORTED; }
    for (uint64_t lo = 0; lo < n_reads; lo += MAX_BATCH) {
      int32_t l = 0;
      PA_TRY(align_batch_device(ix, d_bases, d_quals, d_read_off + lo, std::min(MAX_BATCH, n_reads - lo), max_read_len, prm_in,
                                d_words + lo, d_list, list_cap, d_cursor, d_counters, s, &l, nullptr, 0, 0));
      if (launches) *launches += l;
    }
    return ST_OK;
  }
  AlignParams prm = prm_in;
  const int k = ix.k;
  const bool qual = (prm.has_mrq || prm.has_mkq);
  if (qual && !d_quals) { set_error("align: quality filters requested without quality data"); return ST_INVALID_ARG; }
  int dev = 0, sms = 148;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TableView tv = ix.view();  // k <= 0: the kernels see no windows (kmer.py:91-92) and only apply the read-quality drop
  const bool packed = d_planes != nullptr;
  const ReadInput in{d_bases, d_planes, planes_base0, planes_read0};

  // ---- fast kernel over all reads; what it cannot decide goes to the queue ----
  if (ix.align_queue.bytes < 16 + n_reads * 4) PA_TRY(ix.align_queue.alloc(16 + n_reads * 4 + n_reads / 2));
  unsigned long long* q_count = ix.align_queue.as<unsigned long long>();
  uint32_t* q_items = reinterpret_cast<uint32_t*>(ix.align_queue.as<unsigned char>() + 16);
  PA_CUDA(cudaMemsetAsync(q_count, 0, 8, s));
  {
    // VAL32: see fast_stage_c.
    const bool v32 = tv.val_bits <= 32 && 2 * (tv.k - tv.m) + CHAIN_BITS == 32;
    using FastKernel = void (*)(TableView, ReadInput, const uint8_t*, const uint64_t*, uint64_t, AlignParams, uint64_t*, unsigned long long*,
                                uint32_t*, unsigned long long*);
    FastKernel fk;
    // EXTQUALITY: the filters are evaluated by quality_masks_kernel first and K4 runs in the staggered order on their bits
    // (PA_QUAL_MASKS=0: the kernel that scans the quality bytes itself, in the plain stage order)
    const char* qm_env = getenv("PA_QUAL_MASKS");
    const bool qual_masks = qual && PA_FAST_SPLIT && !(qm_env && *qm_env == '0');
    const uint8_t* fast_quals = d_quals;
    if (qual_masks) {
      if (ix.align_qmasks.bytes < n_reads * 17 + 16) PA_TRY(ix.align_qmasks.alloc(n_reads * 17 + n_reads / 2 + 16));
      uint8_t* qm = ix.align_qmasks.as<uint8_t>();
      const uint64_t qgrid = std::min<uint64_t>((uint64_t)sms * 5, (n_reads + 255) / 256);
      quality_masks_kernel<<<(unsigned)std::max<uint64_t>(qgrid, 1), 256, 0, s>>>(d_quals, d_read_off, n_reads, prm, k,
                                                                                 reinterpret_cast<uint4*>(qm), qm + 16 * n_reads);
      PA_CUDA(cudaGetLastError());
      if (launches) ++*launches;
      fast_quals = qm;
      fk = packed ? (v32 ? align_fast_split_kernel<2, true, true> : align_fast_split_kernel<2, true, false>)
                  : (v32 ? align_fast_split_kernel<2, false, true> : align_fast_split_kernel<2, false, false>);
    }
    else if (qual) fk = packed ? (v32 ? align_fast_kernel<true, true, true> : align_fast_kernel<true, true, false>)
                          : (v32 ? align_fast_kernel<true, false, true> : align_fast_kernel<true, false, false>);
    else if (PA_FAST_SPLIT) fk = packed ? (v32 ? align_fast_split_kernel<0, true, true> : align_fast_split_kernel<0, true, false>)
                                        : (v32 ? align_fast_split_kernel<0, false, true> : align_fast_split_kernel<0, false, false>);
    else fk = packed ? (v32 ? align_fast_kernel<false, true, true> : align_fast_kernel<false, true, false>)
                     : (v32 ? align_fast_kernel<false, false, true> : align_fast_kernel<false, false, false>);
    int occ = 1;
    PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fk, FA_THREADS, 0));
    if (occ < 1) occ = 1;
    uint64_t grid = std::min<uint64_t>((uint64_t)sms * occ, (n_reads + FA_WARPS - 1) / FA_WARPS);
    fk<<<(unsigned)std::max<uint64_t>(grid, 1), FA_THREADS, 0, s>>>(tv, in, fast_quals, d_read_off, n_reads, prm, d_words,
                                                                     d_counters, q_items, q_count);
    PA_CUDA(cudaGetLastError());
    if (launches) ++*launches;
  }

  // ---- general kernel over the queue ----
  const uint32_t G = std::max<uint32_t>(ix.n_genomes, 1);
  const uint64_t Wmax = (k >= 1 && max_read_len >= (uint64_t)k) ? max_read_len - k + 1 : 0;
  uint32_t kset_cap = 256;
  const int kset_in_smem = Wmax <= AL_SUPER;
  if (!kset_in_smem) { uint64_t c = 256; while (c < 2 * Wmax) c <<= 1; if (c > 0x40000000ull) { set_error("align: read too long"); return ST_UNSUPPORTED; } kset_cap = (uint32_t)c; }
  const int gtab_in_smem = G <= 256;
  size_t per_warp_smem = (kset_in_smem ? (size_t)kset_cap * 12 : 0) + (gtab_in_smem ? (size_t)G * 20 : 0);
  per_warp_smem = (per_warp_smem + 15) & ~(size_t)15;
  const size_t dyn_smem = per_warp_smem * AL_WARPS;
  auto kern = qual ? (packed ? align_kernel<true, true, 2> : align_kernel<true, false, 2>)
                   : (packed ? align_kernel<false, true, 2> : align_kernel<false, false, 2>);
  PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(dyn_smem, 1024)));
  int occ = 1;
  PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, AL_THREADS, dyn_smem));
  if (occ < 1) occ = 1;
  uint64_t grid = (uint64_t)sms * occ;
  grid = std::min<uint64_t>(grid, (n_reads + AL_WARPS - 1) / AL_WARPS);
  grid = std::max<uint64_t>(grid, 1);
  const uint64_t n_warps = grid * AL_WARPS;

  uint64_t stride = 0;
  if (!kset_in_smem) stride += ((uint64_t)kset_cap * 12 + 15) & ~15ull;
  if (!gtab_in_smem) stride += ((uint64_t)G * 20 + 15) & ~15ull;
  if (stride) {
    uint64_t key = stride * 1000003ull + kset_cap * 7ull + G;
    if (ix.align_scratch_warps < n_warps || ix.align_scratch_stride != key) {
      PA_TRY(ix.align_scratch.alloc(n_warps * stride));
      scratch_init<<<(unsigned)n_warps, 128, 0, s>>>(ix.align_scratch.as<unsigned char>(), n_warps, stride, G, kset_cap,
                                                     gtab_in_smem, kset_in_smem);
      PA_CUDA(cudaGetLastError());
      ix.align_scratch_warps = n_warps; ix.align_scratch_stride = key;
      if (launches) ++*launches;
    }
  }
  Emit em{d_words, d_list, list_cap, d_cursor};
  kern<<<(unsigned)grid, AL_THREADS, dyn_smem, s>>>(tv, in, d_quals, d_read_off, n_reads, prm, em, d_counters,
                                                    ix.align_scratch.as<unsigned char>(), stride, G, kset_cap,
                                                    gtab_in_smem, kset_in_smem, q_items, q_count);
  PA_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  return ST_OK;
}

int32_t summary_reduce_device(const uint64_t* d_words, const uint32_t* d_list, uint64_t n_reads, uint64_t read_index_base,
                              uint32_t G, unsigned long long* d_stats, unsigned long long* d_unique,
                              unsigned long long* d_ambiguous, unsigned long long* d_first_seen, cudaStream_t s) {
  if (n_reads == 0) return ST_OK;
  int dev = 0, sms = 148;
  PA_CUDA(cudaGetDevice(&dev));
  PA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  uint64_t grid = std::min<uint64_t>((n_reads + SUM_THREADS - 1) / SUM_THREADS, (uint64_t)sms * 8);
  const char* force_global = getenv("PA_SUMMARY_GLOBAL");   // tests: the path for more genomes than the shared counters hold
  if (G <= SUM_SMEM_GENOMES && !(force_global && *force_global == '1'))
    summary_kernel<true><<<(unsigned)grid, SUM_THREADS, 0, s>>>(d_words, d_list, n_reads, read_index_base, G, d_stats, d_unique,
                                                                 d_ambiguous, d_first_seen);
  else
    summary_kernel<false><<<(unsigned)grid, SUM_THREADS, 0, s>>>(d_words, d_list, n_reads, read_index_base, G, d_stats, d_unique,
                                                                  d_ambiguous, d_first_seen);
  PA_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // namespace pa
