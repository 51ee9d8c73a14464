/*
 * pa_b200.h -- C ABI of the B200-native k-mer reference build + read
 * pseudo-alignment path (libpa_b200.so, hand-written sm_100a CUDA).
 *
 * The reference (nyenyu12/BioInformatics-project-for-Shotgun-Metagenomics-
 * Pseudo-alignment-shotgun-) has no FFI layer: its boundary for this path is
 * the Python class surface of src/kmer.py.  Each entry point below names the
 * reference interface (file:line under /root/reference/src) whose work it
 * takes over; the Python shim that binds them with ctypes is the package's
 * _native.py / kmer.py, and INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every function returns int32: PA_OK or a negative PA_ERR_* code;
 *     pa_last_error() returns the message of the calling thread's last error.
 *     No C++ exception crosses the boundary.
 *   - plain pointers and sizes only; the caller owns every buffer it passes.
 *     "host" pointers are ordinary (preferably pinned) host memory, "device"
 *     pointers are CUDA device memory on the index's device (e.g.
 *     torch.Tensor.data_ptr()).
 *   - a pa_index owns one CUDA stream and is not thread-safe; calls on
 *     different handles may run concurrently.  `stream` arguments take a
 *     cudaStream_t cast to void*; NULL means the handle's own stream.
 *   - genome indices are positions in the genome list passed to the build
 *     (FASTA order); after pa_index_drop_genomes they are ranks among the
 *     surviving genomes.
 */
#ifndef PA_B200_H
#define PA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_OK 0
#define PA_ERR_INVALID_ARG (-1) /* -> ValueError / TypeError in the Python shim */
#define PA_ERR_BAD_BASE (-2)    /* genome byte outside ACGTN (records.py:229 admits nothing else) -> ValueError */
#define PA_ERR_CUDA (-3)        /* -> RuntimeError */
#define PA_ERR_NOMEM (-4)       /* -> MemoryError */
#define PA_ERR_CAPACITY (-5)    /* an output buffer is too small; the needed size is reported */
#define PA_ERR_UNSUPPORTED (-6) /* outside the built scope (k > 31, > 2^32-2 bases, ...) -> ValueError */

#define PA_ABI_VERSION 3
#define PA_RANK_MISS UINT64_MAX

typedef struct pa_index pa_index;

typedef struct pa_index_info {
  int32_t k;
  int32_t device;
  uint32_t n_genomes;
  uint32_t blocks_per_digit; /* the lookup table has blocks_per_digit << digit_bits blocks of 512 bytes */
  uint32_t tag_bits;        /* bits of the k-mer stored in a slot (the rest is implied by the line) */
  uint32_t stash_count;     /* k-mers that overflowed their bucket chain */
  uint64_t n_keys;          /* distinct k-mers            (len(KmerReference.kmers)) */
  uint64_t n_runs;          /* (k-mer, genome) pairs      (sum of inner dict sizes) */
  uint64_t n_occ;           /* k-mer occurrences          (sum of position-set sizes) */
  uint64_t total_bases;
  uint64_t n_list_sectors;  /* 32-byte sectors holding multi-genome lists */
  uint64_t device_bytes;
  float build_encode_ms, build_sort_ms, build_rle_ms, build_table_ms; /* CUDA-event times of the last build */
  uint32_t minimizer_len;   /* m: k-mers sharing their minimizer m-mer share a table block */
  uint32_t digit_bits;      /* top bits of the minimizer hash: ownership unit of a multi-GPU build */
  uint64_t n_blocks;        /* blocks of the lookup table */
  uint64_t table_bytes;     /* slots + stash + genome sets: all the alignment kernel reads */
  uint32_t align_only;      /* 1: table-only index (no CSR, no positions): alignment and table lookups only */
  uint32_t reserved;
} pa_index_info;

/* thresholds of Read.pseudo_align (kmer.py:482-489); has_* = "is not None" */
typedef struct pa_align_params {
  int64_t m, p;
  int64_t min_read_quality, min_kmer_quality, max_genomes;
  int32_t has_min_read_quality, has_min_kmer_quality, has_max_genomes;
  int32_t reserved;
} pa_align_params;

int32_t pa_abi_version(void);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its full length */
int32_t pa_last_error(char* buf, size_t n);
int32_t pa_device_count(int32_t* n);
/* The library keeps the big device buffers of finished builds / freed indexes (up to PA_CACHE_GB, default a quarter of
 * the device memory; 0 switches the cache off) and hands them to the next build instead of paying cudaMalloc / cudaFree
 * again; an allocation failure releases them by itself.  pa_trim_memory gives everything back now. */
int32_t pa_trim_memory(void);

/* ---- index build: KmerReference.__init__ / _build_kmer_mapping (kmer.py:113-150) ------------------
 * bases      concatenated genome strings, bytes in ACGTN (host); genome_off[G+1] offsets into it.
 * k <= 0 or k > len(genome) yields no k-mers for that genome (kmer.py:91-92); k > 31 -> PA_ERR_UNSUPPORTED. */
int32_t pa_index_build(const uint8_t* bases, const uint64_t* genome_off, uint32_t n_genomes, int32_t k, int32_t device,
                       pa_index** out);
/* same, with the concatenated bases already resident in device memory (16-byte aligned) */
int32_t pa_index_build_device(const uint8_t* d_bases, const uint64_t* genome_off, uint32_t n_genomes, int32_t k,
                              int32_t device, pa_index** out);
/* rebuild from an exported CSR (KmerReference.load, kmer.py:273-282) */
int32_t pa_index_import(int32_t k, uint32_t n_genomes, const uint64_t* genome_off, uint64_t n_keys, uint64_t n_runs,
                        uint64_t n_occ, const uint64_t* keys, const uint64_t* run_off, const uint32_t* run_genome,
                        const uint64_t* pos_off, const uint32_t* pos, const uint64_t* first_occ /* may be NULL */,
                        int32_t device, pa_index** out);
int32_t pa_index_free(pa_index* idx);
int32_t pa_index_info_get(pa_index* idx, pa_index_info* info);

/* CSR export (KmerReference.kmers / get_summary / save, kmer.py:130, 265-271, 300-329).  Host buffers sized from
 * pa_index_info: keys[n_keys], run_off[n_keys+1], run_genome[n_runs], pos_off[n_runs+1], pos[n_occ],
 * order[n_keys] = distinct k-mers in dict insertion order (first occurrence in genome/position order),
 * first_occ[n_keys] = the order key itself (global base position of the first occurrence in the genome list the
 * index was BUILT from; it survives pa_index_drop_genomes like the reference's dict order survives deletion,
 * kmer.py:237-243) -- pass it back to pa_index_import.
 * keys are private encodings: decode with pa_decode_kmers.  Any pointer may be NULL to skip that array. */
int32_t pa_index_export(pa_index* idx, uint64_t* keys, uint64_t* run_off, uint32_t* run_genome, uint64_t* pos_off,
                        uint32_t* pos, uint32_t* order, uint64_t* first_occ);
/* Content checksum of the CSR: four sums modulo 2^64 over {k-mers, (k-mer, genome) pairs, (k-mer, genome, position)
 * triples, k-mer count}.  Order-independent and additive over disjoint key sets: the checksums of the partitions of a
 * multi-GPU build add up to the checksum of the single-GPU index exactly when the contents agree. */
int32_t pa_index_checksum(pa_index* idx, uint64_t sums[4]);
int32_t pa_decode_kmers(int32_t k, const uint64_t* keys, uint64_t n, uint8_t* ascii /* n*k bytes */);
int32_t pa_encode_kmers(int32_t k, const uint8_t* ascii, uint64_t n, uint64_t* keys /* UINT64_MAX when not ACGT */);

/* KmerReference.get_kmer_references / __getitem__ (kmer.py:284-298): rank of each k-mer (n strings of k bytes)
 * among the exported keys, PA_RANK_MISS when absent. */
int32_t pa_index_lookup(pa_index* idx, const uint8_t* kmers_ascii, uint64_t n, uint64_t* rank);

/* The CSR entries of a few k-mers (point lookups of KmerReference.__getitem__ / get_kmer_references, kmer.py:284-298,
 * and Read.extract_kmer_references, kmer.py:410-429): ranks from pa_index_lookup (no PA_RANK_MISS).  Outputs are a small
 * CSR of their own: run_off[n + 1] into run_genome[*run_total], pos_off[*run_total + 1] into pos[*pos_total].
 * PA_ERR_CAPACITY (with *run_total / *pos_total set) when run_cap / pos_cap are too small or a buffer is NULL. */
int32_t pa_index_entries(pa_index* idx, const uint64_t* ranks, uint64_t n, uint64_t* run_off, uint32_t* run_genome,
                         uint64_t run_cap, uint64_t* pos_off, uint32_t* pos, uint64_t pos_cap, uint64_t* run_total,
                         uint64_t* pos_total);

/* ---- EXTSIM (kmer.py:152-263) ---------------------------------------------------------------------
 * group[g] = identifier class of genome g (kmer.py:162 keys everything by record.identifier).
 * total/unique: _compute_genome_stats (kmer.py:152-177); inter[a*n+b] = number of distinct k-mers shared by
 * classes a and b, the |A & B| of _apply_greedy_filter (kmer.py:206-207).  Host outputs. */
int32_t pa_extsim_stats(pa_index* idx, const uint32_t* group, uint32_t n_groups, uint64_t* total, uint64_t* unique);
int32_t pa_extsim_pairwise(pa_index* idx, const uint32_t* group, uint32_t n_groups, uint64_t* inter);
/* _remove_filtered_genomes_from_kmers + _update_genomes_list (kmer.py:232-250); keep[g] != 0 survives */
int32_t pa_index_drop_genomes(pa_index* idx, const uint8_t* keep);

/* ---- communicator (SURVEY.md 8(b) "multi-GPU: pa_comm_init, pa_comm_allreduce_summary, pa_build_exchange") ----------
 * One process per GPU on one node.  The reference has no counterpart (it is single-threaded); the multi-GPU split sits
 * behind the same two call sites as everything else: KmerReference.__init__ (kmer.py:113-133) and
 * PseudoAlignment.align_reads_from_container (kmer.py:600-620).
 *   pa_comm_unique_id        rank 0 creates the NCCL id; the caller hands the 128 bytes to every rank (file, socket, MPI,
 *                            torch.distributed ... -- the library does not care)
 *   pa_comm_init             ncclCommInitRank on `device`.  NCCL is resolved with dlopen at this call (the copy already in
 *                            the process if there is one, else PA_NCCL_LIB, else libnccl.so.2), so the library itself
 *                            links only the CUDA runtime
 *   pa_comm_init_callbacks   the same communicator over a caller-supplied host all-gather instead of NCCL (tests over
 *                            gloo, MPI bootstraps): allgather(user, in, out, bytes) fills out[r * bytes ..] with rank r's
 *                            `bytes`, returns 0 on success; device data then travels through CUDA IPC peer memory only.
 *                            device < 0: host-only communicator (host collectives, no CUDA device needed)
 *   pa_comm_allreduce_summary  the ONE exchange of read-sharded alignment: SUM over d_sum[n_sum] = {stats[4],
 *                            unique_reads[G], ambiguous_reads[G], counters[3] ...} and MIN over d_min[n_min] =
 *                            first_seen[G] (uint64, device memory, in place; both reductions in one NCCL group on
 *                            `stream`; with the callback transport the call returns after completion)
 *   pa_comm_allreduce_host   SUM (op 0) or MIN (op 1) over host uint64 (EXTSIM: per-class totals and the G x G
 *                            intersection matrix are sums over disjoint key ranges, kmer.py:152-177, 206-207) */
typedef struct pa_comm pa_comm;
typedef struct pa_comm_callbacks {
  void* user;
  int32_t (*allgather)(void* user, const void* in, void* out, uint64_t bytes_per_rank);
} pa_comm_callbacks;
int32_t pa_comm_unique_id(uint8_t id[128]);
int32_t pa_comm_init(int32_t n_ranks, int32_t rank, const uint8_t id[128], int32_t device, pa_comm** out);
int32_t pa_comm_init_callbacks(int32_t n_ranks, int32_t rank, int32_t device, const pa_comm_callbacks* cb, pa_comm** out);
int32_t pa_comm_free(pa_comm* comm);
int32_t pa_comm_info(pa_comm* comm, int32_t* n_ranks, int32_t* rank, int32_t* device, int32_t* has_nccl);
int32_t pa_comm_allreduce_summary(pa_comm* comm, uint64_t* d_sum, uint64_t n_sum, uint64_t* d_min, uint64_t n_min, void* stream);
int32_t pa_comm_allreduce_host(pa_comm* comm, uint64_t* values, uint64_t n, int32_t op);
/* out[r * bytes_per_rank ..] = rank r's `bytes_per_rank` bytes (host memory) */
int32_t pa_comm_allgather_host(pa_comm* comm, const void* in, void* out, uint64_t bytes_per_rank);
int32_t pa_comm_barrier(pa_comm* comm);

/* ---- partitioned / streamed index build (SURVEY.md 8(e) "Build: one exchange step"; kmer.py:135-150 across ranks) -----
 * Rank r holds the genomes [g_lo, g_hi) (FASTA order is kept: genome indices ascend with the rank).  K1 encodes them and
 * names the OWNER of every record: the k-mer space is partitioned by the top bits ("digit") of the k-mer's minimizer
 * hash, which is also what the lookup table is laid out by, so the rank that owns a key range owns a contiguous slice of
 * the table.  One stable scatter pass then stores every record straight into the owner's receive buffer -- peer memory
 * mapped through CUDA IPC, one long coalesced run per (tile, owner), the stores travel over NVLink -- so partition and
 * all-to-all are one kernel (pa_build_exchange); every rank sorts + run-length encodes what it received (K2, K3) into a
 * CSR partition, inserts those k-mers into ITS slice of the (full-size) lookup table, and the slices are all-gathered
 * in place (NCCL broadcasts, or IPC pulls): the replica every rank aligns against is the table alone.
 *   flags  PA_BUILD_TABLE_ONLY   keep no CSR: *partition is not produced, records carry genome indices instead of
 *                                positions (no 2^32-base limit), and the key space may be processed in several ROUNDS
 *                                per rank (n_rounds; 0 = as many as the device memory asks for) -- config E's 2,000
 *                                genomes build on one GPU that way
 *          PA_BUILD_HOST_BASES   `bases` is host memory (uploaded chunk by chunk); default: device memory, 32-byte aligned
 *   bases  the genomes [g_lo, g_hi) concatenated; genome_off[n_genomes + 1] = offsets of ALL genomes (positions and
 *          genome indices are global).  comm == NULL: single GPU.
 *   *partition  CSR of the key range this rank owns, with positions (pa_index_export, pa_extsim_*, pa_index_drop_genomes);
 *   *replica    align-only index of all keys (pa_align_batch*, pa_debug_table_lookup, pa_index_info_get)
 * pa_index_rebuild_replica: the table again from the partitions, after pa_index_drop_genomes on every partition. */
#define PA_BUILD_TABLE_ONLY 1u
#define PA_BUILD_HOST_BASES 2u
int32_t pa_index_build_partitioned(pa_comm* comm, const uint8_t* bases, const uint64_t* genome_off, uint32_t n_genomes,
                                   uint32_t g_lo, uint32_t g_hi, int32_t k, int32_t device, uint32_t flags, uint32_t n_rounds,
                                   pa_index** partition, pa_index** replica);
int32_t pa_index_rebuild_replica(pa_comm* comm, pa_index* partition, pa_index** replica);
/* the genome range [*g_lo, *g_hi) of `rank`: contiguous runs of whole genomes balanced by bases */
int32_t pa_genome_shard(const uint64_t* genome_off, uint32_t n_genomes, int32_t n_ranks, int32_t rank, uint32_t* g_lo,
                        uint32_t* g_hi);
/* The exchange step alone, for callers that drive the phases themselves: n records (device memory) with owner[i] = the
 * rank record i goes to (255: dropped).  On return *recv_keys / *recv_vals (device memory owned by the communicator,
 * valid until its next exchange) hold the *n_recv records this rank owns: sender-major, inside a sender in input order. */
int32_t pa_build_exchange(pa_comm* comm, const uint64_t* d_keys, const uint32_t* d_vals, const uint8_t* d_owner, uint64_t n,
                          uint64_t** recv_keys, uint32_t** recv_vals, uint64_t* n_recv, void* stream);
/* timings (ms) of the phases of the last pa_index_build_partitioned on this rank:
 * [0] encode + count, [1] scatter + exchange, [2] sort, [3] CSR, [4] table slice, [5] table gather, [6] total */
int32_t pa_build_timings(pa_index* replica, float ms[8]);
/* the owner rank of a k-mer (ASCII, k bytes) when the key space is split over n_parts */
int32_t pa_partition_of_kmer(int32_t k, const uint8_t* kmer_ascii, uint32_t n_parts, uint32_t* part);

/* ---- alignment: PseudoAlignment.align_reads_from_container (kmer.py:563-620) over a packed batch ---
 * bases/quals: concatenated read strings (quals may be NULL when no quality filter is on);
 * read_off[n_reads+1].  Outputs, one 64-bit word per read:
 *   bits 63:62  0 = dropped by min_read_quality (kmer.py:587-589), 1/2/3 = ReadMappingType value
 *   bits 61:40  length of genomes_mapped_to
 *   bits 39:0   the genome index when the length is 1, else the offset of the list in out_list
 * counters[3] += {filtered_quality_reads, filtered_quality_kmers, filtered_hr_kmers} (kmer.py:587-597).
 * *list_len receives the number of out_list entries needed; PA_ERR_CAPACITY when it exceeds list_cap
 * (words are then valid, lists are not: call again with a larger out_list). */
int32_t pa_align_batch(pa_index* idx, const uint8_t* bases, const uint8_t* quals, const uint64_t* read_off,
                       uint64_t n_reads, const pa_align_params* params, uint64_t* out_words, uint32_t* out_list,
                       uint64_t list_cap, uint64_t* list_len, uint64_t counters[3]);
/* Reads that were packed once (e.g. when the FASTQ was ingested) instead of on every call: pa_pack_reads turns the
 * concatenated ACGT read strings into the two bit planes K4's ballots would compute -- planes needs
 * 2 * (n_bases / 32 + n_reads + 1) uint32 words; read i, starting at base offset o, owns the words from
 * 2 * ((o - read_off[0]) / 32 + i): ceil(L / 32) low-plane words, then as many high-plane words (bit j of low word c =
 * bit 1 of the ASCII code of base 32 c + j, high word = bit 2; A=0 C=1 T=2 G=3).  *all_acgt = 0 when a base outside ACGT
 * was met: such a batch must go through pa_align_batch (a non-ACGT window can never match but still counts for the
 * quality filter, kmer.py:420-422).  pa_align_batch_packed = pa_align_batch on such planes: a quarter of the bytes
 * cross PCIe and no host core touches the reads again.  quals (may be NULL without quality filters) and read_off as in
 * pa_align_batch (quals is indexed with the absolute offsets). */
int32_t pa_pack_reads(const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t* planes, uint64_t planes_cap,
                      int32_t* all_acgt);
int32_t pa_align_batch_packed(pa_index* idx, const uint32_t* planes, const uint8_t* quals, const uint64_t* read_off,
                              uint64_t n_reads, const pa_align_params* params, uint64_t* out_words, uint32_t* out_list,
                              uint64_t list_cap, uint64_t* list_len, uint64_t counters[3]);
/* device-resident variant: every pointer except params is device memory; d_state is 5 x uint64 of device
 * memory = {list cursor, flag, counters[3]}, zeroed by the caller; asynchronous on `stream`.  flag: 1 = out_list too
 * small (the cursor holds the size needed), 2 = a read was longer than max_read_len (results invalid). */
int32_t pa_align_batch_device(pa_index* idx, const uint8_t* d_bases, const uint8_t* d_quals, const uint64_t* d_read_off,
                              uint64_t n_reads, uint64_t max_read_len, const pa_align_params* params,
                              uint64_t* d_words, uint32_t* d_list, uint64_t list_cap, uint64_t* d_state, void* stream,
                              int32_t* n_launches);

/* ---- summary: PseudoAlignment.get_summary (kmer.py:622-657) -----------------------------------------
 * Device accumulators owned by the caller (so ranks can all-reduce them): stats[4] = {unique, ambiguous,
 * unmapped, dropped} (SUM), unique_reads[G], ambiguous_reads[G] (SUM, one per list element),
 * first_seen[G] (MIN of (read_index_base + i) << 22 | list position; initialise to UINT64_MAX). */
int32_t pa_summary_reduce_device(const uint64_t* d_words, const uint32_t* d_list, uint64_t n_reads,
                                 uint64_t read_index_base, uint32_t n_genomes, uint64_t* d_stats, uint64_t* d_unique_reads,
                                 uint64_t* d_ambiguous_reads, uint64_t* d_first_seen, void* stream);
/* host convenience over the outputs of pa_align_batch */
int32_t pa_summary_reduce(pa_index* idx, const uint64_t* words, const uint32_t* list, uint64_t n_reads, uint64_t list_len,
                          uint64_t read_index_base, uint64_t stats[4], uint64_t* unique_reads, uint64_t* ambiguous_reads,
                          uint64_t* first_seen);

/* ---- ingest: FASTA / FASTQ text -> packed arrays (records.py:141-199, 212-302; pure host code) -------------
 * Parses the canonical form of both formats (4-line FASTQ over ACGT with qualities in ASCII 33..126; FASTA with a
 * '>' line and ACGTN lines) into the arrays pa_index_build / pa_align_batch take.  *canonical = 0 (and *out = NULL)
 * means the text has to go through the regex restatement of the reference's parser instead (records.py of this
 * package), which keeps the reference's acceptance rules, exception types, messages and precedence for every
 * unusual input (blank lines, lower case, duplicates, length mismatches, stray text ...).  ASCII text only.
 * pa_parsed_copy: seq[n_bases], qual[n_bases] (FASTQ), seq_off[n_records + 1], name_beg / name_len[n_records] =
 * the stripped identifier / description inside `text`; plus_beg / plus_len (FASTQ) = the text after '+'. */
typedef struct pa_parsed pa_parsed;
int32_t pa_parse_records(const uint8_t* text, uint64_t n, int32_t fastq, pa_parsed** out, int32_t* canonical,
                         uint64_t* n_records, uint64_t* n_bases);
int32_t pa_parsed_copy(pa_parsed* h, uint8_t* seq, uint8_t* qual, uint64_t* seq_off, uint64_t* name_beg, uint64_t* name_len,
                       uint64_t* plus_beg, uint64_t* plus_len);
int32_t pa_parsed_free(pa_parsed* h);

/* ---- dumpref: the "Kmers" object of KmerReference.get_summary (kmer.py:300-329) as JSON text (pure host code) ----
 * Byte for byte what json.dumps(..., indent=`indent`) prints for that object when it sits `level` levels deep: k-mers in
 * dict insertion order (`order`), per k-mer one entry per description class in order of first appearance (the last
 * genome of a class wins, like the dict assignment), positions ascending.  Inputs: the arrays of pa_index_export;
 * desc_class[g] = class of genome g, desc_json = the classes' descriptions already JSON-escaped (quotes included),
 * concatenated, desc_json_off[n_classes + 1].  *out_text is malloc'ed by the library: release with pa_free_text. */
int32_t pa_format_kmers_json(int32_t k, uint64_t n_keys, const uint64_t* keys, const uint32_t* order, const uint64_t* run_off,
                             const uint32_t* run_genome, const uint64_t* pos_off, const uint32_t* pos,
                             const uint32_t* desc_class, const uint8_t* desc_json, const uint64_t* desc_json_off,
                             int32_t indent, int32_t level, uint8_t** out_text, uint64_t* out_len);
int32_t pa_free_text(uint8_t* p);

/* ---- diagnostics used by the tests ----------------------------------------------------------------- */
/* device pointers of an index's keys[n_keys], run_off[n_keys+1], run_genome[n_runs] (valid until the index changes) */
int32_t pa_index_csr_device(pa_index* idx, uint64_t** d_keys, uint64_t** d_run_off, uint32_t** d_run_genome);
/* stable LSD radix sort of (key, value) pairs on key bits [0, end_bit), host in / host out (K2) */
int32_t pa_debug_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t n, int32_t end_bit, int32_t device);
/* the sort the index builds use for their hashed k-mer keys: digit passes over the top `top_bits` bits only (0 = chosen
 * from n), then a stable repair of the runs of equal top bits that hold different keys; *fell_back = 1 when the keys were
 * not spread evenly enough and the remaining passes ran after all.  Same result as pa_debug_sort_pairs on any input. */
int32_t pa_debug_sort_pairs_hashed(uint64_t* keys, uint32_t* vals, uint64_t n, int32_t end_bit, int32_t top_bits, int32_t device,
                                   int32_t* fell_back);
/* the host-side 2-bit packing of pa_align_batch (pure host code): planes needs 2 * (n_bases / 32 + n_reads + 1) words;
 * read i owns the words from 2 * ((read_off[i] - read_off[0]) / 32 + i): ceil(L / 32) low-plane words, then as many
 * high-plane words.  *all_acgt = 0 when a base outside ACGT was met. */
int32_t pa_debug_pack_reads(const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t* planes, uint64_t planes_cap,
                            int32_t n_threads, int32_t* all_acgt);
/* minimizer of each k-mer (pure host code, the function the table build and K4 share): mhash = bijective hash of its m-mer,
 * m = min(k, 16); offset = position of that m-mer inside the k-mer (leftmost among equal orders) */
int32_t pa_debug_minimizer(int32_t k, const uint8_t* kmers_ascii, uint64_t n, uint32_t* mhash, uint32_t* offset);
/* direct table lookups (K4's lookup step): n_genomes[i] = number of genomes of k-mer i (0 = miss),
 * first_genome[i] = its smallest genome index */
int32_t pa_debug_table_lookup(pa_index* idx, const uint8_t* kmers_ascii, uint64_t n, uint32_t* n_genomes,
                              uint32_t* first_genome);

#ifdef __cplusplus
}
#endif
#endif /* PA_B200_H */
