// table.cuh -- construction of the minimizer-bucketed lookup table (TableView in common.cuh) from a CSR.
// The steps are separate so that a multi-GPU / streamed build (dist.cu) can fix the geometry once, then fill the
// table slice by slice (one CSR partition per rank and round), and build the stash at the end.
#pragma once
#include "index.cuh"
#include <vector>

namespace pa {

struct TableGeom {
  uint32_t bpd = 1, hi_bits = 0, tag_bits = 1, val_bits = 63, gbits = 1, n_inline = 1;
  uint32_t payload_bits = 0;   // val_bits - 3: genome id / inline list / mlist sector
};

// Geometry for (an estimate of) U distinct k-mers of G genomes at the wanted load factor over the slots.
// min_bpd: never fewer blocks per digit than this (a retry after "list references do not fit" doubles it).
int32_t table_geometry(int k, uint32_t G, uint64_t U, double load, uint32_t min_bpd, TableGeom* g);
// the load factor to use when nothing was asked for: PA_TABLE_LOAD, else 1/5, raised (denser) up to 0.42 while the
// table would take more than 40 % of the free device memory
double table_load_factor(int k, uint64_t U, size_t free_bytes);
void apply_geometry(Index& ix, const TableGeom& g);

struct CsrView {
  const uint64_t* ukeys;
  const uint64_t* run_off;
  const uint32_t* run_genome;
  uint64_t U;
};

// de-duplicated genome sets of the k-mers with more than n_inline genomes (see table.cu)
struct GenomeSets {
  DevBuf msec_off;    // [U] first sector of k-mer u's set inside `mlist` (meaningful for k-mers with long lists only)
  DevBuf mlist;       // n_msec sectors of 8 ids
  uint64_t n_msec = 0;
};
int32_t build_genome_sets(cudaStream_t s, const CsrView& c, uint32_t n_inline, GenomeSets* out);

// Inserts the k-mers of `c` into ix.slots (allocated and cleared by the caller).  msec_base is added to the sector
// index of every mlist payload.  K-mers that find no free slot inside their chain are appended to `overflow` as
// {raw k-mer key, value} pairs (for the stash).
int32_t table_insert_csr(Index& ix, const CsrView& c, const GenomeSets& sets, uint64_t msec_base, std::vector<uint64_t>* overflow);
int32_t table_build_stash(Index& ix, const std::vector<uint64_t>& pairs);

// one-shot: geometry, sets, slots, stash of an index from its own CSR
int32_t index_build_tables(Index& ix);

}  // namespace pa
