"""
multi_gpu.py -- multi-GPU paths (one process per GPU): read-sharded alignment and the partitioned index build.

Reads are independent units (the loop of PseudoAlignment.align_reads_from_container,
/root/reference/src/kmer.py:616-620), so rank r aligns one contiguous block of the reads against a
replicated index and the only exchange is the summary of get_summary (kmer.py:622-657):

    SUM  stats[4] + unique_reads[G] + ambiguous_reads[G] + counters[3]      (uint64)
    MIN  first_seen[G] = (global read index << 22) | list position          (orders the "Summary" keys)

The index build (SURVEY.md 8(e) "Build") has one exchange step; it runs entirely behind the C ABI
(pa_index_build_partitioned, csrc/dist.cu): every rank encodes a run of whole genomes, one scatter kernel stores every
record straight into the receive buffer of the rank that owns its key range (peer memory over NVLink), every rank
sorts + run-length encodes its range and fills its slice of the lookup table, and the slices are all-gathered.  EXTSIM
statistics are sums over disjoint key ranges: computed per partition and all-reduced.

All collectives go through pa_comm (_native.Comm): NCCL, or a host all-gather supplied by the caller (the tests use
torch.distributed's gloo backend for that).  `init()` makes a process-wide context that KmerReference and
PseudoAlignment pick up, so the same two calls of the reference's API run on every GPU of the node under torchrun.
"""
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NEVER = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_bounds(n_reads: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; keeps file order inside a rank and across ranks."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def summary_from_accumulators(stats: Sequence[int], unique_reads: Sequence[int], ambiguous_reads: Sequence[int],
                              first_seen: np.ndarray, genome_ids: List[str], flags: Tuple[bool, bool, bool],
                              counters: Sequence[int]) -> Dict[str, Dict]:
    """Builds the reference's get_summary() dict (key order included) from reduced accumulators."""
    statistics = {"unique_mapped_reads": int(stats[0]), "ambiguous_mapped_reads": int(stats[1]), "unmapped_reads": int(stats[2])}
    if flags[0]:
        statistics["filtered_quality_reads"] = int(counters[0])
    if flags[1]:
        statistics["filtered_quality_kmers"] = int(counters[1])
    if flags[2]:
        statistics["filtered_hr_kmers"] = int(counters[2])
    first = np.asarray(first_seen).astype(np.uint64)
    summary: Dict[str, Dict[str, int]] = {}
    for g in np.argsort(first, kind="stable"):
        if first[g] == NEVER:
            break
        row = summary.setdefault(genome_ids[int(g)], {"unique_reads": 0, "ambiguous_reads": 0})
        row["unique_reads"] += int(unique_reads[g])
        row["ambiguous_reads"] += int(ambiguous_reads[g])
    return {"Statistics": statistics, "Summary": summary}


def reduce_summary(comm, stats, unique_reads, ambiguous_reads, counters, first_seen):
    """The one exchange of read-sharded alignment (host arrays): SUM of the counts, MIN of the order keys."""
    G = len(unique_reads)
    acc = comm.allreduce_host(np.concatenate([np.asarray(stats, dtype=np.uint64), np.asarray(unique_reads, dtype=np.uint64),
                                              np.asarray(ambiguous_reads, dtype=np.uint64), np.asarray(counters, dtype=np.uint64)]))
    fs = comm.allreduce_host(np.asarray(first_seen, dtype=np.uint64), take_min=True)
    return acc[:4], acc[4:4 + G], acc[4 + G:4 + 2 * G], acc[4 + 2 * G:4 + 2 * G + 3], fs


# ---------------------------------------------------------------------------
# process-wide context: which communicator KmerReference / PseudoAlignment use
# ---------------------------------------------------------------------------
_CTX: Optional[Dict] = None


def init(device: Optional[int] = None, group=None, comm=None) -> Dict:
    """Attach this process to its ranks.  Under torchrun (RANK / WORLD_SIZE / LOCAL_RANK set) torch.distributed is
    initialised with NCCL when it has not been yet; `comm` may also be handed in directly.  Returns the context."""
    global _CTX
    import _native as nat
    if comm is None:
        import torch
        import torch.distributed as dist
        backend = os.environ.get("PA_DIST_BACKEND", "nccl")   # "gloo": several ranks on fewer GPUs (the 1-GPU test box)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            if backend != "nccl":
                device %= max(torch.cuda.device_count(), 1)
        if not dist.is_initialized():
            torch.cuda.set_device(device)
            if backend == "nccl":
                dist.init_process_group("nccl", device_id=torch.device("cuda", device))
            else:
                dist.init_process_group(backend)
        comm = nat.Comm.from_torch(device, group)
    inf = comm.info()
    _CTX = {"comm": comm, "rank": inf["rank"], "world": inf["n_ranks"], "device": inf["device"]}
    return _CTX


def context() -> Optional[Dict]:
    return _CTX if _CTX is not None and _CTX["world"] > 1 else None


def shutdown() -> None:
    global _CTX
    if _CTX is not None:
        _CTX["comm"].close()
    _CTX = None


# ---------------------------------------------------------------------------
# partitioned index
# ---------------------------------------------------------------------------
class DistributedIndex:
    """partition = CSR of the key range this rank owns (with positions); replica = table-only align index of all keys."""

    def __init__(self, comm, partition, replica, k: int, genome_off: np.ndarray):
        self.comm, self.partition, self.replica = comm, partition, replica
        self.k, self.genome_off = int(k), np.asarray(genome_off, dtype=np.uint64)
        inf = comm.info() if comm is not None else {"rank": 0, "n_ranks": 1}     # comm None: one GPU (streamed builds)
        self.rank, self.world = inf["rank"], inf["n_ranks"]
        self.timings: Dict[str, float] = {}

    def close(self):
        for ix in (self.partition, self.replica):
            if ix is not None:
                ix.close()
        self.partition = self.replica = None

    # -- EXTSIM: per-class sums are additive over key ranges (kmer.py:152-177, 206-207) --
    def extsim_stats(self, group_ids: np.ndarray, n_groups: int):
        total, uniq = self.partition.extsim_stats(group_ids, n_groups)
        a = self._sum(np.concatenate([total, uniq]))
        return a[:n_groups], a[n_groups:]

    def extsim_pairwise(self, group_ids: np.ndarray, n_groups: int) -> np.ndarray:
        inter = self.partition.extsim_pairwise(group_ids, n_groups)
        return self._sum(inter.reshape(-1)).reshape(n_groups, n_groups)

    def _sum(self, a: np.ndarray) -> np.ndarray:
        return self.comm.allreduce_host(a) if self.comm is not None else np.asarray(a, dtype=np.uint64)

    def _gather(self, a: np.ndarray):
        return self.comm.allgather_array(a) if self.comm is not None else [np.asarray(a)]

    def drop_genomes(self, keep: np.ndarray) -> None:
        """_remove_filtered_genomes_from_kmers (kmer.py:232-250) on every partition, then the table is rebuilt."""
        import _native as nat
        self.partition.drop_genomes(keep)
        keep = np.asarray(keep).astype(bool)
        lens = np.diff(self.genome_off.astype(np.int64))[keep]
        self.genome_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        self.replica.close()
        self.replica = nat.rebuild_replica(self.comm, self.partition)

    def sizes(self) -> Tuple[int, int, int]:
        inf = self.replica.info()
        return int(inf.n_keys), int(inf.n_runs), int(inf.n_occ)

    def export_gathered(self):
        """CSR of the whole index on every rank (dumpref / kmers view / pickling): the partitions' exports merged in key
        order; `order` sorts all keys by first occurrence, i.e. the reference's dict insertion order."""
        ex = self.partition.export(with_positions=True, with_order=False)
        parts = {name: self._gather(ex[name]) for name in ("keys", "run_off", "run_genome", "pos_off", "pos", "first_occ")}
        keys = np.concatenate(parts["keys"])
        runs_per_key = np.concatenate([np.diff(r.astype(np.int64)) for r in parts["run_off"]])
        pos_per_run = np.concatenate([np.diff(r.astype(np.int64)) for r in parts["pos_off"]])
        run_genome = np.concatenate(parts["run_genome"])
        pos = np.concatenate(parts["pos"])
        first = np.concatenate(parts["first_occ"])
        # partitions are ranges of the minimizer space, not of the key space: merge by key
        by_key = np.argsort(keys, kind="stable")
        run_start = np.concatenate([[0], np.cumsum(runs_per_key)])[:-1]
        pos_start = np.concatenate([[0], np.cumsum(pos_per_run)])[:-1]
        new_runs = runs_per_key[by_key]
        run_off = np.concatenate([[0], np.cumsum(new_runs)]).astype(np.uint64)
        run_idx = (np.repeat(run_start[by_key], new_runs) + _ramp(new_runs)).astype(np.int64)
        new_pos_counts = pos_per_run[run_idx]
        pos_off = np.concatenate([[0], np.cumsum(new_pos_counts)]).astype(np.uint64)
        pos_idx = (np.repeat(pos_start[run_idx], new_pos_counts) + _ramp(new_pos_counts)).astype(np.int64)
        out = {"keys": keys[by_key], "run_off": run_off, "run_genome": run_genome[run_idx], "pos_off": pos_off,
               "pos": pos[pos_idx], "first_occ": first[by_key]}
        out["order"] = np.argsort(out["first_occ"], kind="stable").astype(np.uint32)
        return out


def _ramp(counts: np.ndarray) -> np.ndarray:
    """0..c-1 for every c of counts, concatenated."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)])[:-1]
    return np.arange(total, dtype=np.int64) - np.repeat(starts, counts)


def build_partitioned(comm, bases, genome_off: np.ndarray, k: int, device: int = 0, table_only: bool = False,
                      n_rounds: int = 0, g_range: Optional[Tuple[int, int]] = None) -> DistributedIndex:
    """Multi-GPU KmerReference build (kmer.py:135-150 across ranks; pa_index_build_partitioned).

    bases       ALL genomes concatenated as a uint8 numpy array (this rank uploads only its share), or -- with g_range --
                this rank's genomes [g_lo, g_hi) as a numpy array or a device pointer (int)
    genome_off  uint64 offsets of ALL genomes (G + 1 entries); positions and genome indices are global
    """
    import _native as nat
    genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
    inf = comm.info() if comm is not None else {"rank": 0, "n_ranks": 1}
    if g_range is None:
        g_lo, g_hi = nat.genome_shard(genome_off, inf["n_ranks"], inf["rank"])
        mine = np.asarray(bases)[int(genome_off[g_lo] - genome_off[0]):int(genome_off[g_hi] - genome_off[0])]
    else:
        (g_lo, g_hi), mine = g_range, bases
    partition, replica = nat.build_partitioned(comm, mine, genome_off, (g_lo, g_hi), k, device=device, table_only=table_only,
                                               n_rounds=n_rounds)
    out = DistributedIndex(comm, partition, replica, k, genome_off - genome_off[0])
    out.timings = nat.build_timings(replica)
    return out


# ---------------------------------------------------------------------------
# read-sharded alignment, end to end
# ---------------------------------------------------------------------------
def align_sharded(index, comm, bases: np.ndarray, quals: Optional[np.ndarray], read_off: np.ndarray, genome_ids: List[str],
                  m: int = 1, p: int = 1, min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                  max_genomes: Optional[int] = None, gather_reads: bool = False):
    """PseudoAlignment.align_reads_from_container + get_summary (kmer.py:600-657) across ranks.

    Every rank holds the whole packed batch (bases / quals / read_off as produced by the native ingest) and a replicated
    align index (`NativeIndex`, e.g. DistributedIndex.replica); rank r aligns the contiguous block shard_bounds(n, world, r),
    the K8 accumulators are reduced through the communicator, and every rank returns the reference's summary dict (key
    order included).  gather_reads=True additionally returns the per-read results of ALL reads in file order on every rank
    as arrays: (types uint8[n], lens int32[n], flat genome indices uint32[sum lens]) plus the three filter counters.
    """
    import _native as nat
    import synth
    inf = comm.info()
    world, rank = inf["n_ranks"], inf["rank"]
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    n = len(read_off) - 1
    lo, hi = shard_bounds(n, world, rank)
    params = nat.make_params(m, p, min_read_quality, min_kmer_quality, max_genomes)
    words, lst, counters = index.align(bases, quals, read_off[lo:hi + 1], params)
    stats, uniq, amb, first = index.summary(words, lst, read_index_base=lo)
    stats, uniq, amb, counters, first = reduce_summary(comm, stats, uniq, amb, counters, first)
    flags = (min_read_quality is not None, min_kmer_quality is not None, max_genomes is not None)
    summary = summary_from_accumulators(stats, uniq, amb, first, genome_ids, flags, counters)
    if not gather_reads:
        return summary
    types, lens, flat = synth.flatten_results(words, lst)
    all_types = np.concatenate(comm.allgather_array(types))
    all_lens = np.concatenate(comm.allgather_array(lens))
    all_flat = np.concatenate(comm.allgather_array(flat))
    return summary, (all_types, all_lens, all_flat, [int(c) for c in counters])
