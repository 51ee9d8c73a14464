"""
multi_gpu.py -- read-sharded alignment across ranks (one process per GPU).

Reads are independent units (the loop of PseudoAlignment.align_reads_from_container,
/root/reference/src/kmer.py:616-620), so rank r aligns one contiguous block of the reads against a
replicated index and the only exchange is the summary of get_summary (kmer.py:622-657):

    SUM  stats[4] + unique_reads[G] + ambiguous_reads[G]      (int64)
    MIN  first_seen[G] = (global read index << 22) | list position   (orders the "Summary" keys)

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NEVER = np.uint64(0xFFFFFFFFFFFFFFFF)
_INT64_MAX = 2 ** 63 - 1


def shard_bounds(n_reads: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; keeps file order inside a rank and across ranks."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def allreduce_summary(acc, first_seen, group=None) -> None:
    """In-place all-reduce of the K8 accumulators (torch tensors, int64 views of the uint64 device arrays).

    first_seen holds uint64 order keys < 2^63 or the all-ones "never seen" marker, i.e. -1 as int64; mapping -1 to
    int64 max turns the unsigned MIN into a signed one."""
    import torch
    import torch.distributed as dist
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    fs = torch.where(first_seen < 0, torch.full_like(first_seen, _INT64_MAX), first_seen)
    dist.all_reduce(fs, op=dist.ReduceOp.MIN, group=group)
    first_seen.copy_(torch.where(fs == _INT64_MAX, torch.full_like(fs, -1), fs))


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
