"""
multi_gpu.py -- multi-GPU paths (one process per GPU): read-sharded alignment and the hash-partitioned index build.

Reads are independent units (the loop of PseudoAlignment.align_reads_from_container,
/root/reference/src/kmer.py:616-620), so rank r aligns one contiguous block of the reads against a
replicated index and the only exchange is the summary of get_summary (kmer.py:622-657):

    SUM  stats[4] + unique_reads[G] + ambiguous_reads[G]      (int64)
    MIN  first_seen[G] = (global read index << 22) | list position   (orders the "Summary" keys)

The index build (SURVEY.md 8(e) "Build") has one exchange step: every rank encodes a run of whole genomes, splits its
(hashed k-mer, position) records by key range, one all-to-all moves every record to the rank that owns its range, and
every rank sorts + run-length encodes its range (`build_partitioned`).  The align index is replicated by gathering
the partitions' keys and genome runs into a replica whose lookup table every rank builds.  EXTSIM statistics are
sums over disjoint key ranges: computed per partition and all-reduced.

torch.distributed is the plumbing (NCCL over NVLink on GPUs; gloo with host staging in the tests).
"""
import ctypes
import os
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


# ---------------------------------------------------------------------------
# hash-partitioned index build
# ---------------------------------------------------------------------------
def genome_shards(genome_lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous runs [g_lo, g_hi) of whole genomes per rank, balanced by bases (FASTA order is kept, so genome
    indices ascend with the rank and equal k-mers arrive at their owner in genome order)."""
    total = int(sum(int(x) for x in genome_lengths))
    bounds, g, acc = [0], 0, 0
    n = len(genome_lengths)
    for r in range(1, world):
        target = total * r // world
        while g < n and acc + int(genome_lengths[g]) // 2 < target:
            acc += int(genome_lengths[g])
            g += 1
        bounds.append(g)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class _DevArray:
    """A raw device pointer dressed as a CUDA array for torch.as_tensor (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _as_tensor(ptr: int, n: int, typestr: str, dev):
    import torch
    dt = {"<i8": torch.int64, "<i4": torch.int32}[typestr]
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dt, device=dev)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=dev)


def _is_nccl(group=None) -> bool:
    import torch.distributed as dist
    return dist.get_backend(group) == "nccl"


def _all_to_all(out, inp, out_splits, in_splits, group=None):
    """all_to_all_single; staged through the host when the backend cannot move device tensors (gloo)."""
    import torch
    import torch.distributed as dist
    if _is_nccl(group):
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return
    world = dist.get_world_size(group)
    src = inp.cpu()
    parts = list(torch.split(src, in_splits)) if world > 0 else []
    got = [None] * world
    gathered = [None] * world
    dist.all_gather_object(gathered, [p.numpy() for p in parts], group=group)
    rank = dist.get_rank(group)
    for r in range(world):
        got[r] = torch.from_numpy(gathered[r][rank].copy()) if len(gathered[r][rank]) else torch.empty(0, dtype=inp.dtype)
    cat = torch.cat(got) if got else torch.empty(0, dtype=inp.dtype)
    assert cat.numel() == out.numel()
    out.copy_(cat)


def _broadcast(t, src: int, group=None):
    import torch.distributed as dist
    if t.numel() == 0:
        return
    if _is_nccl(group):
        dist.broadcast(t, src, group=group)
    else:
        c = t.cpu()
        dist.broadcast(c, src, group=group)
        if dist.get_rank(group) != src:
            t.copy_(c)


class DistributedIndex:
    """partition = CSR of the key range this rank owns (with positions); replica = align-only index of all keys."""

    def __init__(self, partition, replica, k, genome_off, rank, world, group=None):
        self.partition, self.replica = partition, replica
        self.k, self.genome_off, self.rank, self.world, self.group = k, genome_off, rank, world, group
        self.timings: Dict[str, float] = {}

    def close(self):
        for ix in (self.partition, self.replica):
            if ix is not None:
                ix.close()
        self.partition = self.replica = None


def _gather_replica(partition, k: int, genome_off: np.ndarray, dev, group=None):
    """Replicates the align index: keys / run offsets / genome runs of every partition, in key order = rank order."""
    import torch
    import torch.distributed as dist
    import _native as nat
    L = nat.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    inf = partition.info()
    sizes = torch.tensor([inf.n_keys, inf.n_runs, inf.n_occ], dtype=torch.int64)
    all_sizes = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    if _is_nccl(group):
        tmp = [t.to(dev) for t in all_sizes]
        dist.all_gather(tmp, sizes.to(dev), group=group)
        all_sizes = [t.cpu() for t in tmp]
    else:
        dist.all_gather(all_sizes, sizes, group=group)
    U = [int(t[0]) for t in all_sizes]
    R = [int(t[1]) for t in all_sizes]
    N = [int(t[2]) for t in all_sizes]
    U_off = np.concatenate([[0], np.cumsum(U)]).astype(np.int64)
    R_off = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    h = ctypes.c_void_p()
    G = len(genome_off) - 1
    nat.check(L.pa_index_alloc_replica(int(k), G, nat._p(genome_off), int(U_off[-1]), int(R_off[-1]), int(sum(N)),
                                       dev.index or 0, ctypes.byref(h)))
    replica = nat.NativeIndex(h.value)
    pk, po, pg = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    nat.check(L.pa_index_csr_device(replica.handle, ctypes.byref(pk), ctypes.byref(po), ctypes.byref(pg)))
    r_keys = _as_tensor(pk.value, int(U_off[-1]), "<i8", dev)
    r_off = _as_tensor(po.value, int(U_off[-1]), "<i8", dev)
    r_gen = _as_tensor(pg.value, int(R_off[-1]), "<i4", dev)
    nat.check(L.pa_index_csr_device(partition.handle, ctypes.byref(pk), ctypes.byref(po), ctypes.byref(pg)))
    u0, u1, g0, g1 = int(U_off[rank]), int(U_off[rank + 1]), int(R_off[rank]), int(R_off[rank + 1])
    if u1 > u0:
        r_keys[u0:u1].copy_(_as_tensor(pk.value, u1 - u0, "<i8", dev))
        r_off[u0:u1].copy_(_as_tensor(po.value, u1 - u0, "<i8", dev) + g0)   # run offsets become global
    if g1 > g0:
        r_gen[g0:g1].copy_(_as_tensor(pg.value, g1 - g0, "<i4", dev))
    torch.cuda.synchronize(dev)
    for src in range(world):
        _broadcast(r_keys[int(U_off[src]):int(U_off[src + 1])], src, group)
        _broadcast(r_off[int(U_off[src]):int(U_off[src + 1])], src, group)
        _broadcast(r_gen[int(R_off[src]):int(R_off[src + 1])], src, group)
    torch.cuda.synchronize(dev)
    nat.check(L.pa_index_finish_replica(replica.handle))
    return replica


# receive buffers + their peer mappings, kept between builds: cudaIpcOpenMemHandle of a multi-GB buffer costs tens of
# milliseconds (measured: 69 ms open + 12 ms close + 12 ms cudaMalloc for config B on 2 GPUs, against 7 ms for the
# scatter kernel itself), so a process that builds more than once maps them once
_PEER_CACHE: Dict[Tuple[int, int, int], Dict] = {}


def release_peer_buffers() -> None:
    """Unmaps and frees the cached receive buffers of the fused exchange (collective-free; call on every rank)."""
    import _native as nat
    L = nat.lib()
    for (device, _w, _r), ent in list(_PEER_CACHE.items()):
        for p in ent["opened"]:
            L.pa_peer_close(p, device)
        L.pa_peer_free(ent["my_k"], device)
        L.pa_peer_free(ent["my_v"], device)
    _PEER_CACHE.clear()


def _exchange_fused(L, nat, keys, vals, n_slots: int, k: int, device: int, world: int, rank: int, group):
    """Partition + exchange in ONE pass: the stable scatter stores every record into the receive buffer of the rank
    that owns its key range (peer memory through CUDA IPC; NVLink stores).  Returns (recv_keys_ptr, recv_vals_ptr,
    n_recv, n_sent); the receive buffers are cudaMalloc'ed by the library (free with pa_peer_free)."""
    import time
    import torch
    import torch.distributed as dist
    tm = {}
    t0 = time.perf_counter()
    counts = np.zeros(256, dtype=np.uint64)
    begin, tb = ctypes.c_int32(0), ctypes.c_int32(0)
    nat.check(L.pa_records_digit_counts(ctypes.c_void_p(keys.data_ptr()), n_slots, int(k), device, nat._p(counts),
                                        ctypes.byref(begin), ctypes.byref(tb), None))
    tm["digit_counts_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    all_counts = [None] * world
    dist.all_gather_object(all_counts, counts.astype(np.int64), group=group)
    all_counts = np.stack(all_counts)                                  # [sender, digit]
    n_digits = 1 << tb.value
    dest = (np.arange(n_digits, dtype=np.int64) * world) >> tb.value    # part of every real digit (pa_partition_of_key)
    recv_count = [int(all_counts[:, :n_digits][:, dest == r].sum()) for r in range(world)]
    tm["gather_counts_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    # every rank knows every rank's receive count, so all ranks take the same decision without another collective
    key = (device, world, rank)
    ent = _PEER_CACHE.get(key)
    if ent is not None and any(recv_count[r] > ent["cap"][r] for r in range(world)):
        dist.barrier(group=group)      # nobody may still be writing into buffers that are about to go away
        release_peer_buffers()
        ent = None
    tm["reused_mappings"] = ent is not None
    if ent is None:
        cap = [c + c // 16 + 1024 for c in recv_count]
        my_k, my_v = ctypes.c_void_p(), ctypes.c_void_p()
        hk, hv = (ctypes.c_uint8 * 64)(), (ctypes.c_uint8 * 64)()
        peer_k, peer_v = [0] * world, [0] * world
        opened = []
        problem = "peer mapping switched off (PA_TEST_NO_IPC)" if os.environ.get("PA_TEST_NO_IPC") == "1" and rank == world - 1 else None
        try:
            if problem:
                raise RuntimeError(problem)
            nat.check(L.pa_peer_alloc(cap[rank] * 8, device, ctypes.byref(my_k), hk))
            nat.check(L.pa_peer_alloc(cap[rank] * 4, device, ctypes.byref(my_v), hv))
        except (RuntimeError, MemoryError) as e:
            problem = str(e)
        tm["alloc_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
        handles = [None] * world
        dist.all_gather_object(handles, None if problem else (bytes(hk), bytes(hv)), group=group)
        tm["gather_handles_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
        if problem is None and all(h is not None for h in handles):
            try:
                for r in range(world):
                    if r == rank:
                        peer_k[r], peer_v[r] = my_k.value, my_v.value
                    else:
                        pk, pv = ctypes.c_void_p(), ctypes.c_void_p()
                        nat.check(L.pa_peer_open((ctypes.c_uint8 * 64).from_buffer_copy(handles[r][0]), device, ctypes.byref(pk)))
                        opened.append(pk)
                        nat.check(L.pa_peer_open((ctypes.c_uint8 * 64).from_buffer_copy(handles[r][1]), device, ctypes.byref(pv)))
                        opened.append(pv)
                        peer_k[r], peer_v[r] = pk.value, pv.value
            except RuntimeError as e:       # no peer mapping on this system (containers without IPC, no P2P path ...)
                problem = str(e)
        elif problem is None:
            problem = "a peer could not allocate its receive buffer"
        verdicts = [None] * world
        dist.all_gather_object(verdicts, problem, group=group)   # all ranks take the same way out
        if any(v is not None for v in verdicts):
            for p in opened:
                L.pa_peer_close(p, device)
            if my_k.value:
                L.pa_peer_free(my_k, device)
            if my_v.value:
                L.pa_peer_free(my_v, device)
            return None, next(v for v in verdicts if v is not None)
        ent = {"cap": cap, "my_k": my_k, "my_v": my_v, "peer_k": peer_k, "peer_v": peer_v, "opened": opened}
        _PEER_CACHE[key] = ent
    else:
        dist.barrier(group=group)      # the previous build of every rank has consumed its receive buffer
    my_k, my_v, peer_k, peer_v = ent["my_k"], ent["my_v"], ent["peer_k"], ent["peer_v"]
    # receive layout of rank r: sender-major, digit-minor -- equal keys stay in (sender = genome run, position) order
    dst_k = (ctypes.c_void_p * 256)()
    dst_v = (ctypes.c_void_p * 256)()
    for r in range(world):
        digits = np.nonzero(dest == r)[0]
        base = int(all_counts[:rank, :n_digits][:, dest == r].sum())
        for d in digits.tolist():
            dst_k[d] = peer_k[r] + base * 8
            dst_v[d] = peer_v[r] + base * 4
            base += int(all_counts[rank, d])
    torch.cuda.synchronize()
    tm["ipc_open_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    nat.check(L.pa_records_scatter_to_peers(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(vals.data_ptr()), n_slots, int(k),
                                            device, dst_k, dst_v, None))
    tm["scatter_kernel_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    dist.barrier(group=group)          # every rank's stores have landed: the receive buffers are final
    tm["barrier_s"] = time.perf_counter() - t0
    n_sent = int(all_counts[rank, :n_digits].sum())
    return my_k, my_v, recv_count[rank], n_sent, tm


def build_partitioned(my_bases, genome_off: np.ndarray, k: int, g_range: Tuple[int, int], device: int = 0,
                      group=None, fused: Optional[bool] = None) -> DistributedIndex:
    """Multi-GPU KmerReference build (kmer.py:135-150 across ranks).

    my_bases    the genomes [g_lo, g_hi) of this rank concatenated: a uint8 numpy array or a CUDA uint8 tensor
    genome_off  uint64 offsets of ALL genomes (G + 1 entries); positions and genome indices are global
    g_range     (g_lo, g_hi), normally genome_shards(lengths, world)[rank]
    fused       True: the partition pass stores straight into the owners' receive buffers over NVLink (CUDA IPC peer
                memory; one node).  False: partition locally, then all_to_all_single.  Default: fused (PA_FUSED_EXCHANGE=0
                switches it off).
    """
    import time
    import torch
    import torch.distributed as dist
    import _native as nat
    nat.require_device()
    L = nat.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
    G = len(genome_off) - 1
    g_lo, g_hi = g_range
    n_bases = int(genome_off[g_hi] - genome_off[g_lo])
    if isinstance(my_bases, np.ndarray):
        d_bases = torch.from_numpy(np.ascontiguousarray(my_bases, dtype=np.uint8)).to(dev)
    else:
        d_bases = my_bases
    assert d_bases.numel() >= n_bases
    t = {}
    t0 = time.perf_counter()
    keys = torch.empty(max(n_bases, 1), dtype=torch.int64, device=dev)
    vals = torch.empty(max(n_bases, 1), dtype=torch.int32, device=dev)
    n_valid = ctypes.c_uint64(0)
    torch.cuda.synchronize(dev)
    nat.check(L.pa_records_encode_device(ctypes.c_void_p(d_bases.data_ptr()), nat._p(genome_off), G, g_lo, g_hi, int(k), device,
                                         ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                                         ctypes.byref(n_valid), None))
    if fused is None:
        fused = os.environ.get("PA_FUSED_EXCHANGE", "1") != "0"
    if fused and int(k) >= 1:
        torch.cuda.synchronize(dev)
        t["encode_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = _exchange_fused(L, nat, keys, vals, n_bases, k, device, world, rank, group)
        if res[0] is None:
            t["fused_exchange_unavailable"] = res[1]      # every rank got the same verdict: use the all-to-all path
            fused = False
    if fused and int(k) >= 1:
        rk, rv, n_recv, n_sent, t["scatter_exchange_phases"] = res
        assert n_sent == n_valid.value
        del keys, vals
        t["scatter_exchange_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        h = ctypes.c_void_p()
        try:
            nat.check(L.pa_index_build_from_records_device(rk, rv, n_recv, nat._p(genome_off), G, int(k), device, 0,
                                                           ctypes.byref(h)))
        finally:
            pass   # the receive buffers stay mapped for the next build (release_peer_buffers() frees them)
        partition = nat.NativeIndex(h.value)
        t["sort_rle_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        replica = _gather_replica(partition, k, genome_off, dev, group)
        t["replicate_s"] = time.perf_counter() - t0
        out = DistributedIndex(partition, replica, int(k), genome_off, rank, world, group)
        out.timings = t
        out.sent_records, out.received_records, out.fused = n_sent, n_recv, True
        return out
    keys_t = torch.empty_like(keys)
    vals_t = torch.empty_like(vals)
    part_off = np.zeros(world + 1, dtype=np.uint64)
    in_tmp = ctypes.c_int32(0)
    nat.check(L.pa_records_partition_device(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                                            ctypes.c_void_p(keys_t.data_ptr()), ctypes.c_void_p(vals_t.data_ptr()),
                                            n_bases if int(k) >= 1 else 0, int(k), world, device, nat._p(part_off),
                                            ctypes.byref(in_tmp), None))
    if in_tmp.value:
        keys, keys_t, vals, vals_t = keys_t, keys, vals_t, vals
    del keys_t, vals_t
    assert int(part_off[-1]) == n_valid.value
    torch.cuda.synchronize(dev)
    t["encode_partition_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    send = np.diff(part_off.astype(np.int64))
    send_t = torch.from_numpy(send.copy())
    recv_t = torch.empty(world, dtype=torch.int64)
    if _is_nccl(group):
        s_d, r_d = send_t.to(dev), recv_t.to(dev)
        dist.all_to_all_single(r_d, s_d, group=group)
        recv_t = r_d.cpu()
    else:
        _all_to_all(recv_t, send_t, [1] * world, [1] * world, group)
    recv = [int(x) for x in recv_t.tolist()]
    n_recv = sum(recv)
    r_keys = torch.empty(max(n_recv, 1), dtype=torch.int64, device=dev)
    r_vals = torch.empty(max(n_recv, 1), dtype=torch.int32, device=dev)
    send_l = [int(x) for x in send.tolist()]
    _all_to_all(r_keys[:n_recv], keys[:int(part_off[-1])], recv, send_l, group)
    _all_to_all(r_vals[:n_recv], vals[:int(part_off[-1])], recv, send_l, group)
    del keys, vals
    torch.cuda.synchronize(dev)
    t["exchange_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = ctypes.c_void_p()
    nat.check(L.pa_index_build_from_records_device(ctypes.c_void_p(r_keys.data_ptr()), ctypes.c_void_p(r_vals.data_ptr()),
                                                   n_recv, nat._p(genome_off), G, int(k), device, 0, ctypes.byref(h)))
    partition = nat.NativeIndex(h.value)
    del r_keys, r_vals
    t["sort_rle_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    replica = _gather_replica(partition, k, genome_off, dev, group)
    t["replicate_s"] = time.perf_counter() - t0
    out = DistributedIndex(partition, replica, int(k), genome_off, rank, world, group)
    out.timings = t
    out.sent_records = int(part_off[-1])
    out.received_records = n_recv
    out.fused = False
    return out


def extsim_stats_allreduce(dix: DistributedIndex, group_ids: np.ndarray, n_groups: int):
    """_compute_genome_stats (kmer.py:152-177) over all partitions: per-class sums are additive over key ranges."""
    import torch
    import torch.distributed as dist
    total, uniq = dix.partition.extsim_stats(group_ids, n_groups)
    t = torch.from_numpy(np.concatenate([total, uniq]).astype(np.int64))
    if _is_nccl(dix.group):
        d = t.cuda()
        dist.all_reduce(d, group=dix.group)
        t = d.cpu()
    else:
        dist.all_reduce(t, group=dix.group)
    a = t.numpy().astype(np.uint64)
    return a[:n_groups], a[n_groups:]


def extsim_pairwise_allreduce(dix: DistributedIndex, group_ids: np.ndarray, n_groups: int) -> np.ndarray:
    """The |A & B| matrix of _apply_greedy_filter (kmer.py:206-207) summed over all partitions."""
    import torch
    import torch.distributed as dist
    inter = dix.partition.extsim_pairwise(group_ids, n_groups)
    t = torch.from_numpy(inter.astype(np.int64).reshape(-1).copy())
    if _is_nccl(dix.group):
        d = t.cuda()
        dist.all_reduce(d, group=dix.group)
        t = d.cpu()
    else:
        dist.all_reduce(t, group=dix.group)
    return t.numpy().astype(np.uint64).reshape(n_groups, n_groups)


def drop_genomes(dix: DistributedIndex, keep: np.ndarray) -> None:
    """_remove_filtered_genomes_from_kmers (kmer.py:232-250) on every partition, then the replica is re-gathered."""
    import torch
    dix.partition.drop_genomes(keep)
    keep = np.asarray(keep).astype(bool)
    lens = np.diff(dix.genome_off.astype(np.int64))[keep]
    dix.genome_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    dix.replica.close()
    dev = torch.device("cuda", dix.partition.info().device)
    dix.replica = _gather_replica(dix.partition, dix.k, dix.genome_off, dev, dix.group)


def export_gathered(dix: DistributedIndex, dst: int = 0):
    """CSR of the whole index on rank `dst` (dumpref / pickling): the partitions' exports concatenated in rank order
    (= key order); `order` sorts all keys by first occurrence, i.e. the reference's dict insertion order."""
    import torch.distributed as dist
    ex = dix.partition.export(with_positions=True, with_order=False)
    parts = [None] * dix.world
    dist.all_gather_object(parts, ex, group=dix.group)
    if dix.rank != dst:
        return None
    out = {"keys": np.concatenate([p["keys"] for p in parts]), "run_genome": np.concatenate([p["run_genome"] for p in parts]),
           "pos": np.concatenate([p["pos"] for p in parts]), "first_occ": np.concatenate([p["first_occ"] for p in parts])}
    run_off, pos_off, rb, pb = [], [], 0, 0
    for p in parts:
        run_off.append(p["run_off"][:-1].astype(np.uint64) + np.uint64(rb))
        pos_off.append(p["pos_off"][:-1].astype(np.uint64) + np.uint64(pb))
        rb += int(p["run_off"][-1])
        pb += int(p["pos_off"][-1])
    out["run_off"] = np.concatenate(run_off + [np.array([rb], dtype=np.uint64)])
    out["pos_off"] = np.concatenate(pos_off + [np.array([pb], dtype=np.uint64)])
    out["order"] = np.argsort(out["first_occ"], kind="stable").astype(np.uint32)
    return out


# ---------------------------------------------------------------------------
# read-sharded alignment, end to end
# ---------------------------------------------------------------------------
def align_sharded(index, bases: np.ndarray, quals: Optional[np.ndarray], read_off: np.ndarray, genome_ids: List[str],
                  m: int = 1, p: int = 1, min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                  max_genomes: Optional[int] = None, group=None, gather_reads: bool = False):
    """PseudoAlignment.align_reads_from_container + get_summary (kmer.py:600-657) across ranks.

    Every rank holds the whole packed batch (bases / quals / read_off as produced by the native ingest) and a replicated
    align index (`NativeIndex`, e.g. DistributedIndex.replica); rank r aligns the contiguous block shard_bounds(n, world, r),
    the K8 accumulators are all-reduced, and every rank returns the reference's summary dict (key order included).
    gather_reads=True additionally returns, on rank 0, the per-read (type, genome index list) of all reads in file order.
    """
    import torch
    import torch.distributed as dist
    import _native as nat
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    n = len(read_off) - 1
    lo, hi = shard_bounds(n, world, rank)
    params = nat.make_params(m, p, min_read_quality, min_kmer_quality, max_genomes)
    words, lst, counters = index.align(bases, quals, read_off[lo:hi + 1], params)
    stats, uniq, amb, first = index.summary(words, lst, read_index_base=lo)
    G = len(genome_ids)
    acc = torch.from_numpy(np.concatenate([stats, uniq, amb, counters]).astype(np.int64))
    fs = torch.from_numpy(first.astype(np.uint64).view(np.int64).copy())
    if _is_nccl(group):
        dev = torch.device("cuda", index.info().device)
        acc_d, fs_d = acc.to(dev), fs.to(dev)
        allreduce_summary(acc_d, fs_d, group)
        acc, fs = acc_d.cpu(), fs_d.cpu()
    else:
        allreduce_summary(acc, fs, group)
    a = acc.numpy()
    flags = (min_read_quality is not None, min_kmer_quality is not None, max_genomes is not None)
    summary = summary_from_accumulators(a[:4], a[4:4 + G], a[4 + G:4 + 2 * G], fs.numpy().view(np.uint64), genome_ids, flags,
                                        a[4 + 2 * G:4 + 2 * G + 3])
    if not gather_reads:
        return summary
    types, lens, payload = nat.decode_words(words)
    lists = [[int(payload[i])] if lens[i] == 1 else [int(x) for x in lst[int(payload[i]):int(payload[i]) + int(lens[i])]]
             for i in range(len(words))]
    parts = [None] * world
    dist.all_gather_object(parts, (types.tolist(), lists), group=group)
    if rank != 0:
        return summary, None
    all_types = [t for part in parts for t in part[0]]
    all_lists = [l for part in parts for l in part[1]]
    return summary, (all_types, all_lists)

