"""
_native.py -- ctypes binding of libpa_b200.so (the C ABI in include/pa_b200.h).

There is no CPU fallback: if the CUDA library is missing or no CUDA device is
usable, every entry point raises.  Build the library with
`python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PA_B200_LIB") or os.path.join(_HERE, "libpa_b200.so")  # PA_B200_LIB: tuning builds

PA_OK = 0
PA_ERR_INVALID_ARG = -1
PA_ERR_BAD_BASE = -2
PA_ERR_CUDA = -3
PA_ERR_NOMEM = -4
PA_ERR_CAPACITY = -5
PA_ERR_UNSUPPORTED = -6
RANK_MISS = 0xFFFFFFFFFFFFFFFF

EXPORTED_SYMBOLS = [
    "pa_abi_version", "pa_last_error", "pa_device_count", "pa_trim_memory", "pa_index_build", "pa_index_build_device", "pa_index_import",
    "pa_index_free", "pa_index_info_get", "pa_index_export", "pa_decode_kmers", "pa_encode_kmers", "pa_index_lookup", "pa_index_entries", "pa_index_checksum", "pa_index_csr_device",
    "pa_extsim_stats", "pa_extsim_pairwise", "pa_index_drop_genomes", "pa_align_batch", "pa_align_batch_device", "pa_pack_reads", "pa_align_batch_packed",
    "pa_summary_reduce_device", "pa_summary_reduce", "pa_debug_sort_pairs", "pa_debug_sort_pairs_hashed", "pa_debug_table_lookup",
    "pa_comm_unique_id", "pa_comm_init", "pa_comm_init_callbacks", "pa_comm_free", "pa_comm_info", "pa_comm_allreduce_summary",
    "pa_comm_allreduce_host", "pa_comm_allgather_host", "pa_comm_barrier", "pa_index_build_partitioned", "pa_index_rebuild_replica", "pa_genome_shard",
    "pa_build_exchange", "pa_build_timings", "pa_partition_of_kmer",
    "pa_debug_pack_reads", "pa_debug_minimizer",
    "pa_parse_records", "pa_parsed_copy", "pa_parsed_free",
    "pa_format_kmers_json", "pa_free_text",
]


class NativeLibraryMissing(ImportError):
    pass


class IndexInfo(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_int32), ("device", ctypes.c_int32), ("n_genomes", ctypes.c_uint32),
        ("blocks_per_digit", ctypes.c_uint32), ("tag_bits", ctypes.c_uint32), ("stash_count", ctypes.c_uint32),
        ("n_keys", ctypes.c_uint64), ("n_runs", ctypes.c_uint64), ("n_occ", ctypes.c_uint64),
        ("total_bases", ctypes.c_uint64), ("n_list_sectors", ctypes.c_uint64), ("device_bytes", ctypes.c_uint64),
        ("build_encode_ms", ctypes.c_float), ("build_sort_ms", ctypes.c_float), ("build_rle_ms", ctypes.c_float),
        ("build_table_ms", ctypes.c_float), ("minimizer_len", ctypes.c_uint32), ("digit_bits", ctypes.c_uint32),
        ("n_blocks", ctypes.c_uint64), ("table_bytes", ctypes.c_uint64), ("align_only", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
    ]


class AlignParams(ctypes.Structure):
    _fields_ = [
        ("m", ctypes.c_int64), ("p", ctypes.c_int64), ("min_read_quality", ctypes.c_int64),
        ("min_kmer_quality", ctypes.c_int64), ("max_genomes", ctypes.c_int64),
        ("has_min_read_quality", ctypes.c_int32), ("has_min_kmer_quality", ctypes.c_int32),
        ("has_max_genomes", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


def _clamp64(v: int) -> int:
    return max(-(1 << 62), min(1 << 62, int(v)))


def make_params(m: int, p: int, mrq: Optional[int], mkq: Optional[int], mg: Optional[int]) -> AlignParams:
    return AlignParams(_clamp64(m), _clamp64(p), _clamp64(mrq or 0), _clamp64(mkq or 0), _clamp64(mg or 0),
                       int(mrq is not None), int(mkq is not None), int(mg is not None), 0)


_lib = None


def lib() -> ctypes.CDLL:
    """Loads libpa_b200.so; raises NativeLibraryMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run __graft_entry__.build()). "
            "This package has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32
    sig = {
        "pa_abi_version": (i32, []),
        "pa_last_error": (i32, [ctypes.c_char_p, ctypes.c_size_t]),
        "pa_device_count": (i32, [vp]),
        "pa_trim_memory": (i32, []),
        "pa_index_build": (i32, [vp, vp, u32, i32, i32, vp]),
        "pa_index_build_device": (i32, [vp, vp, u32, i32, i32, vp]),
        "pa_index_import": (i32, [i32, u32, vp, u64, u64, u64, vp, vp, vp, vp, vp, vp, i32, vp]),
        "pa_index_free": (i32, [vp]),
        "pa_index_info_get": (i32, [vp, vp]),
        "pa_index_export": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "pa_decode_kmers": (i32, [i32, vp, u64, vp]),
        "pa_encode_kmers": (i32, [i32, vp, u64, vp]),
        "pa_index_lookup": (i32, [vp, vp, u64, vp]),
        "pa_index_entries": (i32, [vp, vp, u64, vp, vp, u64, vp, vp, u64, vp, vp]),
        "pa_index_checksum": (i32, [vp, vp]),
        "pa_index_csr_device": (i32, [vp, vp, vp, vp]),
        "pa_extsim_stats": (i32, [vp, vp, u32, vp, vp]),
        "pa_extsim_pairwise": (i32, [vp, vp, u32, vp]),
        "pa_index_drop_genomes": (i32, [vp, vp]),
        "pa_align_batch": (i32, [vp, vp, vp, vp, u64, vp, vp, vp, u64, vp, vp]),
        "pa_align_batch_device": (i32, [vp, vp, vp, vp, u64, u64, vp, vp, vp, u64, vp, vp, vp]),
        "pa_pack_reads": (i32, [vp, vp, u64, vp, u64, vp]),
        "pa_align_batch_packed": (i32, [vp, vp, vp, vp, u64, vp, vp, vp, u64, vp, vp]),
        "pa_summary_reduce_device": (i32, [vp, vp, u64, u64, u32, vp, vp, vp, vp, vp]),
        "pa_summary_reduce": (i32, [vp, vp, vp, u64, u64, u64, vp, vp, vp, vp]),
        "pa_debug_sort_pairs": (i32, [vp, vp, u64, i32, i32]),
        "pa_debug_sort_pairs_hashed": (i32, [vp, vp, u64, i32, i32, i32, vp]),
        "pa_debug_table_lookup": (i32, [vp, vp, u64, vp, vp]),
        "pa_comm_unique_id": (i32, [vp]),
        "pa_comm_init": (i32, [i32, i32, vp, i32, vp]),
        "pa_comm_init_callbacks": (i32, [i32, i32, i32, vp, vp]),
        "pa_comm_free": (i32, [vp]),
        "pa_comm_info": (i32, [vp, vp, vp, vp, vp]),
        "pa_comm_allreduce_summary": (i32, [vp, vp, u64, vp, u64, vp]),
        "pa_comm_allreduce_host": (i32, [vp, vp, u64, i32]),
        "pa_comm_allgather_host": (i32, [vp, vp, vp, u64]),
        "pa_comm_barrier": (i32, [vp]),
        "pa_index_build_partitioned": (i32, [vp, vp, vp, u32, u32, u32, i32, i32, u32, u32, vp, vp]),
        "pa_index_rebuild_replica": (i32, [vp, vp, vp]),
        "pa_genome_shard": (i32, [vp, u32, i32, i32, vp, vp]),
        "pa_build_exchange": (i32, [vp, vp, vp, vp, u64, vp, vp, vp, vp]),
        "pa_build_timings": (i32, [vp, vp]),
        "pa_partition_of_kmer": (i32, [i32, vp, u32, vp]),
        "pa_debug_pack_reads": (i32, [vp, vp, u64, vp, u64, i32, vp]),
        "pa_debug_minimizer": (i32, [i32, vp, u64, vp, vp]),
        "pa_parse_records": (i32, [vp, u64, i32, vp, vp, vp, vp]),
        "pa_parsed_copy": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "pa_parsed_free": (i32, [vp]),
        "pa_format_kmers_json": (i32, [i32, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]),
        "pa_free_text": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    lib().pa_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(status: int) -> None:
    """Maps the C status codes onto the exception types the reference raises."""
    if status == PA_OK:
        return
    msg = last_error()
    if status in (PA_ERR_INVALID_ARG, PA_ERR_BAD_BASE, PA_ERR_UNSUPPORTED):
        raise ValueError(msg)
    if status == PA_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"pa_b200 error {status}: {msg}")


def device_count() -> int:
    n = ctypes.c_int32(0)
    st = lib().pa_device_count(ctypes.byref(n))
    return n.value if st == PA_OK else 0


def trim_memory() -> None:
    """Gives the device buffers the library keeps for reuse back to the driver (pa_trim_memory)."""
    lib().pa_trim_memory()


def require_device() -> None:
    if device_count() < 1:
        raise RuntimeError("pa_b200: no CUDA device available; this package has no CPU fallback (" + last_error() + ")")


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def pack_strings(strings: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenates str objects into (uint8 bytes, uint64 offsets).  latin-1: one byte per character."""
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    if len(strings):
        off[1:] = np.cumsum(np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings)))
    data = np.frombuffer("".join(strings).encode("latin-1"), dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(data), off


def decode_words(words: np.ndarray):
    """Result words -> (type uint8, list length, payload)."""
    types = (words >> np.uint64(62)).astype(np.uint8)
    lens = ((words >> np.uint64(40)) & np.uint64(0x3FFFFF)).astype(np.int64)
    payload = (words & np.uint64(0xFFFFFFFFFF)).astype(np.int64)
    return types, lens, payload


def flatten_results(words: np.ndarray, lst: np.ndarray):
    """Result words (+ the side list) of pa_align_batch[_device] -> (types uint8[n], lens int32[n], flat genome indices
    uint32[sum lens]) in read order: independent of where the list cursor put a list."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    types, lens, payload = decode_words(words)
    off = np.zeros(len(words) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.zeros(int(off[-1]), dtype=np.uint32)
    single = lens == 1
    flat[off[:-1][single]] = payload[single]
    for i in np.nonzero(lens > 1)[0]:
        flat[off[i]:off[i + 1]] = lst[payload[i]:payload[i] + lens[i]]
    return types, lens.astype(np.int32), flat


def canonical_words(types: np.ndarray, lens: np.ndarray, flat: np.ndarray):
    """The inverse: result words whose lists sit in `flat` in read order (what K8, pa_summary_reduce, reads)."""
    lens64 = np.asarray(lens, dtype=np.int64)
    off = np.zeros(len(lens64) + 1, dtype=np.int64)
    np.cumsum(lens64, out=off[1:])
    payload = off[:-1].astype(np.uint64)
    payload[lens64 == 0] = 0
    single = lens64 == 1
    if single.any():
        payload[single] = np.asarray(flat, dtype=np.uint64)[off[:-1][single]]
    words = (np.asarray(types, dtype=np.uint64) << np.uint64(62)) | (lens64.astype(np.uint64) << np.uint64(40)) | payload
    return words, np.ascontiguousarray(flat, dtype=np.uint32)


class NativeIndex:
    """Owns one pa_index handle."""

    def __init__(self, handle: int):
        self._h = ctypes.c_void_p(handle)

    # -- construction --------------------------------------------------------
    @classmethod
    def build(cls, bases: np.ndarray, genome_off: np.ndarray, k: int, device: int = 0) -> "NativeIndex":
        require_device()
        bases = _u8(bases)
        genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
        h = ctypes.c_void_p()
        check(lib().pa_index_build(_p(bases), _p(genome_off), len(genome_off) - 1, int(k), device, ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def build_device(cls, d_bases_ptr: int, genome_off: np.ndarray, k: int, device: int = 0) -> "NativeIndex":
        require_device()
        genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
        h = ctypes.c_void_p()
        check(lib().pa_index_build_device(ctypes.c_void_p(d_bases_ptr), _p(genome_off), len(genome_off) - 1, int(k),
                                          device, ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def import_csr(cls, k: int, genome_off: np.ndarray, keys, run_off, run_genome, pos_off, pos, first_occ=None,
                   device: int = 0) -> "NativeIndex":
        require_device()
        genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        run_off = np.ascontiguousarray(run_off, dtype=np.uint64)
        run_genome = np.ascontiguousarray(run_genome, dtype=np.uint32)
        pos_off = np.ascontiguousarray(pos_off, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        if first_occ is not None:
            first_occ = np.ascontiguousarray(first_occ, dtype=np.uint64)
        h = ctypes.c_void_p()
        check(lib().pa_index_import(int(k), len(genome_off) - 1, _p(genome_off), len(keys), len(run_genome), len(pos),
                                    _p(keys), _p(run_off), _p(run_genome), _p(pos_off), _p(pos), _p(first_occ), device,
                                    ctypes.byref(h)))
        return cls(h.value)

    def close(self) -> None:
        if self._h is not None and self._h.value and _lib is not None:
            _lib.pa_index_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> ctypes.c_void_p:
        if self._h is None:
            raise RuntimeError("index handle already freed")
        return self._h

    # -- queries ---------------------------------------------------------------
    def info(self) -> IndexInfo:
        inf = IndexInfo()
        check(lib().pa_index_info_get(self.handle, ctypes.byref(inf)))
        return inf

    def export(self, with_positions: bool = True, with_order: bool = True):
        inf = self.info()
        keys = np.zeros(max(inf.n_keys, 1), dtype=np.uint64)
        run_off = np.zeros(inf.n_keys + 1, dtype=np.uint64)
        run_genome = np.zeros(max(inf.n_runs, 1), dtype=np.uint32)
        pos_off = np.zeros(inf.n_runs + 1, dtype=np.uint64)
        pos = np.zeros(max(inf.n_occ, 1), dtype=np.uint32) if with_positions else None
        order = np.zeros(max(inf.n_keys, 1), dtype=np.uint32) if with_order else None
        first = np.zeros(max(inf.n_keys, 1), dtype=np.uint64)
        check(lib().pa_index_export(self.handle, _p(keys), _p(run_off), _p(run_genome), _p(pos_off), _p(pos), _p(order),
                                    _p(first)))
        return {"keys": keys[:inf.n_keys], "run_off": run_off, "run_genome": run_genome[:inf.n_runs], "pos_off": pos_off,
                "pos": None if pos is None else pos[:inf.n_occ], "order": None if order is None else order[:inf.n_keys],
                "first_occ": first[:inf.n_keys]}

    def lookup(self, kmers: Sequence[str]) -> np.ndarray:
        k = self.info().k
        n = len(kmers)
        rank = np.full(max(n, 1), RANK_MISS, dtype=np.uint64)
        good = [i for i, s in enumerate(kmers) if len(s) == k and k >= 1]
        if good:
            try:
                flat = np.frombuffer("".join(kmers[i] for i in good).encode("latin-1"), dtype=np.uint8)
            except UnicodeEncodeError:
                good = [i for i in good if all(ord(c) < 256 for c in kmers[i])]
                flat = np.frombuffer("".join(kmers[i] for i in good).encode("latin-1"), dtype=np.uint8)
            if good:
                sub = np.zeros(len(good), dtype=np.uint64)
                flat = np.ascontiguousarray(flat)
                check(lib().pa_index_lookup(self.handle, _p(flat), len(good), _p(sub)))
                rank[np.asarray(good, dtype=np.int64)] = sub
        return rank[:n]

    def checksum(self) -> np.ndarray:
        """Order-independent content checksum of the CSR (additive over the partitions of a multi-GPU build)."""
        out = np.zeros(4, dtype=np.uint64)
        check(lib().pa_index_checksum(self.handle, _p(out)))
        return out

    def entries(self, ranks: Sequence[int]):
        """CSR entries of the k-mers at `ranks` (no misses): (run_off[n+1], run_genome, pos_off, pos) -- only these k-mers
        leave the device."""
        ranks = np.ascontiguousarray(ranks, dtype=np.uint64)
        n = len(ranks)
        run_off = np.zeros(n + 1, dtype=np.uint64)
        rt, pt = ctypes.c_uint64(0), ctypes.c_uint64(0)
        run_cap, pos_cap = max(4 * n, 16), max(16 * n, 64)
        while True:
            run_genome = np.zeros(run_cap, dtype=np.uint32)
            pos_off = np.zeros(run_cap + 1, dtype=np.uint64)
            pos = np.zeros(pos_cap, dtype=np.uint32)
            st = lib().pa_index_entries(self.handle, _p(ranks if n else np.zeros(1, np.uint64)), n, _p(run_off), _p(run_genome),
                                        run_cap, _p(pos_off), _p(pos), pos_cap, ctypes.byref(rt), ctypes.byref(pt))
            if st == PA_ERR_CAPACITY:
                run_cap, pos_cap = max(run_cap, int(rt.value)), max(pos_cap, int(pt.value))
                continue
            check(st)
            return run_off, run_genome[:rt.value], pos_off[:rt.value + 1], pos[:pt.value]

    def table_lookup(self, kmers: Sequence[str]):
        k = self.info().k
        n = len(kmers)
        assert all(len(s) == k for s in kmers)
        flat = np.ascontiguousarray(np.frombuffer("".join(kmers).encode("latin-1"), dtype=np.uint8)) if n else np.zeros(1, np.uint8)
        ng = np.zeros(max(n, 1), dtype=np.uint32)
        g0 = np.zeros(max(n, 1), dtype=np.uint32)
        check(lib().pa_debug_table_lookup(self.handle, _p(flat), n, _p(ng), _p(g0)))
        return ng[:n], g0[:n]

    # -- EXTSIM ------------------------------------------------------------------
    def extsim_stats(self, group: np.ndarray, n_groups: int):
        group = np.ascontiguousarray(group, dtype=np.uint32)
        total = np.zeros(max(n_groups, 1), dtype=np.uint64)
        uniq = np.zeros(max(n_groups, 1), dtype=np.uint64)
        check(lib().pa_extsim_stats(self.handle, _p(group), n_groups, _p(total), _p(uniq)))
        return total[:n_groups], uniq[:n_groups]

    def extsim_pairwise(self, group: np.ndarray, n_groups: int) -> np.ndarray:
        group = np.ascontiguousarray(group, dtype=np.uint32)
        inter = np.zeros(max(n_groups * n_groups, 1), dtype=np.uint64)
        check(lib().pa_extsim_pairwise(self.handle, _p(group), n_groups, _p(inter)))
        return inter[:n_groups * n_groups].reshape(n_groups, n_groups)

    def drop_genomes(self, keep: np.ndarray) -> None:
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        if keep.size == 0:
            keep = np.zeros(1, dtype=np.uint8)
        check(lib().pa_index_drop_genomes(self.handle, _p(keep)))

    # -- alignment ----------------------------------------------------------------
    def align(self, bases: np.ndarray, quals: Optional[np.ndarray], read_off: np.ndarray, params: AlignParams):
        """Host-buffer batch alignment.  Returns (words, list, counters[3])."""
        bases = _u8(bases)
        quals = None if quals is None else _u8(quals)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        words = np.zeros(max(n, 1), dtype=np.uint64)
        counters = np.zeros(3, dtype=np.uint64)
        cap = max(n // 4, 1024)
        while True:
            lst = np.zeros(cap, dtype=np.uint32)
            need = ctypes.c_uint64(0)
            st = lib().pa_align_batch(self.handle, _p(bases), _p(quals), _p(read_off), n, ctypes.byref(params), _p(words),
                                      _p(lst), cap, ctypes.byref(need), _p(counters))
            if st == PA_ERR_CAPACITY:
                cap = int(need.value) + 16
                counters[:] = 0
                continue
            check(st)
            return words[:n], lst[:need.value], counters

    def align_packed(self, planes: np.ndarray, quals: Optional[np.ndarray], read_off: np.ndarray, params: AlignParams):
        """pa_align_batch_packed: reads already packed with pack_reads()."""
        planes = np.ascontiguousarray(planes, dtype=np.uint32)
        quals = None if quals is None else _u8(quals)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        words = np.zeros(max(n, 1), dtype=np.uint64)
        counters = np.zeros(3, dtype=np.uint64)
        cap = max(n // 4, 1024)
        while True:
            lst = np.zeros(cap, dtype=np.uint32)
            need = ctypes.c_uint64(0)
            st = lib().pa_align_batch_packed(self.handle, _p(planes), _p(quals), _p(read_off), n, ctypes.byref(params), _p(words),
                                             _p(lst), cap, ctypes.byref(need), _p(counters))
            if st == PA_ERR_CAPACITY:
                cap = int(need.value) + 16
                counters[:] = 0
                continue
            check(st)
            return words[:n], lst[:need.value], counters

    def summary(self, words: np.ndarray, lst: np.ndarray, read_index_base: int = 0):
        """K8 on host buffers.  Returns (stats[4], unique_reads[G], ambiguous_reads[G], first_seen[G])."""
        G = self.info().n_genomes
        words = np.ascontiguousarray(words, dtype=np.uint64)
        lst = np.ascontiguousarray(lst, dtype=np.uint32)
        stats = np.zeros(4, dtype=np.uint64)
        uniq = np.zeros(max(G, 1), dtype=np.uint64)
        amb = np.zeros(max(G, 1), dtype=np.uint64)
        first = np.zeros(max(G, 1), dtype=np.uint64)
        check(lib().pa_summary_reduce(self.handle, _p(words if len(words) else np.zeros(1, np.uint64)),
                                      _p(lst if len(lst) else np.zeros(1, np.uint32)), len(words), len(lst),
                                      int(read_index_base), _p(stats), _p(uniq), _p(amb), _p(first)))
        return stats, uniq[:G], amb[:G], first[:G]


PA_BUILD_TABLE_ONLY = 1
PA_BUILD_HOST_BASES = 2
_ALLGATHER_CB = ctypes.CFUNCTYPE(ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64)


class _CommCallbacks(ctypes.Structure):
    _fields_ = [("user", ctypes.c_void_p), ("allgather", _ALLGATHER_CB)]


class Comm:
    """Owns one pa_comm handle: one process per GPU on one node (include/pa_b200.h "communicator").

    Comm.nccl(...)        NCCL transport (the product path): rank 0 calls Comm.unique_id() and hands the 128 bytes to all
    Comm.callbacks(...)   any host all-gather as the control plane (tests over gloo); device data travels through CUDA IPC
    Comm.from_torch(...)  either of the two from an initialised torch.distributed process group
    """

    def __init__(self, handle: int, keep=None):
        self._h = ctypes.c_void_p(handle)
        self._keep = keep          # callback objects must outlive the handle

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_uint8 * 128)()
        check(lib().pa_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def nccl(cls, n_ranks: int, rank: int, unique_id: bytes, device: int) -> "Comm":
        require_device()
        h = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().pa_comm_init(n_ranks, rank, buf, device, ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def callbacks(cls, n_ranks: int, rank: int, device: int, allgather) -> "Comm":
        """allgather(data: bytes) -> list of every rank's bytes (equal lengths), in rank order.
        device < 0: host-only communicator (host collectives only; no CUDA device needed)."""
        if device >= 0:
            require_device()

        def _cb(_user, p_in, p_out, nbytes):
            try:
                parts = allgather(ctypes.string_at(p_in, nbytes))
                blob = b"".join(parts)
                if len(blob) != nbytes * n_ranks:
                    return 2
                ctypes.memmove(p_out, blob, len(blob))
                return 0
            except Exception:      # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        fn = _ALLGATHER_CB(_cb)
        cbs = _CommCallbacks(None, fn)
        h = ctypes.c_void_p()
        check(lib().pa_comm_init_callbacks(n_ranks, rank, device, ctypes.byref(cbs), ctypes.byref(h)))
        return cls(h.value, keep=(fn, cbs))

    @classmethod
    def from_torch(cls, device: int, group=None) -> "Comm":
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if dist.get_backend(group) == "nccl":
            box = [cls.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group,
                                       device=torch.device("cuda", device))
            return cls.nccl(world, rank, box[0], device)

        def allgather(data: bytes):
            mine = torch.frombuffer(bytearray(data), dtype=torch.uint8)
            out = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(out, mine, group=group)
            return [bytes(t.numpy().tobytes()) for t in out]

        return cls.callbacks(world, rank, device, allgather)

    @property
    def handle(self) -> ctypes.c_void_p:
        if self._h is None:
            raise RuntimeError("communicator already freed")
        return self._h

    def close(self) -> None:
        if self._h is not None and self._h.value and _lib is not None:
            _lib.pa_comm_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n, r, d, nc = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        check(lib().pa_comm_info(self.handle, ctypes.byref(n), ctypes.byref(r), ctypes.byref(d), ctypes.byref(nc)))
        return {"n_ranks": n.value, "rank": r.value, "device": d.value, "nccl": bool(nc.value)}

    def barrier(self) -> None:
        check(lib().pa_comm_barrier(self.handle))

    def allgather_bytes(self, data: bytes) -> List[bytes]:
        """Every rank's `data` (any lengths), in rank order."""
        n = self.info()["n_ranks"]
        sizes = np.zeros(n, dtype=np.uint64)
        mine = np.array([len(data)], dtype=np.uint64)
        check(lib().pa_comm_allgather_host(self.handle, _p(mine), _p(sizes), 8))
        mx = int(sizes.max()) if n else 0
        if mx == 0:
            return [b""] * n
        pad = np.zeros(mx, dtype=np.uint8)
        pad[:len(data)] = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(mx * n, dtype=np.uint8)
        check(lib().pa_comm_allgather_host(self.handle, _p(pad), _p(out), mx))
        return [out[r * mx:r * mx + int(sizes[r])].tobytes() for r in range(n)]

    def allgather_array(self, a: np.ndarray) -> List[np.ndarray]:
        a = np.ascontiguousarray(a)
        return [np.frombuffer(b, dtype=a.dtype) for b in self.allgather_bytes(a.tobytes())]

    def allreduce_host(self, values: np.ndarray, take_min: bool = False) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64).copy()
        check(lib().pa_comm_allreduce_host(self.handle, _p(v), v.size, 1 if take_min else 0))
        return v

    def allreduce_summary(self, d_sum_ptr: int, n_sum: int, d_min_ptr: int, n_min: int, stream_ptr: int = 0) -> None:
        check(lib().pa_comm_allreduce_summary(self.handle, ctypes.c_void_p(d_sum_ptr), n_sum, ctypes.c_void_p(d_min_ptr), n_min,
                                              ctypes.c_void_p(stream_ptr) if stream_ptr else None))


def genome_shard(genome_off: np.ndarray, n_ranks: int, rank: int) -> Tuple[int, int]:
    genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
    lo, hi = ctypes.c_uint32(), ctypes.c_uint32()
    check(lib().pa_genome_shard(_p(genome_off), len(genome_off) - 1, n_ranks, rank, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def build_partitioned(comm: Optional[Comm], bases, genome_off: np.ndarray, g_range: Tuple[int, int], k: int, device: int = 0,
                      table_only: bool = False, n_rounds: int = 0):
    """pa_index_build_partitioned.  bases: the genomes [g_lo, g_hi) concatenated -- a uint8 numpy array (host) or an int
    (device pointer).  Returns (partition or None, replica) as NativeIndex objects."""
    require_device()
    genome_off = np.ascontiguousarray(genome_off, dtype=np.uint64)
    flags = PA_BUILD_TABLE_ONLY if table_only else 0
    keep = None
    if isinstance(bases, (int, np.integer)):
        ptr = ctypes.c_void_p(int(bases))
    else:
        keep = _u8(bases)
        if keep.size == 0:
            keep = np.zeros(1, dtype=np.uint8)
        ptr = _p(keep)
        flags |= PA_BUILD_HOST_BASES
    hp, hr = ctypes.c_void_p(), ctypes.c_void_p()
    check(lib().pa_index_build_partitioned(comm.handle if comm is not None else None, ptr, _p(genome_off), len(genome_off) - 1,
                                           int(g_range[0]), int(g_range[1]), int(k), device, flags, int(n_rounds),
                                           None if table_only else ctypes.byref(hp), ctypes.byref(hr)))
    return (NativeIndex(hp.value) if hp.value else None), NativeIndex(hr.value)


def rebuild_replica(comm: Optional[Comm], partition: "NativeIndex") -> "NativeIndex":
    hr = ctypes.c_void_p()
    check(lib().pa_index_rebuild_replica(comm.handle if comm is not None else None, partition.handle, ctypes.byref(hr)))
    return NativeIndex(hr.value)


def build_timings(replica: "NativeIndex") -> dict:
    ms = (ctypes.c_float * 8)()
    check(lib().pa_build_timings(replica.handle, ms))
    names = ["encode_count_ms", "scatter_exchange_ms", "sort_ms", "csr_ms", "table_slice_ms", "table_gather_ms", "total_ms"]
    return {n: float(ms[i]) for i, n in enumerate(names)}


def partition_of_kmer(k: int, kmer: str, n_parts: int) -> int:
    b = np.frombuffer(kmer.encode("latin-1"), dtype=np.uint8).copy()
    out = ctypes.c_uint32()
    check(lib().pa_partition_of_kmer(int(k), _p(b), int(n_parts), ctypes.byref(out)))
    return out.value


def pack_reads(bases: np.ndarray, read_off: np.ndarray):
    """pa_pack_reads: (planes uint32[...], all_acgt).  Pure host code."""
    bases = _u8(bases)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    n = len(read_off) - 1
    words = 2 * (int(read_off[-1] - read_off[0]) // 32 + n + 1)
    planes = np.zeros(words, dtype=np.uint32)
    ok = ctypes.c_int32(0)
    check(lib().pa_pack_reads(_p(bases), _p(read_off), n, _p(planes), words, ctypes.byref(ok)))
    return planes, bool(ok.value)


def decode_kmers(k: int, keys: np.ndarray) -> List[str]:
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    n = len(keys)
    if n == 0 or k <= 0:
        return ["" for _ in range(n)]
    out = np.zeros(n * k, dtype=np.uint8)
    check(lib().pa_decode_kmers(k, _p(keys), n, _p(out)))
    raw = out.tobytes().decode("ascii")
    return [raw[i * k:(i + 1) * k] for i in range(n)]


def debug_sort_pairs(keys: np.ndarray, vals: np.ndarray, end_bit: int = 64, device: int = 0):
    require_device()
    keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    check(lib().pa_debug_sort_pairs(_p(keys), _p(vals), len(keys), end_bit, device))
    return keys, vals


def debug_sort_pairs_hashed(keys: np.ndarray, vals: np.ndarray, end_bit: int = 64, top_bits: int = 0, device: int = 0):
    """Returns (keys, vals, fell_back): the build's top-bits sort + repair (sort.cu), see pa_debug_sort_pairs_hashed."""
    require_device()
    keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    fb = ctypes.c_int32(0)
    check(lib().pa_debug_sort_pairs_hashed(_p(keys), _p(vals), len(keys), end_bit, top_bits, device, ctypes.byref(fb)))
    return keys, vals, bool(fb.value)


def parse_records_native(raw, fastq: bool):
    """Canonical FASTA / FASTQ text (an ASCII str, parsed in place, or bytes) -> packed arrays, or None when the text needs the regex parser (ingest.cpp).

    Returns {"seq": uint8[n_bases], "qual": uint8[n_bases] | None, "off": uint64[n + 1], "name_beg", "name_len",
    "plus_beg", "plus_len": uint64[n] ranges inside `raw`, "raw": raw}."""
    n = len(raw)
    if isinstance(raw, str):
        # an ASCII str is stored one byte per character: parse its buffer in place instead of encoding a copy
        # (the caller checked isascii(); `raw` stays referenced by the result, which keeps the buffer alive)
        size = ctypes.c_ssize_t(0)
        as_utf8 = ctypes.pythonapi.PyUnicode_AsUTF8AndSize
        as_utf8.restype, as_utf8.argtypes = ctypes.c_void_p, [ctypes.py_object, ctypes.POINTER(ctypes.c_ssize_t)]
        ptr = as_utf8(raw, ctypes.byref(size))
        if not ptr or size.value != n:
            raise ValueError("parse_records_native: the text is not ASCII")
        text_ptr = ctypes.c_void_p(ptr)
    else:
        buf = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        text_ptr = _p(buf)
    h = ctypes.c_void_p()
    canonical = ctypes.c_int32(0)
    n_rec, n_bases = ctypes.c_uint64(0), ctypes.c_uint64(0)
    check(lib().pa_parse_records(text_ptr, n, int(bool(fastq)), ctypes.byref(h), ctypes.byref(canonical), ctypes.byref(n_rec),
                                 ctypes.byref(n_bases)))
    if not canonical.value:
        return None
    nr = n_rec.value
    try:
        seq = np.empty(max(n_bases.value, 1), dtype=np.uint8)     # pa_parsed_copy fills every element it reports
        qual = np.empty(max(n_bases.value, 1), dtype=np.uint8) if fastq else None
        off = np.zeros(nr + 1, dtype=np.uint64)
        nb, nl = np.empty(max(nr, 1), dtype=np.uint64), np.empty(max(nr, 1), dtype=np.uint64)
        pb, pl = (np.empty(max(nr, 1), dtype=np.uint64), np.empty(max(nr, 1), dtype=np.uint64)) if fastq else (None, None)
        check(lib().pa_parsed_copy(h, _p(seq), _p(qual), _p(off), _p(nb), _p(nl), _p(pb), _p(pl)))
    finally:
        lib().pa_parsed_free(h)
    return {"seq": seq[:n_bases.value], "qual": None if qual is None else qual[:n_bases.value], "off": off,
            "name_beg": nb[:nr], "name_len": nl[:nr], "plus_beg": None if pb is None else pb[:nr],
            "plus_len": None if pl is None else pl[:nr], "raw": raw, "n": nr}


def parsed_text(raw, beg: int, length: int) -> str:
    """A byte range of the parsed text (`raw` is the ASCII str that was parsed, or bytes)."""
    piece = raw[beg:beg + length]
    return piece if isinstance(piece, str) else piece.decode("ascii")


def parsed_names(packed) -> List[str]:
    raw = packed["raw"]
    if isinstance(raw, str):
        return [raw[b:b + l] for b, l in zip(packed["name_beg"].tolist(), packed["name_len"].tolist())]
    return [raw[b:b + l].decode("ascii") for b, l in zip(packed["name_beg"].tolist(), packed["name_len"].tolist())]


def format_kmers_json(k: int, csr, desc_class: np.ndarray, desc_json: Sequence[str], indent: int = 4, level: int = 1) -> str:
    """The "Kmers" object of get_summary() as JSON text, written natively from an exported CSR (format.cpp)."""
    keys = np.ascontiguousarray(csr["keys"], dtype=np.uint64)
    n = len(keys)
    blobs = [d.encode("ascii") for d in desc_json]       # json.dumps output is ASCII (ensure_ascii)
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    if blobs:
        off[1:] = np.cumsum([len(b) for b in blobs])
    flat = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
    out, out_len = ctypes.c_void_p(), ctypes.c_uint64(0)
    arr = lambda a, t: np.ascontiguousarray(a, dtype=t) if len(a) else np.zeros(1, dtype=t)
    check(lib().pa_format_kmers_json(int(max(k, 0)), n, _p(arr(keys, np.uint64)), _p(arr(csr["order"], np.uint32)),
                                     _p(arr(csr["run_off"], np.uint64)), _p(arr(csr["run_genome"], np.uint32)),
                                     _p(arr(csr["pos_off"], np.uint64)), _p(arr(csr["pos"], np.uint32)),
                                     _p(arr(desc_class, np.uint32)), _p(flat), _p(off), int(indent), int(level),
                                     ctypes.byref(out), ctypes.byref(out_len)))
    try:
        return ctypes.string_at(out, out_len.value).decode("ascii")
    finally:
        lib().pa_free_text(out)

