"""
oracle.py -- ctypes front end of the CPU ORACLE (oracle/pa_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package never does (it fails loudly when its CUDA extension is missing).

Parity status: PINNED (see the header of pa_oracle.c).  The classes here work
on plain data -- genome (identifier, sequence) pairs and read (identifier,
sequence, quality) triples -- and return plain data shaped like the reference's
observable results:

  OracleReference      ~ KmerReference            (/root/reference/src/kmer.py:109-351)
  OracleReference.align ~ PseudoAlignment.align_reads_from_container + get_summary
                                                   (kmer.py:600-657)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
_SRC_PATH = os.path.join(_HERE, "pa_oracle.c")

MAPPING_NAMES = {1: "UNMAPPED", 2: "UNIQUELY_MAPPED", 3: "AMBIGUOUSLY_MAPPED"}


def build_oracle(force: bool = False) -> str:
    """Compile pa_oracle.c into oracle/liborc.so (gcc only; no GPU involved)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC_PATH):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wno-comment",
                               "-o", _LIB_PATH, _SRC_PATH])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = ctypes.CDLL(_LIB_PATH)
        vp, u64, u32, i64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, ctypes.c_int32
        L.orc_index_build.restype = vp
        L.orc_index_build.argtypes = [vp, vp, u32, i32, vp]
        L.orc_index_free.restype = None
        L.orc_index_free.argtypes = [vp]
        L.orc_index_sizes.restype = None
        L.orc_index_sizes.argtypes = [vp, vp, vp, vp]
        L.orc_index_export.restype = None
        L.orc_index_export.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_index_lookup.restype = u32
        L.orc_index_lookup.argtypes = [vp, vp, u64, vp]
        L.orc_index_drop_genomes.restype = vp
        L.orc_index_drop_genomes.argtypes = [vp, vp]
        L.orc_extsim_stats.restype = None
        L.orc_extsim_stats.argtypes = [vp, vp, u32, vp, vp]
        L.orc_extsim_pairwise.restype = None
        L.orc_extsim_pairwise.argtypes = [vp, vp, u32, vp]
        L.orc_align.restype = u64
        L.orc_align.argtypes = [vp, vp, vp, vp, u64, i64, i64, i32, i64, i32, i64, i32, i64, vp, vp, vp, u64, vp, i32]
        L.orc_summary.restype = u32
        L.orc_summary.argtypes = [vp, vp, vp, u64, u32, vp, vp, vp, vp]
        L.orc_max_threads.restype = i32
        L.orc_max_threads.argtypes = []
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pack_strings(strings: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate strings into one uint8 array + uint64 offsets (latin-1 bytes)."""
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    if strings:
        off[1:] = np.cumsum([len(s) for s in strings], dtype=np.uint64)
    data = np.frombuffer("".join(strings).encode("latin-1"), dtype=np.uint8).copy()
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return data, off


class OracleAlignment:
    """Per-read results + summary in the reference's shapes (kmer.py:551-561, 622-657)."""

    def __init__(self, read_ids, types, list_off, genomes, counters, flags, genome_ids):
        self.read_ids = list(read_ids)
        self.types = types
        self.list_off = list_off
        self.genomes = genomes
        self.filtered_quality_reads = int(counters[0])
        self.filtered_quality_kmers = int(counters[1])
        self.filtered_hr_kmers = int(counters[2])
        self.flags = flags            # (read_quality, kmer_quality, max_genomes) requested
        self.genome_ids = list(genome_ids)

    def reads(self) -> Dict[str, Dict[str, object]]:
        """{read id: {"mapping_type": name, "genomes_mapped_to": [genome ids]}} for stored reads, in order."""
        out: Dict[str, Dict[str, object]] = {}
        for i, rid in enumerate(self.read_ids):
            t = int(self.types[i])
            if t == 0:
                continue
            lst = self.genomes[int(self.list_off[i]):int(self.list_off[i + 1])]
            out[rid] = {"mapping_type": MAPPING_NAMES[t], "genomes_mapped_to": [self.genome_ids[g] for g in lst]}
        return out

    def get_summary(self) -> Dict[str, Dict]:
        """PseudoAlignment.get_summary, kmer.py:622-657 (C counts on indices, identifiers merged here)."""
        L = lib()
        G = len(self.genome_ids)
        stats = np.zeros(3, dtype=np.uint64)
        uniq = np.zeros(max(G, 1), dtype=np.uint64)
        amb = np.zeros(max(G, 1), dtype=np.uint64)
        order = np.zeros(max(G, 1), dtype=np.uint32)
        n = L.orc_summary(_ptr(self.types), _ptr(self.list_off), _ptr(self.genomes), len(self.read_ids), G,
                          _ptr(stats), _ptr(uniq), _ptr(amb), _ptr(order))
        statistics = {"unique_mapped_reads": int(stats[0]), "ambiguous_mapped_reads": int(stats[1]),
                      "unmapped_reads": int(stats[2])}
        if self.flags[0]:
            statistics["filtered_quality_reads"] = self.filtered_quality_reads
        if self.flags[1]:
            statistics["filtered_quality_kmers"] = self.filtered_quality_kmers
        if self.flags[2]:
            statistics["filtered_hr_kmers"] = self.filtered_hr_kmers
        summary: Dict[str, Dict[str, int]] = {}
        for g in order[:n]:
            ent = summary.setdefault(self.genome_ids[int(g)], {"unique_reads": 0, "ambiguous_reads": 0})
            ent["unique_reads"] += int(uniq[int(g)])
            ent["ambiguous_reads"] += int(amb[int(g)])
        return {"Statistics": statistics, "Summary": summary}


class OracleReference:
    """KmerReference restated (kmer.py:109-351) over (identifier, sequence) genome pairs."""

    def __init__(self, k: int, genomes: Sequence[Tuple[str, str]], filter_similar: bool = False,
                 similarity_threshold: float = 0.95):
        if filter_similar and not (0 <= similarity_threshold <= 1):  # kmer.py:125-126
            raise ValueError("similarity_threshold must be between 0 and 1")
        self.k = int(k)
        self.genomes: List[Tuple[str, str]] = list(genomes)
        self.similarity_info: Optional[Dict[str, Dict[str, object]]] = None
        L = lib()
        data, off = pack_strings([s for _, s in self.genomes])
        err = ctypes.c_int32(0)
        self._h = L.orc_index_build(_ptr(data), _ptr(off), len(self.genomes), self.k, ctypes.byref(err))
        if not self._h:
            raise ValueError(f"oracle index build failed (code {err.value})")
        if filter_similar:
            self._filter_similar_genomes(similarity_threshold)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.orc_index_free(h)
            self._h = None

    # -- sizes / export ----------------------------------------------------
    def sizes(self) -> Tuple[int, int, int]:
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        lib().orc_index_sizes(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return a.value, b.value, c.value

    def export(self):
        """(kmers[list of str], run_off, run_genome, pos_off, pos) in dict insertion order."""
        nk, nr, no = self.sizes()
        kk = max(self.k, 0)
        keys = np.zeros(max(nk * kk, 1), dtype=np.uint8)
        run_off = np.zeros(nk + 1, dtype=np.uint64)
        run_genome = np.zeros(max(nr, 1), dtype=np.uint32)
        pos_off = np.zeros(nr + 1, dtype=np.uint64)
        pos = np.zeros(max(no, 1), dtype=np.uint32)
        lib().orc_index_export(self._h, _ptr(keys), _ptr(run_off), _ptr(run_genome), _ptr(pos_off), _ptr(pos))
        raw = keys[: nk * kk].tobytes().decode("latin-1")
        kmers = [raw[i * kk:(i + 1) * kk] for i in range(nk)]
        return kmers, run_off, run_genome[:nr], pos_off, pos[:no]

    def kmers_dict(self) -> Dict[str, Dict[int, List[int]]]:
        """{kmer: {genome index: sorted positions}} in the reference's dict order (kmer.py:130, 146-150)."""
        kmers, run_off, run_genome, pos_off, pos = self.export()
        out: Dict[str, Dict[int, List[int]]] = {}
        for e, km in enumerate(kmers):
            inner: Dict[int, List[int]] = {}
            for r in range(int(run_off[e]), int(run_off[e + 1])):
                inner[int(run_genome[r])] = [int(x) for x in pos[int(pos_off[r]):int(pos_off[r + 1])]]
            out[km] = inner
        return out

    def lookup(self, kmer: str) -> int:
        b = np.frombuffer(kmer.encode("latin-1"), dtype=np.uint8).copy() if kmer else np.zeros(1, dtype=np.uint8)
        return int(lib().orc_index_lookup(self._h, _ptr(b), len(kmer), None))

    # -- get_summary, kmer.py:300-329 ----------------------------------------
    def get_summary(self) -> Dict[str, object]:
        kd = self.kmers_dict()
        kmer_details = {km: {self.genomes[g][0]: p for g, p in inner.items()} for km, inner in kd.items()}
        summary: Dict[str, Dict[str, int]] = {}
        per_desc: Dict[str, set] = {}
        for km, inner in kd.items():
            for g in inner:
                d = self.genomes[g][0]
                ent = summary.setdefault(d, {"total_bases": 0, "unique_kmers": 0, "multi_mapping_kmers": 0})
                ent["total_bases"] = len(self.genomes[g][1])
                per_desc.setdefault(d, set()).add(km)
        for d, kms in per_desc.items():
            u = sum(1 for km in kms if len(kd[km]) == 1)
            summary[d]["unique_kmers"] = u
            summary[d]["multi_mapping_kmers"] = len(kms) - u
        out: Dict[str, object] = {"Kmers": kmer_details, "Summary": summary}
        if self.similarity_info is not None:
            out["Similarity"] = self.similarity_info
        return out

    # -- EXTSIM, kmer.py:152-263 ---------------------------------------------
    def _filter_similar_genomes(self, threshold: float) -> None:
        L = lib()
        ids = [g[0] for g in self.genomes]
        classes: Dict[str, int] = {}
        group = np.zeros(max(len(ids), 1), dtype=np.uint32)
        for i, s in enumerate(ids):
            group[i] = classes.setdefault(s, len(classes))
        n = len(classes)
        total = np.zeros(max(n, 1), dtype=np.uint64)
        uniq = np.zeros(max(n, 1), dtype=np.uint64)
        inter = np.zeros(max(n * n, 1), dtype=np.uint64)
        L.orc_extsim_stats(self._h, _ptr(group), n, _ptr(total), _ptr(uniq))
        L.orc_extsim_pairwise(self._h, _ptr(group), n, _ptr(inter))
        # genome_stats (kmer.py:164-176): one entry per identifier, later duplicates overwrite length/order
        stats: Dict[str, Dict[str, int]] = {}
        for order, (gid, seq) in enumerate(self.genomes):
            c = classes[gid]
            stats[gid] = {"unique_kmers": int(uniq[c]), "total_kmers": int(total[c]),
                          "genome_length": len(seq), "order": order}
        ordered = sorted(stats.items(), key=lambda x: (x[1]["unique_kmers"], x[1]["total_kmers"],
                                                       x[1]["genome_length"], x[1]["order"]))  # kmer.py:185-186
        kept: List[str] = []
        info: Dict[str, Dict[str, object]] = {}
        for gid, st in ordered:  # kmer.py:202-228
            ca = classes[gid]
            hit = None
            for other in kept:
                cb = classes[other]
                mn = min(int(total[ca]), int(total[cb]))
                sim = (int(inter[ca * n + cb]) / mn) if mn > 0 else 0
                if sim > threshold:
                    hit = (other, sim)
                    break
            base = {"unique_kmers": st["unique_kmers"], "total_kmers": st["total_kmers"],
                    "genome_length": st["genome_length"]}
            if hit:
                info[gid] = {"kept": "no", **base, "similar_to": hit[0], "similarity_score": hit[1]}
            else:
                info[gid] = {"kept": "yes", **base, "similar_to": "NA", "similarity_score": "NA"}
                kept.append(gid)
        kept_ids = set(kept)
        keep = np.array([1 if g[0] in kept_ids else 0 for g in self.genomes] + [0], dtype=np.uint8)
        new_h = L.orc_index_drop_genomes(self._h, _ptr(keep))
        L.orc_index_free(self._h)
        self._h = new_h
        self.genomes = [g for g in self.genomes if g[0] in kept_ids]
        self.similarity_info = info

    # -- alignment, kmer.py:563-620 -------------------------------------------
    def align(self, reads: Sequence[Tuple[str, str, str]], m: int = 1, p: int = 1,
              min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
              max_genomes: Optional[int] = None, nthreads: int = 1) -> OracleAlignment:
        if m < 0:
            raise ValueError("m must be bigger than or equal to 0")  # kmer.py:509-510
        seqs, off = pack_strings([r[1] for r in reads])
        quals, qoff = pack_strings([r[2] for r in reads])
        assert np.array_equal(off, qoff), "sequence / quality lengths differ"
        return self.align_packed([r[0] for r in reads], seqs, quals, off, m, p, min_read_quality,
                                 min_kmer_quality, max_genomes, nthreads)

    def align_packed(self, read_ids, seqs: np.ndarray, quals: np.ndarray, off: np.ndarray, m: int = 1, p: int = 1,
                     min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                     max_genomes: Optional[int] = None, nthreads: int = 1) -> OracleAlignment:
        L = lib()
        n = len(off) - 1
        types = np.zeros(max(n, 1), dtype=np.uint8)
        list_off = np.zeros(n + 1, dtype=np.uint64)
        counters = np.zeros(3, dtype=np.uint64)
        cap = max(2 * n, 16)
        while True:
            genomes = np.zeros(cap, dtype=np.uint32)
            need = L.orc_align(self._h, _ptr(seqs), _ptr(quals), _ptr(off), n, m, p,
                               int(min_read_quality is not None), int(min_read_quality or 0),
                               int(min_kmer_quality is not None), int(min_kmer_quality or 0),
                               int(max_genomes is not None), int(max_genomes or 0),
                               _ptr(types), _ptr(list_off), _ptr(genomes), cap, _ptr(counters), nthreads)
            if need <= cap:
                break
            cap = int(need)
        flags = (min_read_quality is not None, min_kmer_quality is not None, max_genomes is not None)
        return OracleAlignment(read_ids, types[:n], list_off, genomes[:int(need)], counters, flags,
                               [g[0] for g in self.genomes])


def max_threads() -> int:
    return int(lib().orc_max_threads())
