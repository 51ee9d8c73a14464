"""
kmer.py -- KmerReference / Read / PseudoAlignment with the reference's API
(/root/reference/src/kmer.py), computing on a B200 through libpa_b200.so.

What runs where
  device (hand-written sm_100a CUDA behind the C ABI of include/pa_b200.h)
      index build (K1 encode, K2 radix sort, K3 CSR)          <- _build_kmer_mapping, kmer.py:135-150
      k-mer lookup                                            <- get_kmer_references, kmer.py:292-298
      EXTSIM statistics, pairwise intersections, pruning      <- kmer.py:152-177, 206-207, 232-250
      per-read filters, counting, classification (K4)         <- kmer.py:394-526, 586-597
      summary reduction (K8)                                  <- get_summary, kmer.py:622-657
  host (this file)
      argument checks and exception types, the greedy EXTSIM loop over the device's integer matrix
      (kmer.py:188-230; its only float is Python's int/int, so scores and their JSON are bit-identical),
      dict-shaped views over the device arrays, gzip-pickle persistence.

There is no CPU fallback: every compute entry point needs the CUDA library and a device.
"""
import gzip
import json
import pickle
from collections import namedtuple
from collections.abc import Mapping, Sequence
from enum import Enum
from typing import Any, Dict, Iterator, List, Optional, Set, Tuple, Union

import numpy as np

import constants
import _native as nat
from records import FASTAQRecordContainer, FASTARecordContainer, Record  # noqa: F401  (re-exported like the reference)

IGNORE_AMBIGUOUS_THRESHOLD = 0
M_THRESHOLD = 0


class NotValidatingUniqueMapping(Exception):
    def __init__(self, message: str) -> None:
        super().__init__(message)


class AddingExistingRead(Exception):
    def __init__(self, message: str) -> None:
        super().__init__(message)


class ReadMappingType(Enum):
    UNMAPPED = 1
    UNIQUELY_MAPPED = 2
    AMBIGUOUSLY_MAPPED = 3


class KmerSpecifity(Enum):
    SPECIFIC = 1
    UNSPECIFIC = 2


ReadKmer = namedtuple("ReadKmer", ["specifity", "references"])
ReadMapping = namedtuple("ReadMapping", ["type", "genomes_mapped_to"])

_TYPE_BY_CODE = {1: ReadMappingType.UNMAPPED, 2: ReadMappingType.UNIQUELY_MAPPED, 3: ReadMappingType.AMBIGUOUSLY_MAPPED}


# ---------------------------------------------------------------------------
# host-only helpers kept for API compatibility (dead on every CLI path of the reference)
# ---------------------------------------------------------------------------
def extract_k_max_value_keys_from_dict(d: Dict[str, int], k: int) -> List[str]:
    if not isinstance(d, dict):
        raise ValueError("Input must be a dictionary.")
    return sorted(d, key=d.get, reverse=True)[:k] if d else []


def extract_kmers_from_genome(k: int, genome: str) -> Iterator[Tuple[int, str]]:
    """(position, k-mer) for every window; nothing when k <= 0 or k > len(genome) (kmer.py:84-94)."""
    if k <= 0 or k > len(genome):
        return iter(())
    return ((i, genome[i:i + k]) for i in range(len(genome) - k + 1))


def reverse_complement(seq: str) -> str:
    return seq.translate(str.maketrans("ACGT", "TGCA"))[::-1]


def _pack(strings, what: str):
    try:
        return nat.pack_strings(strings)
    except UnicodeEncodeError:
        raise ValueError(f"{what} contains characters outside the single-byte range")


# ---------------------------------------------------------------------------
# KmerReference
# ---------------------------------------------------------------------------
class _KmerMap(Mapping):
    """Read-only dict view {kmer: {Record: set(positions)}} over the index, in the reference's
    insertion order (first occurrence in FASTA order, kmer.py:146-150)."""

    def __init__(self, owner: "KmerReference") -> None:
        self._owner = owner

    def _entry(self, csr, u: int) -> Dict[Record, Set[int]]:
        genomes = self._owner.genomes
        out: Dict[Record, Set[int]] = {}
        for r in range(int(csr["run_off"][u]), int(csr["run_off"][u + 1])):
            out[genomes[int(csr["run_genome"][r])]] = set(csr["pos"][int(csr["pos_off"][r]):int(csr["pos_off"][r + 1])].tolist())
        return out

    def lookup_many(self, kmers: List[str]) -> List[Optional[Dict[Record, Set[int]]]]:
        """{Record: positions} of each k-mer, None when absent.  Only the looked-up entries leave the device
        (pa_index_lookup + pa_index_entries): the reference answers a lookup in O(1), so must this."""
        owner = self._owner
        k = owner.kmer_len
        out: List[Optional[Dict[Record, Set[int]]]] = [None] * len(kmers)
        good = [i for i, km in enumerate(kmers) if isinstance(km, str) and len(km) == k and k >= 1]
        if not good:
            return out
        if owner._dist is not None:      # partitioned index: the merged host CSR (collective, cached) + binary search
            csr = owner._host_csr()
            keys = owner._hashed_keys([kmers[i] for i in good])
            at = np.searchsorted(csr["keys"], keys)
            for i, a, key in zip(good, at.tolist(), keys.tolist()):
                if a < len(csr["keys"]) and int(csr["keys"][a]) == key and key != nat.RANK_MISS:
                    out[i] = self._entry(csr, a)
            return out
        ranks = owner._index().lookup([kmers[i] for i in good])
        hit = [j for j, r in enumerate(ranks.tolist()) if r != nat.RANK_MISS]
        if hit:
            run_off, run_genome, pos_off, pos = owner._index().entries(ranks[hit])
            small = {"run_off": run_off, "run_genome": run_genome, "pos_off": pos_off, "pos": pos}
            for n, j in enumerate(hit):
                out[good[j]] = self._entry(small, n)
        return out

    def __len__(self) -> int:
        return int(self._owner._index().info().n_keys)   # a replica reports the totals of all partitions

    def __iter__(self) -> Iterator[str]:
        return iter(self._owner._kmers_in_order())

    def __contains__(self, kmer) -> bool:
        return self.lookup_many([kmer])[0] is not None

    def __getitem__(self, kmer: str) -> Dict[Record, Set[int]]:
        entry = self.lookup_many([kmer])[0]
        if entry is None:
            raise KeyError(kmer)
        return entry

    def items(self):
        csr = self._owner._host_csr()
        return [(km, self._entry(csr, int(u))) for km, u in zip(self._owner._kmers_in_order(), csr["order"])]

    def values(self):
        return [v for _, v in self.items()]

    def keys(self):
        return list(iter(self))

    def __repr__(self) -> str:
        return repr(dict(self.items()))


class KmerReference(object):
    """k-mer reference database over FASTA records; the index lives in GPU memory."""

    def __init__(self, k: int, fasta_record_container: FASTARecordContainer, filter_similar: bool = False,
                 similarity_threshold: float = 0.95) -> None:
        if filter_similar and not (0 <= similarity_threshold <= 1):
            raise ValueError("similarity_threshold must be between 0 and 1")
        self.genomes: List[Record] = list(fasta_record_container)
        self.kmer_len: int = k
        self._native: Optional[nat.NativeIndex] = None
        self._dist = None         # multi_gpu.DistributedIndex when the process is one of several ranks (multi_gpu.init)
        self._csr_cache = None
        self._frozen_csr = None
        self._dropped = None      # after EXTSIM: (genome strings of the ORIGINAL list, keep mask) -- see __getstate__
        packed = getattr(fasta_record_container, "packed_batch", lambda: None)()
        self._build_kmer_mapping(self.genomes, k, packed)
        if filter_similar:
            self._filter_similar_genomes(similarity_threshold)

    # -- device index ------------------------------------------------------------
    def _build_kmer_mapping(self, fasta_records: List[Record], k: int, packed=None) -> None:
        if not isinstance(k, int) or isinstance(k, bool):
            raise TypeError("k must be an int")
        if k > 31:
            raise ValueError(f"k = {k} is outside the built scope of the B200 path (k <= 31)")
        if packed is not None and packed["n"] == len(fasta_records):
            data, off = packed["seq"], packed["off"]      # natively parsed FASTA: no Python strings on the way
            if data.size == 0:
                data = np.zeros(1, dtype=np.uint8)
        else:
            data, off = _pack([rec["genome"] for rec in fasta_records], "genome")
        self._build_index(data, off, k)

    def _build_index(self, data, off, k: int) -> None:
        """One GPU: pa_index_build.  Several ranks (multi_gpu.init() was called, e.g. main.py under torchrun): the
        partitioned build -- every rank encodes its share of the genomes, owns a key range, and holds the whole lookup
        table; same results, byte for byte."""
        import multi_gpu
        ctx = multi_gpu.context()
        self._csr_cache = None
        if ctx is None:
            self._native = nat.NativeIndex.build(data, off, k)
            self._dist = None
        else:
            self._dist = multi_gpu.build_partitioned(ctx["comm"], data, off, k, device=ctx["device"])
            self._native = self._dist.replica

    def _csr_index(self):
        """What answers the CSR questions (EXTSIM statistics, genome removal): the index itself, or its partitions."""
        self._index()
        return self._dist if self._dist is not None else self._native

    def _hashed_keys(self, kmers: List[str]) -> np.ndarray:
        flat = np.frombuffer("".join(kmers).encode("latin-1", "replace"), dtype=np.uint8).copy()
        keys = np.zeros(max(len(kmers), 1), dtype=np.uint64)
        nat.check(nat.lib().pa_encode_kmers(self.kmer_len, nat._p(flat), len(kmers), nat._p(keys)))
        return keys[:len(kmers)]

    def _index(self) -> nat.NativeIndex:
        if self._native is None:  # unpickled: re-create the device index
            st = self.__dict__.get("_frozen_csr")
            if st is not None:    # files written before the rebuild format: the CSR itself
                self._native = nat.NativeIndex.import_csr(self.kmer_len, st["genome_off"], st["keys"], st["run_off"],
                                                          st["run_genome"], st["pos_off"], st["pos"], st.get("first_occ"))
            else:
                # Rebuild format: the file holds the genomes, not the 40 bytes per k-mer of the
                # CSR -- the build is deterministic and takes a fraction of a second on the device.  After EXTSIM the
                # ORIGINAL genome list is rebuilt and the same genomes are dropped again, which also restores the
                # dict insertion order (first occurrence over the original list, kmer.py:237-243).
                dropped = self.__dict__.get("_dropped")
                strings = dropped[0] if dropped is not None else [g["genome"] for g in self.genomes]
                data, off = _pack(strings, "genome")
                self._build_index(data, off, self.kmer_len)
                if dropped is not None:
                    self._csr_index().drop_genomes(np.asarray(dropped[1], dtype=np.uint8))
                    if self._dist is not None:
                        self._native = self._dist.replica
        return self._native

    def _host_csr(self):
        """The whole CSR on the host (iteration, items, get_summary, dumpref) -- never needed by a point lookup."""
        if self._csr_cache is None:
            self._index()
            self._csr_cache = self._dist.export_gathered() if self._dist is not None else self._native.export()
        return self._csr_cache

    def _kmers_in_order(self) -> List[str]:
        csr = self._host_csr()
        if "kmers_in_order" not in csr:
            kmers = nat.decode_kmers(max(self.kmer_len, 0), csr["keys"])
            csr["kmers_in_order"] = [kmers[int(u)] for u in csr["order"]]
        return csr["kmers_in_order"]

    @property
    def kmers(self) -> _KmerMap:
        return _KmerMap(self)

    # -- EXTSIM (kmer.py:152-263) ---------------------------------------------------
    def _identifier_classes(self) -> Tuple[Dict[str, int], np.ndarray]:
        classes: Dict[str, int] = {}
        group = np.zeros(max(len(self.genomes), 1), dtype=np.uint32)
        for i, rec in enumerate(self.genomes):
            group[i] = classes.setdefault(rec.identifier, len(classes))
        return classes, group

    def _compute_genome_stats(self):
        classes, group = self._identifier_classes()
        total, unique = self._csr_index().extsim_stats(group, len(classes))
        stats: Dict[str, Dict[str, Union[int, float]]] = {}
        for order, genome in enumerate(self.genomes):
            c = classes[genome.identifier]
            stats[genome.identifier] = {"unique_kmers": int(unique[c]), "total_kmers": int(total[c]),
                                        "genome_length": len(genome["genome"]), "order": order}
        return stats, (classes, group, total)

    def _sort_genomes_for_filtering(self, genome_stats):
        return sorted(genome_stats.items(), key=lambda kv: (kv[1]["unique_kmers"], kv[1]["total_kmers"],
                                                            kv[1]["genome_length"], kv[1]["order"]))

    def _apply_greedy_filter(self, sorted_genomes, class_info, similarity_threshold: float):
        classes, group, total = class_info
        n = len(classes)
        inter = self._csr_index().extsim_pairwise(group, n)  # integer |A & B| for every pair, from the device
        kept: List[str] = []
        info: Dict[str, Dict[str, Union[str, int, float]]] = {}
        for genome_id, stats in sorted_genomes:
            a = classes[genome_id]
            verdict = None
            for other in kept:
                b = classes[other]
                smaller = min(int(total[a]), int(total[b]))
                score = (int(inter[a, b]) / smaller) if smaller > 0 else 0
                if score > similarity_threshold:
                    verdict = (other, score)
                    break
            row = {"kept": "yes" if verdict is None else "no", "unique_kmers": stats["unique_kmers"],
                   "total_kmers": stats["total_kmers"], "genome_length": stats["genome_length"],
                   "similar_to": "NA" if verdict is None else verdict[0],
                   "similarity_score": "NA" if verdict is None else verdict[1]}
            info[genome_id] = row
            if verdict is None:
                kept.append(genome_id)
        return set(kept), info

    def _filter_similar_genomes(self, similarity_threshold: float) -> None:
        stats, class_info = self._compute_genome_stats()
        kept_ids, info = self._apply_greedy_filter(self._sort_genomes_for_filtering(stats), class_info, similarity_threshold)
        keep = np.array([1 if g.identifier in kept_ids else 0 for g in self.genomes], dtype=np.uint8)
        self._csr_index().drop_genomes(keep)  # _remove_filtered_genomes_from_kmers + renumbering
        if self._dist is not None:
            self._native = self._dist.replica
        if not keep.all():
            self._dropped = ([g["genome"] for g in self.genomes], keep.tolist())
        self.genomes = [g for g in self.genomes if g.identifier in kept_ids]
        self._csr_cache = None
        self.similarity_info = info

    # -- persistence (kmer.py:265-282) -------------------------------------------------
    def __getstate__(self):
        # the device index is not stored: _index() rebuilds it from the genomes (and the EXTSIM drop list) on demand
        state = {k: v for k, v in self.__dict__.items() if k not in ("_native", "_csr_cache", "_dist")}
        if state.get("_frozen_csr") is None:
            state.pop("_frozen_csr", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._native = None
        self._dist = None
        self._csr_cache = None

    def save(self, ref_file: str) -> None:
        # gzip-pickle like the reference (kmer.py:265-271); level 1: the payload is genome text, and level 9 runs at
        # a few MB/s for nothing
        with gzip.open(ref_file, "wb", compresslevel=1) as f:
            pickle.dump(self, f, protocol=pickle.HIGHEST_PROTOCOL)

    @classmethod
    def load(cls, ref_file: str) -> "KmerReference":
        with gzip.open(ref_file, "rb") as f:
            return pickle.load(f)

    # -- lookups (kmer.py:284-298, 331-351) ----------------------------------------------
    def __getitem__(self, kmer: str) -> Optional[Dict[Record, Set[int]]]:
        return self.kmers.lookup_many([kmer])[0]

    def get_kmer_references(self, kmer: str) -> Dict[Record, Set[int]]:
        entry = self.kmers.lookup_many([kmer])[0]
        return {} if entry is None else entry

    def get_kmer_and_reverse_references(self, kmer: str) -> Dict[Record, Set[int]]:
        merged = {genome: set(positions) for genome, positions in self.get_kmer_references(kmer).items()}
        reverse = reverse_complement(kmer)
        if reverse != kmer:
            for genome, positions in self.get_kmer_references(reverse).items():
                merged.setdefault(genome, set()).update(positions)
        return merged

    # -- summary (kmer.py:300-329) -----------------------------------------------------------
    def summary_json(self, indent: int = 4) -> str:
        """json.dumps(self.get_summary(), indent=indent), byte for byte, without building the nested dictionaries: the
        "Kmers" object (all of the bulk) is written natively from the exported CSR (csrc/format.cpp), "Summary" comes
        from array reductions.  What `main.py -t dumpref` prints."""
        csr = self._host_csr()
        genomes = self.genomes
        classes: Dict[str, int] = {}
        group = np.zeros(max(len(genomes), 1), dtype=np.uint32)
        for i, rec in enumerate(genomes):
            group[i] = classes.setdefault(rec["description"], len(classes))
        names = list(classes)
        kmers_text = nat.format_kmers_json(self.kmer_len, csr, group[:max(len(genomes), 1)], [json.dumps(d) for d in names],
                                           indent=indent, level=1)
        summary: Dict[str, Dict[str, int]] = {}
        n_keys = len(csr["keys"])
        if n_keys:
            # key order of "Summary": first appearance of a description while scanning the k-mers in insertion order
            # and, inside a k-mer, the genomes in ascending order
            rank_of_key = np.empty(n_keys, dtype=np.int64)
            rank_of_key[csr["order"].astype(np.int64)] = np.arange(n_keys, dtype=np.int64)
            runs_per_key = np.diff(csr["run_off"].astype(np.int64))
            key_of_run = np.repeat(np.arange(n_keys, dtype=np.int64), runs_per_key)
            cls_of_run = group[csr["run_genome"].astype(np.int64)].astype(np.int64)
            when = rank_of_key[key_of_run] * (int(runs_per_key.max()) + 1) + (np.arange(len(key_of_run)) - csr["run_off"].astype(np.int64)[key_of_run])
            first = np.full(len(names), np.iinfo(np.int64).max, dtype=np.int64)
            np.minimum.at(first, cls_of_run, when)
            total, unique = self._csr_index().extsim_stats(group, len(names))
            # the reference assigns total_bases at every (k-mer, genome) visit (kmer.py:315), so the value left is that
            # of the genome of the class visited last: the last k-mer in insertion order holding the class, and its
            # highest such genome
            last = np.full(len(names), -1, dtype=np.int64)
            np.maximum.at(last, cls_of_run, when)
            order_runs = np.argsort(when, kind="stable")
            sorted_when = when[order_runs]
            for c in np.argsort(first, kind="stable"):
                if first[c] == np.iinfo(np.int64).max:
                    continue      # genomes without any valid k-mer do not appear (kmer.py:309-316)
                run = int(order_runs[int(np.searchsorted(sorted_when, last[c]))])
                g = int(csr["run_genome"][run])
                summary[names[int(c)]] = {"total_bases": len(genomes[g]["genome"]), "unique_kmers": int(unique[c]),
                                          "multi_mapping_kmers": int(total[c]) - int(unique[c])}
        pad = " " * indent

        def nested(obj) -> str:
            return json.dumps(obj, indent=indent).replace("\n", "\n" + pad)

        parts = [f'{pad}"Kmers": {kmers_text}', f'{pad}"Summary": {nested(summary)}']
        if hasattr(self, "similarity_info"):
            parts.append(f'{pad}"Similarity": {nested(self.similarity_info)}')
        return "{\n" + ",\n".join(parts) + "\n}"

    def get_summary(self) -> Dict[str, Any]:
        csr = self._host_csr()
        genomes = self.genomes
        run_off, run_genome, pos_off, pos = csr["run_off"], csr["run_genome"], csr["pos_off"], csr["pos"]
        kmer_details: Dict[str, Dict[str, List[int]]] = {}
        summary: Dict[str, Dict[str, int]] = {}
        for km, u in zip(self._kmers_in_order(), csr["order"]):
            u = int(u)
            inner: Dict[str, List[int]] = {}
            for r in range(int(run_off[u]), int(run_off[u + 1])):
                rec = genomes[int(run_genome[r])]
                desc = rec["description"]
                inner[desc] = pos[int(pos_off[r]):int(pos_off[r + 1])].tolist()  # ascending already
                entry = summary.setdefault(desc, {"total_bases": 0, "unique_kmers": 0, "multi_mapping_kmers": 0})
                entry["total_bases"] = len(rec["genome"])
            kmer_details[km] = inner
        if summary:
            classes: Dict[str, int] = {}
            group = np.zeros(max(len(genomes), 1), dtype=np.uint32)
            for i, rec in enumerate(genomes):
                group[i] = classes.setdefault(rec["description"], len(classes))
            total, unique = self._csr_index().extsim_stats(group, len(classes))
            for desc, entry in summary.items():
                c = classes[desc]
                entry["unique_kmers"] = int(unique[c])
                entry["multi_mapping_kmers"] = int(total[c]) - int(unique[c])
        out: Dict[str, Any] = {"Kmers": kmer_details, "Summary": summary}
        if hasattr(self, "similarity_info"):
            out["Similarity"] = self.similarity_info
        return out


# ---------------------------------------------------------------------------
# Read
# ---------------------------------------------------------------------------
def _check_align_args(kmer_reference, m, p, mrq, mkq, mg, debug=False) -> None:
    def opt_int(v):
        return v is None or isinstance(v, int)
    if not (isinstance(kmer_reference, KmerReference) and isinstance(m, int) and isinstance(p, int) and opt_int(mrq)
            and opt_int(mkq) and opt_int(mg) and isinstance(debug, bool)):
        raise TypeError(f"Invalid types given to pseudo align: {type(kmer_reference)}, {type(p)}, {type(m)}, {type(debug)}")
    if m < M_THRESHOLD:
        raise ValueError(f"m must be bigger than or equal to {M_THRESHOLD}")


class Read:
    """One sequencing read (kmer.py:357-526)."""

    def __init__(self, fastaq_record: Record) -> None:
        self.identifier = fastaq_record.identifier
        self.mapping = ReadMapping(ReadMappingType.UNMAPPED, [])
        self.kmers: Dict[str, ReadKmer] = {}
        self.__raw_read: str = fastaq_record["sequence"]
        self.__quality_scores: str = fastaq_record["quality_sequence"]
        self.num_quality_filtered_kmers: int = 0
        self.num_redundant_kmers: int = 0
        self.__genomes_map_count: Optional[Dict[Record, int]] = None

    def __str__(self) -> str:
        rows = [f"Mapping: {self.mapping}"]
        for kmer, info in self.kmers.items():
            rows += [f"k-mer: {kmer}", f"specifity: {info.specifity}", "Genome References:"]
            rows += [f"\t{reference}" for reference in info]
        return "\n".join(rows)

    __repr__ = __str__

    def _packed(self):
        return self.__raw_read, self.__quality_scores

    def mean_quality(self) -> float:
        return sum(map(ord, self.__quality_scores)) / len(self.__quality_scores)

    def kmer_quality(self, start: int, k: int) -> float:
        return sum(map(ord, self.__quality_scores[start:start + k])) / k

    def extract_kmer_references(self, kmer_reference: KmerReference, min_kmer_quality: Optional[int] = None,
                                max_genomes: Optional[int] = None) -> None:
        """Fills self.kmers with the kept k-mers (kmer.py:410-429); the lookups run on the device."""
        k = kmer_reference.kmer_len
        windows = list(extract_kmers_from_genome(k, self.__raw_read))
        if not windows:
            return
        candidates = []
        for start, kmer in windows:
            if min_kmer_quality is not None and self.kmer_quality(start, k) < min_kmer_quality:
                self.num_quality_filtered_kmers += 1
                continue
            candidates.append(kmer)
        if not candidates:
            return
        for kmer, refs in zip(candidates, kmer_reference.kmers.lookup_many(candidates)):
            if refs is None:
                continue
            if max_genomes is not None and len(refs) > max_genomes:
                self.num_redundant_kmers += 1
                continue
            self.kmers[kmer] = ReadKmer(KmerSpecifity.SPECIFIC if len(refs) == 1 else KmerSpecifity.UNSPECIFIC, refs)

    def generate_genome_counts(self, map_count: bool = False) -> Dict[Record, int]:
        counts: Dict[Record, int] = {}
        for info in self.kmers.values():
            if map_count and info.specifity != KmerSpecifity.SPECIFIC:
                continue
            for genome in info.references:
                counts[genome] = counts.get(genome, 0) + 1
        return counts

    def try_to_align_specific(self, m: int) -> bool:
        if m < 0:
            raise ValueError("m must be non-negative")
        counts = self.__genomes_map_count = self.generate_genome_counts(map_count=True)
        if len(counts) == 1:
            self.mapping = ReadMapping(ReadMappingType.UNIQUELY_MAPPED, [next(iter(counts))])
            return True
        if len(counts) > 1:
            ranked = sorted(counts, key=counts.get, reverse=True)
            if counts[ranked[0]] >= counts[ranked[1]] + m:
                self.mapping = ReadMapping(ReadMappingType.UNIQUELY_MAPPED, [ranked[0]])
                return True
        self.mapping = ReadMapping(ReadMappingType.AMBIGUOUSLY_MAPPED, list(counts))
        return False

    def validate_unique_mappings(self, p: int) -> None:
        if self.mapping.type != ReadMappingType.UNIQUELY_MAPPED or p < IGNORE_AMBIGUOUS_THRESHOLD:
            return
        totals = self.generate_genome_counts(map_count=False)
        mapped = self.mapping.genomes_mapped_to[0]
        mine = totals.get(mapped, 0)
        if max(totals.values(), default=0) - mine > p:
            self.mapping = ReadMapping(ReadMappingType.AMBIGUOUSLY_MAPPED,
                                       [mapped] + [g for g, c in totals.items() if c >= mine])

    def pseudo_align(self, kmer_reference: KmerReference, m: int = 1, p: int = 1,
                     min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                     max_genomes: Optional[int] = None, debug: bool = False) -> ReadMappingType:
        """The mapping decision is taken by the device kernel (K4) on a one-read batch; self.kmers and the
        counters are filled for inspection like the reference does."""
        _check_align_args(kmer_reference, m, p, min_read_quality, min_kmer_quality, max_genomes, debug)
        if min_read_quality is not None and self.mean_quality() < min_read_quality:
            return ReadMappingType.UNMAPPED
        seq, off = _pack([self.__raw_read], "read")
        qual, qoff = _pack([self.__quality_scores], "quality")
        if min_kmer_quality is not None and not np.array_equal(off, qoff):
            raise ValueError("sequence and quality lengths differ")   # the kernel reads one quality byte per base
        words, lst, counters = kmer_reference._index().align(
            seq, qual, off, nat.make_params(m, p, None, min_kmer_quality, max_genomes))
        types, lens, payload = nat.decode_words(words)
        self.extract_kmer_references(kmer_reference, min_kmer_quality, max_genomes)
        assert self.num_quality_filtered_kmers >= int(counters[1]) and self.num_redundant_kmers >= int(counters[2])
        code, n = int(types[0]), int(lens[0])
        if code == 1:
            return ReadMappingType.UNMAPPED  # self.mapping keeps its initial value (kmer.py:516-517)
        idx = [int(payload[0])] if n == 1 else [int(x) for x in lst[int(payload[0]):int(payload[0]) + n]]
        self.mapping = ReadMapping(_TYPE_BY_CODE[code], [kmer_reference.genomes[g] for g in idx])
        if debug:
            first = "UNIQUELY_MAPPED" if code == 2 or len(idx) != len(set(idx)) else "AMBIGUOUSLY_MAPPED"
            tail = "" if first == "UNIQUELY_MAPPED" else f", mapped to: {self.mapping}"
            print(f"[DEBUG pseudo_align]: After try_to_align_specific self.mapping: ReadMappingType.{first}{tail}")
        return self.mapping.type


# ---------------------------------------------------------------------------
# PseudoAlignment
# ---------------------------------------------------------------------------
class _Batch:
    """Per-read results of one device batch, array-backed: `types` / `list_off` / `genome_idx` cover the STORED reads
    (a read dropped by min_read_quality is not stored, kmer.py:587-589); `words` / `lst` are canonical result words of
    the whole batch (dropped reads included) for K8."""

    __slots__ = ("ids", "types", "list_off", "genome_idx", "genome_ids", "words", "lst")

    def __init__(self, ids, types, list_off, genome_idx, genome_ids, words, lst):
        self.ids = ids                  # identifiers of the stored reads, in order
        self.types = types              # uint8 mapping type per stored read
        self.list_off = list_off        # int64[n+1] into genome_idx
        self.genome_idx = genome_idx    # genome indices (into genome_ids)
        self.genome_ids = genome_ids    # identifier of every genome of the reference at alignment time
        self.words = words              # canonical result words of the whole batch, for K8
        self.lst = lst

    def entry(self, i: int) -> Dict[str, Any]:
        lo, hi = int(self.list_off[i]), int(self.list_off[i + 1])
        return {"mapping_type": _TYPE_BY_CODE[int(self.types[i])],
                "genomes_mapped_to": [self.genome_ids[int(g)] for g in self.genome_idx[lo:hi]]}

    # the pickle (.aln, kmer.py:659-665) holds the arrays, not one Python dict per read
    def __getstate__(self):
        ids = self.ids if isinstance(self.ids, _LazyIds) else _LazyIds.from_list(list(self.ids))
        return {"ids": ids, "types": self.types, "list_off": self.list_off, "genome_idx": self.genome_idx,
                "genome_ids": self.genome_ids, "words": self.words, "lst": self.lst}

    def __setstate__(self, state):
        for name in self.__slots__:
            setattr(self, name, state[name])


class _LazyIds(Sequence):
    """Identifiers of the stored reads of a batch as one byte blob + offsets; the Python strings are cut on first use (a
    summary never needs them; 10^7 Python strings take seconds).  For a natively parsed FASTQ (csrc/ingest.cpp) the blob is
    gathered from the parsed text with array operations only."""

    def __init__(self, packed, stored, n_total: int) -> None:
        self._packed = packed
        self._stored = None if len(stored) == n_total else stored
        self._n = len(stored)
        self._blob: Optional[bytes] = None
        self._off: Optional[np.ndarray] = None
        self._list: Optional[List[str]] = None

    @classmethod
    def from_list(cls, names: List[str]) -> "_LazyIds":
        self = cls.__new__(cls)
        self._packed = self._stored = None
        self._n = len(names)
        enc = [s.encode("utf-8") for s in names]
        self._blob = b"".join(enc)
        self._off = np.concatenate([[0], np.cumsum([len(e) for e in enc], dtype=np.int64)]).astype(np.int64)
        self._list = list(names)
        return self

    def _compact(self) -> None:
        """(blob, offsets) of the stored identifiers, without building a Python string per read."""
        if self._blob is not None:
            return
        pk = self._packed
        beg, ln = pk["name_beg"].astype(np.int64), pk["name_len"].astype(np.int64)
        if self._stored is not None:
            beg, ln = beg[self._stored], ln[self._stored]
        raw = pk["raw"]
        text = np.frombuffer(raw.encode("ascii") if isinstance(raw, str) else raw, dtype=np.uint8)
        off = np.zeros(len(ln) + 1, dtype=np.int64)
        np.cumsum(ln, out=off[1:])
        total = int(off[-1])
        idx = np.repeat(beg - off[:-1], ln) + np.arange(total, dtype=np.int64)
        self._blob, self._off = text[idx].tobytes(), off
        self._packed = self._stored = None

    def _get(self) -> List[str]:
        if self._list is None:
            self._compact()
            text, off = self._blob.decode("utf-8"), self._off
            if len(text) == len(self._blob):     # ASCII: byte offsets are character offsets
                o = off.tolist()
                self._list = [text[o[i]:o[i + 1]] for i in range(self._n)]
            else:
                o = off.tolist()
                self._list = [self._blob[o[i]:o[i + 1]].decode("utf-8") for i in range(self._n)]
        return self._list

    def __getstate__(self):
        self._compact()
        return {"n": self._n, "blob": self._blob, "off": self._off}

    def __setstate__(self, state):
        self._packed = self._stored = self._list = None
        self._n, self._blob, self._off = state["n"], state["blob"], state["off"]

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())


class _ReadsView(Mapping):
    """Dict-shaped view of PseudoAlignment.reads: {read id: {"mapping_type", "genomes_mapped_to"}} in insertion order.
    Device batches stay as arrays; reads added one at a time (add_read) are kept as plain entries."""

    def __init__(self) -> None:
        self.chunks: List[Union[_Batch, Dict[str, Dict[str, Any]]]] = []
        self._where: Dict[str, Tuple[int, int]] = {}
        self._indexed_chunks = 0

    def _index_ids(self) -> Dict[str, Tuple[int, int]]:
        while self._indexed_chunks < len(self.chunks):
            c = self._indexed_chunks
            chunk = self.chunks[c]
            if isinstance(chunk, _Batch):
                self._where.update((rid, (c, i)) for i, rid in enumerate(chunk.ids))
            self._indexed_chunks += 1
        return self._where

    def __contains__(self, rid) -> bool:
        if any(isinstance(ch, dict) and rid in ch for ch in self.chunks):
            return True
        return rid in self._index_ids()

    def __getitem__(self, rid):
        for ch in self.chunks:
            if isinstance(ch, dict) and rid in ch:
                return ch[rid]
        c, i = self._index_ids()[rid]
        return self.chunks[c].entry(i)

    def __iter__(self):
        for ch in self.chunks:
            yield from (ch if isinstance(ch, dict) else ch.ids)

    def __len__(self) -> int:
        return sum(len(ch) if isinstance(ch, dict) else len(ch.ids) for ch in self.chunks)

    def items(self):
        for ch in self.chunks:
            if isinstance(ch, dict):
                yield from ch.items()
            else:
                for i, rid in enumerate(ch.ids):
                    yield rid, ch.entry(i)

    def add_single(self, rid: str, entry: Dict[str, Any]) -> None:
        if not self.chunks or not isinstance(self.chunks[-1], dict):
            self.chunks.append({})
        self.chunks[-1][rid] = entry

    def __repr__(self) -> str:
        return repr(dict(self.items()))


class PseudoAlignment:
    """Aligns reads against a KmerReference and aggregates the results (kmer.py:532-699)."""

    def __init__(self, kmer_reference: KmerReference) -> None:
        self.kmer_reference: KmerReference = kmer_reference
        self.reads = _ReadsView()
        self.filtered_quality_reads: int = 0
        self.filtered_quality_kmers: int = 0
        self.filtered_hr_kmers: int = 0
        self.filter_read_quality_flag: bool = False
        self.filter_kmer_quality_flag: bool = False
        self.filter_max_genomes_flag: bool = False

    def add_read(self, read: Read) -> None:
        if read.identifier in self.reads:
            raise AddingExistingRead(f"There already exists a read with identifier: {read.identifier}")
        self.reads.add_single(read.identifier, {
            "mapping_type": read.mapping.type,
            "genomes_mapped_to": [genome.identifier for genome in read.mapping.genomes_mapped_to],
        })

    def _set_flags(self, mrq, mkq, mg) -> None:
        self.filter_read_quality_flag |= mrq is not None
        self.filter_kmer_quality_flag |= mkq is not None
        self.filter_max_genomes_flag |= mg is not None

    def add_read_from_read_record(self, read_record: Record, m: int = 1, p: int = 1,
                                  min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                                  max_genomes: Optional[int] = None) -> None:
        self._align_records([read_record], m, p, min_read_quality, min_kmer_quality, max_genomes)

    def align_reads_from_container(self, reads_container: FASTAQRecordContainer, m: int = 1, p: int = 1,
                                   min_read_quality: Optional[int] = None, min_kmer_quality: Optional[int] = None,
                                   max_genomes: Optional[int] = None) -> None:
        packed = getattr(reads_container, "packed_batch", lambda: None)()
        if packed is not None and len(self.reads) == 0:
            # natively parsed FASTQ (csrc/ingest.cpp): the arrays go to the device as they are; identifiers are unique
            # inside one container (records.py:195-198) and nothing has been filed before, so no read can collide
            self._align_arrays(None, packed, packed["seq"], packed["qual"], packed["off"], m, p,
                               min_read_quality, min_kmer_quality, max_genomes)
            return
        self._align_records(list(reads_container), m, p, min_read_quality, min_kmer_quality, max_genomes)

    def _align_records(self, records: List[Record], m, p, mrq, mkq, mg) -> None:
        """One device batch for all records: K4 classifies every read, the host only files the results."""
        if not records:
            return
        self._set_flags(mrq, mkq, mg)
        _check_align_args(self.kmer_reference, m, p, mrq, mkq, mg)
        seqs = [r["sequence"] for r in records]
        quals = [r["quality_sequence"] for r in records]
        if mrq is not None and any(len(q) == 0 for q in quals):
            raise ZeroDivisionError("division by zero")  # Read.mean_quality on an empty read (kmer.py:399)
        seq_bytes, off = _pack(seqs, "read")
        need_q = mrq is not None or mkq is not None
        qual_bytes = None
        if need_q:
            qual_bytes, qoff = _pack(quals, "quality")
            if not np.array_equal(off, qoff):
                raise ValueError("sequence and quality lengths differ")
        self._align_arrays(records, None, seq_bytes, qual_bytes, off, m, p, mrq, mkq, mg)

    def _align_arrays(self, records: Optional[List[Record]], names, seq_bytes, qual_bytes, off,
                      m, p, mrq, mkq, mg) -> None:
        """records is None for a natively parsed container (names = its packed batch; its identifiers are known to be
        collision-free and are only cut out of the text when somebody looks at `reads`)."""
        if records is None:
            self._set_flags(mrq, mkq, mg)
            _check_align_args(self.kmer_reference, m, p, mrq, mkq, mg)
            if seq_bytes.size == 0:
                seq_bytes = np.zeros(1, dtype=np.uint8)
            if not (mrq is not None or mkq is not None):
                qual_bytes = None
        ref = self.kmer_reference
        import multi_gpu
        ctx = multi_gpu.context()
        params = nat.make_params(m, p, mrq, mkq, mg)
        if ctx is None:
            words, lst, counters = ref._index().align(seq_bytes, qual_bytes, off, params)
            types_all, lens_all, flat = nat.flatten_results(words, lst)
        else:
            # reads are independent units (kmer.py:616-620): rank r aligns one contiguous block against its replica of the
            # index; the per-read arrays are gathered so that `reads`, save() and get_summary() agree on every rank
            comm, n_reads = ctx["comm"], len(off) - 1
            lo, hi = multi_gpu.shard_bounds(n_reads, ctx["world"], ctx["rank"])
            w, l, c = ref._index().align(seq_bytes, qual_bytes, np.ascontiguousarray(off[lo:hi + 1]), params)
            t, ln, fl = nat.flatten_results(w, l)
            types_all = np.concatenate(comm.allgather_array(t))
            lens_all = np.concatenate(comm.allgather_array(ln))
            flat = np.concatenate(comm.allgather_array(fl))
            counters = comm.allreduce_host(np.asarray(c, dtype=np.uint64))
        words, lst = nat.canonical_words(types_all, lens_all, flat)
        types, lens = types_all, lens_all.astype(np.int64)
        stored = np.nonzero(types != 0)[0]
        if records is None:
            ids = _LazyIds(names, stored, int(names["n"]))
        else:
            ids = [records[int(i)].identifier for i in stored]
        # duplicate identifiers: the reference raises at the first one, after filing everything before it
        seen_here: Set[str] = set()
        dup_at = None
        for j, rid in enumerate(ids if records is not None else ()):
            if rid in seen_here or rid in self.reads:
                dup_at = j
                break
            seen_here.add(rid)
        if dup_at is not None:
            cut = int(stored[dup_at])
            if cut > 0:
                self._align_records(records[:cut], m, p, mrq, mkq, mg)
            if mrq is not None:
                pass  # the duplicate itself passed the read filter, nothing to count
            dup = records[cut]
            extra = PseudoAlignment(ref)
            extra._align_records([dup], m, p, None, mkq, mg)
            self.filtered_quality_kmers += extra.filtered_quality_kmers
            self.filtered_hr_kmers += extra.filtered_hr_kmers
            raise AddingExistingRead(f"There already exists a read with identifier: {ids[dup_at]}")
        self.filtered_quality_reads += int(counters[0])
        if mkq is not None:
            self.filtered_quality_kmers += int(counters[1])
        if mg is not None:
            self.filtered_hr_kmers += int(counters[2])
        # a dropped read has an empty list, so the lists of the stored reads are `flat` as it is
        s_lens = lens[stored]
        list_off = np.zeros(len(stored) + 1, dtype=np.int64)
        np.cumsum(s_lens, out=list_off[1:])
        batch = _Batch(ids, types[stored].astype(np.uint8), list_off, flat, [g.identifier for g in ref.genomes], words, lst)
        self.reads.chunks.append(batch)

    def get_summary(self) -> Dict[str, Dict[str, Union[int, Dict[str, int]]]]:
        statistics: Dict[str, int] = {"unique_mapped_reads": 0, "ambiguous_mapped_reads": 0, "unmapped_reads": 0}
        if self.filter_read_quality_flag:
            statistics["filtered_quality_reads"] = self.filtered_quality_reads
        if self.filter_kmer_quality_flag:
            statistics["filtered_quality_kmers"] = self.filtered_quality_kmers
        if self.filter_max_genomes_flag:
            statistics["filtered_hr_kmers"] = self.filtered_hr_kmers
        genome_mapping: Dict[str, Dict[str, int]] = {}
        for chunk in self.reads.chunks:
            if isinstance(chunk, _Batch):  # K8 on the device, merged here in first-appearance order
                stats, uniq, amb, first = self.kmer_reference._index().summary(chunk.words, chunk.lst)
                statistics["unique_mapped_reads"] += int(stats[0])
                statistics["ambiguous_mapped_reads"] += int(stats[1])
                statistics["unmapped_reads"] += int(stats[2])
                never = np.uint64(0xFFFFFFFFFFFFFFFF)
                for g in np.argsort(first, kind="stable"):
                    if first[g] == never:
                        break
                    row = genome_mapping.setdefault(chunk.genome_ids[int(g)], {"unique_reads": 0, "ambiguous_reads": 0})
                    row["unique_reads"] += int(uniq[g])
                    row["ambiguous_reads"] += int(amb[g])
            else:
                for details in chunk.values():
                    kind = details["mapping_type"]
                    if kind == ReadMappingType.UNMAPPED:
                        statistics["unmapped_reads"] += 1
                        continue
                    column = None
                    if kind == ReadMappingType.UNIQUELY_MAPPED:
                        statistics["unique_mapped_reads"] += 1
                        column = "unique_reads"
                    elif kind == ReadMappingType.AMBIGUOUSLY_MAPPED:
                        statistics["ambiguous_mapped_reads"] += 1
                        column = "ambiguous_reads"
                    if column:
                        for genome in details["genomes_mapped_to"]:
                            genome_mapping.setdefault(genome, {"unique_reads": 0, "ambiguous_reads": 0})[column] += 1
        return {"Statistics": statistics, "Summary": genome_mapping}

    # -- persistence and reporting (kmer.py:659-699) ------------------------------------------------
    def __getstate__(self):
        # the view's chunks travel as they are: array-backed batches (with their identifiers as one blob) and the plain
        # dict entries of reads added one at a time -- never 10^7 Python dict entries
        state = dict(self.__dict__)
        state["reads"] = {"chunks": self.reads.chunks}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        view = _ReadsView()
        stored = state["reads"]
        if isinstance(stored, dict) and set(stored.keys()) == {"chunks"} and isinstance(stored["chunks"], list):
            view.chunks = stored["chunks"]
        elif stored:                      # files written before the array format: a plain dict of every read
            view.chunks.append(dict(stored))
        self.reads = view

    def save(self, align_file: str) -> None:
        with gzip.open(align_file, "wb", compresslevel=1) as f:
            pickle.dump(self, f, protocol=pickle.HIGHEST_PROTOCOL)

    def __repr__(self) -> str:
        return json.dumps(self.get_summary(), indent=4)

    @classmethod
    def load(cls, align_file: str) -> "PseudoAlignment":
        with gzip.open(align_file, "rb") as f:
            return pickle.load(f)

    def export_summary_to_json(self, json_file: str) -> None:
        with open(json_file, "w") as f:
            json.dump(self.get_summary(), f, indent=4)

    def get_reads_by_mapping_type(self, mapping_type: ReadMappingType) -> List[str]:
        return [rid for rid, details in self.reads.items() if details["mapping_type"] == mapping_type]
