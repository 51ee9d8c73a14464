"""
Loads the UNMODIFIED Python reference from /root/reference/src under private
names so it can be compared with the oracle in this container.

/root/reference does not exist on the GPU box: every test that needs it is
skipped there (see `reference_available`).  Nothing in -m gpu tests, smoke()
or bench.py uses this module.
"""
import importlib
import os
import sys
from types import SimpleNamespace

REFERENCE_SRC = "/root/reference/src"
_NAMES = ("constants", "records", "data_file", "kmer", "main")
_cached = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "kmer.py"))


def load_reference():
    """Returns a namespace with the reference's modules (constants, records, data_file, kmer)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        return None
    saved = {n: sys.modules.pop(n, None) for n in _NAMES}
    sys.path.insert(0, REFERENCE_SRC)
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        mods = {n: importlib.import_module(n) for n in ("constants", "records", "data_file", "kmer")}
    finally:
        sys.dont_write_bytecode = old_flag
        sys.path.remove(REFERENCE_SRC)
        for n in _NAMES:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    _cached = SimpleNamespace(**mods)
    return _cached


def ref_fasta_records(ref, genomes):
    """[(id, seq)] -> reference Record objects, bypassing the regex parser."""
    R, S = ref.records.Record, ref.records.Section
    return [R([S("description", gid), S("genome", seq)]) for gid, seq in genomes]


def ref_fastq_records(ref, reads):
    R, S = ref.records.Record, ref.records.Section
    return [R([S("identifier", rid), S("sequence", seq), S("space", ""), S("quality_sequence", q)])
            for rid, seq, q in reads]


def ref_run(ref, k, genomes, reads, m=1, p=1, mrq=None, mkq=None, mg=None, filter_similar=False, threshold=0.95):
    """Runs the reference end to end; returns plain data comparable with the oracle / product."""
    recs = ref_fasta_records(ref, genomes)
    kr = ref.kmer.KmerReference(k, recs, filter_similar=filter_similar, similarity_threshold=threshold)
    index_of = {id(r): i for i, r in enumerate(kr.genomes)}
    kmers = {km: {index_of[id(r)]: sorted(pos) for r, pos in inner.items()} for km, inner in kr.kmers.items()}
    out = {
        "genomes": [r.identifier for r in kr.genomes],
        "kmers": kmers,
        "ref_summary": kr.get_summary(),
        "similarity_info": getattr(kr, "similarity_info", None),
    }
    if reads is not None:
        pa = ref.kmer.PseudoAlignment(kr)
        pa.align_reads_from_container(ref_fastq_records(ref, reads), m, p, mrq, mkq, mg)
        out["reads"] = {rid: {"mapping_type": d["mapping_type"].name, "genomes_mapped_to": list(d["genomes_mapped_to"])}
                        for rid, d in pa.reads.items()}
        out["align_summary"] = pa.get_summary()
    return out
