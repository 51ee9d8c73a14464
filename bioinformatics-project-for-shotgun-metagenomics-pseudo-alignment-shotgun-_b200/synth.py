"""
synth.py -- seeded synthetic genomes / reads of the shapes BASELINE.json names.

Workload generation only: used by bench.py, the parity tests
and tests/golden/make_golden.py.  Everything is a pure function of its seed and
numpy's PCG64 stream, so the same inputs can be regenerated on the GPU box.

Genomes: i.i.d. uniform ACGT; genomes are grouped into clusters whose members
share a block (a fraction of the genome) copied from a cluster ancestor with a
per-base substitution rate, which makes multi-genome k-mers and ambiguous reads
live; a few runs of 'N' keep the N path live.
Reads: uniform genome, uniform start, copy L bases ('N' -> 'A'), substitute
each base with probability sub_rate; a fraction of reads is fully random.
Qualities: chr(33 + clip(round(Normal(30, 8)), 2, 41)).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _substitute(rng: np.random.Generator, seq: np.ndarray, rate: float) -> np.ndarray:
    """Replace each base with a different base with probability `rate` (N stays N)."""
    if rate <= 0:
        return seq
    hit = rng.random(seq.size) < rate
    idx = np.nonzero(hit)[0]
    if idx.size:
        cur = seq[idx]
        shift = rng.integers(1, 4, size=idx.size)
        code = np.searchsorted(ACGT, cur)  # ACGT is sorted in ASCII
        ok = (code < 4) & (ACGT[np.minimum(code, 3)] == cur)
        new = ACGT[(code + shift) % 4]
        seq = seq.copy()
        seq[idx] = np.where(ok, new, cur)
    return seq


def make_genomes(n_genomes: int, length: int, seed: int, cluster_size: int = 4, shared_frac: float = 0.3,
                 sub_rate: float = 0.01, n_every: int = 1_000_000, n_run: int = 40,
                 length_jitter: float = 0.0) -> List[np.ndarray]:
    """Returns one uint8 ASCII array per genome."""
    rng = np.random.default_rng(seed)
    genomes: List[np.ndarray] = []
    ancestor_block: Optional[np.ndarray] = None
    for g in range(n_genomes):
        glen = length if length_jitter <= 0 else int(length * (1 + length_jitter * (rng.random() - 0.5)))
        seq = ACGT[rng.integers(0, 4, size=glen)]
        blk = int(glen * shared_frac)
        if cluster_size > 1 and blk > 0:
            if g % cluster_size == 0 or ancestor_block is None or ancestor_block.size != blk:
                ancestor_block = ACGT[rng.integers(0, 4, size=blk)]
            start = int(rng.integers(0, glen - blk + 1))
            seq[start:start + blk] = _substitute(rng, ancestor_block, sub_rate)
        if n_every > 0 and n_run > 0:
            for at in range(n_every // 2, glen - n_run, n_every):
                seq[at:at + n_run] = ord("N")
        genomes.append(seq)
    return genomes


def make_qualities(rng: np.random.Generator, total: int) -> np.ndarray:
    q = np.clip(np.rint(rng.normal(30.0, 8.0, size=total)), 2, 41).astype(np.uint8)
    return (q + 33).astype(np.uint8)


def make_reads(genomes: Sequence[np.ndarray], n_reads: int, read_len: int, seed: int, sub_rate: float = 0.01,
               random_frac: float = 0.02) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (bases uint8[n*L], quals uint8[n*L], offsets uint64[n+1]); fixed-length reads."""
    rng = np.random.default_rng(seed)
    G = len(genomes)
    lens = np.array([g.size for g in genomes], dtype=np.int64)
    which = rng.integers(0, G, size=n_reads)
    starts = (rng.random(n_reads) * np.maximum(lens[which] - read_len + 1, 1)).astype(np.int64)
    bases = np.empty((n_reads, read_len), dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    for g in range(G):
        sel = np.nonzero(which == g)[0]
        if sel.size == 0:
            continue
        src = genomes[g]
        if src.size < read_len:
            bases[sel] = ACGT[rng.integers(0, 4, size=(sel.size, read_len))]
        else:
            bases[sel] = src[starts[sel, None] + ar[None, :]]
    bases[bases == ord("N")] = ord("A")
    flat = _substitute(rng, bases.reshape(-1), sub_rate).reshape(n_reads, read_len)
    rnd = rng.random(n_reads) < random_frac
    nr = int(rnd.sum())
    if nr:
        flat = flat.copy()
        flat[rnd] = ACGT[rng.integers(0, 4, size=(nr, read_len))]
    quals = make_qualities(rng, n_reads * read_len)
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)).astype(np.uint64)
    return np.ascontiguousarray(flat.reshape(-1)), quals, off


def genomes_as_pairs(genomes: Sequence[np.ndarray], prefix: str = "genome") -> List[Tuple[str, str]]:
    return [(f"{prefix}{i}", g.tobytes().decode("ascii")) for i, g in enumerate(genomes)]


def reads_as_triples(bases: np.ndarray, quals: np.ndarray, off: np.ndarray, prefix: str = "read") -> List[Tuple[str, str, str]]:
    b = bases.tobytes().decode("ascii")
    q = quals.tobytes().decode("ascii")
    return [(f"{prefix}{i}", b[int(off[i]):int(off[i + 1])], q[int(off[i]):int(off[i + 1])])
            for i in range(len(off) - 1)]


def write_fasta(path: str, pairs: Sequence[Tuple[str, str]], width: int = 80) -> None:
    with open(path, "w") as f:
        for gid, seq in pairs:
            f.write(f">{gid}\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")


def write_fastq(path: str, triples: Sequence[Tuple[str, str, str]]) -> None:
    with open(path, "w") as f:
        for rid, seq, q in triples:
            f.write(f"@{rid}\n{seq}\n+\n{q}\n")


# ---------------------------------------------------------------------------
# Small adversarial cases for differential fuzzing (at k = 2..8 on
# short genomes every branch -- ties, flips, empty-list ambiguous, in-read
# duplicate k-mers, N windows -- fires constantly).
# ---------------------------------------------------------------------------
def fuzz_case(seed: int, k_range=(2, 8), max_genomes: int = 6, with_n: bool = True, dup_ids: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    k = int(rng.integers(k_range[0], k_range[1] + 1))
    G = int(rng.integers(1, max_genomes + 1))
    alphabet = np.frombuffer(b"ACGT" if rng.random() < 0.5 else b"AC", dtype=np.uint8)
    genomes = []
    base = alphabet[rng.integers(0, alphabet.size, size=int(rng.integers(1, 60)))]
    for g in range(G):
        mode = rng.random()
        if mode < 0.35 and base.size:
            seq = base.copy()
            nmut = int(rng.integers(0, 4))
            for _ in range(nmut):
                seq[int(rng.integers(0, seq.size))] = alphabet[int(rng.integers(0, alphabet.size))]
            if rng.random() < 0.5:
                extra = alphabet[rng.integers(0, alphabet.size, size=int(rng.integers(0, 20)))]
                seq = np.concatenate([seq, extra]) if rng.random() < 0.5 else np.concatenate([extra, seq])
        else:
            seq = alphabet[rng.integers(0, alphabet.size, size=int(rng.integers(1, 70)))]
        if with_n and rng.random() < 0.3 and seq.size > 2:
            for _ in range(int(rng.integers(1, 3))):
                at = int(rng.integers(0, seq.size))
                seq[at:at + int(rng.integers(1, 3))] = ord("N")
        gid = f"G{g}"
        if dup_ids and g > 0 and rng.random() < 0.3:
            gid = f"G{int(rng.integers(0, g))}"
        genomes.append((gid, seq.tobytes().decode("ascii")))
    reads = []
    n_reads = int(rng.integers(1, 12))
    for r in range(n_reads):
        L = int(rng.integers(1, 40))
        if rng.random() < 0.75:
            src = genomes[int(rng.integers(0, G))][1].replace("N", "A")
            if rng.random() < 0.4:
                src = src + genomes[int(rng.integers(0, G))][1].replace("N", "C")
            if len(src) >= L:
                s = int(rng.integers(0, len(src) - L + 1))
                seq = np.frombuffer(src[s:s + L].encode(), dtype=np.uint8).copy()
            else:
                seq = alphabet[rng.integers(0, alphabet.size, size=L)]
            for _ in range(int(rng.integers(0, 3))):
                seq[int(rng.integers(0, L))] = alphabet[int(rng.integers(0, alphabet.size))]
        else:
            seq = alphabet[rng.integers(0, alphabet.size, size=L)]
        qual = (33 + rng.integers(0, 60, size=L)).astype(np.uint8)
        if rng.random() < 0.3:
            qual[:] = 33 + int(rng.integers(0, 60))
        reads.append((f"r{r}", seq.tobytes().decode("ascii"), qual.tobytes().decode("ascii")))
    params = {
        "m": int(rng.integers(0, 4)),
        "p": int(rng.integers(-1, 3)),
        "mrq": None if rng.random() < 0.5 else int(rng.integers(30, 95)),
        "mkq": None if rng.random() < 0.5 else int(rng.integers(30, 95)),
        "mg": None if rng.random() < 0.5 else int(rng.integers(0, 4)),
        "filter_similar": bool(rng.random() < 0.4),
        "threshold": float(rng.choice([0.0, 0.3, 0.5, 0.75, 0.95, 1.0])),
    }
    return {"seed": seed, "k": k, "genomes": genomes, "reads": reads, "params": params}


# ---------------------------------------------------------------------------
# Canonical digest of per-read results (tools/fullsize_parity.py, bench.py's parity leg, tests/golden/fullsize_digest.json)
# ---------------------------------------------------------------------------
def flatten_results(words: np.ndarray, lst: np.ndarray):
    """See _native.flatten_results (kept here for the tools that import it from synth)."""
    import _native
    return _native.flatten_results(words, lst)


def result_digest(types: np.ndarray, lens: np.ndarray, flat: np.ndarray) -> str:
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(types, dtype=np.uint8).tobytes())
    h.update(np.ascontiguousarray(lens, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(flat, dtype=np.uint32).tobytes())
    return h.hexdigest()
