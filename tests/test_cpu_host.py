"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host-side record / CLI logic,
and the world_size-2 summary exchange over gloo.  No compute call needs a GPU here."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import conftest
import synth
from oracle import oracle as orc

PKG = conftest.PKG_DIR
ROOT = conftest.ROOT


def test_abi_library_exports_every_declared_symbol():
    import _native as nat
    header = open(os.path.join(ROOT, "include", "pa_b200.h")).read()
    declared = re.findall(r"^int32_t\s+(pa_\w+)\s*\(", header, flags=re.M)
    assert len(declared) >= 20
    lib = nat.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/pa_b200.h but not exported"
    assert sorted(declared) == sorted(nat.EXPORTED_SYMBOLS)
    assert lib.pa_abi_version() == 3


def test_encode_decode_kmers_round_trip_on_host():
    import _native as nat
    kmers = ["ACGTACGTACGTACGTACGTACGTACGTACG", "T" * 31, "G" * 31]
    keys = np.zeros(3, dtype=np.uint64)
    flat = np.frombuffer("".join(kmers).encode(), dtype=np.uint8).copy()
    nat.check(nat.lib().pa_encode_kmers(31, nat._p(flat), 3, nat._p(keys)))
    assert nat.decode_kmers(31, keys) == kmers
    assert len(set(keys.tolist())) == 3 and int(keys.max()) < (1 << 62)
    bad = np.frombuffer(("ACGN" + "A" * 27).encode(), dtype=np.uint8).copy()
    nat.check(nat.lib().pa_encode_kmers(31, nat._p(bad), 1, nat._p(keys)))
    assert int(keys[0]) == 0xFFFFFFFFFFFFFFFF


def test_product_fails_loudly_without_a_device():
    import _native as nat
    if nat.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from kmer import KmerReference
    from records import FASTARecordContainer
    c = FASTARecordContainer()
    c.parse_records(">g\nACGTACGT\n")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        KmerReference(3, c)


# ---------------------------------------------------------------------------
# records / data_file (API-compatible host shims)
# ---------------------------------------------------------------------------
def test_fasta_and_fastq_parsing_rules():
    from records import (FASTAQRecordContainer, FASTARecordContainer, DuplicateRecordError, InvalidRecordData,
                         NoRecordsInData, Record, Section, UnparsedDataError)
    c = FASTARecordContainer()
    c.parse_records(">g1 some description\nACGT\nNNAC\r\n>g2\nTTTT")
    recs = list(c)
    assert [r.identifier for r in recs] == ["g1 some description", "g2"]
    assert recs[0]["genome"] == "ACGTNNAC" and recs[1]["genome"] == "TTTT" and recs[0]["description"] == recs[0].identifier
    with pytest.raises(UnparsedDataError):
        FASTARecordContainer().parse_records(">g1\nACGT\n>g2\nacgt\n")
    with pytest.raises(NoRecordsInData, match="No valid records found in the data."):
        FASTARecordContainer().parse_records("")
    q = FASTAQRecordContainer()
    q.parse_records("@r1\nACGT\n+\nIIII\n@r2\nGG\n+..\n!~\n")
    assert [(r.identifier, r["sequence"], r["quality_sequence"]) for r in q] == [("r1", "ACGT", "IIII"), ("r2", "GG", "!~")]
    with pytest.raises(NoRecordsInData):
        FASTAQRecordContainer().parse_records("@r1\nACGN\n+\nIIII\n")          # N is not allowed in reads
    with pytest.raises(InvalidRecordData):
        FASTAQRecordContainer().parse_records("@r1\nACGT\n+\nIII\n")
    with pytest.raises(DuplicateRecordError):
        FASTAQRecordContainer().parse_records("@r1\nACGT\n+\nIIII\n@r1\nACGT\n+\nIIII\n")
    with pytest.raises(InvalidRecordData, match="has no sections"):
        Record([])
    with pytest.raises(InvalidRecordData, match="has appeared twice"):
        Record([Section("a", "x"), Section("a", "y")])


def test_parsers_agree_with_the_reference_when_it_is_present():
    import refimpl
    if not refimpl.reference_available():
        pytest.skip("/root/reference not present")
    ref = refimpl.load_reference()
    import records as mine
    rng = np.random.default_rng(5)
    fasta_alpha = list("ACGTN\n\r >@+acgt \t")
    fastq_alpha = list("ACGT\n\n\r@+I!~N ")
    for trial in range(400):
        for alpha, mine_cls, ref_cls in ((fasta_alpha, mine.FASTARecordContainer, ref.records.FASTARecordContainer),
                                         (fastq_alpha, mine.FASTAQRecordContainer, ref.records.FASTAQRecordContainer)):
            if trial % 2 == 0:
                text = "".join(rng.choice(alpha, size=int(rng.integers(0, 60))))
            elif mine_cls is mine.FASTARecordContainer:
                text = "".join(f">g{i}\n" + "".join(rng.choice(list("ACGTN\n"), size=int(rng.integers(1, 30)))) + "\n" for i in range(3))
            else:
                text = "".join(f"@r{i % 2 if trial % 7 == 0 else i}\n{s}\n+\n{'I' * (len(s) - (trial % 5 == 0))}\n"
                               for i, s in enumerate("".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 9)))) for _ in range(3)))
            outcomes = []
            for cls in (mine_cls, ref_cls):
                cont = cls()
                try:
                    cont.parse_records(text)
                    outcomes.append([str(r) for r in cont])
                except Exception as e:  # same exception type and message
                    outcomes.append((type(e).__name__, str(e)))
            assert outcomes[0] == outcomes[1], repr(text)


def test_data_file_errors(tmp_path):
    from data_file import FASTAFile, FASTAQFile, InvalidExtensionError, NoRecordsInDataFile
    p = tmp_path / "x.txt"
    p.write_text(">g\nACGT\n")
    with pytest.raises(InvalidExtensionError, match="Invalid file extension"):
        FASTAFile(str(p))
    e = tmp_path / "e.fa"
    e.write_text("")
    with pytest.raises(NoRecordsInDataFile):
        FASTAFile(str(e))
    import gzip
    g = tmp_path / "r.fq.gz"
    with gzip.open(g, "wt") as f:
        f.write("@r\nACGT\n+\nIIII\n")
    assert [r.identifier for r in FASTAQFile(str(g)).container] == ["r"]


def run_cli(*argv):
    r = subprocess.run([sys.executable, os.path.join(PKG, "main.py"), *argv], capture_output=True, text=True, cwd=PKG)
    return r.stdout, r.stderr, r.returncode


def test_cli_error_paths_that_never_reach_the_device(tmp_path):
    out, err, rc = run_cli("-t", "reference", "-g", str(tmp_path / "missing.fa"), "-k", "4", "-r", str(tmp_path / "o.kdb"))
    assert rc != 0 and "does not exist" in err
    bad = tmp_path / "genome.txt"
    bad.write_text(">g\nACGT\n")
    out, err, rc = run_cli("-t", "reference", "-g", str(bad), "-k", "4", "-r", str(tmp_path / "o.kdb"))
    assert rc != 0 and "Invalid file extension" in err
    out, err, rc = run_cli("-t", "nonsense")
    assert rc != 0 and "Unsupported task" in err
    fa = tmp_path / "g.fa"
    fa.write_text(">g\nACGT\n")
    out, err, rc = run_cli("-t", "reference", "-g", str(fa), "-k", "4", "-r", str(tmp_path / "o.kdb"), "-m", "2")
    assert rc != 0 and "For task 'reference'" in err
    out, err, rc = run_cli("-t", "align", "-g", str(fa))
    assert rc != 0 and "For task 'align'" in err
    out, err, rc = run_cli("-t", "dumpalign")
    assert rc != 0 and "For task 'dumpalign'" in err
    notgz = tmp_path / "x.kdb"
    notgz.write_text("not a gzip file")
    out, err, rc = run_cli("-t", "dumpref", "-r", str(notgz))
    assert rc != 0 and "Incorrect format of input file" in err


def test_cli_argument_parsing_keeps_the_reference_flags():
    import main
    a = main.parse_arguments(["-t", "dumpalign", "-r", "x.kdb", "--reads", "r.fq", "-m", "2", "-p", "3", "--reverse-complement",
                              "--min-read-quality", "10", "--min-kmer-quality", "11", "--max-genomes", "4",
                              "--filter-similar", "--similarity-threshold", "0.5"])
    assert (a.task, a.unique_threshold, a.ambiguous_threhold, a.reverse_complement) == ("dumpalign", 2, 3, True)
    assert (a.min_read_quality, a.min_kmer_quality, a.max_genomes, a.filter_similar, a.similarity_threshold) == (10, 11, 4, True, 0.5)


# ---------------------------------------------------------------------------
# world_size-2 summary exchange (gloo)
# ---------------------------------------------------------------------------
WORKER = r'''
import json, os, sys
sys.path[:0] = [{root!r}, {pkg!r}, os.path.join({root!r}, "tests")]
import numpy as np, torch, torch.distributed as dist
import synth, multi_gpu
from oracle import oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
case = synth.fuzz_case(int(sys.argv[1]), max_genomes=5)
pr = case["params"]
o = orc.OracleReference(case["k"], case["genomes"])
reads = case["reads"] * 3
reads = [(f"{{r[0]}}_{{i}}", r[1], r[2]) for i, r in enumerate(reads)]
lo, hi = multi_gpu.shard_bounds(len(reads), world, rank)
al = o.align(reads[lo:hi], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
G = len(case["genomes"])
# per-rank accumulators in the layout K8 produces on the device
acc = np.zeros(4 + 2 * G, dtype=np.int64)
first = np.full(G, -1, dtype=np.int64)
for i in range(hi - lo):
    t = int(al.types[i])
    if t == 0: acc[3] += 1; continue
    if t == 1: acc[2] += 1; continue
    acc[0 if t == 2 else 1] += 1
    for j, g in enumerate(al.genomes[int(al.list_off[i]):int(al.list_off[i + 1])]):
        acc[4 + (0 if t == 2 else G) + int(g)] += 1
        key = ((lo + i) << 22) | j
        first[g] = key if first[g] < 0 else min(first[g], key)
import _native as nat
# the product's communicator (pa_comm, csrc/comm.cu) over a gloo all-gather: host-only, no CUDA device needed
comm = nat.Comm.callbacks(world, rank, -1, lambda data: _allgather(data))
def _allgather(data):
    mine = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.numpy().tobytes()) for t in out]
counters = np.array([al.filtered_quality_reads, al.filtered_quality_kmers, al.filtered_hr_kmers], dtype=np.uint64)
stats, uniq, amb, counters, fs = multi_gpu.reduce_summary(comm, acc[:4].astype(np.uint64), acc[4:4 + G].astype(np.uint64),
                                                          acc[4 + G:].astype(np.uint64), counters, first.view(np.uint64))
assert comm.allgather_bytes(bytes([rank]) * (rank + 1)) == [bytes([r]) * (r + 1) for r in range(world)]
if rank == 0:
    flags = (pr["mrq"] is not None, pr["mkq"] is not None, pr["mg"] is not None)
    got = multi_gpu.summary_from_accumulators(stats, uniq, amb, fs, [g[0] for g in case["genomes"]], flags, counters.tolist())
    want = o.align(reads, pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"]).get_summary()
    assert json.dumps(got) == json.dumps(want), (got, want)
    print("OK")
comm.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("seed", [3, 17, 64, 301])
def test_world_size_2_summary_exchange_matches_single_process(tmp_path, seed):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, pkg=PKG))
    port = 29500 + (os.getpid() + seed) % 2000
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(seed)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    assert "OK" in outs[0][0]


def test_shard_bounds_cover_everything_once():
    import multi_gpu
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            blocks = [multi_gpu.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


# ---------------------------------------------------------------------------
# multi-GPU build: host-side partitioning logic (the kernels are covered by tests/test_gpu_multi.py)
# ---------------------------------------------------------------------------
def test_genome_shards_are_contiguous_and_balanced():
    import _native as nat
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 100):
            lengths = rng.integers(1, 1000, size=n)
            off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
            shards = [nat.genome_shard(off, world, r) for r in range(world)]
            assert len(shards) == world and shards[0][0] == 0 and shards[-1][1] == n
            assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
            if n == 100:
                loads = [int(lengths[a:b].sum()) for a, b in shards]
                assert max(loads) - min(loads) <= 2 * int(lengths.max())


def test_partition_of_kmer_follows_the_minimizer_and_is_balanced():
    """pa_partition_of_kmer is pure host code: the owner of a k-mer is a function of the top bits of its minimizer hash
    (so k-mers sharing minimizer and table block share their owner), and the owners carry equal load although the
    orders of minimizers crowd near zero (the hash spreads them)."""
    import ctypes
    import _native as nat
    L = nat.lib()
    rng = np.random.default_rng(9)
    for k in (4, 11, 31):
        n = 20000
        kmers = ["".join(x) for x in np.frombuffer(b"ACGT", dtype="S1")[rng.integers(0, 4, size=(n, k))].astype(str)]
        flat = np.frombuffer("".join(kmers).encode(), dtype=np.uint8).copy()
        mh = np.zeros(n, np.uint32); off = np.zeros(n, np.uint32)
        assert L.pa_debug_minimizer(k, nat._p(flat), n, nat._p(mh), nat._p(off)) == 0
        m = min(k, 16)
        tb = min(8, 2 * m)
        digit = mh.astype(np.uint64) >> np.uint64(2 * m - tb)
        for parts in (1, 2, 3, 8):
            got = np.array([nat.partition_of_kmer(k, km, parts) for km in kmers[:3000]])
            want = (digit[:3000] * np.uint64(parts)) >> np.uint64(tb)
            assert np.array_equal(got, want.astype(np.int64))
            assert got.max() == parts - 1 and got.min() == 0
            if k >= 11:
                counts = np.bincount(((digit * np.uint64(parts)) >> np.uint64(tb)).astype(np.int64), minlength=parts)
                assert counts.max() < 1.15 * counts.mean() + 50
        if k >= 11:   # the digits themselves: no digit holds more than a few times its share
            counts = np.bincount(digit.astype(np.int64), minlength=1 << tb)
            assert counts.max() < 4 * counts.mean() + 20


@pytest.mark.parametrize("scalar", ["0", "1"])
def test_host_packing_matches_a_numpy_restatement(scalar, monkeypatch):
    """hostpack.cpp (AVX2 and the portable scalar path, threaded) against the plane definition: bit i of word c = ASCII
    bit 1 / bit 2 of base 32c + i."""
    import ctypes
    import _native as nat
    monkeypatch.setenv("PA_PACK_SCALAR", scalar)
    L = nat.lib()
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(1, 4000)) if trial < 6 else int(rng.integers(1, 60))
        lens = rng.integers(0, 330, size=n) if trial % 2 else np.full(n, 150)
        off = np.concatenate([[7], 7 + np.cumsum(lens)]).astype(np.uint64)      # offsets need not start at 0
        total = int(off[-1])
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=total)].copy()
        bad = trial in (4, 5) or trial >= 10
        if bad:   # one byte outside ACGT anywhere (full blocks and tails): same low nibble as a base, lower case, >= 0x80, NUL
            where = int(off[n // 2]) if trial == 5 else int(rng.integers(7, total))
            bases[where] = [ord("N"), ord("Q"), ord("D"), ord("a"), 0xC1, 0xD4, 0x00, 0xFF, ord("S"), ord("W")][int(rng.integers(0, 10))]
        cap = 2 * ((total - 7) // 32 + n + 1)
        for threads in (1, 4):
            planes = np.full(cap, 0xDEADBEEF, dtype=np.uint32)
            ok = ctypes.c_int32(-1)
            nat.check(L.pa_debug_pack_reads(nat._p(bases), nat._p(off), n, nat._p(planes), cap, threads, ctypes.byref(ok)))
            assert ok.value == (0 if bad else 1)
            if bad:
                continue
            for i in rng.choice(n, size=min(n, 200), replace=False):
                o, ln = int(off[i]), int(lens[i])
                nw = (ln + 31) // 32
                w0 = 2 * ((o - 7) // 32 + int(i))
                seq = bases[o:o + ln]
                for c in range(nw):
                    blk = seq[32 * c:32 * c + 32].astype(np.uint32)
                    lo = int(sum(((int(b) >> 1) & 1) << j for j, b in enumerate(blk)))
                    hi = int(sum(((int(b) >> 2) & 1) << j for j, b in enumerate(blk)))
                    assert int(planes[w0 + c]) == lo and int(planes[w0 + nw + c]) == hi


# ---------------------------------------------------------------------------
# native ingest (csrc/ingest.cpp) against the regular-expression parser, and against the reference when present
# ---------------------------------------------------------------------------
def _random_text(rng, fastq: bool) -> str:
    n = int(rng.integers(1, 6))
    eol = "\r\n" if rng.random() < 0.25 else "\n"
    out = []
    for i in range(n):
        name = f"rec{i}" + (" extra words" if rng.random() < 0.3 else "") + ("\t" if rng.random() < 0.1 else "")
        L = int(rng.integers(1, 30))
        if fastq:
            seq = "".join(rng.choice(list("ACGT"), size=L))
            qual = "".join(chr(int(c)) for c in rng.integers(33, 127, size=L))
            plus = "+" + (".." if rng.random() < 0.3 else "") + ("comment" if rng.random() < 0.05 else "")
            out.append(f"@{name}{eol}{seq}{eol}{plus}{eol}{qual}")
        else:
            seq = "".join(rng.choice(list("ACGTN"), size=L))
            w = int(rng.integers(3, 12))
            out.append(f">{name}{eol}" + eol.join(seq[j:j + w] for j in range(0, L, w)))
    text = eol.join(out) + (eol if rng.random() < 0.7 else "")
    # mutations that leave the canonical form (the native parser must hand these to the regex)
    r = rng.random()
    if r < 0.08:
        text = text.replace("A", "a", 1)
    elif r < 0.16:
        text = text.replace(eol, eol + eol, 1)
    elif r < 0.22:
        text = " " + text
    elif r < 0.28:
        text = text + eol + eol
    elif r < 0.34 and fastq:
        text = text.replace("rec1", "rec0")            # duplicate identifier
    elif r < 0.40 and fastq:
        k = text.rfind(eol, 0, len(text) - 2)
        text = text[:k + len(eol)] + text[k + len(eol) + 1:]   # quality one short
    elif r < 0.46:
        text = text[: max(1, len(text) // 2)]
    elif r < 0.50:
        text = text.replace(eol, eol + "garbage line" + eol, 1)
    elif r < 0.54:
        text = text + "é"
    elif r < 0.58:
        text = text.replace("rec0", "rec0\x0b", 1)
    return text


def _parse_outcome(records_module, fastq: bool, text: str):
    c = records_module.FASTAQRecordContainer() if fastq else records_module.FASTARecordContainer()
    names = ("identifier", "sequence", "space", "quality_sequence") if fastq else ("description", "genome")
    try:
        c.parse_records(text)
    except Exception as e:   # noqa: BLE001 -- the type and the message are what is compared
        return ("error", type(e).__name__, str(e))
    return ("ok", [(r.identifier,) + tuple(r[n] for n in names) for r in c])


@pytest.mark.parametrize("piece_bytes", ["1048576", "40"])
@pytest.mark.parametrize("fastq", [True, False])
def test_native_ingest_equals_the_regex_parser(fastq, piece_bytes, monkeypatch):
    import records
    import _native as nat
    monkeypatch.setenv("PA_INGEST_PIECE_BYTES", piece_bytes)      # "40": texts are cut into up to 16 parallel pieces
    rng = np.random.default_rng((77 if fastq else 78) + int(piece_bytes))
    n_native = 0
    ref_mod = None
    if os.path.isdir("/root/reference/src"):
        import importlib.util
        spec_c = importlib.util.spec_from_file_location("constants", "/root/reference/src/constants.py")
        saved = {k: sys.modules.get(k) for k in ("constants", "records")}
        try:
            mod_c = importlib.util.module_from_spec(spec_c); spec_c.loader.exec_module(mod_c)
            sys.modules["constants"] = mod_c
            spec_r = importlib.util.spec_from_file_location("ref_records", "/root/reference/src/records.py")
            ref_mod = importlib.util.module_from_spec(spec_r); spec_r.loader.exec_module(ref_mod)
        finally:
            for k, v in saved.items():
                if v is not None:
                    sys.modules[k] = v
    for trial in range(600):
        text = _random_text(rng, fastq)
        records.NATIVE_INGEST = True
        got = _parse_outcome(records, fastq, text)
        if text.isascii() and nat.parse_records_native(text.encode("ascii"), fastq) is not None:
            n_native += 1
        records.NATIVE_INGEST = False
        try:
            want = _parse_outcome(records, fastq, text)
        finally:
            records.NATIVE_INGEST = True
        assert got == want, (text, got, want)
        if ref_mod is not None:
            assert got == _parse_outcome(ref_mod, fastq, text), text
    assert 150 < n_native < 560      # both paths are exercised


def test_native_ingest_feeds_the_packed_arrays():
    import records
    c = records.FASTAQRecordContainer()
    c.parse_records("@r1 x\nACGT\n+..\nIIII\n@r2\nGG\n+\n!~\n")
    pk = c.packed_batch()
    assert pk is not None and pk["n"] == 2 and len(c) == 2
    assert pk["seq"].tobytes() == b"ACGTGG" and pk["qual"].tobytes() == b"IIII!~" and pk["off"].tolist() == [0, 4, 6]
    recs = list(c)                    # Records appear only now
    assert [(r.identifier, r["sequence"], r["space"], r["quality_sequence"]) for r in recs] == \
        [("r1 x", "ACGT", "..", "IIII"), ("r2", "GG", "", "!~")]
    assert c.packed_batch() is pk
    with pytest.raises(records.DuplicateRecordError):
        c.parse_records("@r2\nA\n+\nI\n")           # the unique index survives the lazy path
    f = records.FASTARecordContainer()
    f.parse_records(">g1 d\nACGT\nNNAC\r\n>g2\nTTTT")
    assert f.packed_batch()["seq"].tobytes() == b"ACGTNNACTTTT"
    assert [(r.identifier, r["genome"]) for r in f] == [("g1 d", "ACGTNNAC"), ("g2", "TTTT")]


def test_native_dumpref_writer_against_json_dumps():
    """format.cpp on a hand-made CSR (pure host code): byte-identical to json.dumps(indent=4) of the dictionary that
    kmer.py:300-329 builds, including descriptions that need escaping and genomes sharing a description."""
    import _native as nat
    rng = np.random.default_rng(21)
    for k in (1, 5, 31):
        n = 40 if k > 1 else 4
        kmers = sorted({"".join(rng.choice(list("ACGT"), size=k)) for _ in range(n)})
        flat = np.frombuffer("".join(kmers).encode(), dtype=np.uint8).copy()
        keys = np.zeros(len(kmers), dtype=np.uint64)
        nat.check(nat.lib().pa_encode_kmers(k, nat._p(flat), len(kmers), nat._p(keys)))
        srt = np.argsort(keys)
        keys, kmers = keys[srt], [kmers[i] for i in srt]
        descs = ['g "zero"', "tab\there", "g\\two", 'g "zero"', "ünï"]          # genomes 0 and 3 share a description
        cls = {}
        group = np.array([cls.setdefault(d, len(cls)) for d in descs], dtype=np.uint32)
        run_off, run_genome, pos_off, pos = [0], [], [0], []
        for _ in kmers:
            gs = sorted(rng.choice(len(descs), size=int(rng.integers(1, len(descs) + 1)), replace=False).tolist())
            for g in gs:
                run_genome.append(g)
                pos += sorted(rng.choice(1000, size=int(rng.integers(1, 4)), replace=False).tolist())
                pos_off.append(len(pos))
            run_off.append(len(run_genome))
        order = rng.permutation(len(kmers)).astype(np.uint32)
        csr = {"keys": keys, "order": order, "run_off": np.array(run_off, np.uint64), "run_genome": np.array(run_genome, np.uint32),
               "pos_off": np.array(pos_off, np.uint64), "pos": np.array(pos, np.uint32)}
        assert nat.decode_kmers(k, keys) == kmers
        want = {}
        for u in order.tolist():
            inner = {}
            for r in range(run_off[u], run_off[u + 1]):
                inner[descs[run_genome[r]]] = pos[pos_off[r]:pos_off[r + 1]]
            want[kmers[u]] = inner
        got = nat.format_kmers_json(k, csr, group, [json.dumps(d) for d in cls], indent=4, level=1)
        ref = json.dumps({"Kmers": want}, indent=4)
        assert "{\n    \"Kmers\": " + got + "\n}" == ref
    empty = {"keys": np.zeros(0, np.uint64), "order": np.zeros(0, np.uint32), "run_off": np.zeros(1, np.uint64),
             "run_genome": np.zeros(0, np.uint32), "pos_off": np.zeros(1, np.uint64), "pos": np.zeros(0, np.uint32)}
    assert nat.format_kmers_json(3, empty, np.zeros(1, np.uint32), [], indent=4, level=1) == "{}"


def _minimizer_restated(kmer):
    """The definition in DESIGN.md §3: m = min(k,16); the m-mer x at offset j is (high plane << m) | low plane with bit i =
    base j+i; hdrop = max(0, 2m-28); order(x) = xorshift-multiply-xorshift of x >> hdrop on 2m-hdrop bits; the minimizer is the
    leftmost m-mer of smallest order and its hash is [low hdrop bits of x][order * 0x9E3779B1 mod 2^(2m-hdrop)] (the
    multiplication spreads the orders of minimizers, which crowd near zero; the raw bits go on top so that m-mers of equal
    order land in different table blocks)."""
    k = len(kmer)
    m = min(k, 16)
    code = {"A": 0, "C": 1, "T": 2, "G": 3}
    hdrop = max(0, 2 * m - 28)
    ybits = 2 * m - hdrop
    ymask, yshift = (1 << ybits) - 1, (ybits + 1) // 2
    best = None
    for j in range(k - m + 1):
        lo = sum((code[c] & 1) << i for i, c in enumerate(kmer[j:j + m]))
        hi = sum((code[c] >> 1) << i for i, c in enumerate(kmer[j:j + m]))
        x = (hi << m) | lo
        y = x >> hdrop
        y ^= y >> yshift
        y = (y * 0x7FEB352D) & ymask
        y ^= y >> yshift
        if best is None or y < best[0]:
            best = (y, j, x)
    y, j, x = best
    return ((x & ((1 << hdrop) - 1)) << ybits) | (((y * 0x9E3779B1) & 0xFFFFFFFF) & ymask), j


def test_minimizer_hash_is_a_bijection_and_matches_its_definition():
    """Block, bucket and tag identify a k-mer only if the m-mer hash is a bijection: exhaustive for m <= 8, and for m = 16 on
    m-mers that differ only in the bits the hash keeps raw / only in the mixed bits.  The minimizer (hash, offset) of random
    k-mers equals the restated definition, ties included (low-complexity k-mers)."""
    import itertools
    import _native as nat
    L = nat.lib()

    def native(kmers, k):
        flat = np.frombuffer("".join(kmers).encode(), dtype=np.uint8).copy()
        mh = np.zeros(len(kmers), np.uint32); off = np.zeros(len(kmers), np.uint32)
        assert L.pa_debug_minimizer(k, nat._p(flat), len(kmers), nat._p(mh), nat._p(off)) == 0
        return mh, off

    for k in (1, 2, 5, 8):   # k <= 16: the k-mer is its own m-mer
        kmers = ["".join(p) for p in itertools.product("ACGT", repeat=k)]
        mh, off = native(kmers, k)
        assert len(set(mh.tolist())) == 4 ** k and int(mh.max()) < 4 ** k and not off.any()
    rng = np.random.default_rng(5)
    base = "".join(rng.choice(list("ACGT"), size=16))
    variants = {base}
    for _ in range(4000):
        s = list(base)
        for pos in rng.integers(0, 16, size=int(rng.integers(1, 4))):
            s[int(pos)] = "ACGT"[int(rng.integers(0, 4))]
        variants.add("".join(s))
    variants = sorted(variants)
    mh, _ = native(variants, 16)
    assert len(set(mh.tolist())) == len(variants)
    for k in (3, 15, 16, 17, 24, 31):
        kmers = ["".join(rng.choice(list("ACGT"), size=k)) for _ in range(300)]
        kmers += ["A" * k, "AC" * (k // 2) + "A" * (k % 2), "ACG" * (k // 3) + "T" * (k % 3), "T" * (k - 1) + "G"]
        mh, off = native(kmers, k)
        for s, h, o in zip(kmers, mh, off):
            assert (int(h), int(o)) == _minimizer_restated(s), (k, s)


def test_host_pool_serves_concurrent_callers():
    """ctypes releases the GIL, so two Python threads (one per GPU, say) may pack and parse at the same time: the worker
    pool must run one job at a time and every caller must get its own complete result (Pool::parallel_for)."""
    import ctypes
    import threading
    import _native as nat
    L = nat.lib()
    rng = np.random.default_rng(21)
    jobs = []
    for t in range(6):
        n = 20_000 + 1000 * t
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(150))
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * 150)].copy()
        want, ok = nat.pack_reads(bases, off)      # single-threaded caller: the reference result
        assert ok
        jobs.append((bases, off, want))
    text = "".join(f"@r{i}\nACGTACGTAC\n+\nIIIIIIIIII\n" for i in range(30_000))
    errors = []

    def pack_worker(bases, off, want):
        for _ in range(6):
            got, ok = nat.pack_reads(bases, off)
            if not ok or not np.array_equal(got, want):
                errors.append("pack")

    def parse_worker():
        for _ in range(6):
            pk = nat.parse_records_native(text, True)
            if pk is None or pk["n"] != 30_000 or pk["seq"].size != 300_000:
                errors.append("parse")

    threads = [threading.Thread(target=pack_worker, args=j) for j in jobs] + [threading.Thread(target=parse_worker) for _ in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors



def test_committed_bench_lines_keep_the_driver_contract():
    """The bench lines committed under profiles/ (written by bench.py on the B200 boxes) carry every key the driver and the
    judge read: metric / value / unit, e2e with its byte counts, roofline with the measured traffic, the CPU baseline at
    N = 1, clocks, launches, and the sub-records of configs[2] / [3] / [4]."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, n in (("r02_bench_1gpu_v3.json", 1), ("r02_bench_4gpu_v3.json", 4), ("r02_bench_8gpu_v3.json", 8)):
        line = open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1]
        d = json.loads(line)
        assert d["metric"].startswith("reads/s") and d["unit"] == "reads/s" and d["higher_is_better"] is True
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "u64"
        assert d["value"] > 0 and abs(d["value"] - n * d["config"]["reads_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        assert d["config"]["workload"].startswith("configs[1]") and d["vs_baseline"] is None
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] and r["traffic"] > r["algorithmic_bytes_per_read"] * d["config"]["reads_per_gpu"]
        assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0 and not d["clocks"]["reasons"]
        assert d["parity_fullsize_digest"]["equals_oracle_digest"] is True
        if n == 1:
            c = d["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and d["parity"].startswith("bit-exact")
        else:
            bp = d["build_partitioned"]
            assert bp["checksum_equals_single_gpu_build"] is True and bp["alignment_through_replica_equals_single_gpu_index"] is True
        cfg = d["configs"]
        assert cfg["extquality"]["parity_digest"]["equals_oracle_digest"] is True
        assert cfg["config_e"]["parity"]["equal_to_oracle"] is True and cfg["config_e"]["index"]["genomes"] == 2000
        assert cfg["extsim"]["value"] > 0 and cfg["extsim"]["genomes_kept"] == 100
