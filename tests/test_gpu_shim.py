"""
GPU parity tests through the reference-facing Python API (kmer.KmerReference / Read / PseudoAlignment):
committed golden vectors (generated from the Python reference), the oracle on seeded inputs, and the
behaviours the reference's own tests pin (/root/reference/src/test_kmer.py, restated here, not copied).
"""
import gzip
import hashlib
import json
import pickle
import random

import numpy as np
import pytest

import goldencheck
import synth
import kmer
from kmer import AddingExistingRead, KmerReference, PseudoAlignment, Read, ReadMapping, ReadMappingType
from records import FASTAQRecordContainer, FASTARecordContainer, Record, Section

pytestmark = pytest.mark.gpu


def fasta_records(genomes):
    return [Record([Section("description", gid), Section("genome", seq)]) for gid, seq in genomes]


def fastq_records(reads):
    return [Record([Section("identifier", rid), Section("sequence", s), Section("space", ""), Section("quality_sequence", q)])
            for rid, s, q in reads]


def canonical_from_product(case):
    """Same canonical shape as goldencheck.canonical_from_oracle, computed by the product."""
    pr = case["params"]
    ref = KmerReference(case["k"], fasta_records([tuple(g) for g in case["genomes"]]),
                        filter_similar=pr.get("filter_similar", False), similarity_threshold=pr.get("threshold", 0.95))
    if case.get("seed", 1) % 2 == 1:
        # every other case goes through the on-disk form: the unpickled object rebuilds its device index from the
        # genomes (and re-applies the EXTSIM drop), and must be indistinguishable -- dict insertion order included
        ref = pickle.loads(pickle.dumps(ref))
    index_of = {id(r): i for i, r in enumerate(ref.genomes)}
    # the native dumpref writer must print exactly what json.dumps prints for the dictionary version
    assert ref.summary_json(indent=4) == json.dumps(ref.get_summary(), indent=4)
    out = {"genomes": [g.identifier for g in ref.genomes],
           "kmers": [[km, [[index_of[id(r)], sorted(pos)] for r, pos in inner.items()]] for km, inner in ref.kmers.items()],
           "ref_summary_json": json.dumps(ref.get_summary())}
    if hasattr(ref, "similarity_info"):
        out["similarity_info_json"] = json.dumps(ref.similarity_info)
    if case.get("reads") is not None:
        pa = PseudoAlignment(ref)
        pa.align_reads_from_container(fastq_records([tuple(r) for r in case["reads"]]), pr.get("m", 1), pr.get("p", 1),
                                      pr.get("mrq"), pr.get("mkq"), pr.get("mg"))
        out["reads"] = [[rid, d["mapping_type"].name, list(d["genomes_mapped_to"])] for rid, d in pa.reads.items()]
        out["align_summary_json"] = json.dumps(pa.get_summary())
    return out


# ---------------------------------------------------------------------------
# golden vectors produced by the Python reference itself
# ---------------------------------------------------------------------------
def test_golden_kat_reference_fixtures():
    for item in goldencheck.load_golden("kat_reference_tests.json"):
        goldencheck.assert_matches(canonical_from_product(item["case"]), item["expect"], item["case"]["name"])


def test_golden_fuzz_small_k():
    for item in goldencheck.load_golden("fuzz_small_k.json"):
        goldencheck.assert_matches(canonical_from_product(item["case"]), item["expect"], f"seed {item['case']['seed']}")


def test_golden_config_a():
    """BASELINE.json configs[0]: 3 x 50 kb genomes, 10,000 x 100 bp reads, k = 31."""
    gold = goldencheck.load_golden("config_a.json")
    genomes = synth.make_genomes(3, 50_000, seed=1234, cluster_size=3, shared_frac=0.3, sub_rate=0.01, n_every=20_000, n_run=40)
    b, q, off = synth.make_reads(genomes, 10_000, 100, seed=4321, sub_rate=0.01, random_frac=0.02)
    inp = hashlib.sha256(b"".join(g.tobytes() for g in genomes) + b.tobytes() + q.tobytes()).hexdigest()
    assert inp == gold["input_sha256"]
    ref = KmerReference(31, fasta_records(synth.genomes_as_pairs(genomes)))
    reads = fastq_records(synth.reads_as_triples(b, q, off))
    for run in gold["runs"]:
        pr = run["params"]
        pa = PseudoAlignment(ref)
        pa.align_reads_from_container(reads, pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        assert len(ref.kmers) == run["n_distinct_kmers"]
        assert json.dumps(pa.get_summary()) == run["align_summary_json"], run["name"]
        got = [[rid, d["mapping_type"].name, list(d["genomes_mapped_to"])] for rid, d in pa.reads.items()]
        assert len(got) == run["n_stored_reads"]
        assert got[:300] == run["reads_sample"]
        assert goldencheck.digest_reads(got) == run["reads_digest"], run["name"]


# ---------------------------------------------------------------------------
# seeded differential tests against the oracle (cases the fixtures do not hold)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("block", range(4))
def test_fuzz_against_oracle_with_extsim(block):
    for seed in range(300_000 + block * 60, 300_000 + block * 60 + 60):
        case = synth.fuzz_case(seed, dup_ids=(seed % 3 == 0))
        goldencheck.assert_matches(canonical_from_product(case), goldencheck.canonical_from_oracle(case), f"seed {seed}")


def test_extsim_clusters_k31_against_oracle():
    genomes = synth.make_genomes(12, 20_000, seed=21, cluster_size=4, shared_frac=0.9, sub_rate=0.002, n_every=0, n_run=0)
    b, q, off = synth.make_reads(genomes, 1500, 150, seed=22, sub_rate=0.01, random_frac=0.03)
    for thr in (0.5, 0.95):
        case = {"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                "params": {"m": 1, "p": 1, "mrq": None, "mkq": None, "mg": None, "filter_similar": True, "threshold": thr}}
        got = canonical_from_product(case)
        got.pop("kmers"); got.pop("ref_summary_json")          # O(index) text: compared on the small cases
        want = goldencheck.canonical_from_oracle(case)
        want.pop("kmers"); want.pop("ref_summary_json")
        goldencheck.assert_matches(got, want, f"thr {thr}")


# ---------------------------------------------------------------------------
# behaviours pinned by the reference's own test_kmer.py
# ---------------------------------------------------------------------------
FOUR = (">Genome1\nAGCTAGCTAGCTAGCTAGCT\n>Genome2\nTGCATGCATGCATGCATGCA\n"
        ">Genome3\nAGCTTGCATGCAGCTAGCTA\n>Genome4\nCCGGAAGCTTGCATGCAGCTA\n")
THREE_READS = "@Read1\nAGCTAGCT\n+\nIIIIIIII\n@Read2\nTGCATGCA\n+\n!!!!!!!!\n@Read3\nGGGGGGGG\n+\n!!IIIIII\n"


def parse_fasta(text):
    c = FASTARecordContainer()
    c.parse_records(text)
    return c


def parse_fastq(text):
    c = FASTAQRecordContainer()
    c.parse_records(text)
    return c


def one_read(text):
    return Read(list(parse_fastq(text))[0])


def test_extract_kmers_from_genome_kat():
    got = list(kmer.extract_kmers_from_genome(3, "AGCTAGCTAGCT"))
    assert got == [(i, "AGCTAGCTAGCT"[i:i + 3]) for i in range(10)]
    assert list(kmer.extract_kmers_from_genome(0, "ACGT")) == [] and list(kmer.extract_kmers_from_genome(5, "ACGT")) == []


def test_reference_building_presence_and_absence():
    ref = KmerReference(3, parse_fasta(FOUR))
    for present in ("AGC", "TGC", "GCT", "CCG"):
        assert ref.get_kmer_references(present)
    assert not ref.get_kmer_references("GGG")
    assert ref["GGG"] is None and ref["AGC"] is not None
    inner = ref.get_kmer_references("AGC")
    assert all(isinstance(r, Record) for r in inner) and all(isinstance(p, set) for p in inner.values())


def test_mapping_classes():
    unmapped = KmerReference(4, parse_fasta(">Genome1\nAACCGGTTAACC\n>Genome2\nGGTTCCAAGGTT\n"))
    assert one_read("@Read1\nTAGGCAT\n+\nIIIIIII\n").pseudo_align(unmapped) == ReadMappingType.UNMAPPED
    unique = KmerReference(4, parse_fasta(">Genome1\nATGGCTATGCTA\n>Genome2\nCTATGGCAGGCA\n"))
    assert one_read("@Read2\nATGGCTAT\n+\nIIIIIIII\n").pseudo_align(unique) == ReadMappingType.UNIQUELY_MAPPED
    amb = KmerReference(4, parse_fasta(">Genome1\nATCGACGGTCGTTA\n>Genome2\nCGATGATCAGTACGA\n"
                                       ">Genome3\nATCCACCTAACGTACGGT\n>Genome4\nCTAGGGACTGCACTA\n"))
    r = one_read("@Read3\nATCGATCCTAG\n+\nIIIIIIIIIII\n")
    assert r.pseudo_align(amb) == ReadMappingType.AMBIGUOUSLY_MAPPED
    assert len(r.mapping.genomes_mapped_to) == 4 and all(isinstance(g, Record) for g in r.mapping.genomes_mapped_to)


def test_unique_flips_to_ambiguous_under_default_p_only():
    ref = KmerReference(4, parse_fasta(">Genome1\nATGCCTTTTCGGGG\n>Genome2\nGCCGTTTTCGGGGCTA\n>Genome3\nCCGG\n"
                                       ">Genome4\nAAAAAAAAGGGCT\n>Genome5\nTTTTTTTTGCTAA\n"))
    text = "@Read4\nATGCCGGGGCTAA\n+\nIIIIIIIIIIIII\n"
    r = one_read(text)
    assert r.pseudo_align(ref) == ReadMappingType.AMBIGUOUSLY_MAPPED
    assert [g.identifier for g in r.mapping.genomes_mapped_to] == ["Genome1", "Genome1", "Genome2"]
    assert one_read(text).pseudo_align(ref, p=5) == ReadMappingType.UNIQUELY_MAPPED
    assert one_read(text).pseudo_align(ref, p=-1) == ReadMappingType.UNIQUELY_MAPPED


def test_argument_checks():
    ref = KmerReference(4, parse_fasta(">Genome1\nATGGCTATGCTA\n"))
    with pytest.raises(TypeError):
        one_read("@r\nATGGC\n+\nIIIII\n").pseudo_align(ref, m=1.5)
    with pytest.raises(TypeError):
        one_read("@r\nATGGC\n+\nIIIII\n").pseudo_align("not a reference")
    with pytest.raises(ValueError):
        one_read("@r\nATGGC\n+\nIIIII\n").pseudo_align(ref, m=-1)
    with pytest.raises(ValueError):
        KmerReference(4, parse_fasta(">Genome1\nATGGCTATGCTA\n"), filter_similar=True, similarity_threshold=1.5)
    with pytest.raises(ValueError):
        KmerReference(32, parse_fasta(">Genome1\n" + "ACGT" * 20 + "\n"))


@pytest.mark.parametrize("execution_number", range(10))
def test_random_31mers_scattered_over_four_genomes(execution_number):
    """Differential property of test_kmer.py:364-421: scatter the 33 31-mers of a fixed read over 4 genomes joined by
    NN, recompute specific / total counts independently, check the decision for m = p = 1."""
    rnd = random.Random(1000 + execution_number)
    seq = "AGCTAGCTAGAGGTCCTAATCCTAGCTAGCTAGCTAGCTAGCTAGCTGGTCATCAAAACCTTT"
    read = one_read(f"@BigRead\n{seq}\n+\n{'I' * len(seq)}\n")
    k = 31
    kmers = list(kmer.extract_kmers_from_genome(k, read._Read__raw_read))
    members = {f"Genome{i + 1}": [] for i in range(4)}
    for _, km in kmers:
        for name in rnd.sample(list(members), k=rnd.randint(1, 4)):
            members[name].append(km)
    fasta = "".join(f">{name}\n{'NN'.join(parts)}\n" for name, parts in members.items() if parts)
    ref = KmerReference(k, parse_fasta(fasta))
    specific, total = {}, {}
    for km, genome_map in ref.kmers.items():
        owners = list(genome_map.keys())
        if len(owners) == 1:
            specific[owners[0]] = specific.get(owners[0], 0) + len(genome_map[owners[0]])
        for genome, positions in genome_map.items():
            total[genome] = total.get(genome, 0) + len(positions)
    result = read.pseudo_align(ref, m=1, p=1)
    ranked = sorted(specific, key=specific.get, reverse=True)
    if not ranked:
        assert result == ReadMappingType.AMBIGUOUSLY_MAPPED
        return
    runner_up = specific[ranked[1]] if len(ranked) > 1 else 0
    m_ok = specific[ranked[0]] - runner_up >= 1
    p_ok = not (max(total.values()) - total[ranked[0]] > 1)
    assert result == (ReadMappingType.UNIQUELY_MAPPED if (m_ok and p_ok) else ReadMappingType.AMBIGUOUSLY_MAPPED)


def make_read(identifier, sequence, quality, genomes=None, kind="UNMAPPED"):
    r = one_read(f"@{identifier}\n{sequence}\n+\n{quality}\n")
    if genomes is not None:
        r.mapping = ReadMapping(ReadMappingType[kind], genomes)
    return r


def test_pseudo_alignment_add_summary_and_pickle(tmp_path):
    ref = KmerReference(3, parse_fasta(FOUR))
    pa = PseudoAlignment(ref)
    g1 = list(parse_fasta(">Genome1\nAGCTAGCTAG\n"))[0]
    g2 = list(parse_fasta(">Genome2\nAGCTAGCTAG\n"))[0]
    r1 = make_read("read1", "AGCTAGCT", "IIIIIIII", [g1], "UNIQUELY_MAPPED")
    pa.add_read(r1)
    assert "read1" in pa.reads
    assert pa.reads["read1"]["mapping_type"] == ReadMappingType.UNIQUELY_MAPPED
    assert pa.reads["read1"]["genomes_mapped_to"] == ["Genome1"]
    with pytest.raises(AddingExistingRead):
        pa.add_read(r1)
    pa.add_read(make_read("read2", "TGCATGCA", "IIIIIIII", [g1, g2], "AMBIGUOUSLY_MAPPED"))
    pa.add_read(make_read("read3", "GGGGGGGG", "IIIIIIII", [], "UNMAPPED"))
    s = pa.get_summary()
    assert s["Statistics"] == {"unique_mapped_reads": 1, "ambiguous_mapped_reads": 1, "unmapped_reads": 1}
    assert s["Summary"] == {"Genome1": {"unique_reads": 1, "ambiguous_reads": 1}, "Genome2": {"unique_reads": 0, "ambiguous_reads": 1}}
    path = tmp_path / "t.aln"
    pa.save(str(path))
    with gzip.open(str(path), "rb") as f:
        loaded = pickle.load(f)
    assert "read1" in loaded.reads and loaded.reads["read1"]["genomes_mapped_to"] == ["Genome1"]
    assert PseudoAlignment.load(str(path)).get_summary() == s
    assert pa.get_reads_by_mapping_type(ReadMappingType.UNMAPPED) == ["read3"]


def test_extquality_counters_kat():
    """Counters asserted literally by test_kmer.py:523-545."""
    ref = KmerReference(3, parse_fasta(FOUR))
    pa = PseudoAlignment(ref)
    pa.align_reads_from_container(parse_fastq(THREE_READS), min_read_quality=40, min_kmer_quality=50, max_genomes=2)
    st = pa.get_summary()["Statistics"]
    assert (st["filtered_quality_reads"], st["filtered_quality_kmers"], st["filtered_hr_kmers"]) == (1, 1, 5)
    pa = PseudoAlignment(ref)
    pa.align_reads_from_container(parse_fastq(THREE_READS), min_read_quality=30, min_kmer_quality=30, max_genomes=3)
    st = pa.get_summary()["Statistics"]
    assert st == {"unique_mapped_reads": 0, "ambiguous_mapped_reads": 2, "unmapped_reads": 1,
                  "filtered_quality_reads": 0, "filtered_quality_kmers": 0, "filtered_hr_kmers": 0}
    pa = PseudoAlignment(ref)
    pa.align_reads_from_container(parse_fastq(THREE_READS))
    assert set(pa.get_summary()["Statistics"]) == {"unique_mapped_reads", "ambiguous_mapped_reads", "unmapped_reads"}


def test_duplicate_read_identifier_in_a_batch():
    ref = KmerReference(3, parse_fasta(FOUR))
    pa = PseudoAlignment(ref)
    pa.align_reads_from_container(parse_fastq(THREE_READS))
    with pytest.raises(AddingExistingRead):
        pa.align_reads_from_container(parse_fastq(THREE_READS))
    assert len(pa.reads) == 3


def test_extsim_identical_genomes():
    text = ">GenomeA\nAGCTAGCTAGCT\n>GenomeB\nAGCTAGCTAGCT\n>GenomeC\nTGCATGCATGCA\n"
    ref = KmerReference(4, parse_fasta(text), filter_similar=True, similarity_threshold=0.95)
    assert [g.identifier for g in ref.genomes] == ["GenomeA", "GenomeC"]
    assert ref.similarity_info["GenomeB"]["kept"] == "no" and ref.similarity_info["GenomeB"]["similar_to"] == "GenomeA"
    assert ref.similarity_info["GenomeB"]["similarity_score"] == 1.0
    assert all(ref.similarity_info[g]["similar_to"] == "NA" for g in ("GenomeA", "GenomeC"))
    assert "Similarity" in ref.get_summary()
    plain = KmerReference(4, parse_fasta(text), filter_similar=False)
    assert not hasattr(plain, "similarity_info") and "Similarity" not in plain.get_summary()
    single = KmerReference(4, parse_fasta(">GenomeA\nAGCTAGCTAGCT\n"), filter_similar=True)
    assert single.get_summary()["Similarity"]["GenomeA"]["kept"] == "yes"


def test_reference_pickle_round_trip(tmp_path):
    ref = KmerReference(3, parse_fasta(FOUR), filter_similar=True, similarity_threshold=0.5)
    path = tmp_path / "r.kdb"
    ref.save(str(path))
    with gzip.open(str(path), "rb") as f:
        loaded = pickle.load(f)
    assert [g.identifier for g in loaded.genomes] == [g.identifier for g in ref.genomes]
    assert loaded.similarity_info == ref.similarity_info
    assert json.dumps(loaded.get_summary()) == json.dumps(ref.get_summary())
    pa1, pa2 = PseudoAlignment(ref), PseudoAlignment(KmerReference.load(str(path)))
    for pa in (pa1, pa2):
        pa.align_reads_from_container(parse_fastq(THREE_READS), 1, 1, None, 40, 2)
    assert json.dumps(pa1.get_summary()) == json.dumps(pa2.get_summary())
    assert dict(pa1.reads.items()) == dict(pa2.reads.items())


def test_degenerate_k_and_empty_inputs():
    for k in (0, -2):
        ref = KmerReference(k, parse_fasta(FOUR))
        assert len(ref.kmers) == 0 and ref.get_summary() == {"Kmers": {}, "Summary": {}}
        pa = PseudoAlignment(ref)
        pa.align_reads_from_container(parse_fastq(THREE_READS), min_read_quality=40)
        assert pa.get_summary()["Statistics"] == {"unique_mapped_reads": 0, "ambiguous_mapped_reads": 0, "unmapped_reads": 2,
                                                  "filtered_quality_reads": 1}
    short = KmerReference(31, parse_fasta(">a\nACGT\n>b\nNNNN\n"))
    assert len(short.kmers) == 0
    pa = PseudoAlignment(short)
    pa.align_reads_from_container(parse_fastq(THREE_READS))
    assert pa.get_summary()["Statistics"]["unmapped_reads"] == 3
