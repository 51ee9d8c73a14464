"""The C oracle against the committed golden vectors (generated from the Python reference by
tests/golden/make_golden.py).  CPU only; runs on any box."""
import json

import numpy as np
import pytest

import goldencheck
import synth
from oracle.oracle import OracleReference


def test_kat_reference_fixtures():
    for item in goldencheck.load_golden("kat_reference_tests.json"):
        got = goldencheck.canonical_from_oracle(item["case"])
        goldencheck.assert_matches(got, item["expect"], item["case"]["name"])


def test_kat_values_quoted_by_the_reference_tests():
    """Counters asserted literally in /root/reference/src/test_kmer.py:523-545."""
    items = {i["case"]["name"]: i for i in goldencheck.load_golden("kat_reference_tests.json")}
    s = json.loads(items["four_genomes_k3_combined_40_50_2"]["expect"]["align_summary_json"])["Statistics"]
    assert (s["filtered_quality_reads"], s["filtered_quality_kmers"], s["filtered_hr_kmers"]) == (1, 1, 5)
    s = json.loads(items["four_genomes_k3_combined_30_30_3"]["expect"]["align_summary_json"])["Statistics"]
    assert (s["unique_mapped_reads"], s["ambiguous_mapped_reads"], s["unmapped_reads"]) == (0, 2, 1)
    assert (s["filtered_quality_reads"], s["filtered_quality_kmers"], s["filtered_hr_kmers"]) == (0, 0, 0)
    flip = items["flip_p1"]["expect"]["reads"][0]
    assert flip == ["Read4", "AMBIGUOUSLY_MAPPED", ["Genome1", "Genome1", "Genome2"]]
    assert items["flip_p5"]["expect"]["reads"][0][1] == "UNIQUELY_MAPPED"


def test_fuzz_small_k_golden():
    for item in goldencheck.load_golden("fuzz_small_k.json"):
        got = goldencheck.canonical_from_oracle(item["case"])
        goldencheck.assert_matches(got, item["expect"], f"seed {item['case']['seed']}")


@pytest.mark.parametrize("nthreads", [1, 4])
def test_config_a_golden(nthreads):
    gold = goldencheck.load_golden("config_a.json")
    genomes = synth.make_genomes(3, 50_000, seed=1234, cluster_size=3, shared_frac=0.3, sub_rate=0.01,
                                 n_every=20_000, n_run=40)
    b, q, off = synth.make_reads(genomes, 10_000, 100, seed=4321, sub_rate=0.01, random_frac=0.02)
    import hashlib
    inp = hashlib.sha256(b"".join(g.tobytes() for g in genomes) + b.tobytes() + q.tobytes()).hexdigest()
    assert inp == gold["input_sha256"], "synthetic generator drifted from the one that produced the fixture"
    o = OracleReference(31, synth.genomes_as_pairs(genomes))
    ids = [f"read{i}" for i in range(10_000)]
    for run in gold["runs"]:
        pr = run["params"]
        al = o.align_packed(ids, b, q, off, pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"], nthreads=nthreads)
        reads = [[rid, d["mapping_type"], d["genomes_mapped_to"]] for rid, d in al.reads().items()]
        assert o.sizes()[0] == run["n_distinct_kmers"]
        assert json.dumps(al.get_summary()) == run["align_summary_json"], run["name"]
        assert len(reads) == run["n_stored_reads"]
        assert reads[:300] == run["reads_sample"]
        assert goldencheck.digest_reads(reads) == run["reads_digest"], run["name"]
