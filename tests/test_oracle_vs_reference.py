"""
Pins the C oracle (oracle/pa_oracle.c) against the UNMODIFIED Python reference
imported from /root/reference/src.  Runs only where the reference exists (this
container); skipped on the GPU box.  The committed fixtures in tests/golden/
carry the same pin to places the reference cannot travel to.
"""
import json

import pytest

import refimpl
import synth
from oracle.oracle import OracleReference

pytestmark = pytest.mark.skipif(not refimpl.reference_available(), reason="/root/reference not present")


def oracle_run(case):
    pr = case["params"]
    o = OracleReference(case["k"], case["genomes"], filter_similar=pr["filter_similar"],
                        similarity_threshold=pr["threshold"])
    al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
    return o, al


def compare(case):
    ref = refimpl.load_reference()
    pr = case["params"]
    want = refimpl.ref_run(ref, case["k"], case["genomes"], case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"],
                           pr["mg"], pr["filter_similar"], pr["threshold"])
    o, al = oracle_run(case)
    assert [g[0] for g in o.genomes] == want["genomes"]
    got_kmers = o.kmers_dict()
    assert list(got_kmers.keys()) == list(want["kmers"].keys())
    assert got_kmers == want["kmers"]
    assert [list(v.keys()) for v in got_kmers.values()] == [list(v.keys()) for v in want["kmers"].values()]
    assert json.dumps(o.get_summary()) == json.dumps(want["ref_summary"])
    assert o.similarity_info == want["similarity_info"]
    if want["similarity_info"] is not None:
        assert json.dumps(o.similarity_info) == json.dumps(want["similarity_info"])
    assert al.reads() == want["reads"]
    assert list(al.reads().keys()) == list(want["reads"].keys())
    assert json.dumps(al.get_summary()) == json.dumps(want["align_summary"])


@pytest.mark.parametrize("block", range(8))
def test_fuzz_small_k(block):
    for seed in range(block * 150, block * 150 + 150):
        compare(synth.fuzz_case(seed))


@pytest.mark.parametrize("block", range(2))
def test_fuzz_duplicate_identifiers(block):
    for seed in range(10_000 + block * 100, 10_000 + block * 100 + 100):
        compare(synth.fuzz_case(seed, dup_ids=True))


def test_k31_config_a_scaled():
    genomes = synth.make_genomes(3, 6000, seed=11, cluster_size=3, shared_frac=0.4, n_every=2500, n_run=7)
    b, q, off = synth.make_reads(genomes, 400, 100, seed=12, sub_rate=0.02, random_frac=0.05)
    case = {"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
            "params": {"m": 1, "p": 1, "mrq": None, "mkq": None, "mg": None, "filter_similar": False, "threshold": 0.95}}
    compare(case)
    case["params"].update(mrq=62, mkq=60, mg=1)
    compare(case)
    case["params"].update(filter_similar=True, threshold=0.2, mrq=None)
    compare(case)


def test_degenerate_k():
    for k in (0, -1, 1, 31):
        case = {"k": k, "genomes": [("a", "ACGTNACGT"), ("b", "AC")], "reads": [("r", "ACGTA", "IIIII")],
                "params": {"m": 1, "p": 1, "mrq": None, "mkq": None, "mg": None, "filter_similar": True, "threshold": 0.5}}
        compare(case)
