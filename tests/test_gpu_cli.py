"""
CLI integration on the GPU: the four tasks of main.py (reference / dumpref / align / dumpalign) produce the same files
and JSON as the reference CLI (/root/reference/src/main.py:317-402; behaviours pinned by test_main.py:76-171, 246-279).
Expected JSON comes from the oracle, which is pinned to the Python reference.
"""
import gzip
import json
import os
import pickle
import subprocess
import sys

import pytest

import conftest
import synth
from oracle.oracle import OracleReference

pytestmark = pytest.mark.gpu
PKG = conftest.PKG_DIR


def cli(*argv):
    r = subprocess.run([sys.executable, "main.py", *map(str, argv)], capture_output=True, text=True, cwd=PKG)
    return r.stdout, r.stderr, r.returncode


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    genomes = synth.make_genomes(5, 3000, seed=31, cluster_size=5, shared_frac=0.6, sub_rate=0.004, n_every=1200, n_run=6)
    b, q, off = synth.make_reads(genomes, 300, 80, seed=32, sub_rate=0.01, random_frac=0.05)
    pairs, triples = synth.genomes_as_pairs(genomes), synth.reads_as_triples(b, q, off)
    synth.write_fasta(str(d / "g.fa"), pairs, width=70)
    synth.write_fastq(str(d / "r.fq"), triples)
    with gzip.open(d / "g2.fa.gz", "wt") as f:
        f.write(open(d / "g.fa").read())
    return {"dir": d, "pairs": pairs, "triples": triples}


def test_reference_then_dumpref(data):
    d = data["dir"]
    out, err, rc = cli("-t", "reference", "-g", d / "g.fa", "-k", 15, "-r", d / "ref.kdb")
    assert rc == 0, err
    with gzip.open(d / "ref.kdb", "rb") as f:   # test_main.py:91-93: the .kdb is a gzip pickle
        sys.path.insert(0, PKG)
        obj = pickle.load(f)
    assert [g.identifier for g in obj.genomes] == [p[0] for p in data["pairs"]]
    out, err, rc = cli("-t", "dumpref", "-r", d / "ref.kdb")
    assert rc == 0, err
    want = json.dumps(OracleReference(15, data["pairs"]).get_summary(), indent=4)
    assert out.rstrip("\n") == want
    out2, err, rc = cli("-t", "dumpref", "-g", d / "g2.fa.gz", "-k", 15)
    assert rc == 0 and out2 == out


def test_align_and_dumpalign_variants(data):
    d = data["dir"]
    o = OracleReference(15, data["pairs"])
    want_plain = json.dumps(o.align(data["triples"], 1, 1).get_summary(), indent=4)
    assert cli("-t", "reference", "-g", d / "g.fa", "-k", 15, "-r", d / "ref2.kdb")[2] == 0
    out, err, rc = cli("-t", "dumpalign", "-r", d / "ref2.kdb", "--reads", d / "r.fq")
    assert rc == 0, err
    assert out.rstrip("\n") == want_plain
    out, err, rc = cli("-t", "align", "-r", d / "ref2.kdb", "--reads", d / "r.fq", "-a", d / "a.aln")
    assert rc == 0 and (d / "a.aln").exists(), err
    out, err, rc = cli("-t", "dumpalign", "-a", d / "a.aln")
    assert rc == 0 and out.rstrip("\n") == want_plain
    out, err, rc = cli("-t", "dumpalign", "-g", d / "g.fa", "-k", 15, "--reads", d / "r.fq", "-m", 2, "-p", 3,
                       "--min-read-quality", 62, "--min-kmer-quality", 61, "--max-genomes", 2)
    assert rc == 0, err
    want = json.dumps(o.align(data["triples"], 2, 3, 62, 61, 2).get_summary(), indent=4)
    assert out.rstrip("\n") == want
    # "-m 0" falls back to the default 1 (main.py:337-342)
    out, err, rc = cli("-t", "dumpalign", "-r", d / "ref2.kdb", "--reads", d / "r.fq", "-m", 0)
    assert rc == 0 and out.rstrip("\n") == want_plain
    # align from FASTA without -r: the reference crashes (main.py:372); here it just works (see main.py docstring)
    out, err, rc = cli("-t", "align", "-g", d / "g.fa", "-k", 15, "--reads", d / "r.fq", "-a", d / "b.aln")
    assert rc == 0, err
    assert cli("-t", "dumpalign", "-a", d / "b.aln")[0].rstrip("\n") == want_plain


def test_extsim_through_the_cli(data):
    d = data["dir"]
    out, err, rc = cli("-t", "reference", "-g", d / "g.fa", "-k", 15, "-r", d / "sim.kdb", "--filter-similar",
                       "--similarity-threshold", 0.5)
    assert rc == 0, err
    o = OracleReference(15, data["pairs"], filter_similar=True, similarity_threshold=0.5)
    assert any(v["kept"] == "no" for v in o.similarity_info.values())
    with gzip.open(d / "sim.kdb", "rb") as f:
        sys.path.insert(0, PKG)
        obj = pickle.load(f)
    assert obj.similarity_info == o.similarity_info                     # test_main.py:270-274
    assert [g.identifier for g in obj.genomes] == [g[0] for g in o.genomes]
    out, err, rc = cli("-t", "dumpref", "-r", d / "sim.kdb")
    assert rc == 0 and out.rstrip("\n") == json.dumps(o.get_summary(), indent=4)
    out, err, rc = cli("-t", "dumpalign", "-r", d / "sim.kdb", "--reads", d / "r.fq")
    assert rc == 0 and out.rstrip("\n") == json.dumps(o.align(data["triples"], 1, 1).get_summary(), indent=4)
    out, err, rc = cli("-t", "reference", "-g", d / "g.fa", "-k", 15, "-r", d / "x.kdb", "--filter-similar",
                       "--similarity-threshold", 1.5)
    assert rc != 0 and "similarity_threshold must be between 0 and 1" in err


def test_k_above_31_exits_with_a_message(data):
    d = data["dir"]
    out, err, rc = cli("-t", "reference", "-g", d / "g.fa", "-k", 40, "-r", d / "big.kdb")
    assert rc != 0 and "k <= 31" in err
