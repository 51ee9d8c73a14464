#!/usr/bin/env python3
"""
Generates tests/golden/*.json by running the UNMODIFIED Python reference
(/root/reference/src, imported through tests/refimpl.py).  Run in the build
container only:   python tests/golden/make_golden.py

Fixtures:
  kat_reference_tests.json  the reference's own known-answer fixtures
                            (inputs of /root/reference/src/test_kmer.py:36-65, 86-111, 119-141, 149-176,
                            184-212, 556-560 and SURVEY.md 8(c)'s order / EXTSIM cases) with the
                            reference's full outputs for several parameter sets.
  fuzz_small_k.json         240 seeded adversarial cases (synth.fuzz_case) with full outputs.
  config_a.json             BASELINE.json configs[0]: 3 x 50 kb genomes, 10,000 x 100 bp reads, k=31,
                            plain and EXTQUALITY runs: summaries, counters, per-read digest + sample.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT,
                os.path.join(ROOT, "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")]

import refimpl  # noqa: E402
import synth  # noqa: E402


def run(ref, case):
    pr = case["params"]
    out = refimpl.ref_run(ref, case["k"], case["genomes"], case.get("reads"), pr.get("m", 1), pr.get("p", 1),
                          pr.get("mrq"), pr.get("mkq"), pr.get("mg"), pr.get("filter_similar", False),
                          pr.get("threshold", 0.95))
    # JSON object keys must be strings: store kmers as ordered lists
    out["kmers"] = [[km, [[g, pos] for g, pos in inner.items()]] for km, inner in out["kmers"].items()]
    out["ref_summary_json"] = json.dumps(out.pop("ref_summary"))
    if "reads" in out:
        out["reads"] = [[rid, d["mapping_type"], d["genomes_mapped_to"]] for rid, d in out["reads"].items()]
        out["align_summary_json"] = json.dumps(out.pop("align_summary"))
    if out["similarity_info"] is not None:
        out["similarity_info_json"] = json.dumps(out["similarity_info"])
    out.pop("similarity_info")
    return out


def P(**kw):
    base = {"m": 1, "p": 1, "mrq": None, "mkq": None, "mg": None, "filter_similar": False, "threshold": 0.95}
    base.update(kw)
    return base


def kat_cases():
    four = [("Genome1", "AGCTAGCTAGCTAGCTAGCT"), ("Genome2", "TGCATGCATGCATGCATGCA"),
            ("Genome3", "AGCTTGCATGCAGCTAGCTA"), ("Genome4", "CCGGAAGCTTGCATGCAGCTA")]
    three_reads = [("Read1", "AGCTAGCT", "IIIIIIII"), ("Read2", "TGCATGCA", "!!!!!!!!"), ("Read3", "GGGGGGGG", "!!IIIIII")]
    cases = []
    for name, pr in [("plain", P()), ("mrq40", P(mrq=40)), ("mkq60", P(mkq=60)), ("mg2", P(mg=2)),
                     ("combined_40_50_2", P(mrq=40, mkq=50, mg=2)), ("combined_30_30_3", P(mrq=30, mkq=30, mg=3))]:
        cases.append({"name": f"four_genomes_k3_{name}", "k": 3, "genomes": four, "reads": three_reads, "params": pr})
    cases.append({"name": "unmapped", "k": 4, "genomes": [("Genome1", "AACCGGTTAACC"), ("Genome2", "GGTTCCAAGGTT")],
                  "reads": [("Read1", "TAGGCAT", "IIIIIII")], "params": P()})
    cases.append({"name": "unique", "k": 4, "genomes": [("Genome1", "ATGGCTATGCTA"), ("Genome2", "CTATGGCAGGCA")],
                  "reads": [("Read2", "ATGGCTAT", "IIIIIIII")], "params": P()})
    cases.append({"name": "ambiguous4", "k": 4,
                  "genomes": [("Genome1", "ATCGACGGTCGTTA"), ("Genome2", "CGATGATCAGTACGA"),
                              ("Genome3", "ATCCACCTAACGTACGGT"), ("Genome4", "CTAGGGACTGCACTA")],
                  "reads": [("Read3", "ATCGATCCTAG", "IIIIIIIIIII")], "params": P()})
    flip = [("Genome1", "ATGCCTTTTCGGGG"), ("Genome2", "GCCGTTTTCGGGGCTA"), ("Genome3", "CCGG"),
            ("Genome4", "AAAAAAAAGGGCT"), ("Genome5", "TTTTTTTTGCTAA")]
    flip_read = [("Read4", "ATGCCGGGGCTAA", "IIIIIIIIIIIII")]
    for name, pr in [("p1", P()), ("p5", P(p=5)), ("pneg", P(p=-1)), ("m0", P(m=0)), ("m3", P(m=3))]:
        cases.append({"name": f"flip_{name}", "k": 4, "genomes": flip, "reads": flip_read, "params": pr})
    order = [("G1", "AAAACCCC"), ("G2", "GGGGTTTT"), ("G3", "ACGTACGA")]
    order_reads = [("a", "GGGGTAAAAC", "IIIIIIIIII"), ("b", "AAAACGGGGT", "IIIIIIIIII")]
    cases.append({"name": "order_m1", "k": 4, "genomes": order, "reads": order_reads, "params": P()})
    cases.append({"name": "order_m0", "k": 4, "genomes": order, "reads": order_reads, "params": P(m=0)})
    sim3 = [("GenomeA", "AGCTAGCTAGCT"), ("GenomeB", "AGCTAGCTAGCT"), ("GenomeC", "TGCATGCATGCA")]
    cases.append({"name": "extsim_abc_095", "k": 4, "genomes": sim3, "reads": three_reads,
                  "params": P(filter_similar=True)})
    cases.append({"name": "extsim_single", "k": 4, "genomes": sim3[:1], "reads": None, "params": P(filter_similar=True)})
    sim5 = [("A", "AGCTAGCTAGCT"), ("B", "AGCTAGCTAGCT"), ("C", "TGCATGCATGCA"), ("D", "AGCTAGCTAGCTTTTTGGGA"), ("E", "AC")]
    cases.append({"name": "extsim_abcde_06", "k": 4, "genomes": sim5, "reads": three_reads,
                  "params": P(filter_similar=True, threshold=0.6)})
    cases.append({"name": "with_n_and_dups", "k": 3, "genomes": [("x", "ACGNNACGTACG"), ("y", "NNACGN"), ("x", "TTTACG")],
                  "reads": [("r1", "ACGTACGACG", "IIIIII!!!!"), ("r2", "AC", "II")], "params": P(mkq=50, mg=2)})
    return cases


def digest_reads(reads):
    h = hashlib.sha256()
    for rid, ty, lst in reads:
        h.update(f"{rid}\t{ty}\t{','.join(lst)}\n".encode())
    return h.hexdigest()


def config_a(ref):
    genomes = synth.make_genomes(3, 50_000, seed=1234, cluster_size=3, shared_frac=0.3, sub_rate=0.01,
                                 n_every=20_000, n_run=40)
    b, q, off = synth.make_reads(genomes, 10_000, 100, seed=4321, sub_rate=0.01, random_frac=0.02)
    pairs = synth.genomes_as_pairs(genomes)
    triples = synth.reads_as_triples(b, q, off)
    inp = hashlib.sha256(b"".join(g.tobytes() for g in genomes) + b.tobytes() + q.tobytes()).hexdigest()
    out = {"generator": {"genomes": "synth.make_genomes(3, 50000, seed=1234, cluster_size=3, shared_frac=0.3, "
                                    "sub_rate=0.01, n_every=20000, n_run=40)",
                         "reads": "synth.make_reads(genomes, 10000, 100, seed=4321, sub_rate=0.01, random_frac=0.02)"},
           "k": 31, "input_sha256": inp, "runs": []}
    for name, pr in [("plain", P()), ("extquality", P(mrq=62, mkq=60, mg=1)), ("p0_m2", P(m=2, p=0))]:
        r = run(ref, {"k": 31, "genomes": pairs, "reads": triples, "params": pr})
        n_kmers = len(r["kmers"])
        out["runs"].append({"name": name, "params": pr, "n_distinct_kmers": n_kmers,
                            "align_summary_json": r["align_summary_json"], "reads_digest": digest_reads(r["reads"]),
                            "n_stored_reads": len(r["reads"]), "reads_sample": r["reads"][:300]})
    return out


def main():
    ref = refimpl.load_reference()
    assert ref is not None, "needs /root/reference"
    kats = [{"case": c, "expect": run(ref, c)} for c in kat_cases()]
    with open(os.path.join(HERE, "kat_reference_tests.json"), "w") as f:
        json.dump(kats, f, indent=1)
    fuzz = []
    for seed in list(range(100_000, 100_200)) + list(range(200_000, 200_040)):
        c = synth.fuzz_case(seed, dup_ids=seed >= 200_000)
        fuzz.append({"case": c, "expect": run(ref, c)})
    with open(os.path.join(HERE, "fuzz_small_k.json"), "w") as f:
        json.dump(fuzz, f, separators=(",", ":"))
    with open(os.path.join(HERE, "config_a.json"), "w") as f:
        json.dump(config_a(ref), f, indent=1)
    print("wrote", len(kats), "KAT cases,", len(fuzz), "fuzz cases, config A")


if __name__ == "__main__":
    main()
