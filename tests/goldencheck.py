"""Shared helpers: run a case through the oracle (or the product) and compare with golden expectations."""
import hashlib
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def digest_reads(reads):
    h = hashlib.sha256()
    for rid, ty, lst in reads:
        h.update(f"{rid}\t{ty}\t{','.join(lst)}\n".encode())
    return h.hexdigest()


def canonical_from_oracle(case, nthreads=1):
    """Same canonical shape tests/golden/make_golden.py stores for the reference."""
    from oracle.oracle import OracleReference
    pr = case["params"]
    o = OracleReference(case["k"], [tuple(g) for g in case["genomes"]], filter_similar=pr.get("filter_similar", False),
                        similarity_threshold=pr.get("threshold", 0.95))
    out = {"genomes": [g[0] for g in o.genomes],
           "kmers": [[km, [[g, pos] for g, pos in inner.items()]] for km, inner in o.kmers_dict().items()],
           "ref_summary_json": json.dumps(o.get_summary())}
    if o.similarity_info is not None:
        out["similarity_info_json"] = json.dumps(o.similarity_info)
    if case.get("reads") is not None:
        al = o.align([tuple(r) for r in case["reads"]], pr.get("m", 1), pr.get("p", 1), pr.get("mrq"), pr.get("mkq"),
                     pr.get("mg"), nthreads=nthreads)
        out["reads"] = [[rid, d["mapping_type"], d["genomes_mapped_to"]] for rid, d in al.reads().items()]
        out["align_summary_json"] = json.dumps(al.get_summary())
    return out


def assert_matches(got, expect, label=""):
    for key in ("genomes", "kmers", "ref_summary_json", "similarity_info_json", "reads", "align_summary_json"):
        if key in expect or key in got:
            assert got.get(key) == expect.get(key), f"{label}: mismatch in {key}"
