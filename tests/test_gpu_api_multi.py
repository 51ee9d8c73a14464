"""
The reference's two call sites -- KmerReference(...) and PseudoAlignment(ref).align_reads_from_container(...)
(/root/reference/src/main.py:100-102, 176-178) -- on several ranks: same inputs, byte-identical JSON at every world size
(SURVEY.md section 4 item 3).  One process per rank; on the 1-GPU test box every rank uses cuda:0 with a gloo control
plane (PA_DIST_BACKEND=gloo), on a multi-GPU box the same code runs over NCCL.
"""
import json
import os
import subprocess
import sys

import pytest

import conftest
import synth
from oracle.oracle import OracleReference

pytestmark = pytest.mark.gpu
PKG = conftest.PKG_DIR

WORKER = r'''
import json, os, sys, pickle
sys.path[:0] = [{root!r}, {pkg!r}, os.path.join({root!r}, "tests")]
import numpy as np
import synth, multi_gpu
import kmer
from records import FASTARecordContainer, FASTAQRecordContainer

ctx = multi_gpu.init()
rank, world = ctx["rank"], ctx["world"]
out = {{}}
for seed in [int(x) for x in sys.argv[1].split(",")]:
    case = synth.fuzz_case(seed, dup_ids=seed % 5 == 0) if seed >= 0 else None
    if case is None:
        genomes = synth.make_genomes(5, 20_000, seed=-seed, cluster_size=5, shared_frac=0.5, n_every=7000, n_run=9)
        b, q, off = synth.make_reads(genomes, 1200, 120, seed=-seed + 1, sub_rate=0.01, random_frac=0.03)
        case = {{"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                "params": dict(m=1, p=1, mrq=55 if seed % 2 else None, mkq=60 if seed % 2 else None, mg=2 if seed % 2 else None,
                               filter_similar=seed % 2 == 0, threshold=0.3)}}
    pr = case["params"]
    fasta = "".join(f">{{g}}\n{{s}}\n" for g, s in case["genomes"])
    fastq = "".join(f"@{{r}}\n{{s}}\n+\n{{q}}\n" for r, s, q in case["reads"])
    try:
        fc = FASTARecordContainer()
        fc.parse_records(fasta)
        ref = kmer.KmerReference(case["k"], fc, filter_similar=pr.get("filter_similar", False),
                                 similarity_threshold=pr.get("threshold", 0.95))
    except ValueError as e:   # the build refuses the input (k > 31 ...): every rank must refuse it the same way
        out[seed] = ["error", type(e).__name__, str(e)]
        continue
    assert ref._dist is not None or world == 1
    res = {{"ref": ref.summary_json(), "ref_dict": json.dumps(ref.get_summary(), indent=4), "genomes": [g.identifier for g in ref.genomes]}}
    if hasattr(ref, "similarity_info"):
        res["sim"] = json.dumps(ref.similarity_info)
    some = [km for km in list(ref.kmers)[:5]]
    res["lookups"] = [[km, sorted((g.identifier, sorted(p)) for g, p in ref.get_kmer_references(km).items())] for km in some]
    try:
        reads = FASTAQRecordContainer()
        reads.parse_records(fastq)
        al = kmer.PseudoAlignment(ref)
        al.align_reads_from_container(reads, pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        res["align"] = json.dumps(al.get_summary(), indent=4)
        res["reads"] = [[rid, d["mapping_type"].name, d["genomes_mapped_to"]] for rid, d in al.reads.items()]
        al2 = pickle.loads(pickle.dumps(al))          # .aln round trip (array-backed)
        res["align_reloaded"] = json.dumps(al2.get_summary(), indent=4)
        res["reads_reloaded"] = [[rid, d["mapping_type"].name, d["genomes_mapped_to"]] for rid, d in al2.reads.items()]
    except (ValueError, ZeroDivisionError) as e:
        res["align_error"] = [type(e).__name__, str(e)]
    out[seed] = res
if rank == 0:
    with open(sys.argv[2], "w") as f:
        json.dump(out, f)
multi_gpu.shutdown()
'''


def _run_world(world, seeds, tmp_path):
    script = tmp_path / f"worker{world}.py"
    script.write_text(WORKER.format(root=conftest.ROOT, pkg=PKG))
    out = tmp_path / f"out{world}.json"
    port = 24000 + (os.getpid() * 13 + world * 57) % 4000
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PA_DIST_BACKEND="gloo")
        procs.append(subprocess.Popen([sys.executable, str(script), ",".join(map(str, seeds)), str(out)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=1200) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    return json.loads(out.read_text())


def test_same_json_at_every_world_size(tmp_path):
    seeds = list(range(7000, 7024)) + [-31, -32]
    results = {w: _run_world(w, seeds, tmp_path) for w in (1, 2, 3)}
    for seed in map(str, seeds):
        assert results[2][seed] == results[1][seed], seed
        assert results[3][seed] == results[1][seed], seed
    # and world size 1 is the reference's answer (oracle pinned to the reference)
    checked = 0
    for seed in seeds:
        got = results[1][str(seed)]
        if isinstance(got, list) or "align" not in got:
            continue
        case = synth.fuzz_case(seed, dup_ids=seed % 5 == 0) if seed >= 0 else None
        if case is None:
            continue
        pr = case["params"]
        o = OracleReference(case["k"], [tuple(g) for g in case["genomes"]], filter_similar=pr.get("filter_similar", False),
                            similarity_threshold=pr.get("threshold", 0.95))
        assert got["ref_dict"] == json.dumps(o.get_summary(), indent=4)
        assert got["ref"] == got["ref_dict"]
        al = o.align([tuple(r) for r in case["reads"]], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        assert got["align"] == json.dumps(al.get_summary(), indent=4) == got["align_reloaded"]
        assert got["reads"] == got["reads_reloaded"] == [[rid, d["mapping_type"], d["genomes_mapped_to"]] for rid, d in al.reads().items()]
        checked += 1
    assert checked >= 10


def test_cli_under_torchrun_prints_the_single_process_output(tmp_path):
    genomes = synth.make_genomes(5, 6000, seed=41, cluster_size=5, shared_frac=0.6, sub_rate=0.004, n_every=2500, n_run=6)
    b, q, off = synth.make_reads(genomes, 400, 90, seed=42, sub_rate=0.01, random_frac=0.05)
    synth.write_fasta(str(tmp_path / "g.fa"), synth.genomes_as_pairs(genomes), width=70)
    synth.write_fastq(str(tmp_path / "r.fq"), synth.reads_as_triples(b, q, off))
    tasks = [["-t", "dumpalign", "-g", str(tmp_path / "g.fa"), "-k", "21", "--reads", str(tmp_path / "r.fq"), "--min-kmer-quality", "58",
              "--max-genomes", "3"],
             ["-t", "dumpref", "-g", str(tmp_path / "g.fa"), "-k", "21", "--filter-similar", "--similarity-threshold", "0.4"]]
    for n, argv in enumerate(tasks):
        single = subprocess.run([sys.executable, "main.py", *argv], capture_output=True, text=True, cwd=PKG)
        assert single.returncode == 0, single.stderr
        port = 25000 + (os.getpid() * 3 + n) % 3000
        multi = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                                "127.0.0.1", "--master-port", str(port), "main.py", *argv], capture_output=True, text=True, cwd=PKG,
                               env=dict(os.environ, PA_DIST_BACKEND="gloo"), timeout=900)
        assert multi.returncode == 0, multi.stderr[-3000:]
        assert multi.stdout == single.stdout
    # reference / align write their files on rank 0 only; one process reads them back
    for n, argv in enumerate([["-t", "reference", "-g", str(tmp_path / "g.fa"), "-k", "21", "-r", str(tmp_path / "m.kdb")],
                              ["-t", "align", "-r", str(tmp_path / "m.kdb"), "--reads", str(tmp_path / "r.fq"), "-a", str(tmp_path / "m.aln")]]):
        multi = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                                "127.0.0.1", "--master-port", str(25700 + n), "main.py", *argv], capture_output=True, text=True, cwd=PKG,
                               env=dict(os.environ, PA_DIST_BACKEND="gloo"), timeout=900)
        assert multi.returncode == 0, multi.stderr[-3000:]
    a = subprocess.run([sys.executable, "main.py", "-t", "dumpalign", "-a", str(tmp_path / "m.aln")], capture_output=True, text=True, cwd=PKG)
    b2 = subprocess.run([sys.executable, "main.py", "-t", "dumpalign", "-r", str(tmp_path / "m.kdb"), "--reads", str(tmp_path / "r.fq")],
                        capture_output=True, text=True, cwd=PKG)
    assert a.returncode == 0 and b2.returncode == 0, a.stderr + b2.stderr
    assert a.stdout == b2.stdout and '"Statistics"' in a.stdout
