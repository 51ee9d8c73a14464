"""
Multi-process tests of the hash-partitioned index build (multi_gpu.build_partitioned) on real kernels.

World size 2 or 3, one process per rank.  With fewer GPUs than ranks (the driver's GPU box has one) every rank uses
cuda:0 and torch.distributed runs over gloo with host staging; with enough GPUs the same worker runs over NCCL
(`PA_TEST_NCCL=1`).  Checks, against the CPU oracle: the gathered CSR in dict insertion order, alignment through the
replica, all-reduced EXTSIM statistics, and genome removal followed by a re-gather.
"""
import os
import subprocess
import sys

import pytest

import conftest

pytestmark = pytest.mark.gpu

WORKER = r'''
import json, os, sys
sys.path[:0] = [{root!r}, {pkg!r}, os.path.join({root!r}, "tests")]
import numpy as np, torch, torch.distributed as dist
import synth, multi_gpu
import _native as nat
from oracle import oracle as orc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
use_nccl = os.environ.get("PA_TEST_NCCL") == "1" and torch.cuda.device_count() >= world
device = rank if use_nccl else 0
torch.cuda.set_device(device)
dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world,
                        **({{"device_id": torch.device("cuda", device)}} if use_nccl else {{}}))
NAMES = {{1: "UNMAPPED", 2: "UNIQUELY_MAPPED", 3: "AMBIGUOUSLY_MAPPED"}}
FUSED = os.environ.get("PA_TEST_FUSED", "1") == "1"   # scatter into peer memory (CUDA IPC) vs partition + all_to_all

def kmers_dict_from_export(ex, k):
    kmers = nat.decode_kmers(k, ex["keys"])
    out = {{}}
    for u in ex["order"]:
        u = int(u)
        inner = {{}}
        for r in range(int(ex["run_off"][u]), int(ex["run_off"][u + 1])):
            inner[int(ex["run_genome"][r])] = [int(x) for x in ex["pos"][int(ex["pos_off"][r]):int(ex["pos_off"][r + 1])]]
        out[kmers[u]] = inner
    return out

def align_reads(ix, reads, pr, genome_ids):
    seqs, off = nat.pack_strings([r[1] for r in reads])
    quals, _ = nat.pack_strings([r[2] for r in reads])
    words, lst, counters = ix.align(seqs, quals, off, nat.make_params(pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"]))
    types, lens, payload = nat.decode_words(words)
    out = {{}}
    for i, r in enumerate(reads):
        t = int(types[i])
        if t == 0: continue
        n = int(lens[i])
        gl = [int(payload[i])] if n == 1 else [int(x) for x in lst[int(payload[i]):int(payload[i]) + n]]
        out[r[0]] = {{"mapping_type": NAMES[t], "genomes_mapped_to": [genome_ids[g] for g in gl]}}
    return out, [int(c) for c in counters]

def run_case(case):
    k, genomes, pr = case["k"], case["genomes"], case["params"]
    data, goff = nat.pack_strings([g[1] for g in genomes])
    lengths = np.diff(goff.astype(np.int64))
    g_lo, g_hi = multi_gpu.genome_shards(lengths, world)[rank]
    mine = np.zeros(int(goff[g_hi] - goff[g_lo]) + 64, dtype=np.uint8)
    mine[:int(goff[g_hi] - goff[g_lo])] = data[int(goff[g_lo]):int(goff[g_hi])]
    dix = multi_gpu.build_partitioned(mine, goff, k, (g_lo, g_hi), device=device, fused=FUSED)
    if os.environ.get("PA_TEST_NO_IPC") == "1" and FUSED and k >= 1:
        assert not dix.fused and "fused_exchange_unavailable" in dix.timings    # every rank fell back together
    else:
        assert dix.fused == (FUSED and k >= 1) or "fused_exchange_unavailable" in dix.timings
    o = orc.OracleReference(k, genomes)
    try:
        # every record went to exactly one owner, and keys are partitioned by range
        tot = torch.tensor([dix.sent_records, dix.received_records, dix.partition.info().n_occ], dtype=torch.int64)
        if use_nccl:
            d = tot.cuda(); dist.all_reduce(d); tot = d.cpu()
        else:
            dist.all_reduce(tot)
        assert tot[0] == tot[1] == tot[2] == o.sizes()[2], (tot, o.sizes())
        rinf = dix.replica.info()
        assert (rinf.n_keys, rinf.n_runs, rinf.n_occ) == o.sizes()
        ex = multi_gpu.export_gathered(dix, dst=0)
        if rank == 0:
            want = o.kmers_dict()
            got = kmers_dict_from_export(ex, k)
            assert list(got.keys()) == list(want.keys())
            assert got == want
            assert np.all(np.diff(ex["keys"].astype(np.uint64)) > 0) if len(ex["keys"]) > 1 else True   # rank order = key order
        # alignment through the replica, every rank
        ids = [g[0] for g in genomes]
        al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        got_reads, counters = align_reads(dix.replica, case["reads"], pr, ids)
        assert got_reads == al.reads(), case.get("seed")
        assert counters == [al.filtered_quality_reads, al.filtered_quality_kmers if pr["mkq"] is not None else 0,
                            al.filtered_hr_kmers if pr["mg"] is not None else 0]
        # read-sharded alignment end to end: summary on every rank, per-read results gathered on rank 0
        seqs, roff = nat.pack_strings([r[1] for r in case["reads"]])
        qls, _ = nat.pack_strings([r[2] for r in case["reads"]])
        summary, reads = multi_gpu.align_sharded(dix.replica, seqs, qls, roff, ids, pr["m"], pr["p"], pr["mrq"], pr["mkq"],
                                                 pr["mg"], gather_reads=True)
        assert json.dumps(summary) == json.dumps(al.get_summary()), (summary, al.get_summary())
        if rank == 0:
            want_types = [int(t) for t in al.types]
            assert reads[0] == want_types
            for i in range(len(want_types)):
                if want_types[i] != 0:
                    assert reads[1][i] == [int(g) for g in al.genomes[int(al.list_off[i]):int(al.list_off[i + 1])]]
        # EXTSIM: all-reduced statistics and intersections
        classes = {{}}
        group = np.array([classes.setdefault(s, len(classes)) for s in ids], dtype=np.uint32)
        n = len(classes)
        total, uniq = multi_gpu.extsim_stats_allreduce(dix, group, n)
        inter = multi_gpu.extsim_pairwise_allreduce(dix, group, n)
        L = orc.lib()
        ot = np.zeros(max(n, 1), np.uint64); ou = np.zeros(max(n, 1), np.uint64); oi = np.zeros(max(n * n, 1), np.uint64)
        L.orc_extsim_stats(o._h, orc._ptr(group), n, orc._ptr(ot), orc._ptr(ou))
        L.orc_extsim_pairwise(o._h, orc._ptr(group), n, orc._ptr(oi))
        assert np.array_equal(total, ot[:n]) and np.array_equal(uniq, ou[:n])
        assert np.array_equal(inter.reshape(-1), oi[:n * n])
        # genome removal on every partition + re-gathered replica
        rng = np.random.default_rng(case.get("seed", 0))
        keep = (rng.random(len(ids)) < 0.6).astype(np.uint8)
        multi_gpu.drop_genomes(dix, keep)
        h2 = L.orc_index_drop_genomes(o._h, orc._ptr(np.concatenate([keep, [0]]).astype(np.uint8)))
        L.orc_index_free(o._h)
        o._h = h2
        o.genomes = [g for g, kp in zip(o.genomes, keep) if kp]
        ex = multi_gpu.export_gathered(dix, dst=0)
        if rank == 0:
            want = o.kmers_dict()
            got = kmers_dict_from_export(ex, k)
            assert list(got.keys()) == list(want.keys()) and got == want
        if o.genomes:
            al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
            got_reads, _ = align_reads(dix.replica, case["reads"], pr, [g[0] for g in o.genomes])
            assert got_reads == al.reads()
    finally:
        dix.close()

seeds = [int(x) for x in sys.argv[1].split(",")]
for seed in seeds:
    if seed >= 0:
        run_case(synth.fuzz_case(seed, dup_ids=seed % 5 == 0))
    else:
        # k = 31 on clustered genomes with N runs; reads with errors
        genomes = synth.make_genomes(6, 30_000, seed=-seed, cluster_size=3, shared_frac=0.4, n_every=7000, n_run=9)
        b, q, off = synth.make_reads(genomes, 1500, 150, seed=-seed + 1, sub_rate=0.01, random_frac=0.03)
        case = {{"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                "params": dict(m=1, p=1, mrq=None, mkq=60 if seed % 2 else None, mg=2 if seed % 2 else None), "seed": -seed}}
        run_case(case)
dist.barrier()
multi_gpu.release_peer_buffers()
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def _run(world, seeds, tmp_path, nccl=False, fused=True, no_ipc=False):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=conftest.ROOT, pkg=conftest.PKG_DIR))
    port = 23000 + (os.getpid() * 7 + world * 131 + len(seeds)) % 4000
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PA_TEST_NCCL="1" if nccl else "0", PA_TEST_FUSED="1" if fused else "0",
                   PA_TEST_NO_IPC="1" if no_ipc else "0")
        procs.append(subprocess.Popen([sys.executable, str(script), ",".join(str(s) for s in seeds)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-3000:]
    assert "OK" in outs[0][0]


def test_partitioned_build_world_2_fuzz(tmp_path):
    _run(2, list(range(5000, 5030)), tmp_path)


def test_partitioned_build_world_3_fuzz(tmp_path):
    _run(3, list(range(6000, 6016)), tmp_path)


def test_partitioned_build_falls_back_when_peer_mapping_fails(tmp_path):
    _run(3, [5200, 5201, 5202, -15], tmp_path, no_ipc=True)


def test_partitioned_build_all_to_all_exchange(tmp_path):
    _run(2, list(range(5100, 5112)) + [-13], tmp_path, fused=False)


def test_partitioned_build_world_2_k31(tmp_path):
    _run(2, [-11, -12], tmp_path)


def test_partitioned_build_over_nccl_when_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, [5000, 5001, -11], tmp_path, nccl=True)
