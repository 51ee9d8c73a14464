"""
Multi-process tests of the partitioned index build (pa_index_build_partitioned through multi_gpu.build_partitioned) on
real kernels.

World size 1 to 3, one process per rank.  With fewer GPUs than ranks (the driver's GPU box has one) every rank uses
cuda:0, the communicator's control plane is a gloo all-gather (pa_comm_init_callbacks) and records / table slices
travel through CUDA IPC between the processes; with enough GPUs the same worker runs over NCCL (`PA_TEST_NCCL=1`).
Checks, against the CPU oracle: the replicated table (every k-mer of the index, absent k-mers), which rank owns which
k-mer, the gathered CSR in dict insertion order, alignment through the replica, read-sharded alignment with the reduced
summary, all-reduced EXTSIM statistics, and genome removal followed by a rebuilt replica; table-only and multi-round
(streamed) builds; chunked encoding.
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
comm = nat.Comm.from_torch(device)      # NCCL transport over the nccl backend, gloo all-gather as the control plane otherwise
assert comm.info()["nccl"] == use_nccl
MODE = os.environ.get("PA_TEST_MODE", "partition")   # partition | table_only | rounds

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

def run_case(case, table_only=False, rounds=0):
    k, genomes, pr = case["k"], case["genomes"], case["params"]
    data, goff = nat.pack_strings([g[1] for g in genomes])
    dix = multi_gpu.build_partitioned(comm, data, goff, k, device=device, table_only=table_only, n_rounds=rounds)
    o = orc.OracleReference(k, genomes)
    try:
        rinf = dix.replica.info()
        assert (rinf.n_keys, rinf.n_runs, rinf.n_occ) == o.sizes(), ((rinf.n_keys, rinf.n_runs, rinf.n_occ), o.sizes())
        assert rinf.align_only == 1
        ids = [g[0] for g in genomes]
        want = o.kmers_dict()
        if want:   # the replicated table answers every k-mer of the index (and misses the ones it does not hold)
            kms = list(want.keys())
            ng, g0 = dix.replica.table_lookup(kms)
            assert [int(x) for x in ng] == [len(want[km]) for km in kms]
            assert [int(x) for x in g0] == [min(want[km]) for km in kms]
            rng = np.random.default_rng(1)
            absent = ["".join(rng.choice(list("ACGT"), size=k)) for _ in range(64)] if k >= 1 else []
            absent = [a for a in absent if a not in want]
            if absent:
                ng, _ = dix.replica.table_lookup(absent)
                assert not ng.any()
            # a k-mer lives in the partition of the rank that owns its minimizer digit, and nowhere else
            if not table_only:
                pinf = dix.partition.info()
                mine = sum(1 for km in kms if nat.partition_of_kmer(k, km, world) == rank)
                assert pinf.n_keys == mine, (pinf.n_keys, mine)
        if not table_only:
            ex = dix.export_gathered()
            got = kmers_dict_from_export(ex, k)
            assert list(got.keys()) == list(want.keys())
            assert got == want
            assert np.all(np.diff(ex["keys"].astype(np.uint64)) > 0) if len(ex["keys"]) > 1 else True
        # alignment through the replica, every rank
        al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        got_reads, counters = align_reads(dix.replica, case["reads"], pr, ids)
        assert got_reads == al.reads(), case.get("seed")
        assert counters == [al.filtered_quality_reads, al.filtered_quality_kmers if pr["mkq"] is not None else 0,
                            al.filtered_hr_kmers if pr["mg"] is not None else 0]
        # read-sharded alignment end to end: summary and per-read results on every rank
        seqs, roff = nat.pack_strings([r[1] for r in case["reads"]])
        qls, _ = nat.pack_strings([r[2] for r in case["reads"]])
        summary, (types, lens, flat, cnt) = multi_gpu.align_sharded(dix.replica, comm, seqs, qls, roff, ids, pr["m"], pr["p"],
                                                                    pr["mrq"], pr["mkq"], pr["mg"], gather_reads=True)
        assert json.dumps(summary) == json.dumps(al.get_summary()), (summary, al.get_summary())
        assert np.array_equal(types, al.types)
        assert np.array_equal(lens, np.diff(al.list_off.astype(np.int64)))
        assert np.array_equal(flat, al.genomes[:len(flat)])
        if table_only:
            return
        # EXTSIM: all-reduced statistics and intersections
        classes = {{}}
        group = np.array([classes.setdefault(s, len(classes)) for s in ids], dtype=np.uint32)
        n = len(classes)
        total, uniq = dix.extsim_stats(group, n)
        inter = dix.extsim_pairwise(group, n)
        L = orc.lib()
        ot = np.zeros(max(n, 1), np.uint64); ou = np.zeros(max(n, 1), np.uint64); oi = np.zeros(max(n * n, 1), np.uint64)
        L.orc_extsim_stats(o._h, orc._ptr(group), n, orc._ptr(ot), orc._ptr(ou))
        L.orc_extsim_pairwise(o._h, orc._ptr(group), n, orc._ptr(oi))
        assert np.array_equal(total, ot[:n]) and np.array_equal(uniq, ou[:n])
        assert np.array_equal(inter.reshape(-1), oi[:n * n])
        # genome removal on every partition + rebuilt replica
        rng = np.random.default_rng(case.get("seed", 0))
        keep = (rng.random(len(ids)) < 0.6).astype(np.uint8)
        dix.drop_genomes(keep)
        h2 = L.orc_index_drop_genomes(o._h, orc._ptr(np.concatenate([keep, [0]]).astype(np.uint8)))
        L.orc_index_free(o._h)
        o._h = h2
        o.genomes = [g for g, kp in zip(o.genomes, keep) if kp]
        ex = dix.export_gathered()
        want = o.kmers_dict()
        got = kmers_dict_from_export(ex, k)
        assert list(got.keys()) == list(want.keys()) and got == want
        assert dix.sizes() == o.sizes()
        if o.genomes:
            al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
            got_reads, _ = align_reads(dix.replica, case["reads"], pr, [g[0] for g in o.genomes])
            assert got_reads == al.reads()
    finally:
        dix.close()

seeds = [int(x) for x in sys.argv[1].split(",")]
for seed in seeds:
    if seed >= 0:
        case = synth.fuzz_case(seed, dup_ids=seed % 5 == 0)
    else:
        # k = 31 on clustered genomes with N runs; reads with errors
        genomes = synth.make_genomes(6, 30_000, seed=-seed, cluster_size=3, shared_frac=0.4, n_every=7000, n_run=9)
        b, q, off = synth.make_reads(genomes, 1500, 150, seed=-seed + 1, sub_rate=0.01, random_frac=0.03)
        case = {{"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                "params": dict(m=1, p=1, mrq=None, mkq=60 if seed % 2 else None, mg=2 if seed % 2 else None), "seed": -seed}}
    if MODE == "partition":
        run_case(case)
    elif MODE == "table_only":
        run_case(case, table_only=True)
    else:   # several rounds per rank, where the digit space of k has room for them
        room = (1 << min(8, 2 * min(max(case["k"], 1), 16))) // world
        run_case(case, table_only=True, rounds=max(1, min(3, room)))
comm.barrier()
comm.close()
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def _run(world, seeds, tmp_path, nccl=False, mode="partition", no_ipc=False, env_extra=None):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=conftest.ROOT, pkg=conftest.PKG_DIR))
    port = 23000 + (os.getpid() * 7 + world * 131 + len(seeds)) % 4000
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PA_TEST_NCCL="1" if nccl else "0", PA_TEST_MODE=mode, PA_TEST_NO_IPC="1" if no_ipc else "0",
                   **(env_extra or {}))
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


def test_partitioned_build_world_2_k31(tmp_path):
    _run(2, [-11, -12], tmp_path)


def test_table_only_build_world_2(tmp_path):
    _run(2, list(range(5100, 5112)) + [-13], tmp_path, mode="table_only")


def test_streamed_rounds_world_1_and_2(tmp_path):
    # several rounds per rank: the key space of a rank is built slice by slice (config E on few GPUs)
    _run(1, list(range(5300, 5312)) + [-17], tmp_path, mode="rounds")
    _run(2, list(range(5320, 5328)) + [-19], tmp_path, mode="rounds")


def test_chunked_encode_small_chunks(tmp_path):
    # genomes walked in 4096-base chunks (windows spanning a chunk end come from the overlap), several chunks per round
    _run(2, [-21, -22], tmp_path, env_extra={"PA_BUILD_CHUNK": "4096"})
    _run(1, [-23], tmp_path, mode="rounds", env_extra={"PA_BUILD_CHUNK": "8192"})


def test_partitioned_build_without_peer_mapping_needs_nccl(tmp_path):
    # PA_TEST_NO_IPC=1 makes one rank refuse the peer mapping; over gloo there is no other device data plane, so every
    # rank must fail together with the same clear error instead of hanging
    with pytest.raises(AssertionError, match="peer memory cannot be mapped"):
        _run(2, [5200], tmp_path, no_ipc=True)


def test_partitioned_build_over_nccl_when_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, [5000, 5001, -11], tmp_path, nccl=True)
    _run(2, [5002, -12], tmp_path, nccl=True, no_ipc=True)              # exchange through NCCL send / recv
    _run(2, [-14], tmp_path, nccl=True, env_extra={"PA_TABLE_GATHER": "ipc"})   # table slices pulled over IPC
