"""
Full-size checks (BASELINE.json configs[1] / configs[2]: 100 genomes x 5 Mb, 10^7 x 150-bp reads, k = 31) through
size-independent properties -- the oracle cannot run at this size:

  * CSR invariants and checksums: keys strictly ascending, genome runs ascending, positions ascending, occurrences =
    number of N-free windows counted independently on the device with torch;
  * lookup round trip: k-mers cut from known genome positions are found with their genome in the set;
  * ground truth by construction: an error-free read cut from genome g only carries k-mers of g, so it is UNIQUE [g]
    or AMBIGUOUS [] -- never another genome, never unmapped;
  * consistency: device-resident call = host-buffer call (packed and raw), any chunking, any read order;
  * counters: reads dropped by min-read-quality = the exact integer test evaluated with torch; statistics sum to n.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu


class _DevArray:
    """A raw device pointer dressed as a CUDA array for torch.as_tensor (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _as_tensor(ptr, n, typestr, dev):
    import torch
    dt = {"<i8": torch.int64, "<i4": torch.int32}[typestr]
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dt, device=dev)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=dev)

G, GL, NR, RL, K = 100, 5_000_000, 10_000_000, 150, 31


@pytest.fixture(scope="module")
def world():
    import torch
    sys.path.insert(0, conftest.ROOT)
    import bench
    import _native as nat
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    bases = bench.device_genomes(torch, dev, G, GL, seed=1000)
    goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL))
    ix = nat.NativeIndex.build_device(bases.data_ptr(), goff, K, device=0)
    rb, rq, roff = bench.device_reads(torch, dev, bases, G, GL, NR, RL, seed=2000)
    yield {"torch": torch, "nat": nat, "dev": dev, "bases": bases, "goff": goff, "ix": ix, "rb": rb, "rq": rq, "roff": roff}
    ix.close()


def _align_device(w, rb, rq, roff, n, params, need_q):
    torch, nat = w["torch"], w["nat"]
    L = nat.lib()
    words = torch.empty(n, dtype=torch.int64, device=w["dev"])
    lst = torch.empty(max(n // 2, 1024), dtype=torch.int32, device=w["dev"])
    state = torch.zeros(5, dtype=torch.int64, device=w["dev"])
    nl = ctypes.c_int32(0)
    torch.cuda.synchronize()
    nat.check(L.pa_align_batch_device(w["ix"].handle, ctypes.c_void_p(rb.data_ptr()), ctypes.c_void_p(rq.data_ptr()) if need_q else None,
                                      ctypes.c_void_p(roff.data_ptr()), n, RL, ctypes.byref(params), ctypes.c_void_p(words.data_ptr()),
                                      ctypes.c_void_p(lst.data_ptr()), lst.numel(), ctypes.c_void_p(state.data_ptr()), None,
                                      ctypes.byref(nl)))
    torch.cuda.synchronize()
    assert int(state[1]) == 0
    return words, lst, state


def test_index_invariants_and_checksums(world):
    torch, nat, ix = world["torch"], world["nat"], world["ix"]
    inf = ix.info()
    # occurrences = windows of 31 bases without N, counted independently
    is_n = (world["bases"].view(G, GL) == 78).to(torch.int32)
    c = torch.cumsum(is_n, dim=1)
    win = c[:, K - 1:] - torch.cat([torch.zeros(G, 1, dtype=c.dtype, device=c.device), c[:, :GL - K]], dim=1)
    assert int((win == 0).sum()) == inf.n_occ
    del is_n, c, win
    L = nat.lib()
    pk, po, pg = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    nat.check(L.pa_index_csr_device(ix.handle, ctypes.byref(pk), ctypes.byref(po), ctypes.byref(pg)))
    keys = _as_tensor(pk.value, inf.n_keys, "<i8", world["dev"])
    run_off = _as_tensor(po.value, inf.n_keys + 1, "<i8", world["dev"])
    run_genome = _as_tensor(pg.value, inf.n_runs, "<i4", world["dev"])
    assert bool((keys[1:] > keys[:-1]).all()) and int(keys.min()) >= 0 and int(keys.max()) < (1 << 62)
    d = run_off[1:] - run_off[:-1]
    assert bool((d >= 1).all()) and int(run_off[0]) == 0 and int(run_off[-1]) == inf.n_runs
    same_key = torch.ones(inf.n_runs - 1, dtype=torch.bool, device=world["dev"])
    same_key[(run_off[1:-1] - 1)] = False                    # the last run of a key has no successor inside the key
    assert bool((run_genome[1:][same_key] > run_genome[:-1][same_key]).all())
    assert 0 <= int(run_genome.min()) and int(run_genome.max()) < G
    assert inf.n_keys <= inf.n_runs <= inf.n_occ
    # EXTSIM statistics are consistent with the CSR: sum of totals = runs, unique = keys with one run
    total, uniq = ix.extsim_stats(np.arange(G, dtype=np.uint32), G)
    assert int(total.sum()) == inf.n_runs and int(uniq.sum()) == int((d == 1).sum())


def test_lookup_round_trip_at_known_positions(world):
    torch, nat, ix = world["torch"], world["nat"], world["ix"]
    gen = torch.Generator(device=world["dev"]); gen.manual_seed(7)
    n = 200_000
    g = torch.randint(0, G, (n,), generator=gen, device=world["dev"])
    p = torch.randint(0, GL - K + 1, (n,), generator=gen, device=world["dev"])
    idx = (g * GL + p)[:, None] + torch.arange(K, device=world["dev"])[None, :]
    kmers = world["bases"][idx]
    ok = ~(kmers == 78).any(dim=1)
    host = kmers[ok].cpu().numpy()
    gs = g[ok].cpu().numpy()
    strings = [row.tobytes().decode() for row in host[:50_000]]
    ng, g0 = ix.table_lookup(strings)
    assert (ng >= 1).all()
    assert (g0 <= gs[:50_000]).all()                          # the smallest genome of the set cannot exceed the source
    spec = ng == 1
    assert (g0[spec] == gs[:50_000][spec]).all()
    ranks = ix.lookup(strings[:5000])
    assert (ranks != nat.RANK_MISS).all()
    rnd = ["".join(x) for x in np.random.default_rng(1).choice(list("ACGT"), size=(2000, K))]
    ng2, _ = ix.table_lookup(rnd)
    assert (ng2 == 0).mean() > 0.99                           # random 31-mers are not in a 5e8-k-mer index


def test_error_free_reads_never_name_another_genome(world):
    torch, nat = world["torch"], world["nat"]
    gen = torch.Generator(device=world["dev"]); gen.manual_seed(11)
    n = 1_000_000
    g = torch.randint(0, G, (n,), generator=gen, device=world["dev"])
    p = torch.randint(0, GL - RL + 1, (n,), generator=gen, device=world["dev"])
    idx = (g * GL + p)[:, None] + torch.arange(RL, device=world["dev"])[None, :]
    reads = world["bases"][idx]
    has_n = (reads == 78).any(dim=1)
    reads = reads.reshape(-1).contiguous()
    off = torch.arange(n + 1, device=world["dev"], dtype=torch.int64) * RL
    words, lst, _ = _align_device(world, reads, reads, off, n, nat.make_params(1, 1, None, None, None), False)
    types = (words >> 62) & 3
    lens = (words >> 40) & 0x3FFFFF
    payload = words & 0xFFFFFFFFFF
    clean = ~has_n
    assert bool((types[clean] >= 2).all())                                       # never unmapped
    uniq = clean & (types == 2)
    assert bool((payload[uniq] == g[uniq]).all()) and bool((lens[uniq] == 1).all())
    amb = clean & (types == 3)
    assert bool((lens[amb] == 0).all())                                          # no specific k-mer at all: AMBIGUOUS []
    assert int(uniq.sum()) > 0.6 * n


@pytest.mark.parametrize("extq", [False, True])
def test_consistency_across_call_paths_and_orders(world, extq):
    torch, nat, ix = world["torch"], world["nat"], world["ix"]
    mrq, mkq, mg = (62, 60, 3) if extq else (None, None, None)
    params = nat.make_params(1, 1, mrq, mkq, mg)
    rb, rq, roff = world["rb"], world["rq"], world["roff"]
    words, lst, state = _align_device(world, rb, rq, roff, NR, params, extq)
    tl = words >> 40                                     # type + length; payloads of length-1 results are genomes
    single = ((words >> 40) & 0x3FFFFF) == 1
    # statistics sum to n; dropped reads = the exact integer test of kmer.py:587
    types = (words >> 62) & 3
    counts = torch.bincount(types, minlength=4)
    assert int(counts.sum()) == NR
    if extq:
        qsum = rq.view(NR, RL).to(torch.int64).sum(dim=1)
        assert int(counts[0]) == int((qsum < mrq * RL).sum()) == int(state[2])
    else:
        assert int(counts[0]) == 0
    # host-buffer call, packed and raw, on the first 2M reads
    n = 2_000_000
    hb = rb[:n * RL].cpu().numpy(); hq = rq[:n * RL].cpu().numpy(); ho = roff[:n + 1].cpu().numpy().astype(np.uint64)
    ref_tl = tl[:n].cpu().numpy(); ref_w = words[:n].cpu().numpy(); ref_single = single[:n].cpu().numpy()
    for mode, chunk in (("1", "300000"), ("0", "700000"), ("2", "")):
        os.environ["PA_HOST_PACK"] = mode
        if chunk:
            os.environ["PA_CHUNK_READS"] = chunk
        else:
            os.environ.pop("PA_CHUNK_READS", None)
        try:
            w2, l2, c2 = ix.align(hb, hq if extq else None, ho, params)
        finally:
            os.environ.pop("PA_HOST_PACK", None); os.environ.pop("PA_CHUNK_READS", None)
        w2 = w2.view(np.int64)
        assert np.array_equal(w2 >> 40, ref_tl)
        assert np.array_equal(w2[ref_single], ref_w[ref_single])
    # a permutation of the reads permutes the results
    perm = torch.randperm(1_000_000, device=world["dev"], generator=torch.Generator(device=world["dev"]).manual_seed(3))
    prb = rb[:1_000_000 * RL].view(-1, RL)[perm].reshape(-1).contiguous()
    prq = rq[:1_000_000 * RL].view(-1, RL)[perm].reshape(-1).contiguous()
    poff = torch.arange(1_000_001, device=world["dev"], dtype=torch.int64) * RL
    pw, _, _ = _align_device(world, prb, prq, poff, 1_000_000, params, extq)
    assert bool(((pw >> 40) == tl[:1_000_000][perm]).all())
    s1 = (((pw >> 40) & 0x3FFFFF) == 1)
    assert bool((pw[s1] == words[:1_000_000][perm][s1]).all())


@pytest.mark.parametrize("name,filters", [("configs[1] plain", (None, None, None)), ("configs[2] extquality", (62, 60, 3))])
def test_all_reads_equal_the_oracle_digest(world, name, filters):
    """Every one of the 10^7 per-read results (type + ordered genome list) equals what the C oracle produced for this exact
    workload: tools/fullsize_parity.py ran both on the GPU box, diffed them read by read and committed the digest."""
    import json
    import synth
    gold = json.load(open(os.path.join(conftest.ROOT, "tests", "golden", "fullsize_digest.json")))
    wl = gold["workload"]
    assert (wl["genomes"], wl["genome_len"], wl["reads"], wl["read_len"], wl["k"], wl["genome_seed"], wl["read_seed"]) == (G, GL, NR, RL, K, 1000, 2000)
    case = gold["cases"][name]
    assert all(case["equal_to_oracle"].values())
    nat = world["nat"]
    mrq, mkq, mg = filters
    words, lst, state = _align_device(world, world["rb"], world["rq"], world["roff"], NR, nat.make_params(1, 1, mrq, mkq, mg), mrq is not None)
    n_list = int(state[0])
    types, lens, flat = nat.flatten_results(words.cpu().numpy().view(np.uint64), lst[:max(n_list, 1)].cpu().numpy().view(np.uint32))
    assert synth.result_digest(types, lens, flat) == case["oracle_digest_sha256"]
    assert [int(x) for x in state[2:5].cpu().tolist()] == case["counters"]
    inf = world["ix"].info()
    assert (int(inf.n_keys), int(inf.n_runs), int(inf.n_occ)) == (gold["index"]["n_keys"], gold["index"]["n_runs"], gold["index"]["n_occ"])
