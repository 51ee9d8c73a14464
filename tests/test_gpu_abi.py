"""
GPU parity tests through the raw C ABI (libpa_b200.so via _native.py) against the CPU oracle.
Bit-exact: everything on this path is integer / index work.
"""
import ctypes

import numpy as np
import pytest

import synth
import _native as nat
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

NAMES = {1: "UNMAPPED", 2: "UNIQUELY_MAPPED", 3: "AMBIGUOUSLY_MAPPED"}


def build_native(case):
    data, off = nat.pack_strings([g[1] for g in case["genomes"]])
    return nat.NativeIndex.build(data, off, case["k"])


def native_kmers_dict(ix):
    inf = ix.info()
    ex = ix.export()
    kmers = nat.decode_kmers(inf.k, ex["keys"])
    out = {}
    for u in ex["order"]:
        u = int(u)
        inner = {}
        for r in range(int(ex["run_off"][u]), int(ex["run_off"][u + 1])):
            inner[int(ex["run_genome"][r])] = [int(x) for x in ex["pos"][int(ex["pos_off"][r]):int(ex["pos_off"][r + 1])]]
        out[kmers[u]] = inner
    return out


def native_reads(ix, case, genome_ids):
    pr = case["params"]
    reads = case["reads"]
    seqs, off = nat.pack_strings([r[1] for r in reads])
    quals, _ = nat.pack_strings([r[2] for r in reads])
    words, lst, counters = ix.align(seqs, quals, off, nat.make_params(pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"]))
    types, lens, payload = nat.decode_words(words)
    out = {}
    for i, r in enumerate(reads):
        t = int(types[i])
        if t == 0:
            continue
        n = int(lens[i])
        gl = [int(payload[i])] if n == 1 else [int(x) for x in lst[int(payload[i]):int(payload[i]) + n]]
        out[r[0]] = {"mapping_type": NAMES[t], "genomes_mapped_to": [genome_ids[g] for g in gl]}
    return out, words, lst, counters


def check_case(case):
    pr = dict(case["params"])
    pr["filter_similar"] = False  # EXTSIM's host logic is covered through the Python shim tests
    o = orc.OracleReference(case["k"], case["genomes"])
    ix = build_native(case)
    try:
        want_kmers = o.kmers_dict()
        got_kmers = native_kmers_dict(ix)
        assert list(got_kmers.keys()) == list(want_kmers.keys())
        assert got_kmers == want_kmers
        inf = ix.info()
        assert (inf.n_keys, inf.n_runs, inf.n_occ) == o.sizes()
        # table lookups agree with the CSR for every present k-mer and for some absent ones
        if inf.k >= 1 and want_kmers:
            probe = list(want_kmers.keys())
            rng = np.random.default_rng(case.get("seed", 0))
            probe += ["".join(rng.choice(list("ACGT"), size=inf.k)) for _ in range(20)]
            ng, g0 = ix.table_lookup(probe)
            ranks = ix.lookup(probe)
            for i, km in enumerate(probe):
                inner = want_kmers.get(km)
                assert int(ng[i]) == (len(inner) if inner else 0), km
                assert (ranks[i] != nat.RANK_MISS) == (inner is not None)
                if inner:
                    assert int(g0[i]) == min(inner.keys())
        # alignment
        al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
        got_reads, words, lst, counters = native_reads(ix, case, [g[0] for g in case["genomes"]])
        want_reads = al.reads()
        assert list(got_reads.keys()) == list(want_reads.keys())
        assert got_reads == want_reads, (case.get("seed"), pr)
        assert int(counters[0]) == al.filtered_quality_reads
        assert int(counters[1]) == (al.filtered_quality_kmers if pr["mkq"] is not None else 0)
        assert int(counters[2]) == (al.filtered_hr_kmers if pr["mg"] is not None else 0)
        # summary (K8) against orc_summary on the oracle's own per-read output
        G = len(case["genomes"])
        stats, uniq, amb, first = ix.summary(words, lst)
        L = orc.lib()
        ostats = np.zeros(3, np.uint64); ou = np.zeros(max(G, 1), np.uint64); oa = np.zeros(max(G, 1), np.uint64)
        oorder = np.zeros(max(G, 1), np.uint32)
        n_seen = L.orc_summary(orc._ptr(al.types), orc._ptr(al.list_off), orc._ptr(al.genomes), len(case["reads"]), G,
                               orc._ptr(ostats), orc._ptr(ou), orc._ptr(oa), orc._ptr(oorder))
        assert [int(x) for x in stats[:3]] == [int(x) for x in ostats]
        assert int(stats[3]) == al.filtered_quality_reads
        assert np.array_equal(uniq, ou[:G]) and np.array_equal(amb, oa[:G])
        seen = [g for g in np.argsort(first, kind="stable") if first[g] != np.uint64(0xFFFFFFFFFFFFFFFF)]
        assert [int(g) for g in seen] == [int(g) for g in oorder[:n_seen]]
    finally:
        ix.close()


@pytest.mark.parametrize("n,end_bit", [(0, 64), (1, 64), (31, 8), (4095, 64), (4096, 17), (4097, 64), (100_003, 63),
                                        (1_000_000, 40), (3_000_001, 64)])
def test_radix_sort_pairs(n, end_bit):
    rng = np.random.default_rng(n + end_bit)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    if n > 10:
        keys[rng.integers(0, n, size=n // 3)] = keys[0]          # heavy duplicates
        keys[rng.integers(0, n, size=n // 50 + 1)] = np.uint64(0xFFFFFFFFFFFFFFFF)
    if end_bit < 64:
        keys &= np.uint64((1 << end_bit) - 1)
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = nat.debug_sort_pairs(keys, vals, end_bit)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])      # stability


def _check_hashed_sort(keys, end_bit, top_bits, expect_fallback=None):
    vals = np.arange(len(keys), dtype=np.uint32)
    k2, v2, fb = nat.debug_sort_pairs_hashed(keys, vals, end_bit, top_bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])      # stability inside equal keys
    if expect_fallback is not None:
        assert fb == expect_fallback
    return fb


@pytest.mark.parametrize("n,end_bit,top_bits", [(2, 63, 8), (1000, 63, 0), (4097, 62, 16), (100_003, 63, 24), (1_000_000, 63, 24),
                                                 (1_000_000, 63, 0), (3_000_001, 62, 32), (2_000_000, 40, 24)])
def test_hashed_sort_repairs_mixed_runs(n, end_bit, top_bits):
    """Sort on the top bits + repair = the full stable sort: spread keys with duplicates (the occurrences of a k-mer)."""
    rng = np.random.default_rng(n + end_bit + top_bits)
    keys = rng.integers(0, 1 << (end_bit - 1), size=n, dtype=np.uint64)
    if n > 10:
        dup = rng.integers(0, n, size=n // 2)
        keys[dup] = keys[rng.integers(0, n, size=n // 2)]                      # many repeated keys
        keys[rng.integers(0, n, size=n // 50 + 1)] = np.uint64((1 << end_bit) - 1)  # the sentinel of invalid windows
    _check_hashed_sort(keys, end_bit, top_bits)


def test_hashed_sort_long_runs_and_fallback():
    rng = np.random.default_rng(77)
    end_bit, top_bits = 63, 16
    low = end_bit - top_bits
    # (a) a k-mer with 50,000 occurrences sharing its top bits with three strangers, interleaved: one big group
    top = np.uint64(0x1234) << np.uint64(low)
    a, b, c, d = (top | np.uint64(x) for x in (500, 20, 900, 7))
    group = np.full(50_000, a, dtype=np.uint64)
    group[[0, 17, 30_000, 49_999]] = [c, b, d, b]
    rest = rng.integers(0, 1 << 62, size=200_000, dtype=np.uint64)
    rest = rest[(rest >> np.uint64(low)) != np.uint64(0x1234)]
    keys = np.concatenate([rest[:100_000], group[:25_000], rest[100_000:], group[25_000:]])
    assert _check_hashed_sort(keys, end_bit, top_bits) in (False, True)
    # (b) two heavy keys interleaved record by record (every position is a boundary)
    inter = np.where(np.arange(40_000) % 2 == 0, a, b).astype(np.uint64)
    keys = np.concatenate([rest[:50_000], inter, rest[50_000:]])
    _check_hashed_sort(keys, end_bit, top_bits)
    # (c) keys that are not spread at all (all in one run of top bits, thousands of distinct keys): the fallback passes
    dense = rng.integers(0, 1 << 20, size=300_000, dtype=np.uint64)
    _check_hashed_sort(dense, end_bit, top_bits, expect_fallback=True)
    # (d) nothing to repair
    spread = (np.arange(100_000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64((1 << 62) - 1)
    _check_hashed_sort(spread, end_bit, 32, expect_fallback=False)


@pytest.mark.parametrize("block", range(6))
def test_fuzz_small_k_against_oracle(block):
    for seed in range(block * 100, block * 100 + 100):
        check_case(synth.fuzz_case(seed))


def test_summary_global_counter_path(monkeypatch):
    """K8 counts per genome in shared memory; indexes with more genomes than the shared counters hold use global atomics."""
    monkeypatch.setenv("PA_SUMMARY_GLOBAL", "1")
    for seed in range(8000, 8040):
        check_case(synth.fuzz_case(seed))


def test_fuzz_wider_k():
    for seed in range(5000, 5080):
        check_case(synth.fuzz_case(seed, k_range=(9, 31), max_genomes=8))


def test_k31_moderate():
    genomes = synth.make_genomes(6, 40_000, seed=5, cluster_size=3, shared_frac=0.35, n_every=9000, n_run=11)
    b, q, off = synth.make_reads(genomes, 4000, 150, seed=6, sub_rate=0.02, random_frac=0.05)
    for pr in [dict(m=1, p=1, mrq=None, mkq=None, mg=None), dict(m=1, p=1, mrq=62, mkq=60, mg=1),
               dict(m=0, p=0, mrq=None, mkq=63, mg=3), dict(m=3, p=-1, mrq=None, mkq=None, mg=2)]:
        check_case({"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                    "params": pr, "seed": 1})


@pytest.mark.parametrize("masks", ["1", "0"])
def test_quality_filters_by_masks_and_in_kernel(masks, monkeypatch):
    """EXTQUALITY has two implementations: quality_masks_kernel + the staggered K4 reading one bit per window (default), and
    the K4 variant that scans the quality bytes itself (PA_QUAL_MASKS=0).  Both against the oracle: read drop, window filter
    counted per occurrence, together with max-genomes, on reads shorter than k, of ragged lengths and longer than 128 windows."""
    monkeypatch.setenv("PA_QUAL_MASKS", masks)
    genomes = synth.make_genomes(6, 40_000, seed=5, cluster_size=3, shared_frac=0.35, n_every=9000, n_run=11)
    for read_len, n in ((150, 1500), (40, 300), (31, 100), (20, 50), (158, 200), (159, 200), (300, 150)):
        b, q, off = synth.make_reads(genomes, n, read_len, seed=60 + read_len, sub_rate=0.02, random_frac=0.05)
        for pr in [dict(m=1, p=1, mrq=62, mkq=60, mg=1), dict(m=0, p=0, mrq=None, mkq=63, mg=3),
                   dict(m=2, p=-1, mrq=61, mkq=None, mg=None)]:
            check_case({"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                        "params": pr, "seed": 1})
    for seed in range(9300, 9340):   # small k, ragged reads, random thresholds
        check_case(synth.fuzz_case(seed))


@pytest.mark.parametrize("dense", ["2", "3", "5"])
def test_overloaded_table_walks_chains_and_stash(dense, monkeypatch):
    """PA_TABLE_DENSE doubles the load factor per step: buckets overflow into the next blocks (CONT) and into the stash,
    many strain variants share minimizer and offset -- results must not change."""
    monkeypatch.setenv("PA_TABLE_DENSE", dense)
    genomes = synth.make_genomes(12, 30_000, seed=15, cluster_size=6, shared_frac=0.6, sub_rate=0.03, n_every=9000, n_run=11)
    b, q, off = synth.make_reads(genomes, 3000, 150, seed=16, sub_rate=0.02, random_frac=0.05)
    case = {"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
            "params": dict(m=1, p=1, mrq=None, mkq=None, mg=None), "seed": 2}
    data, goff = nat.pack_strings([g[1] for g in case["genomes"]])
    ix = nat.NativeIndex.build(data, goff, 31)
    inf = ix.info()
    ix.close()
    if dense == "5":
        assert inf.stash_count > 0, "the test is meant to reach the stash"
    check_case(case)
    check_case(dict(case, params=dict(m=2, p=0, mrq=None, mkq=60, mg=4)))
    for k in (7, 13, 20):   # w = 1 (k <= 16) and w = 5: short minimizer windows
        check_case(dict(case, k=k))


@pytest.mark.parametrize("mode", ["1", "0"])
def test_host_packed_reads_give_the_same_results(mode, monkeypatch):
    """PA_HOST_PACK=1: every chunk of pa_align_batch is turned into 2-bit planes on the host and aligned by the packed
    kernel variants (fast + general, plain + EXTQUALITY, long reads over several super-rounds); a chunk holding a base
    outside ACGT falls back to ASCII.  PA_CHUNK_READS makes the batches span several chunks and both slots."""
    monkeypatch.setenv("PA_HOST_PACK", mode)
    monkeypatch.setenv("PA_CHUNK_READS", "700")
    genomes = synth.make_genomes(6, 40_000, seed=5, cluster_size=3, shared_frac=0.35, n_every=9000, n_run=11)
    b, q, off = synth.make_reads(genomes, 4000, 150, seed=6, sub_rate=0.02, random_frac=0.05)
    reads = synth.reads_as_triples(b, q, off)
    reads[1234] = (reads[1234][0], reads[1234][1][:70] + "N" + reads[1234][1][71:], reads[1234][2])   # one ASCII chunk
    for pr in [dict(m=1, p=1, mrq=None, mkq=None, mg=None), dict(m=1, p=1, mrq=62, mkq=60, mg=1),
               dict(m=0, p=0, mrq=None, mkq=63, mg=3)]:
        check_case({"k": 31, "genomes": synth.genomes_as_pairs(genomes), "reads": reads, "params": pr, "seed": 1})
    lg = synth.make_genomes(5, 6000, seed=9, cluster_size=5, shared_frac=0.5, n_every=2500, n_run=5)
    lb, lq, loff = synth.make_reads(lg, 300, 700, seed=10, sub_rate=0.03, random_frac=0.1)
    for k in (5, 12, 31):
        check_case({"k": k, "genomes": synth.genomes_as_pairs(lg), "reads": synth.reads_as_triples(lb, lq, loff),
                    "params": dict(m=2, p=0, mrq=60, mkq=61, mg=2), "seed": k})
    for seed in range(40):   # ragged read lengths, tiny k
        check_case(synth.fuzz_case(900 + seed))


def test_prepacked_reads_through_pa_align_batch_packed(monkeypatch):
    """pa_pack_reads once, then pa_align_batch_packed: same words, lists and counters as pa_align_batch on the ASCII reads --
    fixed-length and ragged reads, plain and EXTQUALITY, long reads, several chunks (chunk starts that are not multiples
    of 32 bases inside the batch-wide plane layout); a batch with a base outside ACGT is reported by pa_pack_reads."""
    import _native as nat
    monkeypatch.setenv("PA_CHUNK_READS", "333")
    genomes = synth.make_genomes(6, 40_000, seed=15, cluster_size=3, shared_frac=0.35, n_every=9000, n_run=11)
    data, goff = nat.pack_strings([s for _, s in synth.genomes_as_pairs(genomes)])
    ix = nat.NativeIndex.build(data, goff, 31)
    rng = np.random.default_rng(16)
    for trial, (n, L, ragged) in enumerate([(3000, 150, False), (2500, 101, False), (1500, 150, True), (200, 700, False)]):
        b, q, off = synth.make_reads(genomes, n, L, seed=17 + trial, sub_rate=0.02, random_frac=0.05)
        if ragged:   # cut every read to a random length: offsets stop being an arithmetic sequence
            lens = rng.integers(1, L + 1, size=n)
            keep = np.concatenate([np.arange(int(o), int(o) + int(l)) for o, l in zip(off[:-1], lens)])
            b, q = b[keep], q[keep]
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        shift = 7 * trial                      # offsets need not start at 0
        bb = np.concatenate([np.full(shift, ord("A"), np.uint8), b]); qq = np.concatenate([np.full(shift, 33, np.uint8), q])
        off = off + np.uint64(shift)
        planes, ok = nat.pack_reads(bb, off)
        assert ok
        for pr in [(None, None, None), (62, 60, 1), (None, 63, 3)]:
            params = nat.make_params(1, 1, *pr)
            quals = qq if (pr[0] is not None or pr[1] is not None) else None
            w1, l1, c1 = ix.align(bb, quals, off, params)
            w2, l2, c2 = ix.align_packed(planes, quals, off, params)
            t1, n1, f1 = nat.flatten_results(w1, l1)
            t2, n2, f2 = nat.flatten_results(w2, l2)
            assert np.array_equal(t1, t2) and np.array_equal(n1, n2) and np.array_equal(f1, f2)
            assert np.array_equal(c1, c2)
    bad = bb.copy()
    bad[int(off[5]) + 3] = ord("N")
    assert nat.pack_reads(bad, off)[1] is False
    ix.close()


def test_long_reads_take_the_multi_round_path():
    genomes = synth.make_genomes(5, 6000, seed=9, cluster_size=5, shared_frac=0.5, n_every=2500, n_run=5)
    b, q, off = synth.make_reads(genomes, 300, 700, seed=10, sub_rate=0.03, random_frac=0.1)
    for k in (5, 12, 31):
        for pr in [dict(m=1, p=1, mrq=None, mkq=None, mg=None), dict(m=2, p=0, mrq=60, mkq=61, mg=2)]:
            check_case({"k": k, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off),
                        "params": pr, "seed": k})


def test_many_genomes_use_the_global_scratch():
    rng = np.random.default_rng(3)
    base = synth.ACGT[rng.integers(0, 4, size=300)]
    genomes = []
    for g in range(300):   # > 256 genomes: per-genome table in global memory; most k-mers shared by many genomes
        seq = base.copy()
        idx = rng.integers(0, seq.size, size=6)
        seq[idx] = synth.ACGT[rng.integers(0, 4, size=6)]
        genomes.append((f"g{g}", seq.tobytes().decode()))
    reads = []
    for r in range(60):
        src = genomes[int(rng.integers(0, 300))][1]
        s = int(rng.integers(0, 200))
        reads.append((f"r{r}", src[s:s + 90], "I" * 90))
    for pr in [dict(m=1, p=1, mrq=None, mkq=None, mg=None), dict(m=1, p=1, mrq=None, mkq=None, mg=50)]:
        check_case({"k": 11, "genomes": genomes, "reads": reads, "params": pr, "seed": 3})


def test_extsim_kernels_against_oracle():
    for seed in list(range(7000, 7060)) + list(range(10_000, 10_040)):
        case = synth.fuzz_case(seed, dup_ids=seed >= 10_000)
        o = orc.OracleReference(case["k"], case["genomes"])
        ix = build_native(case)
        try:
            ids = [g[0] for g in case["genomes"]]
            classes = {}
            group = np.array([classes.setdefault(s, len(classes)) for s in ids], dtype=np.uint32)
            n = len(classes)
            total, uniq = ix.extsim_stats(group, n)
            inter = ix.extsim_pairwise(group, n)
            L = orc.lib()
            ot = np.zeros(n, np.uint64); ou = np.zeros(n, np.uint64); oi = np.zeros(n * n, np.uint64)
            L.orc_extsim_stats(o._h, orc._ptr(group), n, orc._ptr(ot), orc._ptr(ou))
            L.orc_extsim_pairwise(o._h, orc._ptr(group), n, orc._ptr(oi))
            assert np.array_equal(total, ot) and np.array_equal(uniq, ou)
            assert np.array_equal(inter.reshape(-1), oi)
            # drop a pseudo-random subset of genomes and compare the surviving index
            rng = np.random.default_rng(seed)
            keep = (rng.random(len(ids)) < 0.6).astype(np.uint8)
            ix.drop_genomes(keep)
            h2 = L.orc_index_drop_genomes(o._h, orc._ptr(np.concatenate([keep, [0]]).astype(np.uint8)))
            L.orc_index_free(o._h)
            o._h = h2
            o.genomes = [g for g, kp in zip(o.genomes, keep) if kp]
            assert native_kmers_dict(ix) == o.kmers_dict()
            assert list(native_kmers_dict(ix).keys()) == list(o.kmers_dict().keys())
            case2 = dict(case, genomes=o.genomes)
            if o.genomes:
                pr = case["params"]
                al = o.align(case["reads"], pr["m"], pr["p"], pr["mrq"], pr["mkq"], pr["mg"])
                got, _, _, _ = native_reads(ix, case2, [g[0] for g in o.genomes])
                assert got == al.reads()
        finally:
            ix.close()


@pytest.mark.parametrize("env", [{}, {"PA_K6_SETS": "0"}, {"PA_K6_SETS": "4"}, {"PA_K6_WEAK_HASH": "1"}])
def test_extsim_pairwise_paths(env, monkeypatch):
    """K6 counts class lists through a hash table; without the table, with a table far too small (probe windows full) and with
    a hash under which all lists collide (detected, then redone pair by pair) the matrix must be the same."""
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    for seed in list(range(7100, 7115)) + list(range(10_100, 10_115)):
        case = synth.fuzz_case(seed, dup_ids=seed >= 10_000)
        o = orc.OracleReference(case["k"], case["genomes"])
        ix = build_native(case)
        try:
            classes = {}
            group = np.array([classes.setdefault(g[0], len(classes)) for g in case["genomes"]], dtype=np.uint32)
            n = len(classes)
            oi = np.zeros(n * n, np.uint64)
            orc.lib().orc_extsim_pairwise(o._h, orc._ptr(group), n, orc._ptr(oi))
            assert np.array_equal(ix.extsim_pairwise(group, n).reshape(-1), oi)
        finally:
            ix.close()


def test_extsim_pairwise_long_genome_lists():
    """k-mers shared by more genomes than a thread takes alone (the warp-cooperative branch of K6), with and without repeated
    identifier classes."""
    rng = np.random.default_rng(77)
    genomes = [(f"g{g}", "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(30, 400))))) for g in range(45)]
    for k in (3, 5):
        o = orc.OracleReference(k, genomes)
        ix = build_native({"k": k, "genomes": genomes})
        try:
            for n, group in ((45, np.arange(45, dtype=np.uint32)), (7, (np.arange(45) % 7).astype(np.uint32)),
                             (20, rng.integers(0, 20, 45).astype(np.uint32))):
                total, uniq = ix.extsim_stats(group, n)
                inter = ix.extsim_pairwise(group, n)
                L = orc.lib()
                ot = np.zeros(n, np.uint64); ou = np.zeros(n, np.uint64); oi = np.zeros(n * n, np.uint64)
                L.orc_extsim_stats(o._h, orc._ptr(group), n, orc._ptr(ot), orc._ptr(ou))
                L.orc_extsim_pairwise(o._h, orc._ptr(group), n, orc._ptr(oi))
                assert np.array_equal(total, ot) and np.array_equal(uniq, ou)
                assert np.array_equal(inter.reshape(-1), oi)
        finally:
            ix.close()


def test_bad_genome_character_is_rejected():
    data, off = nat.pack_strings(["ACGTNNACGT", "ACGXACGT"])
    with pytest.raises(ValueError):
        nat.NativeIndex.build(data, off, 3)


def test_k_above_31_is_out_of_scope():
    data, off = nat.pack_strings(["ACGT" * 20])
    with pytest.raises(ValueError):
        nat.NativeIndex.build(data, off, 32)


def test_builds_on_two_devices_in_one_process():
    """The > 48 KB shared-memory opt-in of the sort and scatter kernels is per device (a static flag made the
    first build on a second device fail); the buffer cache is per device too."""
    import ctypes
    n = ctypes.c_int32(0)
    nat.lib().pa_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    genomes = synth.make_genomes(4, 30_000, seed=51, cluster_size=2, shared_frac=0.3, n_every=9000, n_run=7)
    data, goff = nat.pack_strings([s for _, s in synth.genomes_as_pairs(genomes)])
    b, q, off = synth.make_reads(genomes, 2000, 150, seed=52, sub_rate=0.01, random_frac=0.03)
    results = []
    for device in (0, 1, 0, 1):
        ix = nat.NativeIndex.build(data, goff, 31, device=device)
        w, l, c = ix.align(b, q, off, nat.make_params(1, 1, 60, 58, 2))
        results.append((ix.info().n_keys, nat.flatten_results(w, l), [int(x) for x in c]))
        ix.close()
    for r in results[1:]:
        assert r[0] == results[0][0] and r[2] == results[0][2]
        assert all(np.array_equal(a, b_) for a, b_ in zip(r[1], results[0][1]))
    nat.trim_memory()
