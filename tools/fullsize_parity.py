#!/usr/bin/env python3
"""
fullsize_parity.py -- configs[1] and configs[2] at FULL size, GPU path against the C oracle, read by read.

Runs on the GPU box (python tools/fullsize_parity.py [--out tests/golden/fullsize_digest.json]).  The workload is the
one bench.py times (same device generator, same seeds: genomes 1000, reads 2000): 100 genomes x 5 Mb, 10^7 x 150 bp
reads, k = 31.  The device-generated genomes and reads are copied to the host ONCE, the oracle (oracle/pa_oracle.c,
OpenMP over reads) builds the same index and aligns the same reads, and every per-read result (type, ordered genome
list), the three filter counters, the index sizes and the K8 summary are compared.  On success a digest of the
per-read results is written; bench.py's parity leg re-checks its own results against that file, and
tests/test_gpu_fullsize.py does the same through the C ABI.
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=100)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench
    import synth
    import _native as nat
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    G, GL, NR, RL, k = args.genomes, args.genome_len, args.reads, args.read_len, args.k
    bases = bench.device_genomes(torch, dev, G, GL, seed=1000)
    goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL)).astype(np.uint64)
    rb, rq, roff = bench.device_reads(torch, dev, bases, G, GL, NR, RL, seed=2000)
    torch.cuda.synchronize()
    ix = nat.NativeIndex.build_device(bases.data_ptr(), goff, k, device=0)
    inf = ix.info()
    h_genomes = bases.cpu().numpy()
    h_b, h_q, h_off = rb.cpu().numpy(), rq.cpu().numpy(), roff.cpu().numpy().astype(np.uint64)
    del bases, rb, rq, roff
    nthreads = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    pairs = [(f"genome{g}", h_genomes[g * GL:(g + 1) * GL].tobytes().decode("ascii")) for g in range(G)]
    o = orc.OracleReference(k, pairs)
    t_build = time.perf_counter() - t0
    sizes_ok = (inf.n_keys, inf.n_runs, inf.n_occ) == o.sizes()
    out = {"workload": {"genomes": G, "genome_len": GL, "reads": NR, "read_len": RL, "k": k, "genome_seed": 1000, "read_seed": 2000},
           "index": {"n_keys": int(inf.n_keys), "n_runs": int(inf.n_runs), "n_occ": int(inf.n_occ), "equal_to_oracle": bool(sizes_ok)},
           "oracle": {"threads": nthreads, "build_s": t_build}, "cases": {}}
    ok_all = sizes_ok
    for name, (mrq, mkq, mg) in {"configs[1] plain": (None, None, None), "configs[2] extquality": (62, 60, 3)}.items():
        params = nat.make_params(1, 1, mrq, mkq, mg)
        words, lst, counters = ix.align(h_b, h_q if mrq is not None else None, h_off, params)
        stats, uniq, amb, first = ix.summary(words, lst)
        t0 = time.perf_counter()
        al = o.align_packed([], h_b, h_q, h_off, 1, 1, mrq, mkq, mg, nthreads=nthreads)
        t_al = time.perf_counter() - t0
        types, lens, flat = synth.flatten_results(words, lst)
        want_lens = np.diff(al.list_off.astype(np.int64)).astype(np.int32)
        same_types = bool(np.array_equal(types, al.types))
        same_lens = bool(np.array_equal(lens, want_lens))
        same_lists = bool(same_lens and np.array_equal(flat, al.genomes[:len(flat)]))
        want_c = [al.filtered_quality_reads, al.filtered_quality_kmers if mkq is not None else 0,
                  al.filtered_hr_kmers if mg is not None else 0]
        same_counters = [int(c) for c in counters] == want_c
        # K8 against the oracle's summary (genome ids are their indices here)
        al.read_ids = [""] * NR
        osum = al.get_summary()
        gsum = {}
        never = np.uint64(0xFFFFFFFFFFFFFFFF)
        for g in np.argsort(first, kind="stable"):
            if first[g] == never:
                break
            gsum[f"genome{int(g)}"] = {"unique_reads": int(uniq[g]), "ambiguous_reads": int(amb[g])}
        same_summary = (json.dumps(gsum) == json.dumps(osum["Summary"]) and
                        [int(stats[0]), int(stats[1]), int(stats[2])] ==
                        [osum["Statistics"]["unique_mapped_reads"], osum["Statistics"]["ambiguous_mapped_reads"],
                         osum["Statistics"]["unmapped_reads"]])
        ok = same_types and same_lens and same_lists and same_counters and same_summary
        ok_all = ok_all and ok
        out["cases"][name] = {
            "params": {"m": 1, "p": 1, "mrq": mrq, "mkq": mkq, "mg": mg},
            "equal_to_oracle": {"types": same_types, "list_lengths": same_lens, "ordered_lists": same_lists,
                                "counters": same_counters, "summary_json": same_summary},
            "digest_sha256": synth.result_digest(types, lens, flat),
            "oracle_digest_sha256": synth.result_digest(al.types, want_lens, al.genomes[:int(al.list_off[-1])]),
            "stats": {"unique": int(stats[0]), "ambiguous": int(stats[1]), "unmapped": int(stats[2]), "dropped": int(stats[3])},
            "counters": [int(c) for c in counters], "multi_genome_lists": int((lens > 1).sum()),
            "oracle_align_s": t_al, "oracle_reads_per_s": NR / t_al,
        }
        print(name, json.dumps(out["cases"][name]), flush=True)
    out["all_equal"] = bool(ok_all)
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    ix.close()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
