#!/usr/bin/env python3
"""
bench.py -- reads/s pseudo-aligned (k=31, 150 bp) on B200, with the reference-build k-mers/s beside it.

    python bench.py --gpus N --steps K --warmup W            # this implementation (one rank per GPU, weak scaling)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on all host cores

A step = one pass of the hot path over one batch of synthetic reads: K4 (align) + K8 (summary reduction) per
rank, plus one NCCL all-reduce of the per-genome summary when N > 1 (pa_comm_allreduce_summary).  Workload =
BASELINE.json configs[1] ("1xB200: 100 synthetic bacterial genomes (~5 Mb each), 10M 150-bp reads, k=31, plain
pseudo-alignment"); `--extquality` makes configs[2] the headline instead.  Inputs are synthetic (seeded, generated on
the device) and far larger than L2 (1.5 GB of reads against a multi-GB table), so no explicit L2 flush is needed
between iterations.  At N > 1 the index the timed steps align against is the replica of the PARTITIONED build
(pa_index_build_partitioned), and its content is checked against the single-GPU build (CSR checksums of the
partitions, per-read results).

Next to the headline the line carries one sub-record per other BASELINE config, each driver-timed with its own roofline:
  "configs"."extquality"   configs[2]: the same reads with min-read-quality / min-kmer-quality / max-genomes on
  "configs"."extsim"       configs[3]: EXTSIM build, 1,000 genomes in clusters of 10 at ~99 % identity
  "configs"."config_e"     configs[4]: 2,000-genome index (table-only, built partitioned over the ranks or, on one GPU,
                           streamed in rounds), reads/s device-resident and end to end, bytes of index per k-mer
(--no-configs skips them).

Prints ONE JSON line (see the keys at the bottom).  Only the `cpu_baseline` leg and `--impl reference` touch
oracle/ (the CPU checker); the timed GPU path goes through the C ABI of libpa_b200.so only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

ALG_BYTES_PLAIN = 3998      # per 150-bp read at k=31: 150 bases + 32 B x 120 lookups + 8 B result
ALG_BYTES_QUAL = 4148       # + 150 quality bytes
BUILD_BYTES_PER_KMER = 17   # 1 base in + 16 B record out


def random_line_peak():
    """Distinct 128-byte lines per second a B200 serves under uniformly random 32-byte requests into a 16 GiB table: the
    random-access lookup roofline, read from the committed microbenchmark (tools/gather_bench.cu)."""
    best, src = None, os.path.join(ROOT, "profiles", "r01_gather_roofline.jsonl")
    try:
        for line in open(src):
            line = line.strip()
            if not line.startswith("{"):
                continue
            r = json.loads(line)
            if r.get("table_GiB") == 16.0 and r.get("mode") == "nc32(256b)":
                best = max(best or 0.0, float(r["lookups_per_s"]))
    except OSError:
        pass
    return (best, "profiles/r01_gather_roofline.jsonl (16 GiB table, 256-bit loads)") if best else (4.36e10, "fallback constant")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=100)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("--extquality", action="store_true", help="configs[2]: min-read-quality 62, min-kmer-quality 60, max-genomes 3")
    ap.add_argument("--sample-genomes", type=int, default=4, help="CPU baseline / parity sample: genomes")
    ap.add_argument("--sample-reads", type=int, default=1_000_000, help="CPU baseline / parity sample: reads")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] / [3] / [4] sub-records")
    ap.add_argument("--e-genomes", type=int, default=2000, help="configs[4]: genomes of the index")
    ap.add_argument("--e-genome-len", type=int, default=2_000_000, help="configs[4]: bases per genome")
    ap.add_argument("--e-reads", type=int, default=10_000_000, help="configs[4]: reads per GPU per step")
    ap.add_argument("--d-genomes", type=int, default=1000, help="configs[3]: genomes (clusters of 10)")
    ap.add_argument("--d-genome-len", type=int, default=1_000_000, help="configs[3]: bases per genome")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# synthetic workload on the device (torch is plumbing: allocation, RNG, streams, NCCL)
# ---------------------------------------------------------------------------
def device_genomes(torch, dev, n_genomes, length, seed, cluster=4, shared_frac=0.3, sub=0.01, n_every=1_000_000, n_run=40):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    bases = torch.empty(n_genomes * length, dtype=torch.uint8, device=dev)
    blk = int(length * shared_frac)
    anc = None
    for g in range(n_genomes):
        seq = lut[torch.randint(0, 4, (length,), generator=gen, device=dev)]
        if blk > 0 and cluster > 1:
            if g % cluster == 0 or anc is None:
                anc = lut[torch.randint(0, 4, (blk,), generator=gen, device=dev)]
            shared = anc.clone()
            hit = torch.rand(blk, generator=gen, device=dev) < sub
            shared[hit] = lut[torch.randint(0, 4, (int(hit.sum().item()),), generator=gen, device=dev)]
            start = int(torch.randint(0, length - blk + 1, (1,), generator=gen, device=dev).item())
            seq[start:start + blk] = shared
        if n_every > 0:
            for at in range(n_every // 2, length - n_run, n_every):
                seq[at:at + n_run] = 78  # 'N'
        bases[g * length:(g + 1) * length] = seq
    return bases


def device_reads(torch, dev, bases, n_genomes, length, n_reads, read_len, seed, sub=0.01, random_frac=0.02, chunk=1_000_000):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    quals = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for lo in range(0, n_reads, chunk):
        n = min(chunk, n_reads - lo)
        which = torch.randint(0, n_genomes, (n,), generator=gen, device=dev)
        start = (torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (length - read_len + 1)).long()
        idx = (which * length + start)[:, None] + ar[None, :]
        r = bases[idx]
        r[r == 78] = 65
        hit = torch.rand(n, read_len, generator=gen, device=dev) < sub
        r[hit] = lut[torch.randint(0, 4, (int(hit.sum().item()),), generator=gen, device=dev)]
        rnd = torch.rand(n, generator=gen, device=dev) < random_frac
        nr = int(rnd.sum().item())
        if nr:
            r[rnd] = lut[torch.randint(0, 4, (nr, read_len), generator=gen, device=dev)]
        out[lo * read_len:(lo + n) * read_len] = r.reshape(-1)
        q = torch.clamp(torch.round(torch.randn(n * read_len, generator=gen, device=dev) * 8.0 + 30.0), 2, 41) + 33
        quals[lo * read_len:(lo + n) * read_len] = q.to(torch.uint8)
    off = torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * read_len
    return out, quals, off


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and s[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU arm (oracle port) -- also the parity checker of the GPU arm's sample
# ---------------------------------------------------------------------------
def cpu_sample_workload(args):
    """The bounded sample: `sample_genomes` genomes of the workload's shape and `sample_reads` reads from them."""
    import synth
    genomes = synth.make_genomes(args.sample_genomes, args.genome_len, seed=4242, cluster_size=4, shared_frac=0.3,
                                 sub_rate=0.01, n_every=1_000_000, n_run=40)
    b, q, off = synth.make_reads(genomes, args.sample_reads, args.read_len, seed=4243, sub_rate=0.01, random_frac=0.02)
    return genomes, b, q, off


def filters(args):
    return (62, 60, 3) if args.extquality else (None, None, None)


def host_threads():
    """The host cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which is advice for its workers, not a
    statement about the machine: the CPU arm uses every core of the affinity mask."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_cpu_oracle(args, genomes, b, q, off, steps=1, warmup=0):
    from oracle import oracle as orc
    import synth
    nthreads = host_threads()
    t0 = time.perf_counter()
    o = orc.OracleReference(args.k, synth.genomes_as_pairs(genomes))
    t_build = time.perf_counter() - t0
    mrq, mkq, mg = filters(args)
    ids = None
    times = []
    al = None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        al = o.align_packed(ids or [], b, q, off, 1, 1, mrq, mkq, mg, nthreads=nthreads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    n_occ = o.sizes()[2]
    return o, al, {"build_s": t_build, "build_kmers_per_s": n_occ / t_build if t_build > 0 else None,
                   "align_s": float(np.mean(times)), "reads_per_s": (len(off) - 1) / float(np.mean(times)), "threads": nthreads}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    genomes, b, q, off = cpu_sample_workload(args)
    _, _, r = run_cpu_oracle(args, genomes, b, q, off, steps=args.steps, warmup=args.warmup)
    sample = (f"{args.sample_genomes} genomes x {args.genome_len} bp index, {args.sample_reads} x {args.read_len} bp reads of "
              f"the same generator; oracle/pa_oracle.c (C port of kmer.py), OpenMP over reads")
    cfg = workload_config(args)
    # the CPU arm runs a bounded SAMPLE of the workload (the oracle needs ~70 s for the full index): say so where the sizes are
    cfg.update({"workload": cfg["workload"] + f" -- CPU arm on a bounded sample: {args.sample_genomes} genomes x {args.genome_len} bp, "
                                              f"{args.sample_reads} reads per step",
                "genomes": args.sample_genomes, "reads_per_gpu": args.sample_reads, "reads_per_step": args.sample_reads,
                "full_workload": {"genomes": args.genomes, "genome_len": args.genome_len, "reads_per_gpu": args.reads},
                "parallelism": f"{r['threads']} host threads (OpenMP over reads)"})
    line = {
        "impl": "reference", "metric": "reads/s pseudo-aligned (k=31,150bp)", "value": r["reads_per_s"], "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["align_s"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["reads_per_s"], "unit": "reads/s", "cores": r["threads"], "kind": "port", "sample": sample,
                         "build_kmers_per_s": r["build_kmers_per_s"]},
        "e2e": {"value": r["reads_per_s"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    name = "configs[2]: configs[1] + EXTQUALITY (mrq=62,mkq=60,mg=3)" if args.extquality else \
        "configs[1]: 100 genomes x 5 Mb, 10M x 150 bp reads, k=31, plain"
    return {"workload": name, "genomes": args.genomes, "genome_len": args.genome_len, "reads_per_gpu": args.reads,
            "read_len": args.read_len, "k": args.k, "m": 1, "p": 1, "l2": "inputs larger than L2 (no flush needed)",
            "parallelism": f"reads sharded over {args.gpus} GPU(s), index replicated"}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class AlignRunner:
    """One rank's alignment step on device-resident reads: K4 + K8 (+ the summary all-reduce when there are several ranks),
    all through the C ABI on one dedicated stream (CUDA events on that stream see exactly these launches)."""

    def __init__(self, torch, nat, ix, comm, rank, world, dev, stream, G, rbases, rquals, roff, NR, RL, params, need_q):
        self.torch, self.nat, self.ix, self.comm, self.rank, self.world, self.dev, self.stream = torch, nat, ix, comm, rank, world, dev, stream
        self.G, self.rbases, self.rquals, self.roff, self.NR, self.RL, self.params, self.need_q = G, rbases, rquals, roff, NR, RL, params, need_q
        self.L = nat.lib()
        self.sptr = ctypes.c_void_p(stream.cuda_stream)
        self.words = torch.empty(NR, dtype=torch.int64, device=dev)
        self.list_cap = max(NR // 2, 1024)
        self.lst = torch.empty(self.list_cap, dtype=torch.int32, device=dev)
        self.state = torch.zeros(5, dtype=torch.int64, device=dev)
        # SUM block: stats[4], unique[G], ambiguous[G]; MIN block: first_seen[G] (uint64 max = never)
        self.acc = torch.zeros(4 + 2 * G, dtype=torch.int64, device=dev)
        self.first_seen = torch.full((max(G, 1),), -1, dtype=torch.int64, device=dev)
        self.n_launch = ctypes.c_int32(0)
        self.launches = 0

    def align(self):
        nat, L = self.nat, self.L
        nat.check(L.pa_align_batch_device(self.ix.handle, ctypes.c_void_p(self.rbases.data_ptr()),
                                          ctypes.c_void_p(self.rquals.data_ptr()) if self.need_q else None,
                                          ctypes.c_void_p(self.roff.data_ptr()), self.NR, self.RL, ctypes.byref(self.params),
                                          ctypes.c_void_p(self.words.data_ptr()), ctypes.c_void_p(self.lst.data_ptr()), self.list_cap,
                                          ctypes.c_void_p(self.state.data_ptr()), self.sptr, ctypes.byref(self.n_launch)))
        self.launches += self.n_launch.value

    def step(self):
        nat, L, G = self.nat, self.L, self.G
        self.state.zero_()
        self.acc.zero_()
        self.first_seen.fill_(-1)
        self.align()
        nat.check(L.pa_summary_reduce_device(ctypes.c_void_p(self.words.data_ptr()), ctypes.c_void_p(self.lst.data_ptr()), self.NR,
                                             self.rank * self.NR, G, ctypes.c_void_p(self.acc.data_ptr()),
                                             ctypes.c_void_p(self.acc.data_ptr() + 32), ctypes.c_void_p(self.acc.data_ptr() + 32 + 8 * G),
                                             ctypes.c_void_p(self.first_seen.data_ptr()), self.sptr))
        self.launches += 1
        if self.world > 1:   # the one exchange of read-sharded alignment: SUM + MIN in one NCCL group on this stream
            self.comm.allreduce_summary(self.acc.data_ptr(), 4 + 2 * G, self.first_seen.data_ptr(), G, self.stream.cuda_stream)

    def kernel_only_ms(self, reps):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):
            self.state.zero_()
            e0.record(self.stream)
            self.align()
            e1.record(self.stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total / reps

    def timed_steps(self, steps, warmup, dist):
        """(ms per step as the max over ranks, launches inside the timed region)"""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        if int(self.state[1].item()) != 0:
            raise RuntimeError("list buffer overflow / oversized read in the benchmark step")
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        self.launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            self.step()
        e1.record(self.stream)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, self.launches

    def e2e(self, iters, dist):
        """End to end through the host-buffer C ABI call: pinned host inputs, H2D + D2H inside the timed region."""
        torch, nat, L, NR, RL = self.torch, self.nat, self.L, self.NR, self.RL
        hb = torch.empty(NR * RL, dtype=torch.uint8).pin_memory()
        hb.copy_(self.rbases)
        hq = None
        if self.need_q:
            hq = torch.empty(NR * RL, dtype=torch.uint8).pin_memory()
            hq.copy_(self.rquals)
        hoff = torch.empty(NR + 1, dtype=torch.int64).pin_memory()
        hoff.copy_(self.roff)
        hwords = torch.empty(NR, dtype=torch.int64).pin_memory()
        hlist = torch.empty(self.list_cap, dtype=torch.int32).pin_memory()
        need = ctypes.c_uint64(0)
        counters = np.zeros(3, dtype=np.uint64)
        times = []
        for it in range(2 + iters):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            nat.check(L.pa_align_batch(self.ix.handle, ctypes.c_void_p(hb.data_ptr()), ctypes.c_void_p(hq.data_ptr()) if self.need_q else None,
                                       ctypes.c_void_p(hoff.data_ptr()), NR, ctypes.byref(self.params), ctypes.c_void_p(hwords.data_ptr()),
                                       ctypes.c_void_p(hlist.data_ptr()), self.list_cap, ctypes.byref(need),
                                       counters.ctypes.data_as(ctypes.c_void_p)))
            dt = time.perf_counter() - t0
            if it >= 2:
                times.append(dt)
        te = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # list offsets are handed out by an atomic cursor, so only type / length / single-genome payloads are comparable
        hw, dw = hwords.to(self.dev), self.words
        same = torch.equal(hw >> 40, dw >> 40) and torch.equal(torch.where(((hw >> 40) & 0x3FFFFF) == 1, hw, 0),
                                                                torch.where(((dw >> 40) & 0x3FFFFF) == 1, dw, 0))
        assert same, "host-buffer call disagrees with the device-resident call"
        h2d = NR * RL * (2 if self.need_q else 1) + (NR + 1) * 8
        d2h = NR * 8 + int(need.value) * 4 + 40
        out = {"value": self.world * NR / float(te.item()), "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": float(te.item()) * 1e3,
               "note": "h2d = the pinned host buffers handed to pa_align_batch (ASCII bases [+ qualities] + offsets); the call "
                       "packs chunks to 2-bit planes on the host cores before the PCIe copy when that is faster than the link"}
        # the same reads packed ONCE (pa_pack_reads: what an ingest does when it parses the FASTQ), then pa_align_batch_packed per
        # step: a quarter of the bytes cross PCIe and no host core touches the reads inside the timed region
        n_words = 2 * (NR * RL // 32 + NR + 1)
        hp = torch.empty(n_words, dtype=torch.int32).pin_memory()
        ok = ctypes.c_int32(0)
        nat.check(L.pa_pack_reads(ctypes.c_void_p(hb.data_ptr()), ctypes.c_void_p(hoff.data_ptr()), NR, ctypes.c_void_p(hp.data_ptr()),
                                  n_words, ctypes.byref(ok)))
        if ok.value:
            ptimes = []
            for it in range(2 + iters):
                torch.cuda.synchronize()
                if self.world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                nat.check(L.pa_align_batch_packed(self.ix.handle, ctypes.c_void_p(hp.data_ptr()), ctypes.c_void_p(hq.data_ptr()) if self.need_q else None,
                                                  ctypes.c_void_p(hoff.data_ptr()), NR, ctypes.byref(self.params), ctypes.c_void_p(hwords.data_ptr()),
                                                  ctypes.c_void_p(hlist.data_ptr()), self.list_cap, ctypes.byref(need),
                                                  counters.ctypes.data_as(ctypes.c_void_p)))
                dt = time.perf_counter() - t0
                if it >= 2:
                    ptimes.append(dt)
            tp = torch.tensor([float(np.mean(ptimes))], dtype=torch.float64, device=self.dev)
            if self.world > 1:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            hw = hwords.to(self.dev)
            same = torch.equal(hw >> 40, dw >> 40) and torch.equal(torch.where(((hw >> 40) & 0x3FFFFF) == 1, hw, 0),
                                                                    torch.where(((dw >> 40) & 0x3FFFFF) == 1, dw, 0))
            assert same, "pa_align_batch_packed disagrees with the device-resident call"
            out["packed_once"] = {"value": self.world * NR / float(tp.item()), "unit": "reads/s", "ms_per_step": float(tp.item()) * 1e3,
                                  "h2d_bytes_per_step": n_words * 4 + (NR * RL if self.need_q else 0), "d2h_bytes_per_step": d2h,
                                  "note": "pa_align_batch_packed on host buffers packed once outside the timed region (pa_pack_reads)"}
        return out

    def host_results(self):
        """(types, lens, flat genome lists) of the last step on the host, in read order."""
        words = self.words.cpu().numpy().view(np.uint64)
        n_list = int(self.state[0].item())
        lst = self.lst[:max(n_list, 1)].cpu().numpy().view(np.uint32)
        return self.nat.flatten_results(words, lst)


def align_roofline(NR, RL, k, need_q, k4_ms, hbm_peak, peaks_found, traffic=None):
    alg = ALG_BYTES_QUAL if need_q else ALG_BYTES_PLAIN
    if RL != 150 or k != 31:
        alg = RL * (2 if need_q else 1) + 32 * max(RL - k + 1, 0) + 8
    achieved = NR * alg / (k4_ms * 1e-3) / 1e9
    line_peak, line_src = random_line_peak()
    lookups = NR * max(RL - k + 1, 0) / (k4_ms * 1e-3)
    return {"bound": "hbm", "kernel": ("align_fast_kernel (K4)" if need_q else "align_fast_split_kernel (K4)"), "achieved": achieved,
            "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks_found else "fallback 6650 (of fallback)",
            "traffic": traffic, "kernel_ms": k4_ms, "algorithmic_bytes_per_read": alg,
            "random_access": {
                "note": "a B200 serves random table reads at a fixed rate of distinct 128-B lines; lanes of one load instruction "
                        "that share a line are served together (profiles/r01_locality_roofline.jsonl)",
                "peak_lines_per_s": line_peak, "unit": "128-B lines/s", "peak_source": line_src,
                "achieved_lines_per_s": (traffic / 128.0 / (k4_ms * 1e-3)) if traffic else None,
                "frac": (traffic / 128.0 / (k4_ms * 1e-3) / line_peak) if traffic else None,
                "window_lookups_per_s": lookups, "lookups_vs_one_line_per_lookup": lookups / line_peak}}


def device_cluster_genomes(torch, dev, n_genomes, length, seed, cluster=10, sub=0.01):
    """configs[3]: clusters of `cluster` genomes, each member = the cluster ancestor with `sub` substitutions."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    bases = torch.empty(n_genomes * length, dtype=torch.uint8, device=dev)
    anc = None
    for g in range(n_genomes):
        if g % cluster == 0:
            anc = lut[torch.randint(0, 4, (length,), generator=gen, device=dev)]
        seq = anc.clone()
        hit = torch.rand(length, generator=gen, device=dev) < sub
        seq[hit] = lut[torch.randint(0, 4, (int(hit.sum().item()),), generator=gen, device=dev)]
        bases[g * length:(g + 1) * length] = seq
    return bases


def extsim_record(args, torch, dist, nat, comm, rank, world, local, dev, hbm_peak):
    """configs[3]: EXTSIM reference build -- index, per-genome statistics (K5), pairwise intersections (K6), the greedy
    filter (host, on the device's integers) and the removal of the filtered genomes (K7), at threshold 0.5 (siblings at
    99 % identity share ~0.62-0.66 of their 31-mers, so the default 0.95 filters nothing: SURVEY.md 8(d))."""
    import multi_gpu
    G, GL, k = args.d_genomes, args.d_genome_len, args.k
    bases = device_cluster_genomes(torch, dev, G, GL, seed=3000)
    goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL)).astype(np.uint64)
    def one_pass():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dix = None
        if world > 1:
            g_lo, g_hi = nat.genome_shard(goff, world, rank)
            dix = multi_gpu.build_partitioned(comm, bases[g_lo * GL:].data_ptr(), goff, k, device=local, g_range=(g_lo, g_hi))
            ix, sizes = dix, dix.sizes()
        else:
            ix = nat.NativeIndex.build_device(bases.data_ptr(), goff, k, device=local)
            inf = ix.info()
            sizes = (int(inf.n_keys), int(inf.n_runs), int(inf.n_occ))
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        group = np.arange(G, dtype=np.uint32)
        t0 = time.perf_counter()
        total, unique = ix.extsim_stats(group, G)
        t_stats = time.perf_counter() - t0
        t0 = time.perf_counter()
        inter = ix.extsim_pairwise(group, G)
        t_pair = time.perf_counter() - t0
        t0 = time.perf_counter()
        order = sorted(range(G), key=lambda g: (int(unique[g]), int(total[g]), GL, g))     # kmer.py:179-186
        kept, keep = [], np.zeros(G, dtype=np.uint8)
        for g in order:                                                                      # kmer.py:188-230
            hit = False
            for o in kept:
                smaller = min(int(total[g]), int(total[o]))
                if smaller > 0 and int(inter[g, o]) / smaller > 0.5:
                    hit = True
                    break
            if not hit:
                kept.append(g)
                keep[g] = 1
        t_greedy = time.perf_counter() - t0
        t0 = time.perf_counter()
        ix.drop_genomes(keep)
        torch.cuda.synchronize()
        t_drop = time.perf_counter() - t0
        after = ix.sizes() if world > 1 else (lambda i: (int(i.n_keys), int(i.n_runs), int(i.n_occ)))(ix.info())
        tms = [t_build, t_stats, t_pair, t_drop]
        if world > 1:
            tt = torch.tensor(tms, dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tms = tt.tolist()
        ix.close()
        return tms, t_greedy, sizes, after, keep

    # two passes: the first pays the process's first large allocations (hundreds of ms on some boxes, see the buffer cache in
    # index.cuh), the second is what a process that rebuilds sees; both are reported
    first = one_pass()
    tms, t_greedy, sizes, after, keep = one_pass()
    first_total = sum(first[0]) + first[1]
    del bases
    torch.cuda.empty_cache()
    nat.trim_memory()
    runs = sizes[1]
    stream_bytes = 4.0 * runs      # EXTSIM = one streaming pass over the genome-run array per kernel
    return {"workload": f"configs[3]: EXTSIM build, {G} genomes x {GL} bp in clusters of 10 (1 % substitutions), k={k}, threshold 0.5",
            "metric": "ref-build k-mers/s (EXTSIM build: index + K5 + K6 + greedy + K7)", "unit": "k-mers/s",
            "value": sizes[2] / (tms[0] + tms[1] + tms[2] + t_greedy + tms[3]), "n_gpus": world,
            "seconds": {"build": tms[0], "stats_k5": tms[1], "pairwise_k6": tms[2], "greedy_host": t_greedy, "drop_k7_and_table": tms[3]},
            "first_pass_seconds": first_total, "note": "second of two passes in one process (buffer cache warm); first_pass_seconds = the same sequence with the process's first allocations",
            "kmer_occurrences": sizes[2], "distinct_kmers": sizes[0], "genome_runs": runs, "genomes_kept": int(keep.sum()),
            "distinct_kmers_after": after[0],
            "roofline": {"bound": "hbm", "kernel": "extsim_pairwise_kernel (K6) + verify", "achieved": stream_bytes * 2 / tms[2] / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": stream_bytes * 2 / tms[2] / 1e9 / hbm_peak,
                         "algorithmic_bytes": "4 B x genome runs per pass, two passes (count, verify); call time incl. launches and D2H of the matrix"}}


def config_e_record(args, torch, dist, nat, comm, rank, world, local, dev, stream, hbm_peak, peaks_found):
    """configs[4]: the 2,000-genome index, table-only.  Built partitioned over the ranks (pa_index_build_partitioned; on
    one GPU the key space is streamed in rounds), then every rank aligns its own reads against the replicated table."""
    import multi_gpu
    G, GL, NR, RL, k = args.e_genomes, args.e_genome_len, args.e_reads, args.read_len, args.k
    goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL)).astype(np.uint64)
    g_lo, g_hi = nat.genome_shard(goff, world, rank)
    # every rank generates the genomes it encodes; reads are cut from a window of genomes that is generated on every rank
    # (the first 32), so that each rank has reads without holding all 4 Gb
    n_local = g_hi - g_lo
    src_genomes = min(32, G)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = device_genomes(torch, dev, n_local, GL, seed=5000 + g_lo) if n_local else torch.zeros(16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dix = multi_gpu.build_partitioned(comm, mine.data_ptr(), goff, k, device=local, table_only=True, g_range=(g_lo, g_hi))
    torch.cuda.synchronize()
    tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    t_build = float(tb.item())
    inf = dix.replica.info()
    # the source genomes of the reads = the genomes rank 0 generated first (seed 5000 + 0 walks the same stream)
    src = mine[:src_genomes * GL] if g_lo == 0 else device_genomes(torch, dev, src_genomes, GL, seed=5000)
    if g_lo != 0:
        del mine
    rb, rq, roff = device_reads(torch, dev, src, src_genomes, GL, NR, RL, seed=6000 + rank)
    params = nat.make_params(1, 1, None, None, None)
    run = AlignRunner(torch, nat, dix.replica, comm, rank, world, dev, stream, G, rb, rq, roff, NR, RL, params, False)
    ms_step, launches = run.timed_steps(max(2, min(args.steps, 5)), 3, dist)
    k4_ms = run.kernel_only_ms(2)
    stats = run.acc[:4].cpu().tolist()
    e2e = None if args.no_e2e else run.e2e(2, dist)
    # parity on a sample: reads cut from the first 8 genomes, aligned against the 2,000-genome table, must get what the
    # oracle gets on an 8-genome index (a random 31-mer of another genome coincides with one of theirs with probability
    # ~1e-9 per read; clusters are 4 genomes, so the 8 hold two whole clusters)
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        ns, nr_s = 8, 200_000
        h_src = src[:ns * GL].cpu().numpy()
        sb, sq, soff = device_reads(torch, dev, src, ns, GL, nr_s, RL, seed=6500)
        hb, hoff = sb.cpu().numpy(), soff.cpu().numpy().astype(np.uint64)
        words, lst, counters = dix.replica.align(hb, None, hoff, params)
        t_, l_, f_ = nat.flatten_results(words, lst)
        o = orc.OracleReference(k, [(f"genome{g}", h_src[g * GL:(g + 1) * GL].tobytes().decode("ascii")) for g in range(ns)])
        al = o.align_packed([], hb, sq.cpu().numpy(), hoff, 1, 1, None, None, None, nthreads=host_threads())
        ok = (np.array_equal(t_, al.types) and np.array_equal(l_, np.diff(al.list_off.astype(np.int64))) and
              np.array_equal(f_, al.genomes[:len(f_)]))
        import synth
        parity = {"equal_to_oracle": bool(ok), "sample": f"{nr_s} reads cut from the first {ns} genomes; oracle index of those {ns} genomes",
                  "digest_sha256": synth.result_digest(t_, l_, f_)}
    rec = {"workload": f"configs[4]: {G} genomes x {GL} bp index (table-only, replicated), {NR} x {RL} bp reads per GPU, k={k}, plain",
           "metric": "reads/s pseudo-aligned (k=31,150bp)", "unit": "reads/s", "value": world * NR / (ms_step * 1e-3), "n_gpus": world,
           "ms_per_step": ms_step, "gpu_launches": launches, "e2e": e2e,
           "index": {"genomes": G, "genome_len": GL, "kmer_occurrences": int(inf.n_occ), "distinct_kmers": int(inf.n_keys),
                     "table_bytes": int(inf.table_bytes), "device_bytes": int(inf.device_bytes),
                     "bytes_per_kmer": inf.device_bytes / max(int(inf.n_keys), 1), "stash_count": int(inf.stash_count),
                     "table_blocks": int(inf.n_blocks), "build_seconds": t_build, "build_kmers_per_s": int(inf.n_occ) / t_build,
                     "build_phases_ms_rank0": dix.timings, "genome_generation_seconds": t_gen,
                     "how": "pa_index_build_partitioned(PA_BUILD_TABLE_ONLY): partitioned over the ranks" if world > 1 else
                            "pa_index_build_partitioned(PA_BUILD_TABLE_ONLY): one GPU, key space streamed in rounds"},
           "roofline": align_roofline(NR, RL, k, False, k4_ms, hbm_peak, peaks_found),
           "result": {"unique": stats[0], "ambiguous": stats[1], "unmapped": stats[2], "dropped": stats[3]}, "parity": parity}
    dix.close()
    del rb, rq, roff, src, run
    torch.cuda.empty_cache()
    nat.trim_memory()
    return rec


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import _native as nat
    import multi_gpu
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = nat.Comm.from_torch(local)      # pa_comm over NCCL (its own communicator; torch's carries the unique id)
    nat.require_device()
    G, GL, NR, RL, k = args.genomes, args.genome_len, args.reads, args.read_len, args.k
    mrq, mkq, mg = filters(args)
    params = nat.make_params(1, 1, mrq, mkq, mg)
    need_q = args.extquality
    # a dedicated (non-default) stream: its handle is non-NULL, so the C ABI launches on it and not on the index's own
    # stream, and the CUDA events see exactly the kernels being timed
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- index build (single GPU, every rank), device-resident input ----
    bases = device_genomes(torch, dev, G, GL, seed=1000)
    goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL)).astype(np.uint64)
    torch.cuda.synchronize()
    build_times = []
    ix = None
    phase_ms = lambda i: i.build_encode_ms + i.build_sort_ms + i.build_rle_ms + i.build_table_ms
    inf = None
    for _ in range(3):   # the phases are timed with events around host-side allocations too: keep the build with the least of them
        if ix is not None:
            ix.close()
        t0 = time.perf_counter()
        ix = nat.NativeIndex.build_device(bases.data_ptr(), goff, k, device=local)
        build_times.append(time.perf_counter() - t0)
        cur = ix.info()
        if inf is None or phase_ms(cur) < phase_ms(inf):
            inf = cur
    build_kernel_ms = phase_ms(inf)

    # ---- multi-GPU build (N > 1): partitioned by minimizer digit, one fused scatter + exchange, table slices gathered ----
    build_part, dix = None, None
    align_ix = ix
    if world > 1:
        g_lo, g_hi = nat.genome_shard(goff, world, rank)
        part_times = []
        for attempt in range(2):      # the first call also maps the peers' receive buffers (cudaIpcOpenMemHandle)
            if dix is not None:
                dix.close()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            dix = multi_gpu.build_partitioned(comm, bases[g_lo * GL:].data_ptr(), goff, k, device=local, g_range=(g_lo, g_hi))
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            part_times.append(float(dt.item()))
        # content check: the checksums of the partitions (k-mers, genome runs, positions) add up to the single-GPU index's
        sums = comm.allreduce_host(dix.partition.checksum())
        want = ix.checksum()
        same_content = bool(np.array_equal(sums, want))
        assert same_content, f"partitioned build differs from the single-GPU build: {sums} != {want}"
        assert dix.sizes() == (int(inf.n_keys), int(inf.n_runs), int(inf.n_occ))
        pinf, rinf = dix.partition.info(), dix.replica.info()
        build_part = {"kmers_per_s": inf.n_occ / part_times[-1], "seconds": part_times[-1], "first_call_seconds": part_times[0],
                      "phases_ms_rank0": dix.timings, "partition_keys_rank0": int(pinf.n_keys),
                      "replica_table_bytes": int(rinf.table_bytes), "replica_stash": int(rinf.stash_count),
                      "checksum_equals_single_gpu_build": same_content,
                      "how": "K1 names the owner of every record (digit of its minimizer hash); one scatter kernel stores the "
                             "records into the owners' receive buffers over NVLink (CUDA IPC peer memory); K2 + K3 per rank; "
                             "every rank fills its slice of the lookup table; slices all-gathered in place"}
        align_ix = dix.replica      # the timed steps align against the replica of the partitioned build

    # ---- reads of this rank (weak scaling: every rank aligns its own `reads` reads) ----
    rbases, rquals, roff = device_reads(torch, dev, bases, G, GL, NR, RL, seed=2000 + rank)
    run = AlignRunner(torch, nat, align_ix, comm, rank, world, dev, stream, G, rbases, rquals, roff, NR, RL, params, need_q)
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, launches = run.timed_steps(args.steps, max(args.warmup, 3), dist)
    k4_ms = run.kernel_only_ms(3)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    value = world * NR / (ms_step * 1e-3)
    stats_host = run.acc[:4].cpu().tolist()
    replica_equals_single = None
    if world > 1:   # the replica must answer exactly like the single-GPU index: same result words for this rank's reads
        w_rep = run.words.clone()
        run.ix = ix
        run.state.zero_()
        run.align()
        torch.cuda.synchronize()
        hw, dw = w_rep, run.words
        replica_equals_single = bool(torch.equal(hw >> 40, dw >> 40) and
                                     torch.equal(torch.where(((hw >> 40) & 0x3FFFFF) == 1, hw, 0), torch.where(((dw >> 40) & 0x3FFFFF) == 1, dw, 0)))
        assert replica_equals_single, "alignment against the replica differs from the single-GPU index"
        run.ix = align_ix
        run.step()
        build_part["alignment_through_replica_equals_single_gpu_index"] = True

    # ---- full-size parity: this run's per-read results against the digest of the oracle's (tests/golden/fullsize_digest.json) ----
    digest = None
    if rank == 0:
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_digest.json")))
            w = gold["workload"]
            if (w["genomes"], w["genome_len"], w["reads"], w["read_len"], w["k"]) == (G, GL, NR, RL, k):
                case = gold["cases"]["configs[2] extquality" if need_q else "configs[1] plain"]
                got = synth.result_digest(*run.host_results())
                digest = {"sha256": got, "equals_oracle_digest": got == case["oracle_digest_sha256"],
                          "source": "tests/golden/fullsize_digest.json (tools/fullsize_parity.py: all 10^7 reads against the C oracle)"}
                assert digest["equals_oracle_digest"], "per-read results differ from the committed full-size oracle digest"
        except (OSError, KeyError):
            pass

    e2e = None if args.no_e2e else run.e2e(max(2, min(args.steps, 5)), dist)

    # ---- sub-records: configs[2] on the same index and reads ----
    sub = {}
    if not args.no_configs and not need_q:
        qparams = nat.make_params(1, 1, 62, 60, 3)
        qrun = AlignRunner(torch, nat, align_ix, comm, rank, world, dev, stream, G, rbases, rquals, roff, NR, RL, qparams, True)
        q_ms, q_launch = qrun.timed_steps(max(2, min(args.steps, 5)), 3, dist)
        q_k4 = qrun.kernel_only_ms(2)
        qstats = qrun.acc[:4].cpu().tolist()
        qcount = qrun.state[2:5].cpu().tolist()
        qdig = None
        if rank == 0:
            try:
                gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_digest.json")))
                w = gold["workload"]
                if (w["genomes"], w["genome_len"], w["reads"], w["read_len"], w["k"]) == (G, GL, NR, RL, k):
                    got = synth.result_digest(*qrun.host_results())
                    qdig = {"sha256": got, "equals_oracle_digest": got == gold["cases"]["configs[2] extquality"]["oracle_digest_sha256"]}
                    assert qdig["equals_oracle_digest"], "configs[2]: per-read results differ from the committed oracle digest"
            except (OSError, KeyError):
                pass
        q_e2e = None if args.no_e2e else qrun.e2e(2, dist)
        sub["extquality"] = {"workload": "configs[2]: configs[1] + EXTQUALITY (min-read-quality 62, min-kmer-quality 60, max-genomes 3)",
                             "metric": "reads/s pseudo-aligned (k=31,150bp)", "unit": "reads/s", "value": world * NR / (q_ms * 1e-3),
                             "n_gpus": world, "ms_per_step": q_ms, "gpu_launches": q_launch, "e2e": q_e2e,
                             "roofline": align_roofline(NR, RL, k, True, q_k4, hbm_peak, bool(peaks)),
                             "result": {"unique": qstats[0], "ambiguous": qstats[1], "unmapped": qstats[2], "dropped": qstats[3]},
                             "filter_counters_this_rank": {"filtered_quality_reads": qcount[0], "filtered_quality_kmers": qcount[1],
                                                           "filtered_hr_kmers": qcount[2]},
                             "parity_digest": qdig}
        del qrun

    # ---- CPU baseline + parity of the GPU path on the same bounded sample (rank 0, N = 1) ----
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        genomes, b, q, off = cpu_sample_workload(args)
        o, al, r = run_cpu_oracle(args, genomes, b, q, off, steps=1, warmup=0)
        data, sgoff = nat.pack_strings([g.tobytes().decode("ascii") for g in genomes])
        six = nat.NativeIndex.build(data, sgoff, k, device=local)
        sinf = six.info()
        w2, l2, c2 = six.align(b, q, off, params)
        t2, n2, f2 = nat.flatten_results(w2, l2)
        ok = (sinf.n_keys, sinf.n_runs, sinf.n_occ) == o.sizes() and np.array_equal(t2, al.types)
        ok = ok and np.array_equal(n2, np.diff(al.list_off.astype(np.int64))) and np.array_equal(f2, al.genomes[:len(f2)])
        ok = ok and [int(x) for x in c2] == [al.filtered_quality_reads, al.filtered_quality_kmers if mkq is not None else 0,
                                            al.filtered_hr_kmers if mg is not None else 0]
        parity = "bit-exact vs oracle on the sample" if ok else "MISMATCH vs oracle on the sample"
        six.close()
        cpu_baseline = {"value": r["reads_per_s"], "unit": "reads/s", "cores": r["threads"], "kind": "port",
                        "sample": f"{args.sample_genomes} genomes x {GL} bp index, {args.sample_reads} x {RL} bp reads; "
                                  f"oracle/pa_oracle.c with OpenMP over reads",
                        "build_kmers_per_s": r["build_kmers_per_s"]}

    traffic = None
    try:  # DRAM bytes per launch from the committed ncu capture of this exact workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "align_traffic.json")))
        w = tr["workload"]
        if (w["genomes"], w["genome_len"], w["reads"], w["read_len"], w["k"], w["extquality"]) == (G, GL, NR, RL, k, need_q):
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    clocks = sampler.summary()

    # ---- the other configs: free the configs[1] state first (config E needs the memory) ----
    del run, rbases, rquals, roff
    if dix is not None:
        dix.close()
    ix.close()
    del bases
    torch.cuda.empty_cache()
    nat.trim_memory()      # the library's buffer cache: configs[3] / [4] start from an empty device
    if not args.no_configs and not need_q:
        sub["extsim"] = extsim_record(args, torch, dist, nat, comm, rank, world, local, dev, hbm_peak)
        sub["config_e"] = config_e_record(args, torch, dist, nat, comm, rank, world, local, dev, stream, hbm_peak, bool(peaks))

    if rank == 0:
        line = {
            "metric": "reads/s pseudo-aligned (k=31,150bp)", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": align_roofline(NR, RL, k, need_q, k4_ms, hbm_peak, bool(peaks), traffic),
            "cpu_baseline": cpu_baseline, "parity": parity, "parity_fullsize_digest": digest,
            "build": {"kmers_per_s_kernels": inf.n_occ / (build_kernel_ms * 1e-3) if build_kernel_ms > 0 else None,
                      "kmers_per_s_call": inf.n_occ / min(build_times), "kmer_occurrences": int(inf.n_occ),
                      "distinct_kmers": int(inf.n_keys), "encode_ms": inf.build_encode_ms, "sort_ms": inf.build_sort_ms,
                      "rle_ms": inf.build_rle_ms, "table_ms": inf.build_table_ms, "index_bytes": int(inf.device_bytes),
                      "table_bytes": int(inf.table_bytes), "table_bytes_per_kmer": inf.table_bytes / max(int(inf.n_keys), 1),
                      "stash_count": int(inf.stash_count), "table_blocks": int(inf.n_blocks), "minimizer_len": int(inf.minimizer_len),
                      "roofline_frac_17B": (inf.n_occ * BUILD_BYTES_PER_KMER / (build_kernel_ms * 1e-3) / 1e9 / hbm_peak) if build_kernel_ms > 0 else None},
            "build_partitioned": build_part,
            "result": {"unique": stats_host[0], "ambiguous": stats_host[1], "unmapped": stats_host[2], "dropped": stats_host[3]},
            "configs": sub or None,
        }
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
