"""Extended differential fuzzing of the GPU path against the oracle (a soak run, not part of the suite).
Usage: python tests/soak.py [first_seed=100000] [n=2000]   -- prints the number of cases and the first failure, if any.
Imports the checker of tests/ (which uses oracle/): run it like a test."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ lives next to the package
sys.path[:0] = [ROOT, [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0], os.path.join(ROOT, "tests")]
import synth
import test_gpu_abi as T

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
done = 0
for seed in range(first, first + n):
    for dense in ("", "3"):
        if dense:
            os.environ["PA_TABLE_DENSE"] = dense
        else:
            os.environ.pop("PA_TABLE_DENSE", None)
        os.environ["PA_HOST_PACK"] = "1" if seed % 2 else "0"
        if seed % 3 == 0:
            case = synth.fuzz_case(seed, k_range=(9, 31), max_genomes=8)
        else:
            case = synth.fuzz_case(seed, dup_ids=seed % 7 == 0)
        try:
            T.check_case(case)
        except Exception as e:   # noqa: BLE001
            print("FAILED seed", seed, "dense", dense, type(e).__name__, str(e)[:500])
            sys.exit(1)
        done += 1
print("soak ok:", done, "cases")

# ---- second half: k = 17..31 on clustered genomes (strain variants: chains, CONT, stash, general kernel) ----
import numpy as np
rng = np.random.default_rng(first)
big = 0
for it in range(max(n // 25, 1)):
    G = int(rng.integers(2, 24))
    L = int(rng.integers(2_000, 15_000))
    k = int(rng.integers(17, 32))
    genomes = synth.make_genomes(G, L, seed=int(rng.integers(1 << 30)), cluster_size=int(rng.integers(2, 9)),
                                 shared_frac=float(rng.uniform(0.3, 0.95)), sub_rate=float(rng.uniform(0.002, 0.05)),
                                 n_every=int(rng.integers(800, 5000)), n_run=int(rng.integers(1, 12)))
    b, q, off = synth.make_reads(genomes, int(rng.integers(200, 1200)), int(rng.integers(k, 260)), seed=int(rng.integers(1 << 30)),
                                 sub_rate=float(rng.uniform(0.0, 0.04)), random_frac=float(rng.uniform(0.0, 0.1)))
    pr = dict(m=int(rng.integers(0, 4)), p=int(rng.integers(-1, 4)),
              mrq=int(rng.integers(55, 66)) if rng.random() < 0.4 else None,
              mkq=int(rng.integers(55, 66)) if rng.random() < 0.4 else None,
              mg=int(rng.integers(1, 6)) if rng.random() < 0.4 else None)
    case = {"k": k, "genomes": synth.genomes_as_pairs(genomes), "reads": synth.reads_as_triples(b, q, off), "params": pr, "seed": it}
    for dense in ("", "2", "4"):
        if dense:
            os.environ["PA_TABLE_DENSE"] = dense
        else:
            os.environ.pop("PA_TABLE_DENSE", None)
        os.environ["PA_HOST_PACK"] = "1" if it % 2 else "0"
        os.environ["PA_CHUNK_READS"] = "257"
        try:
            T.check_case(case)
        except Exception as e:   # noqa: BLE001
            print("FAILED clustered case", it, "dense", dense, k, G, L, pr, type(e).__name__, str(e)[:500])
            sys.exit(1)
        big += 1
print("soak ok:", big, "clustered cases")
