"""Debug helper: run fuzz seeds through the raw ABI one at a time, printing progress (use under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")]
import synth
import test_gpu_abi as T

lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    case = synth.fuzz_case(seed)
    print("seed", seed, "k", case["k"], "G", len(case["genomes"]), "reads", len(case["reads"]), case["params"], flush=True)
    T.check_case(case)
print("all ok")
