"""dumpref timing: native JSON writer (KmerReference.summary_json) vs json.dumps(get_summary()).  Needs a GPU."""
import glob, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0]]
import synth
from kmer import KmerReference
from records import Record, Section

G, GL = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10, 200_000)
genomes = synth.make_genomes(G, GL, seed=5, cluster_size=5, shared_frac=0.3, sub_rate=0.01, n_every=100_000, n_run=20)
recs = [Record([Section("description", f"genome{i}"), Section("genome", g.tobytes().decode())]) for i, g in enumerate(genomes)]
ref = KmerReference(31, recs)
t0 = time.perf_counter(); a = ref.summary_json(indent=4); t_native = time.perf_counter() - t0
t0 = time.perf_counter(); b = json.dumps(ref.get_summary(), indent=4); t_dict = time.perf_counter() - t0
print(json.dumps({"genomes": G, "genome_len": GL, "distinct_kmers": len(ref.kmers), "json_MB": len(a) / 1e6, "identical": a == b,
                  "native_s": t_native, "dict_s": t_dict, "speedup": t_dict / t_native}))
