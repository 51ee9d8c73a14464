"""End-to-end CLI timing on synthetic files: `main.py -t reference`, `-t dumpalign` (ingest + build / align + save).
Usage: python tools/cli_bench.py [genomes=20] [genome_len=2000000] [reads=2000000]"""
import glob, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0]
sys.path[:0] = [ROOT, PKG]
import numpy as np
import synth

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20
GL = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
NR = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
tmp = tempfile.mkdtemp(prefix="pa_cli_")
fa, fq, kdb = os.path.join(tmp, "g.fa"), os.path.join(tmp, "r.fq"), os.path.join(tmp, "ref.kdb")
genomes = synth.make_genomes(G, GL, seed=5, cluster_size=4, shared_frac=0.3, sub_rate=0.01, n_every=1_000_000, n_run=40)
with open(fa, "w") as f:
    for i, g in enumerate(genomes):
        s = g.tobytes().decode()
        f.write(f">genome{i}\n")
        f.write("\n".join(s[j:j + 80] for j in range(0, len(s), 80)) + "\n")
b, q, off = synth.make_reads(genomes, NR, 150, seed=6, sub_rate=0.01, random_frac=0.02)
bb, qq = b.reshape(NR, 150), q.reshape(NR, 150)
with open(fq, "wb") as f:
    chunk = 100_000
    for lo in range(0, NR, chunk):
        hi = min(NR, lo + chunk)
        rows = [b"@read%d\n%s\n+\n%s\n" % (i, bb[i].tobytes(), qq[i].tobytes()) for i in range(lo, hi)]
        f.write(b"".join(rows))
out = {"genomes": G, "genome_len": GL, "reads": NR, "fasta_MB": os.path.getsize(fa) / 1e6, "fastq_MB": os.path.getsize(fq) / 1e6}
env = dict(os.environ, PYTHONPATH=PKG)
t0 = time.perf_counter()
subprocess.run([sys.executable, os.path.join(PKG, "main.py"), "-t", "reference", "-g", fa, "-k", "31", "-r", kdb], check=True, env=env)
out["reference_s"] = time.perf_counter() - t0
out["kdb_MB"] = os.path.getsize(kdb) / 1e6
t0 = time.perf_counter()
r = subprocess.run([sys.executable, os.path.join(PKG, "main.py"), "-t", "dumpalign", "-r", kdb, "--reads", fq], check=True, env=env,
                   capture_output=True, text=True)
out["dumpalign_s"] = time.perf_counter() - t0
out["statistics"] = json.loads(r.stdout)["Statistics"]
print(json.dumps(out))
