"""Times the single-GPU index build of the bench workload (PA_TRACE=1 prints the table-build phases)."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0]]
import numpy as np, torch
import bench, _native as nat
dev = torch.device("cuda", 0)
G, GL = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
bases = bench.device_genomes(torch, dev, G, GL, seed=1000)
goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL))
torch.cuda.synchronize()
best = None
for i in range(int(os.environ.get('PA_BUILD_REPS', '3'))):
    t0 = time.perf_counter()
    ix = nat.NativeIndex.build_device(bases.data_ptr(), goff, 31, device=0)
    dt = time.perf_counter() - t0
    inf = ix.info()
    print(f"build {dt*1e3:.1f} ms  enc {inf.build_encode_ms:.1f} sort {inf.build_sort_ms:.1f} rle {inf.build_rle_ms:.1f} "
          f"table {inf.build_table_ms:.1f}  sets_sectors {inf.n_list_sectors} stash {inf.stash_count}", file=sys.stderr)
    cur = (inf.build_sort_ms, inf.build_rle_ms, inf.build_table_ms)
    best = cur if best is None else tuple(min(a, b) for a, b in zip(best, cur))
    ix.close()
print('best sort %.1f rle %.1f table %.1f' % best, file=sys.stderr)
