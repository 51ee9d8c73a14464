"""Host-only timing of the 2-bit read packing (hostpack.cpp) on this machine's cores."""
import ctypes, glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0]]
import numpy as np
import _native as nat
L = nat.lib()
n, rl = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 150
rng = np.random.default_rng(1)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * rl)].copy()
off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(rl))
cap = 2 * (n * rl // 32 + n + 1)
planes = np.zeros(cap, dtype=np.uint32)
ok = ctypes.c_int32(0)
for threads in (1, 2, 4, 8, 16, 32):
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        nat.check(L.pa_debug_pack_reads(nat._p(bases), nat._p(off), n, nat._p(planes), cap, threads, ctypes.byref(ok)))
        best = min(best, time.perf_counter() - t0)
    print(f"threads {threads:3d}: {best*1e3:8.2f} ms  {n*rl/best/1e9:6.2f} GB/s of ASCII  ok={ok.value}")
