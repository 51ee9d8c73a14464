"""Host-only timing of FASTQ / FASTA ingest: native parser (csrc/ingest.cpp) vs the regex restatement vs, when
/root/reference is present, the reference's own parser.  Prints one JSON line."""
import glob, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, [p for p in glob.glob(os.path.join(ROOT, "bio*")) if os.path.isdir(p)][0]]
import numpy as np
import records

rng = np.random.default_rng(3)


def fastq_text(n, L=150):
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
    qual = rng.integers(35, 75, size=(n, L)).astype(np.uint8)
    parts = []
    for i in range(n):
        parts.append(f"@read{i}\n{seq[i].tobytes().decode()}\n+\n{qual[i].tobytes().decode()}\n")
    return "".join(parts)


def timed(mod, text, native):
    records.NATIVE_INGEST = native
    c = mod.FASTAQRecordContainer()
    t0 = time.perf_counter()
    c.parse_records(text)
    dt = time.perf_counter() - t0
    records.NATIVE_INGEST = True
    return dt


out = {}
big = fastq_text(int(sys.argv[1]) if len(sys.argv) > 1 else 400_000)
small = big[: big.find("@read20000\n")]
tiny = big[: big.find("@read4000\n")]
dt = min(timed(records, big, True) for _ in range(3))
out["native"] = {"MB": len(big) / 1e6, "seconds": dt, "MB_per_s": len(big) / 1e6 / dt}
dt = timed(records, small, False)
out["regex_restatement"] = {"MB": len(small) / 1e6, "seconds": dt, "MB_per_s": len(small) / 1e6 / dt}
if os.path.isdir("/root/reference/src"):
    import importlib.util
    saved = sys.modules.get("constants")
    spec_c = importlib.util.spec_from_file_location("constants", "/root/reference/src/constants.py")
    mod_c = importlib.util.module_from_spec(spec_c); spec_c.loader.exec_module(mod_c)
    sys.modules["constants"] = mod_c
    spec_r = importlib.util.spec_from_file_location("ref_records", "/root/reference/src/records.py")
    ref = importlib.util.module_from_spec(spec_r); spec_r.loader.exec_module(ref)
    if saved is not None:
        sys.modules["constants"] = saved
    c = ref.FASTAQRecordContainer()
    t0 = time.perf_counter()
    c.parse_records(tiny)
    dt = time.perf_counter() - t0
    out["reference"] = {"MB": len(tiny) / 1e6, "seconds": dt, "MB_per_s": len(tiny) / 1e6 / dt}
print(json.dumps(out))
