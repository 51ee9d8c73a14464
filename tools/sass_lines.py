#!/usr/bin/env python3
"""Static SASS instructions per CUDA source line inside one kernel (nvdisasm -g): where the instructions of a kernel come
from, without a GPU.  usage: tools/sass_lines.py build/obj/align.cu.o 'align_fast_split_kernelILb0ELb0ELb1E' [top]"""
import os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
inside, cur, counts, srcs = False, None, {}, {}
total = 0
for line in txt:
    if line.startswith("//--------------------- .text."):
        inside = pat in line
        continue
    if line.startswith("//--------------------- ") and inside:
        inside = False
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and cur:
        counts[cur] = counts.get(cur, 0) + 1
        total += 1
print("kernel", pat, "instructions with line info:", total)
cache = {}
def src(f, n):
    for base in ("bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200/csrc",):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in cache:
                cache[p] = open(p).read().splitlines()
            return cache[p][n - 1].strip()[:110] if n <= len(cache[p]) else ""
    return ""
for (f, n), c in sorted(counts.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{c:5d} {100 * c / total:5.1f}%  {f}:{n:<5d} {src(f, n)}")
