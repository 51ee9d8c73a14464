#!/usr/bin/env python3
"""Static SASS statistics per kernel of an object file: instructions, and the body of the biggest loop (the widest backward
branch) -- a proxy for dynamic instruction counts while iterating on an issue-bound kernel without a GPU.
usage: tools/sass_count.py build/obj/align.cu.o [name-filter]"""
import re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "build/obj/align.cu.o"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
kern, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and cur:
        kern[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in sorted(kern.items()):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem.replace("pa::(anonymous namespace)::", "").replace("void ", ""))
    if flt and flt not in dem:
        continue
    best = (0, 0, 0)
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and addr - tgt > best[0]:
                best = (addr - tgt, tgt, addr)
    body = [t for a, t in ins if best[1] <= a <= best[2]]
    ops = {}
    for t in body:
        op = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10])
    print(f"{dem:45s} total {len(ins):5d}  main loop {len(body):5d}  [{top}]")
