#!/usr/bin/env python3
"""ptxas -v summary of one .cu file: registers / stack / spills per kernel (no GPU needed)."""
import re, subprocess, sys, os
src = sys.argv[1]
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v", "-c", src,
                      "-o", "/tmp/ptxas_info.o"] + sys.argv[2:], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(src))).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("pa::(anonymous namespace)::", "").replace("void ", ""))
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print(f"{name:55s} regs {m.group(1):>3s}  stack {stack[0]:>4s}  spill st/ld {stack[1]:>4s}/{stack[2]:>4s}")
        name = None
