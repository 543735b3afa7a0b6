#!/usr/bin/env python
"""Compile one .cu with -Xptxas -v and print registers / spills / smem per kernel (CPU only)."""
import re, subprocess, sys
src = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
       "-diag-suppress", "128", "-Xptxas", "-v", "-c", src, "-o", "/tmp/ptxas_report.o"]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    if "error" in line:
        print(line)
    if cur and re.search(pat, cur):
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m:
            print(f"{cur[:90]:90s} regs={m.group(1):>3s} stack={stack[0]} spill_st={stack[1]} spill_ld={stack[2]}")
