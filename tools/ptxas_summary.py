#!/usr/bin/env python3
"""Summarise `-Xptxas -v` logs (csrc/*.ptxas.log): kernel, registers, stack, spills, shared memory."""
import glob
import os
import re
import subprocess
import sys

root = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "gaussiansplatting.jl_b200", "csrc")
rows = []
for path in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for '(\S+)'\s*\n.*?Function properties for \S+\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n.*?Used (\d+) registers(.*)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        smem = re.search(r"(\d+) bytes smem", m.group(7))
        rows.append((os.path.basename(path).split(".")[0], name.replace("void ", ""), int(m.group(6)), int(m.group(3)), int(m.group(4)), int(smem.group(1)) if smem else 0))
print(f"{'file':<20}{'kernel':<46}{'regs':>5}{'stack':>7}{'spill':>7}{'smem':>8}")
for r in rows:
    print(f"{r[0]:<20}{r[1]:<46}{r[2]:>5}{r[3]:>7}{r[4]:>7}{r[5]:>8}")
