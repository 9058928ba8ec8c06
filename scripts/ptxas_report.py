#!/usr/bin/env python
"""Registers / spills of every kernel in one CUDA source (ptxas -v), one line per kernel.

    python scripts/ptxas_report.py iodine_b200/csrc/conv_tc.cu [filter]
"""
import re
import subprocess
import sys

src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ''
cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
       '--expt-relaxed-constexpr', '-Xptxas', '-v', '-c', src, '-o', '/dev/null']
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        spill = None
        continue
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
    if m:
        spill = m.groups()
        continue
    m = re.search(r'Used (\d+) registers', line)
    if m and name:
        rows.append((name, int(m.group(1)), spill))
        name = None
for n, r, sp in sorted(rows):
    if flt in n:
        short = re.sub(r'^void iod::', '', n)
        short = re.sub(r'\(.*\)$', '', short)
        print('%-70s regs %3d  stack %s spill st/ld %s/%s' % (short[:70], r, sp[0], sp[1], sp[2]))
