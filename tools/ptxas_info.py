#!/usr/bin/env python
"""Compile the CUDA source with -Xptxas -v and print registers / spills per kernel
(development tool).  usage: tools/ptxas_info.py [substring-filter]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import build  # noqa: E402


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ''
    cmd = [build._nvcc(), *build.NVCC_FLAGS, '-Xptxas', '-v', '-o', '/tmp/_ptxas_info.so',
           build.SOURCE]
    out = subprocess.run(cmd, capture_output=True, text=True)
    text = out.stderr + out.stdout
    name = None
    spill = ''
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True,
                                  text=True).stdout.strip()
            continue
        if 'spill' in line:
            spill = line.strip()
        m = re.search(r'Used (\d+) registers', line)
        if m and name and flt in name:
            short = re.sub(r'\(anonymous namespace\)::', '', name)
            short = re.sub(r'\(.*\)$', '', short).replace('void ', '')
            print(f'{short:60s} regs {m.group(1):>3s}  {spill}')
    if out.returncode:
        print(text[-3000:])
        sys.exit(out.returncode)


if __name__ == '__main__':
    main()
