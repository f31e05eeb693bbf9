#!/usr/bin/env python3
"""Registers / stack / spills of every kernel, from the ptxas -v log the build keeps
(smartpy_b200/csrc/ptxas_info.txt).  Usage: python tools/ptxas_table.py [substring]"""
import re
import subprocess
import sys

log = open(sys.argv[2] if len(sys.argv) > 2 else "smartpy_b200/csrc/ptxas_info.txt").read()
want = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                     r"(\d+) bytes spill loads\n.*?Used (\d+) registers", log):
    rows.append(m.groups())
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for name, r in zip(names, rows):
    short = re.sub(r"\(anonymous namespace\)::|void |\(.*\)$", "", name).replace("(int)", "").replace("(bool)", "")
    if want in short:
        print("%-70s regs %3s  stack %4s  spill st/ld %4s/%4s" % (short, r[4], r[1], r[2], r[3]))
