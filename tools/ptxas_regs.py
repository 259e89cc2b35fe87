#!/usr/bin/env python
"""Summarise `ptxas -v` output (make EXTRA="-Xptxas -v" > log): registers and spills per kernel."""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None
spill = ""
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        spill = ""
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m and int(m.group(2)) > 0:
        spill = " SPILL %s B" % m.group(2)
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        if pat in name:
            print(m.group(1), name[:100], spill)
