#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel: tools/sass_hist.py <substring-of-mangled-name> [top]"""
import collections, re, subprocess, sys
pat = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lib = sys.argv[3] if len(sys.argv) > 3 else "pde_opt_b200/libpdeopt_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False; c = collections.Counter()
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line; continue
    if on:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m: c[m.group(2)] += 1
tot = sum(c.values())
print("TOTAL", tot)
for k, v in c.most_common(top): print(f"{v:6d} {k}")
