#!/usr/bin/env python
"""Print the positions of memory/shuffle/branch instructions of one kernel (to see where ptxas put the loads).
usage: python profiles/sass_order.py <lib.so> <mangled-substring>"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name = None; k = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); k = 0; show = sys.argv[2] in name
        if show: print("==", name)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and name:
        k += 1
        if show and re.search(r"LDG|STG|SHFL|BRA|BAR|ATOM|RED|MUFU.RCP64H|EXIT", m.group(2)):
            print(k, m.group(1), m.group(2)[:90])
