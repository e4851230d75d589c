#!/usr/bin/env python
"""Static SASS instruction mix per kernel of a built library (cuobjdump -sass).

usage: python profiles/sass_mix.py <lib.so> [substring-filter]
Prints, for every kernel whose name contains the filter: total instructions, FP64-pipe instructions
(DADD/DMUL/DFMA/DSETP/DMNMX), MUFU, LDG/STG, LDS/STS, BAR.
"""
import collections, re, subprocess, sys
lib = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = None
mix = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); mix[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        mix[name][m.group(1)] += 1
for k, c in mix.items():
    if flt not in k: continue
    dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    fp64 = sum(v for op, v in c.items() if op in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
    print(f"{dem[:110]}\n   total={sum(c.values())} fp64={fp64} (DFMA={c['DFMA']} DMUL={c['DMUL']} DADD={c['DADD']} DSETP={c['DSETP']}) "
          f"MUFU={c['MUFU']} LDG={c['LDG']} STG={c['STG']} LDS={c['LDS']} STS={c['STS']} BAR={c['BAR']} CALL={c['CALL']}")
