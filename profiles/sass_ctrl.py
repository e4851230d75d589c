#!/usr/bin/env python
"""Decode the scheduling control bits of a `cuobjdump -sass` listing (sm_100a, 128-bit instructions):
stall count, yield, write / read scoreboard slot, wait mask.  usage: sass_ctrl.py file.sass [from_hex] [to_hex]"""
import re, sys
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
lines = open(sys.argv[1]).read().splitlines()
pat = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* 0x([0-9a-f]{16}) \*/')
pat2 = re.compile(r'^\s+/\* 0x([0-9a-f]{16}) \*/')
i = 0
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16)
            if lo <= addr < hi:
                w = int(m2.group(1), 16)          # bits 64..127
                ctrl = (w >> (105 - 64)) & ((1 << 17) - 1)
                stall = ctrl & 15; yld = (ctrl >> 4) & 1; wbar = (ctrl >> 5) & 7; rbar = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 63
                print(f"{addr:05x} st={stall:2d} y={yld} w={'-' if wbar == 7 else wbar} r={'-' if rbar == 7 else rbar} wait={wait:06b}  {m.group(2)[:80]}")
            i += 2
            continue
    i += 1
