#!/usr/bin/env python
"""Static instruction count of the main loop of a kernel (cuobjdump -sass of an object file or library).

usage: python profiles/sass_loop.py <file.o|lib.so> <kernel-name-substring> [rows-per-trip]
The main loop is the longest backward branch that does not span an unconditional EXIT (the spin stubs of the mbarrier
waits sit behind it); blocks that a forward branch jumps over and that contain a CALL
(the out-of-line cold paths) are left out.  Prints instructions per row: FP64-pipe (two issue slots each on sm_100a),
others, and the issue-slot estimate 2*FP64 + others."""
import collections, re, subprocess, sys
path, flt = sys.argv[1], sys.argv[2]
rows_per_trip = int(sys.argv[3]) if len(sys.argv) > 3 else 2
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
name, ker = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1); ker[name] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and name:
        ker[name].append((int(m.group(1), 16), m.group(2).strip()))
for k, ins in ker.items():
    if flt not in k: continue
    back = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and "U.ANY" not in t:
            tgt = int(m.group(1), 16)
            if not any(tgt <= b <= a and re.search(r"\bEXIT\b", u) and not u.startswith("@") for b, u in ins): back.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    if not back: continue
    _, lo, hi = max(back)
    cold = []
    for a, t in ins:
        if not (lo <= a <= hi): continue
        m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m and a < int(m.group(1), 16) <= hi:
            tgt = int(m.group(1), 16)
            if any(a < b < tgt and "CALL" in u for b, u in ins): cold.append((a + 16, tgt))
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi and not any(x <= a < y for x, y in cold):
            c[re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]] += 1
    tot = sum(c.values()); fp = sum(v for o, v in c.items() if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    n = rows_per_trip
    print(subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:90])
    print(f"  loop 0x{lo:x}..0x{hi:x}, per row: {tot/n:.1f} instructions = {fp/n:.1f} FP64 + {(tot-fp)/n:.1f} others -> {(2*fp+tot-fp)/n:.1f} issue slots")
    print("  " + ", ".join(f"{o} {v/n:g}" for o, v in c.most_common()))
