#!/usr/bin/env python
"""Per-instruction view of one launch of an .ncu-rep (source page): executed warp instructions and stall samples summed over
address windows, plus the hottest instructions.  usage: ncu_hot.py report.ncu-rep [launch_index] [window_instructions]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 0; win = int(sys.argv[3]) if len(sys.argv) > 3 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]          # one table per profiled launch ...
if len(heads) >= 2 and rows[heads[0] + 1:heads[1] - 1] == rows[heads[1] + 1:heads[1] + (heads[1] - heads[0]) - 1]:
    heads = heads[::2]                                                      # ... which this ncu version prints twice
hi = heads[li]
hdr = rows[hi]
allh = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
end = min([i for i in allh if i > hi] + [len(rows)])
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
base = int(body[0][0], 16)
tot_s = sum(int(r[ix["# Samples"]]) for r in body); tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"instructions {len(body)}, executed warp instructions {tot_i}, samples {tot_s}")
print("overall stall mix:", ", ".join(f"{h[6:]}={sum(int(r[ix[h]]) for r in body)/tot_s:.3f}" for h in sorted(stalls, key=lambda h: -sum(int(r[ix[h]]) for r in body))[:9]))
print(f"{'window':>14s} {'exec %':>7s} {'samples %':>9s}  top stalls / opcode mix")
for w in range(0, len(body), win):
    seg = body[w:w + win]
    ex = sum(int(r[ix["Instructions Executed"]]) for r in seg); sm = sum(int(r[ix["# Samples"]]) for r in seg)
    if sm == 0 and ex == 0: continue
    st = sorted(((sum(int(r[ix[h]]) for r in seg), h[6:]) for h in stalls), reverse=True)[:3]
    ops = {}
    for r in seg:
        op = r[1].split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"{int(seg[0][0],16)-base:#8x}-{int(seg[-1][0],16)-base:#6x} {100*ex/tot_i:7.2f} {100*sm/tot_s:9.2f}  " +
          " ".join(f"{n}:{100*v/max(sm,1):.0f}%" for v, n in st) + " | " + " ".join(f"{k}:{100*v/max(ex,1):.0f}%" for k, v in top))
print("hottest instructions:")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  {int(r[0],16)-base:#7x} {100*int(r[ix['# Samples']])/tot_s:5.2f}%  {r[1].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in st))
ops = {}
for r in body:
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
print("executed opcode mix:", ", ".join(f"{k}:{100*v/tot_i:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
