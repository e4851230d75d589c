#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.
usage: python profiles/ncu_summary.py <report.ncu-rep> [cells_per_launch]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def g(r, k):
    return r[hdr.index(k)] if k in hdr else "n/a"
keys = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__cycles_elapsed.avg.per_second", "sm clock"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %")]
for r in rows[2:]:
    print("kernel:", g(r, "Kernel Name"), " grid", g(r, "Grid Size"), " block", g(r, "Block Size"))
    for k, name in keys:
        if k in hdr:
            print(f"  {name:26s} {g(r,k):>18s} {units[hdr.index(k)]}")
    if cells:
        def val(k):
            v = float(g(r, k)); u = units[hdr.index(k)].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
        tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        print(f"  dram bytes per cell         {tot / cells:18.2f} B  (read {val('dram__bytes_read.sum')/cells:.2f} + write {val('dram__bytes_write.sum')/cells:.2f})")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
