#!/usr/bin/env python
"""Copy a reference Fortran file omitting the line ranges of the routines the GPU library replaces.

    python tools/split_reference.py <reference.f90> <out.f90> 221-260 264-279 465-618
    python tools/split_reference.py <reference.f90> <out.f90> 188-251 "56-76=  call wb_fvm1d_time_loop(u,t,dt,iter)"

The reference keeps `program` and all subroutines in one translation unit, so link-time symbol override is
unreliable (intra-file calls bind locally); dropping the replaced routines' source lines is the only edit needed.
Ranges are 1-based and inclusive (the file:line citations of include/wbeuler.h).  `A-B=text` replaces the range by
one line of text: fvm.f90 and dg_with_source.f90 keep their time loop inside the main program, where it is replaced
by a call into the shim."""
import sys


def parse(ranges):
    drop, insert = set(), {}
    for r in ranges:
        text = None
        if "=" in r:
            r, text = r.split("=", 1)
        a, b = (int(v) for v in r.split("-"))
        drop.update(range(a, b + 1))
        if text is not None:
            insert[a] = text
    return drop, insert


def split(lines, ranges):
    drop, insert = parse(ranges)
    out = []
    for k, line in enumerate(lines, 1):
        if k in insert:
            out.append(insert[k].rstrip("\n") + "      ! [libwbeuler]\n")
        out.append(("! [replaced by libwbeuler] " + line) if k in drop else line)
    return out


def main():
    src, dst, *ranges = sys.argv[1:]
    with open(src) as f:
        lines = f.readlines()
    with open(dst, "w") as g:
        g.writelines(split(lines, ranges))


if __name__ == "__main__":
    main()
