#!/usr/bin/env python
"""Copy a reference Fortran file omitting the line ranges of the routines the GPU library replaces.

    python tools/split_reference.py <reference.f90> <out.f90> 221-260 264-279 465-618

The reference keeps `program` and all subroutines in one translation unit, so link-time symbol override is
unreliable (intra-file calls bind locally); dropping the replaced routines' source lines is the only edit needed.
Ranges are 1-based and inclusive (the file:line citations of include/wbeuler.h)."""
import sys


def main():
    src, dst, *ranges = sys.argv[1:]
    drop = set()
    for r in ranges:
        a, b = (int(v) for v in r.split("-"))
        drop.update(range(a, b + 1))
    with open(src) as f, open(dst, "w") as g:
        for k, line in enumerate(f, 1):
            g.write(("! [replaced by libwbeuler] " + line) if k in drop else line)


if __name__ == "__main__":
    main()
