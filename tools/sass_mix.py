#!/usr/bin/env python3
"""Opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv` output: share of executed warp instructions and
stall samples per SASS opcode.  usage: tools/sass_mix.py source_page.csv [top]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr = None
    ops, stall, tot = collections.Counter(), collections.Counter(), 0
    for r in rows:
        if "Instructions Executed" in r:
            hdr = r
            ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr is None or len(r) <= ie:
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
        if not m or not r[ie].isdigit():
            continue
        op = m.group(2).split(".")[0]
        n = int(r[ie])
        ops[op] += n
        tot += n
        stall[op] += int(r[iss]) if r[iss].isdigit() else 0
    print("total warp instructions", tot)
    st = sum(stall.values()) or 1
    for op, n in ops.most_common(top):
        print("%-10s %6.2f%% of instructions   %6.2f%% of stall samples" % (op, 100.0 * n / tot, 100.0 * stall[op] / st))


if __name__ == "__main__":
    main()
