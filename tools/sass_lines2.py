#!/usr/bin/env python3
"""Executed warp instructions and stall samples per (file, line) of one kernel: joins `ncu --page source --csv` rows with
`nvdisasm --print-line-info` of the same build, instruction by instruction, and prints the source text from the right file.
usage: tools/sass_lines2.py source_page.csv cubin mangled_name_fragment [top]"""
import collections
import csv
import os
import re
import subprocess
import sys


def main():
    page, cubin, frag = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, infun = [], None, False
    for s in dis:
        m = re.match(r"\s*\.text\.(\S+):", s)
        if m:
            infun = frag in m.group(1)
            continue
        if s.strip().startswith(".section"):
            infun = False
        if not infun:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', s)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", s):
            lines.append(cur)
    rows = list(csv.reader(open(page)))
    hdr = None
    ex, st, n = [], [], 0
    for r in rows:
        if "Instructions Executed" in r:
            hdr = r
            ie, iss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr is None or len(r) <= ie or not r[ie].isdigit():
            continue
        ex.append(int(r[ie])); st.append(int(r[iss]) if r[iss].isdigit() else 0)
    print("sass instructions: nvdisasm %d, ncu %d" % (len(lines), len(ex)))
    agg_i, agg_s = collections.Counter(), collections.Counter()
    for k, (a, b) in zip(lines, zip(ex, st)):
        agg_i[k] += a; agg_s[k] += b
    ti, ts = sum(agg_i.values()) or 1, sum(agg_s.values()) or 1
    cache = {}
    def text(k):
        if not k: return "?"
        f, l = k
        if f not in cache:
            try: cache[f] = open(f).read().split("\n")
            except Exception: cache[f] = []
        t = cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
        return "%s:%d  %s" % (os.path.basename(f), l, t)
    for k, v in agg_i.most_common(top):
        print("%6.2f%% inst %6.2f%% stall | %s" % (100.0 * v / ti, 100.0 * agg_s[k] / ts, text(k)))


if __name__ == "__main__":
    main()
