#!/usr/bin/env python3
"""Where a kernel's warps wait: stall samples per reason, and the SASS instructions / source lines that carry them.
Reads the source page of an .ncu-rep captured with `--set full --import-source on` (no GPU needed here):

    tools/ncu_stalls.py X.ncu-rep <kernel-name-regex> [launch-index] [top]

For every stall reason ncu samples (long scoreboard, wait, math pipe, ...) it prints the share of all samples and the
instructions with the most samples of that reason, with the CUDA source line when the cubin was built with -lineinfo and
the library in peleanalysis_b200/lib is the same build (instruction order is matched against `nvdisasm --print-line-info`)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_page(rep, kernel, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and "Instructions Executed" in r:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]):
            cur["rows"].append(r)
    if not blocks:
        raise SystemExit("no kernel matches " + kernel)
    return blocks[min(launch, len(blocks) - 1)]


def line_table(mangled_fragment):
    """source line of every SASS instruction of the kernel, in address order (None if unavailable)"""
    lib = os.path.join(ROOT, "peleanalysis_b200", "lib", "libpelestencil_b200.so")
    if not os.path.exists(lib):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
        for f in os.listdir(tmp):
            if not f.endswith(".cubin"):
                continue
            dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
            if mangled_fragment not in dis:
                continue
            lines, cur, infun = [], None, False
            for s in dis.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", s)
                if m:
                    infun = mangled_fragment in m.group(1)
                    continue
                if s.strip().startswith(".section"):
                    infun = False
                if not infun:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', s)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                elif re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", s):
                    lines.append(cur)
            return lines
    return None


def mangled_guess(name):
    """k_stencil_tma<(int)4, (int)16, (bool)0, (bool)0> -> k_stencil_tmaILi4ELi16ELb0ELb0E"""
    m = re.search(r"(k_\w+)<([^>]*)>", name)
    if not m:
        m2 = re.search(r"(k_\w+)", name)
        return m2.group(1) if m2 else name
    parts = []
    for a in m.group(2).split(","):
        a = a.strip()
        v = re.sub(r"\(\w+\)", "", a)
        parts.append(("Lb" if "(bool)" in a else "Li") + v + "E")
    return m.group(1) + "I" + "".join(parts) + "E"


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    blk = source_page(rep, kernel, launch)
    hdr, rows = blk["hdr"], blk["rows"]
    ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
    reasons = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    lines = line_table(mangled_guess(blk["name"]))
    if lines is not None and len(lines) != len(rows):
        lines = None
    tot = {c: sum(int(r[hdr.index(c)]) for r in rows if r[hdr.index(c)].isdigit()) for c in reasons}
    allsamp = sum(tot.values()) or 1
    print("kernel:", blk["name"][:140])
    print("SASS instructions %d, executed warp instructions %d, stall samples %d%s" % (
        len(rows), sum(int(r[ie]) for r in rows if r[ie].isdigit()), allsamp, "" if lines else "  (no source-line table: library differs from the profiled build)"))
    src_cache = {}
    for c, n in sorted(tot.items(), key=lambda kv: -kv[1]):
        if n * 100 < allsamp:
            continue
        print("== %-28s %5.1f %% of samples" % (c, 100.0 * n / allsamp))
        i = hdr.index(c)
        order = sorted(range(len(rows)), key=lambda k: -(int(rows[k][i]) if rows[k][i].isdigit() else 0))[:top]
        for k in order:
            v = int(rows[k][i]) if rows[k][i].isdigit() else 0
            if v == 0:
                break
            where = ""
            if lines and lines[k]:
                f, ln = lines[k]
                if f not in src_cache:
                    p = os.path.join(ROOT, "peleanalysis_b200", "csrc", f)
                    src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
                text = src_cache[f][ln - 1].strip()[:60] if 0 < ln <= len(src_cache[f]) else ""
                where = "  | %s:%d  %s" % (f, ln, text)
            print("   %5.1f %%  %-58s%s" % (100.0 * v / allsamp, rows[k][ia].strip()[:58], where))


if __name__ == "__main__":
    main()
