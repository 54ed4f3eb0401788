#!/usr/bin/env python3
"""Generate tests/golden/filter_*.npz by running the compiled, unmodified reference filterPlt tool
(oracle/_ref/filterPlt3d.ref.ex, built by oracle/build_ref.py from /root/reference) on the cases of
tests/cases.py FILTER_CASES / FILTER_TYPE_SWEEP.

Run in the build container (needs oracle/_ref):  python tests/golden/make_golden_filter.py [case ...]
A fixture stores the input hierarchy like the grad / curvature fixtures, the tool options, the output grids (the input
grids re-chopped to max_grid_size) and every output variable as the flat concatenation (level-major, output box order,
[k][j][i]) of the valid regions -- bit-exact float64."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import FILTER_CASES, FILTER_TYPE_SWEEP  # noqa: E402
from oracle import oracle as O  # noqa: E402
from peleanalysis_b200 import plotfile  # noqa: E402


def flat(pf, name):
    c = pf.comp(name)
    return np.concatenate([f[c].ravel() for l in pf.levels for f in l.fabs])


def hier_record(pf):
    rec = dict(
        names=np.array(pf.names), prob_lo=np.array(pf.prob_lo), prob_hi=np.array(pf.prob_hi), nlev=len(pf.levels),
        domains=np.array([list(l.domain_lo) + list(l.domain_hi) for l in pf.levels]),
        dx=np.array([l.dx for l in pf.levels]),
        nboxes=np.array([len(l.boxes) for l in pf.levels]),
        boxes=np.array([list(lo) + list(hi) for l in pf.levels for lo, hi in l.boxes]),
    )
    for n in pf.names:
        rec["in_" + n] = flat(pf, n)
    return rec


def run(d, opts):
    out = os.path.join(os.path.dirname(d), os.path.basename(d) + "_filtered")     # the tool's fixed output name, in its cwd
    O.run_ref("filterPlt", d, out, **opts)
    return plotfile.read_plotfile(out)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    want = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, (builder, opts) in FILTER_CASES.items():
            if want and name not in want:
                continue
            pf = builder()
            d = os.path.join(tmp, "plt_" + name)
            plotfile.write_plotfile(d, pf, clean="remove")
            r = run(d, opts)
            rec = hier_record(pf)
            rec["opts"] = np.array([str(k) + "=" + str(v) for k, v in opts.items()])
            rec["out_names"] = np.array(r.names)
            rec["out_nboxes"] = np.array([len(l.boxes) for l in r.levels])
            rec["out_boxes"] = np.array([list(lo) + list(hi) for l in r.levels for lo, hi in l.boxes])
            for n in r.names:
                rec["out_" + n] = flat(r, n)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
            print(name, opts, "levels", len(r.levels), "cells", rec["out_" + r.names[0]].size)
        name, builder, combos = FILTER_TYPE_SWEEP
        if not want or name in want:
            pf = builder()
            d = os.path.join(tmp, "plt_" + name)
            plotfile.write_plotfile(d, pf, clean="remove")
            rec = hier_record(pf)
            rec["combos"] = np.array(combos)
            for t, f in combos:
                r = run(d, dict(filter_type=t, base_fgr=f, same_fgr_all_levels=1))
                rec["out_t%d_f%d" % (t, f)] = flat(r, r.names[0])
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
            print(name, len(combos), "runs")


if __name__ == "__main__":
    main()
