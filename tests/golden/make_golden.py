#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the compiled, unmodified reference (oracle/_ref, built by
oracle/build_ref.py from /root/reference) on the synthetic cases of tests/cases.py.

Run in the build container (needs oracle/_ref):  python tests/golden/make_golden.py [case ...]
Each fixture stores the hierarchy metadata, the input component(s) and the reference's output components as the
flat concatenation (level-major, box order, [k][j][i]) of the valid regions -- bit-exact float64."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES  # noqa: E402
from oracle import oracle as O  # noqa: E402
from peleanalysis_b200 import plotfile  # noqa: E402


def flat(pf, name):
    c = pf.comp(name)
    return np.concatenate([f[c].ravel() for l in pf.levels for f in l.fabs])


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        for name, (builder, is_per, sym, tools, ckw) in CASES.items():
            if len(sys.argv) > 1 and name not in sys.argv[1:]:
                continue
            pf = builder()
            d = os.path.join(tmp, "plt_" + name)
            plotfile.write_plotfile(d, pf, clean="remove")
            rec = dict(
                names=np.array(pf.names), prob_lo=np.array(pf.prob_lo), prob_hi=np.array(pf.prob_hi),
                nlev=len(pf.levels), is_per=np.array(is_per), sym_dir=np.array(sym),
                domains=np.array([list(l.domain_lo) + list(l.domain_hi) for l in pf.levels]),
                dx=np.array([l.dx for l in pf.levels]),
                nboxes=np.array([len(l.boxes) for l in pf.levels]),
                boxes=np.array([list(lo) + list(hi) for l in pf.levels for lo, hi in l.boxes]),
            )
            for n in pf.names:
                rec["in_" + n] = flat(pf, n)
            if "grad" in tools:
                O.run_ref("grad", d, d + "_gt", gradVar="temp", is_per=is_per, sym_dir=sym)
                r = plotfile.read_plotfile(d + "_gt")
                for k, n in zip(["gx", "gy", "gz", "mag"], ["temp_gx", "temp_gy", "temp_gz", "||gradtemp||"]):
                    rec["grad_" + k] = flat(r, n)
            if "curvature" in tools:
                O.run_ref("curvature", d, d + "_K", progressName="temp", is_per=is_per, sym_dir=sym, **ckw)
                r = plotfile.read_plotfile(d + "_K")
                pmin, pmax = plotfile.file_min_max(d, "temp", len(pf.levels))
                rec["prog_min"], rec["prog_max"] = pmin, pmax
                rec["curv_opts"] = np.array([str(k) + "=" + str(v) for k, v in ckw.items()])
                skip = {"SmoothedProgress"} | ({"GaussianCurvature_temp"} if not ckw.get("do_gaussCurv") else set())
                for n in r.names:
                    if n in pf.names or n in skip:      # uninitialised in the reference when the option is off
                        continue
                    rec["curv_" + n] = flat(r, n)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
            print(name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape") and v.ndim == 1 and v.size > 64})


if __name__ == "__main__":
    main()
