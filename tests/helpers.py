"""Shared helpers for the tests: golden fixtures <-> hierarchy objects, flat-field conversions."""
import os

import numpy as np

from peleanalysis_b200.plotfile import Level, Plotfile

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    nlev = int(z["nlev"])
    levels, o = [], 0
    for l in range(nlev):
        d = z["domains"][l]
        nb = int(z["nboxes"][l])
        boxes = [(tuple(int(v) for v in b[:3]), tuple(int(v) for v in b[3:])) for b in z["boxes"][o:o + nb]]
        o += nb
        levels.append(Level(tuple(int(v) for v in d[:3]), tuple(int(v) for v in d[3:]), tuple(float(v) for v in z["dx"][l]), boxes))
    names = [str(n) for n in z["names"]]
    ratios = [(levels[l + 1].domain_hi[0] + 1) // (levels[l].domain_hi[0] + 1) for l in range(nlev - 1)]
    pf = Plotfile(names, tuple(float(v) for v in z["prob_lo"]), tuple(float(v) for v in z["prob_hi"]), ratios, levels)
    # attach fabs from the flat inputs
    flats = [z["in_" + n] for n in names]
    o = 0
    for lv in levels:
        lv.fabs = []
        for lo, hi in lv.boxes:
            n = [hi[d] - lo[d] + 1 for d in range(3)]
            m = n[0] * n[1] * n[2]
            lv.fabs.append(np.stack([f[o:o + m].reshape(n[2], n[1], n[0]) for f in flats]))
            o += m
    return pf, z


def flat_from_fabs(fabs_per_level):
    return np.concatenate([np.asarray(f).ravel() for lv in fabs_per_level for f in lv])


def fabs_from_flat(pf, flat):
    out, o = [], 0
    for lv in pf.levels:
        lst = []
        for lo, hi in lv.boxes:
            n = [hi[d] - lo[d] + 1 for d in range(3)]
            m = n[0] * n[1] * n[2]
            lst.append(flat[o:o + m].reshape(n[2], n[1], n[0]))
            o += m
        out.append(lst)
    return out


def bit_equal(a, b):
    """Bit-for-bit equality of float64 arrays -- the sign of zero included -- treating any NaN == any NaN."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    same = a.view(np.int64) == b.view(np.int64)
    return bool(np.all(same | (np.isnan(a) & np.isnan(b))))


def max_rel(a, b):
    d = np.max(np.abs(np.asarray(a) - np.asarray(b)))
    s = np.max(np.abs(np.asarray(b)))
    return float(d / s) if s > 0 else float(d)
