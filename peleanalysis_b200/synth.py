"""Deterministic synthetic AMR hierarchies (no RNG: analytic fields sampled at cell centres).

These are the inputs of BASELINE.json's five configs plus the branch-coverage cases of SURVEY.md §4
(periodic wrap, Neumann / reflect_odd walls, coarse-fine faces next to the domain edge, L-shaped fine
regions, refinement ratio 4, non-power-of-two dx, 16^3 boxes).  Index-space conventions are AMReX's:
inclusive cell boxes, level l+1 index = ratio * level l index.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence

import numpy as np

from .plotfile import Level, Plotfile

FIELD_NAMES = ["x_velocity", "y_velocity", "z_velocity", "temp", "Y_CH4"]


def chop(lo, hi, max_grid_size: int) -> List[tuple]:
    """Split box [lo,hi] into boxes of at most max_grid_size per side (equal-ish chunks, like BoxArray::maxSize)."""
    cuts = []
    for d in range(3):
        n = hi[d] - lo[d] + 1
        nchunk = (n + max_grid_size - 1) // max_grid_size
        base, rem = divmod(n, nchunk)
        edges = [lo[d]]
        for c in range(nchunk):
            edges.append(edges[-1] + base + (1 if c < rem else 0))
        cuts.append(edges)
    out = []
    for kz in range(len(cuts[2]) - 1):
        for ky in range(len(cuts[1]) - 1):
            for kx in range(len(cuts[0]) - 1):
                out.append(((cuts[0][kx], cuts[1][ky], cuts[2][kz]),
                            (cuts[0][kx + 1] - 1, cuts[1][ky + 1] - 1, cuts[2][kz + 1] - 1)))
    return out


def field_values(name: str, x, y, z, prob_hi=(1.0, 1.0, 1.0), phase: float = 0.0):
    """Analytic fields.  x,y,z are broadcastable arrays of physical cell-centre coordinates."""
    X, Y, Z = x / prob_hi[0], y / prob_hi[1], z / prob_hi[2]
    tp = 2.0 * math.pi
    if name == "temp":
        r = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2)
        return 300.0 + 750.0 * (1.0 + np.tanh((0.25 - r) / 0.05)) + 5.0 * np.sin(tp * X + phase) * np.cos(2 * tp * Y) + 0.0 * Z
    if name == "x_velocity":
        return np.sin(tp * X + phase) * np.cos(tp * Y) * np.cos(tp * Z)
    if name == "y_velocity":
        return -np.cos(tp * X + phase) * np.sin(tp * Y) * np.cos(tp * Z)
    if name == "z_velocity":
        return 0.3 * np.sin(2 * tp * Z + phase) * np.cos(tp * X) + 0.0 * Y
    if name == "Y_CH4":
        r = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2)
        return 0.05 * (1.0 - np.tanh((0.25 - r) / 0.05)) + 0.001 * np.sin(tp * (X + Y + Z) + phase)
    if name.startswith("mix"):          # config-5 style: phase-shifted sin/tanh mixes, mixNN
        m = int(name[3:])
        ph = 0.37 * m + phase
        r = np.sqrt((X - 0.5) ** 2 + (Y - 0.45) ** 2 + (Z - 0.55) ** 2)
        return (np.sin(tp * X * (1 + m % 3) + ph) * np.cos(tp * Y * (1 + m % 2) - ph)
                + np.tanh((0.3 - r) / (0.04 + 0.01 * m)) + 0.25 * np.sin(tp * Z + 2 * ph))
    raise KeyError(name)


def fill_level(level: Level, names: Sequence[str], prob_lo, prob_hi, fn: Callable = field_values) -> None:
    level.fabs = []
    for lo, hi in level.boxes:
        ax = []
        for d in range(3):
            idx = np.arange(lo[d], hi[d] + 1, dtype=np.float64)
            ax.append(prob_lo[d] + (idx - level.domain_lo[d] + 0.5) * level.dx[d])
        x = ax[0][None, None, :]
        y = ax[1][None, :, None]
        z = ax[2][:, None, None]
        fab = np.empty((len(names), hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1))
        for c, n in enumerate(names):
            fab[c] = fn(n, x, y, z, prob_hi)
        level.fabs.append(fab)


def make_hierarchy(base_n, fine_regions: Sequence[Sequence[tuple]] = (), ratios: Sequence[int] = (),
                   max_grid_size: int = 32, names: Sequence[str] = ("temp",), prob_lo=(0.0, 0.0, 0.0),
                   prob_hi=(1.0, 1.0, 1.0), fill: bool = True, header_ratio: int | None = None) -> Plotfile:
    """base_n: cells per side of level 0 (int or 3-tuple).
    fine_regions[l-1]: list of (lo,hi) boxes in level-l index space that make up level l (each is chopped).
    ratios[l-1]: refinement ratio between level l-1 and l."""
    if isinstance(base_n, int):
        base_n = (base_n,) * 3
    levels = []
    n = tuple(base_n)
    dom_lo, dom_hi = (0, 0, 0), tuple(v - 1 for v in n)
    # dx exactly as amrex Geometry: (prob_hi - prob_lo) / N  (AMReX_Geometry.cpp:520)
    dx = tuple((prob_hi[d] - prob_lo[d]) / n[d] for d in range(3))
    levels.append(Level(dom_lo, dom_hi, dx, chop(dom_lo, dom_hi, max_grid_size)))
    for l, regions in enumerate(fine_regions):
        r = ratios[l]
        n = tuple(v * r for v in n)
        dom_hi = tuple(v - 1 for v in n)
        dx = tuple((prob_hi[d] - prob_lo[d]) / n[d] for d in range(3))
        boxes = []
        for lo, hi in regions:
            boxes += chop(tuple(lo), tuple(hi), max_grid_size)
        levels.append(Level(dom_lo, dom_hi, dx, boxes))
    pf = Plotfile(list(names), tuple(prob_lo), tuple(prob_hi),
                  [header_ratio or r for r in ratios[: len(fine_regions)]], levels)
    if fill:
        for lv in levels:
            fill_level(lv, names, prob_lo, prob_hi)
    return pf


def central_half(n_coarse, ratio: int):
    """Fine box (fine index space) refining the central half of a coarse region of n_coarse cells per side
    starting at coarse index 0."""
    lo = tuple((v // 4) * ratio for v in n_coarse)
    hi = tuple((v // 4 + v // 2) * ratio - 1 for v in n_coarse)
    return lo, hi


# ----------------------------------------------------------------------------------------------
# Named cases.  `scale` shrinks config sizes for CPU-side tests (the structure is unchanged).
# ----------------------------------------------------------------------------------------------

def config1(base: int = 64, mgs: int = 32, names=("temp",), corner: bool = False) -> Plotfile:
    """BASELINE config 1: 2 levels, ratio 2, L1 refines the central half (or the low corner)."""
    h = base // 2
    if corner:
        reg = [((0, 0, 0), (2 * h - 1,) * 3)]
    else:
        reg = [((base // 4 * 2,) * 3, ((base // 4 + h) * 2 - 1,) * 3)]
    return make_hierarchy(base, [reg], [2], mgs, names)


def config2(n: int = 512, mgs: int = 128, names=tuple(FIELD_NAMES), fill: bool = True) -> Plotfile:
    """BASELINE config 2: uniform single level, 5 components."""
    return make_hierarchy(n, [], [], mgs, names, fill=fill)


def config3(base: int = 256, mgs: int = 64, names=("temp",), nlev: int = 3, fill: bool = True) -> Plotfile:
    """BASELINE config 3: 3 levels ratio 2, each level refines the central half of the previous one."""
    regs = []
    lo = (0, 0, 0)
    n = (base,) * 3
    for _ in range(nlev - 1):
        # region of previous level occupies [lo, lo+n) in its own index space
        q = tuple(v // 4 for v in n)
        flo = tuple((lo[d] + q[d]) * 2 for d in range(3))
        fn = tuple(v // 2 * 2 for v in n)          # half the cells, refined by 2 -> same count
        fhi = tuple(flo[d] + fn[d] - 1 for d in range(3))
        regs.append([(flo, fhi)])
        lo, n = flo, fn
    return make_hierarchy(base, regs, [2] * (nlev - 1), mgs, names, fill=fill)


def config4(n: int = 1024, mgs: int = 128, names=("temp",), fill: bool = True) -> Plotfile:
    """BASELINE config 4: uniform periodic, one variable (strong-scaling case)."""
    return make_hierarchy(n, [], [], mgs, names, fill=fill)


def config5(base: int = 128, mgs: int = 16, ncomp: int = 12, ratios=(2, 4, 2), fill: bool = True) -> Plotfile:
    """BASELINE config 5: 4 levels, mixed ratios, many small FABs.  Level l+1 is a slab-like block
    inside level l (properly nested with >= 2 coarse cells of margin)."""
    names = ["mix%02d" % i for i in range(ncomp)]
    regs = []
    lo = (0, 0, 0)
    n = (base,) * 3
    frac = [(4, 2), (8, 2), (4, 2)]
    for l, r in enumerate(ratios):
        q = tuple(max(2, v // 4) for v in n)
        ext = (n[0] // 2, n[1] // 2, max(mgs // r, n[2] // 4))
        flo = tuple((lo[d] + q[d]) * r for d in range(3))
        fhi = tuple(flo[d] + ext[d] * r - 1 for d in range(3))
        regs.append([(flo, fhi)])
        lo, n = flo, tuple(e * r for e in ext)
    return make_hierarchy(base, regs, list(ratios), mgs, names, fill=fill)


def case_lshape(base: int = 32, mgs: int = 16, names=("temp",)) -> Plotfile:
    """L-shaped fine level: tangential neighbours of some c-f ghost cells are covered by another fine box."""
    b = base
    q = b // 4 * 2
    h = b // 4 * 2
    regs = [[((q, q, q), (q + 2 * h - 1, q + h - 1, q + h - 1)),
             ((q, q + h, q), (q + h - 1, q + 2 * h - 1, q + h - 1))]]
    return make_hierarchy(base, regs, [2], mgs, names)


def case_ratio4(base: int = 32, mgs: int = 32, names=("temp",)) -> Plotfile:
    q = base // 4
    return make_hierarchy(base, [[((q * 4,) * 3, ((q + base // 2) * 4 - 1,) * 3)]], [4], mgs, names, header_ratio=4)


def case_np2(names=("temp",)) -> Plotfile:
    """Non-cubic, non-power-of-two dx: 48x24x40 cells on a 0.7 x 0.35 x 1.3 domain, one refined block."""
    base = (48, 24, 40)
    reg = [((24, 12, 20), (71, 35, 59))]
    return make_hierarchy(base, [reg], [2], 24, names, prob_hi=(0.7, 0.35, 1.3))


def case_edge(base: int = 32, mgs: int = 16, names=("temp",)) -> Plotfile:
    """Fine region touching the low-x/low-y domain walls and the high-z wall (domain-edge branches)."""
    n2 = 2 * base
    reg = [((0, 0, n2 - base), (base - 1, base - 1, n2 - 1))]
    return make_hierarchy(base, [reg], [2], mgs, names)


def case_thin(names=("temp",)) -> Plotfile:
    """Boxes only 2 cells thick in x on the fine level: exercises NX = min(blen+1, 4) < 4 in the c-f polynomial."""
    base = 16
    reg = [((8, 8, 8), (9, 23, 23)), ((10, 8, 8), (11, 23, 23)), ((12, 8, 8), (23, 23, 23))]
    pf = make_hierarchy(base, [reg], [2], 16, names)
    return pf


def case_mixed(names=("temp",)) -> Plotfile:
    """Boxes of three different sizes on one level (64x16x16, 32x16x16, 8^3, abutting raggedly) over a single 32x16x16
    coarse box: every CTA-shape class of the TMA stencil occurs in one hierarchy, so one stencil pass is several launches."""
    reg = [((0, 0, 0), (63, 15, 15)), ((0, 16, 0), (7, 23, 7)), ((8, 16, 0), (39, 31, 15))]
    return make_hierarchy((32, 16, 16), [reg], [2], 64, names)


CASES: Dict[str, Callable[..., Plotfile]] = {
    "config1": config1, "config2": config2, "config3": config3, "config4": config4, "config5": config5,
    "lshape": case_lshape, "ratio4": case_ratio4, "np2": case_np2, "edge": case_edge, "thin": case_thin,
    "mixed": case_mixed,
}
