"""CPU restatement (NumPy) of the reference's filterPlt path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the
product package never imports it.  Parity is PINNED: tests/test_filter_oracle.py compares this restatement bit
for bit with the compiled, unmodified reference tool (oracle/_ref/filterPlt3d.ref.ex, built by
oracle/build_ref.py) through the fixtures tests/golden/filter_*.npz.

R  = /root/reference, AX = R/Submodules/PelePhysics/Submodules/amrex/Src, PP = R/Submodules/PelePhysics/Source

What the tool does (R/Src/filterPlt.cpp:100-222):
  * the plotfile's grids of every level are re-chopped to max_grid_size (BoxArray::maxSize,
    AX/Base/AMReX_BoxList.cpp:765-815) and their valid cells filled with the plotfile's data of the same
    level (PP/Utility/PltFileManager/PltFileManager.cpp:165-300; the geometry is NON-periodic, :118-120);
  * ghost cells, nGrow = filter-to-grid ratio / 2 (PP/Utility/Filter/Filter.cpp): level 0 by
    FillPatchSingleLevel, finer levels by FillPatchTwoLevels (AX/AmrCore/AMReX_FillPatchUtil_I.H) --
    same-level valid data where a box covers the cell, else interpolation from the VALID cells of the next
    coarser level (MFCellConsLinInterp with mcslope, AX/AmrCore/AMReX_MFInterp_3D_C.H:178-260, or MFPCInterp),
    and first-order extrapolation outside the domain (faces, then edges, then corners:
    AX/Base/AMReX_PhysBCFunct.H:406-640 + AX/Base/AMReX_FilCC_3D_C.H), which is a clamp of the index into
    the domain -- on the coarse patch too;
  * Filter::apply_filter on every box (PP/Utility/Filter/Filter.H:28-50): qh = 0; for n, m, l (k, j, i
    offsets, i fastest): qh += w[l]*w[m]*w[n] * q(i+l, j+m, k+n), separate multiplies and adds.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


# ---------------------------------------------------------------------------------------------- weights
def _box_weights(fgr: int):
    ng = fgr // 2
    n = 2 * ng + 1
    w = [1.0 / fgr] * n
    if fgr > 1:
        w[0] = 0.5 * w[0]
        w[n - 1] = w[0]
    return ng, w


def _box_3pt(fgr: int):
    w0 = fgr * fgr / 24.0
    return 1, [w0, (12.0 - fgr * fgr) / 12.0, w0]


def _gauss_5pt(fgr: int):
    f2 = fgr * fgr
    f4 = f2 * f2
    w0 = (f4 - 4.0 * f2) / 1152.0
    w1 = (16.0 * f2 - f4) / 288.0
    return 2, [w0, w1, (f4 - 20.0 * f2 + 192.0) / 192.0, w1, w0]


_OPT3_BOX = {1: 0.079, 2: 0.274, 3: 1.377, 4: -2.375, 5: -1.000, 6: -0.779, 7: -0.680, 8: -0.627, 9: -0.596, 10: -0.575}
_OPT3_GAUSS = {1: 0.0763, 2: 0.2527, 3: 1.1160, 4: -3.144, 5: -1.102, 6: -0.809, 7: -0.696, 8: -0.638, 9: -0.604, 10: -0.581}
_OPT5_BOX = {1: (0.0886, -0.0169), 2: (0.3178, -0.0130), 3: (1.0237, 0.0368), 4: (2.4414, 0.5559), 5: (0.2949, 0.7096),
             6: (-0.5276, 0.4437), 7: (-0.6708, 0.3302), 8: (-0.7003, 0.2767), 9: (-0.7077, 0.2532), 10: (-0.6996, 0.2222)}
_OPT5_GAUSS = {1: (0.0871, -0.0175), 2: (0.2596, -0.0021), 3: (0.4740, 0.0785), 4: (0.1036, 0.2611), 5: (-0.4252, 0.3007),
               6: (-0.6134, 0.2696), 7: (-0.6679, 0.2419), 8: (-0.6836, 0.2231), 9: (-0.6873, 0.2103), 10: (-0.6870, 0.2014)}


def filter_weights(ftype: int, fgr: int):
    """(ngrow, weights) of Filter(type, fgr) -- PP/Utility/Filter/Filter.H:56-111 and Filter.cpp:3-404."""
    if ftype == 1:
        return _box_weights(fgr)
    if ftype == 2:                                  # Filter.cpp:27-52
        ng = fgr // 2
        n = 2 * ng + 1
        sigma = math.sqrt(1.0 / (2.0 * 6.0)) * fgr
        w = [1.0 / (math.sqrt(2.0 * math.pi) * sigma) * math.exp((-(i - ng) * (i - ng)) / (2 * sigma * sigma)) for i in range(n)]
        s = 0.0
        for x in w:
            s = s + x
        return ng, [x / s for x in w]
    if ftype in (3, 7):
        return _box_3pt(fgr)
    if ftype == 4:                                  # Filter.cpp:74-92
        f2 = fgr * fgr
        f4 = f2 * f2
        w0 = (3.0 * f4 - 20.0 * f2) / 5760.0
        w1 = (80.0 * f2 - 3.0 * f4) / 1440.0
        return 2, [w0, w1, (3.0 * f4 - 100.0 * f2 + 960.0) / 960.0, w1, w0]
    if ftype in (5, 9):
        tab = _OPT3_BOX if ftype == 5 else _OPT3_GAUSS
        if fgr not in tab:
            return _box_weights(fgr) if ftype == 5 else _box_3pt(fgr)
        r = tab[fgr]
        w0 = r / (1 + 2.0 * r)
        return 1, [w0, 1.0 - 2.0 * w0, w0]
    if ftype in (6, 10):
        tab = _OPT5_BOX if ftype == 6 else _OPT5_GAUSS
        if fgr not in tab:
            return _box_weights(fgr) if ftype == 6 else _gauss_5pt(fgr)
        r1, r2 = tab[fgr]
        w0 = r2 / (1 + 2.0 * r1 + 2.0 * r2)
        w1 = r1 / r2 * w0
        return 2, [w0, w1, 1.0 - 2.0 * w0 - 2.0 * w1, w1, w0]
    if ftype == 8:
        return _gauss_5pt(fgr)
    return 0, [1.0]                                 # no_filter / unknown type


# ---------------------------------------------------------------------------------------------- grids
def max_size(boxes: Sequence[tuple], chunk: int) -> List[tuple]:
    """BoxList::maxSize (AX/Base/AMReX_BoxList.cpp:765-815): every box replaced, in place, by its chunks."""
    out = []
    for lo, hi in boxes:
        ln = [hi[d] - lo[d] + 1 for d in range(3)]
        ratio, numblk, extra, sz = [1, 1, 1], [1, 1, 1], [0, 0, 0], list(ln)
        for d in range(3):
            if ln[d] > chunk:
                bs, nlen = chunk, ln[d]
                while bs % 2 == 0 and nlen % 2 == 0:
                    ratio[d] *= 2
                    bs //= 2
                    nlen //= 2
                numblk[d] = (nlen + bs - 1) // bs
                sz[d] = nlen // numblk[d]
                extra[d] = nlen - sz[d] * numblk[d]
        if numblk == [1, 1, 1]:
            out.append((tuple(lo), tuple(hi)))
            continue

        def cut(d, a):
            if a < extra[d]:
                l0 = a * (sz[d] + 1) * ratio[d]
                h0 = l0 + (sz[d] + 1) * ratio[d] - 1
            else:
                l0 = (a * sz[d] + extra[d]) * ratio[d]
                h0 = l0 + sz[d] * ratio[d] - 1
            return l0 + lo[d], h0 + lo[d]
        for k in range(numblk[2]):
            klo, khi = cut(2, k)
            for j in range(numblk[1]):
                jlo, jhi = cut(1, j)
                for i in range(numblk[0]):
                    ilo, ihi = cut(0, i)
                    out.append(((ilo, jlo, klo), (ihi, jhi, khi)))
    return out


def _dense(level, comp: int):
    """valid data of one component of a level on a dense domain-shaped array (NaN where no box) + coverage mask"""
    n = [level.domain_hi[d] - level.domain_lo[d] + 1 for d in range(3)]
    D = np.full((n[2], n[1], n[0]), np.nan)
    M = np.zeros((n[2], n[1], n[0]), dtype=bool)
    for (lo, hi), fab in zip(level.boxes, level.fabs):
        s = tuple(slice(lo[d] - level.domain_lo[d], hi[d] - level.domain_lo[d] + 1) for d in (2, 1, 0))
        D[s] = fab[comp]
        M[s] = True
    return D, M


def _mc_slope(um, u0, up):
    # AX/AmrCore/AMReX_MFInterp_3D_C.H:190-196 (BC foextrap: the centred slope has no one-sided form)
    dc = 0.5 * (up - um)
    df = 2.0 * (up - u0)
    db = 2.0 * (u0 - um)
    with np.errstate(invalid="ignore"):
        s = np.where(df * db >= 0.0, np.minimum(np.abs(df), np.abs(db)), 0.0)
        return np.copysign(1.0, dc) * np.minimum(s, np.abs(dc))


def interp_from_coarse(C: np.ndarray, r: int, interp_type: int) -> np.ndarray:
    """Every fine cell of the refined coarse domain interpolated from the dense coarse array C (NaN where the coarse
    level has no valid cell; outside the domain C is clamped: first-order extrapolation of the coarse patch)."""
    nz, ny, nx = C.shape
    if interp_type != 1:                            # MFPCInterp: fine = crse(coarsen(i))
        return np.repeat(np.repeat(np.repeat(C, r, axis=0), r, axis=1), r, axis=2)
    U = np.pad(C, 1, mode="edge")
    c = U[1:-1, 1:-1, 1:-1]
    sx = _mc_slope(U[1:-1, 1:-1, :-2], c, U[1:-1, 1:-1, 2:])
    sy = _mc_slope(U[1:-1, :-2, 1:-1], c, U[1:-1, 2:, 1:-1])
    sz = _mc_slope(U[:-2, 1:-1, 1:-1], c, U[2:, 1:-1, 1:-1])
    with np.errstate(invalid="ignore", divide="ignore"):
        # :216-238; Real(r-1)/Real(2r) is applied as (|s| * (r-1)) / (2r), left to right
        dumax = (np.abs(sx) * float(r - 1)) / float(2 * r) + (np.abs(sy) * float(r - 1)) / float(2 * r) + (np.abs(sz) * float(r - 1)) / float(2 * r)
        umax = c.copy()
        umin = c.copy()
        for k in range(3):
            for j in range(3):
                for i in range(3):
                    v = U[k:k + nz, j:j + ny, i:i + nx]
                    umin = np.where(v < umin, v, umin)        # amrex::min / max are std::min / max: (v < umin) ? v : umin
                    umax = np.where(v > umax, v, umax)
        alpha = np.ones_like(c)
        nz_ = (sx != 0.0) | (sy != 0.0) | (sz != 0.0)
        a1 = np.where(nz_ & (dumax * alpha > (umax - c)), (umax - c) / dumax, alpha)
        a2 = np.where(nz_ & (dumax * a1 > (c - umin)), (c - umin) / dumax, a1)
        slx, sly, slz = sx * a2, sy * a2, sz * a2
    off = np.array([(float(i) + 0.5) / float(r) - 0.5 for i in range(r)])       # :253-255
    rep = lambda A: np.repeat(np.repeat(np.repeat(A, r, axis=0), r, axis=1), r, axis=2)
    xo = np.tile(off, nx)[None, None, :]
    yo = np.tile(off, ny)[None, :, None]
    zo = np.tile(off, nz)[:, None, None]
    with np.errstate(invalid="ignore"):
        return ((rep(c) + xo * rep(slx)) + yo * rep(sly)) + zo * rep(slz)       # :256-259


def level_field(levels, l: int, comp: int, interp_type: int) -> np.ndarray:
    """V_l on the whole level domain: the level's own valid data where a box covers the cell, else the interpolation
    of the VALID data of level l-1 (FillPatchTwoLevels; NaN where neither exists)."""
    D, M = _dense(levels[l], comp)
    if l == 0:
        return D
    lc = levels[l - 1]
    r = (levels[l].domain_hi[0] - levels[l].domain_lo[0] + 1) // (lc.domain_hi[0] - lc.domain_lo[0] + 1)
    Cd, _ = _dense(lc, comp)
    return np.where(M, D, interp_from_coarse(Cd, r, interp_type))


def grown_fab(V: np.ndarray, level, lo, hi, ng: int) -> np.ndarray:
    """the box grown by ng ghost cells: V at the index clamped into the domain (first-order extrapolation)"""
    P = np.pad(V, ng, mode="edge") if ng > 0 else V
    s = tuple(slice(lo[d] - level.domain_lo[d], hi[d] - level.domain_lo[d] + 1 + 2 * ng) for d in (2, 1, 0))
    return P[s]


def apply_filter(q: np.ndarray, ng: int, w: Sequence[float]) -> np.ndarray:
    """Filter::apply_filter on one grown fab q [nz+2ng][ny+2ng][nx+2ng] (Filter.H:28-50)"""
    nz, ny, nx = (q.shape[0] - 2 * ng, q.shape[1] - 2 * ng, q.shape[2] - 2 * ng)
    out = np.zeros((nz, ny, nx))
    for n in range(-ng, ng + 1):
        for m in range(-ng, ng + 1):
            for l in range(-ng, ng + 1):
                c = (w[l + ng] * w[m + ng]) * w[n + ng]
                out = out + c * q[ng + n:ng + n + nz, ng + m:ng + m + ny, ng + l:ng + l + nx]
    return out


def level_fgr(base_fgr: int, ratios: Sequence[int], same_fgr_all_levels: bool, lev: int) -> int:
    f = base_fgr
    if not same_fgr_all_levels:
        for l in range(1, lev + 1):
            f *= ratios[l - 1]
    return f


def filter_plotfile(pf, filter_type: int = 1, base_fgr: int = 2, same_fgr_all_levels: bool = False, max_grid_size: int = 32,
                    interp_type: int = 1, variables: Sequence[str] | None = None, max_filter_level: int = 1000):
    """filterPlt on a plotfile.Plotfile: returns (names, per level: (boxes, [fab [ncomp][nz][ny][nx] per box]), grown
    input fabs per level for ghost-cell parity)."""
    names = list(variables) if variables else list(pf.names)
    comps = [pf.comp(n) for n in names]
    nlev = min(max_filter_level + 1, len(pf.levels))
    ratios = []
    for l in range(1, len(pf.levels)):
        ratios.append((pf.levels[l].domain_hi[0] - pf.levels[l].domain_lo[0] + 1) // (pf.levels[l - 1].domain_hi[0] - pf.levels[l - 1].domain_lo[0] + 1))
    out_levels, grown_levels = [], []
    for l in range(nlev):
        ng, w = filter_weights(filter_type, level_fgr(base_fgr, ratios, same_fgr_all_levels, l))
        boxes = max_size(pf.levels[l].boxes, max_grid_size)
        fabs = [np.empty((len(comps), hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)) for lo, hi in boxes]
        grown = [np.empty((len(comps), hi[2] - lo[2] + 1 + 2 * ng, hi[1] - lo[1] + 1 + 2 * ng, hi[0] - lo[0] + 1 + 2 * ng)) for lo, hi in boxes]
        for ci, c in enumerate(comps):
            V = level_field(pf.levels, l, c, interp_type)
            for b, (lo, hi) in enumerate(boxes):
                q = grown_fab(V, pf.levels[l], lo, hi, ng)
                grown[b][ci] = q
                fabs[b][ci] = apply_filter(q, ng, w)
        out_levels.append((boxes, fabs))
        grown_levels.append(grown)
    return names, out_levels, grown_levels
