"""CPU: the product's host-built descriptor tables (hier.cpp, reached through the C ABI's debug entry points)
against the oracle's AMReX-style objects, cell by cell -- integer parity, bit-exact."""
import numpy as np
import pytest

from cases import CASES
from oracle import oracle as O
from peleanalysis_b200 import capi


def _hiers(name):
    builder, is_per, sym, _, _ = CASES[name]
    pf = builder()
    return pf, O.OracleHier(pf, is_per, sym), capi.Hierarchy(pf.levels, is_per, sym)


@pytest.mark.parametrize("name", list(CASES))
def test_fill_boundary_descriptors_full(palib, name):
    """Expanded halo tag table == FillBoundary's copy rule for every ghost cell, ng = 1 and 2."""
    pf, OH, PH = _hiers(name)
    for ng in (1, 2):
        for l in range(len(pf.levels)):
            want = OH.fb_source_map(l, ng)
            got = PH.fb_source_map(l, ng, cross=False)
            for b, (w, g) in enumerate(zip(want, got)):
                assert np.array_equal(w, g), (name, ng, l, b, int((w != g).sum()))


@pytest.mark.parametrize("name", list(CASES))
def test_fill_boundary_descriptors_cross(palib, name):
    """The width-1 'cross' table the hot path uses equals the full table restricted to the six face slabs."""
    pf, OH, PH = _hiers(name)
    for l in range(len(pf.levels)):
        want = OH.fb_source_map(l, 1)
        got = PH.fb_source_map(l, 1, cross=True)
        for w, g in zip(want, got):
            face = np.zeros(w.shape, dtype=bool)
            face[0, 1:-1, 1:-1] = face[-1, 1:-1, 1:-1] = True
            face[1:-1, 0, 1:-1] = face[1:-1, -1, 1:-1] = True
            face[1:-1, 1:-1, 0] = face[1:-1, 1:-1, -1] = True
            assert np.array_equal(w[face], g[face])
            assert (g[~face] == -1).all()


@pytest.mark.parametrize("name", list(CASES))
def test_face_masks_and_coefficients(palib, name):
    """Face flags == MultiMask values (both mask sets), records exist exactly where a mask value is > 0,
    and the polynomial coefficients equal poly_interp_coeff's."""
    pf, OH, PH = _hiers(name)
    builder, is_per, sym, _, _ = CASES[name]
    for l, lv in enumerate(pf.levels):
        r = OH.ratios[l - 1] if l > 0 else 1
        for b, (lo, hi) in enumerate(lv.boxes):
            for face in range(6):
                d = face % 3
                _, m0 = OH.mask(l, b, face, 0)            # m_maskvals: out 1, extent 0
                m0 = np.squeeze(m0, axis=2 - d)
                fl = PH.face_flags(l, b, face)
                if fl is None:
                    assert (m0 == 0).all(), (name, l, b, face)
                    continue
                assert (m0 > 0).any()
                assert np.array_equal(fl & 3, m0), (name, l, b, face)
                kind, nx, coef = PH.face_coef(l, b, face)
                at_wall = (lo[d] == lv.domain_lo[d]) if face < 3 else (hi[d] == lv.domain_hi[d])
                if at_wall and not is_per[d]:
                    assert kind == (1 if sym[d] else 0)
                    continue
                assert kind == 2 and l > 0
                blen = hi[d] - lo[d] + 1
                assert nx == min(blen + 1, 4)
                if blen >= 3 and abs(lv.dx[d] * (1.0 / lv.dx[d]) - 1.0) == 0.0:
                    want = {2: [0.45714285714285713, 1.0, -0.6, 0.14285714285714285],
                            4: [0.1523809523809524, 1.8, -1.2857142857142858, 0.3333333333333333]}[r]
                    assert coef == want, (coef, want)
                # BndryData mask (out 2, extent 5), layer adjacent to the box: tangential neighbours at +-r
                pb, m1 = OH.mask(l, b, face, 1)
                layer = 1 if face < 3 else 0               # out_rad 2: low faces store [lo-2, lo-1], high faces [hi+1, hi+2]
                m1 = np.take(m1, layer, axis=2 - d)        # -> [t2, t1] grown by 5
                n2, n1 = fl.shape
                offs = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 1), (1, 1)]
                for bit, (o1, o2) in enumerate(offs):
                    want = (m1[5 + o2 * r: 5 + o2 * r + n2, 5 + o1 * r: 5 + o1 * r + n1] == 1)
                    got = ((fl >> (2 + bit)) & 1).astype(bool)
                    assert np.array_equal(want, got), (name, l, b, face, bit)


@pytest.mark.parametrize("name", list(CASES))
def test_neighbour_links_agree_with_fill_boundary(palib, name):
    """A linked face must name exactly the cells FillBoundary would copy into that face's ghost layer (oracle's
    expanded copy rule), and every face that is one same-shaped neighbour's territory must be linked."""
    pf, OH, PH = _hiers(name)
    nlinked = 0
    for l, lv in enumerate(pf.levels):
        want = OH.fb_source_map(l, 1)
        for b, (lo, hi) in enumerate(lv.boxes):
            n = [hi[d] - lo[d] + 1 for d in range(3)]
            L = PH.links(l, b)
            src = want[b]                                    # [nz+2, ny+2, nx+2], (box<<40 | linear idx) or -1
            for face in range(6):
                d = face % 3
                sl = [slice(1, -1)] * 3                      # numpy axes are (z, y, x)
                sl[2 - d] = 0 if face < 3 else -1
                layer = src[tuple(sl)]
                # ghost cell coordinates (box-relative) of the layer
                idx = [np.arange(n[2]), np.arange(n[1]), np.arange(n[0])]
                idx[2 - d] = np.array([-1 if face < 3 else n[d]])
                K, J, I = np.meshgrid(*idx, indexing="ij")
                K, J, I = [np.squeeze(a, axis=2 - d) for a in (K, J, I)]
                boxes = np.where(layer >= 0, layer >> 40, -1)
                single = (boxes >= 0).all() and (boxes == boxes.flat[0]).all()
                nb = int(L[face][0])
                if nb >= 0:
                    nlinked += 1
                    nlo, nhi = lv.boxes[nb]
                    nn = [nhi[t] - nlo[t] + 1 for t in range(3)]
                    assert nn[0] == n[0] and nn[1] == n[1]
                    rel = L[face][2:5]
                    lin = ((K + rel[2]) * nn[1] + (J + rel[1])) * nn[0] + (I + rel[0])
                    assert single and boxes.flat[0] == nb, (name, l, b, face)
                    assert np.array_equal(layer & ((1 << 40) - 1), lin), (name, l, b, face)
                elif single:
                    nlo, nhi = lv.boxes[int(boxes.flat[0])]
                    same = (nhi[0] - nlo[0] == hi[0] - lo[0]) and (nhi[1] - nlo[1] == hi[1] - lo[1])
                    assert not same or (d == 0 and n[0] < 3), (name, l, b, face, "face should have been linked")
    assert nlinked > 0
    # PA_HIER_NO_LINKS turns them all off
    builder, is_per, sym, _, _ = CASES[name]
    H0 = capi.Hierarchy(pf.levels, is_per, sym, flags=capi.NO_LINKS)
    assert all((H0.links(l, b)[:, 0] == -1).all() for l, lv in enumerate(pf.levels) for b in range(len(lv.boxes)))


def test_hierarchy_validation(palib):
    from peleanalysis_b200 import synth
    pf = synth.config1(16, 8)
    # fine box not aligned to the ratio
    lo, hi = pf.levels[1].boxes[0]
    pf.levels[1].boxes[0] = ((lo[0] + 1, lo[1], lo[2]), hi)
    with pytest.raises(capi.PaError):
        capi.Hierarchy(pf.levels)
    # level 0 not covering the domain
    pf = synth.config1(16, 8)
    pf.levels[0].boxes.pop()
    with pytest.raises(capi.PaError):
        capi.Hierarchy(pf.levels)


def test_sfc_distribute_balanced(palib):
    from peleanalysis_b200 import synth
    pf = synth.make_hierarchy(64, [], [], 16, fill=False)
    for n in (2, 4, 8):
        ow = capi.sfc_distribute(pf.levels[0].boxes, n)
        counts = np.bincount(ow, minlength=n)
        assert counts.min() == counts.max() == len(pf.levels[0].boxes) // n
