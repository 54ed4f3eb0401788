"""filterPlt path on the GPU through the C ABI against the golden vectors of the compiled, unmodified reference tool and the
NumPy restatement: bit for bit (ghost cells included)."""
import numpy as np
import pytest

from cases import FILTER_CASES, FILTER_TYPE_SWEEP
from helpers import bit_equal, load_golden, max_rel
from peleanalysis_b200 import filterplt, synth
from test_filter_oracle import out_boxes, parse_opts

pytestmark = pytest.mark.gpu


def check_filter_case(P, name):
    pf, z = load_golden(name)
    names, levels, run = filterplt.filter_plotfile(P, pf, **parse_opts(z))
    assert names == [str(n) for n in z["out_names"]]
    assert [b for b, _ in levels] == out_boxes(z)
    for ci, n in enumerate(names):
        got = np.concatenate([f[ci].ravel() for _, fabs in levels for f in fabs])
        assert bit_equal(got, z["out_" + n]), (name, n, max_rel(got, z["out_" + n]))


@pytest.mark.parametrize("name", sorted(FILTER_CASES))
def test_filter_matches_reference_golden(gpu, name):
    check_filter_case(gpu, name)


def check_filter_types(P):
    pf, z = load_golden(FILTER_TYPE_SWEEP[0])
    for t, f in z["combos"]:
        _, levels, _ = filterplt.filter_plotfile(P, pf, filter_type=int(t), base_fgr=int(f), same_fgr_all_levels=True)
        got = np.concatenate([fab[0].ravel() for _, fabs in levels for fab in fabs])
        assert bit_equal(got, z["out_t%d_f%d" % (t, f)]), (int(t), int(f))


def test_filter_types_match_reference_golden(gpu):
    check_filter_types(gpu)


def check_ghost_cells(P, name):
    """every ghost cell the filter reads, cell by cell, against the restatement's grown FABs"""
    from oracle import filter_oracle as FO
    pf, z = load_golden(name)
    kw = parse_opts(z)
    _, _, grown = FO.filter_plotfile(pf, **kw)
    run = filterplt.FilterRun(P, pf, **kw)
    run.step()
    P.sync()
    G = max(run.ngrow)
    for l in range(run.nlev):
        g = run.ngrow[l]
        for b in range(len(run.levels[l].boxes)):
            for c in range(run.ncomp):
                q = run.grown_input(l, b, c)
                q = q[G - g:q.shape[0] - (G - g), G - g:q.shape[1] - (G - g), G - g:q.shape[2] - (G - g)]
                assert bit_equal(q, grown[l][b][c]), (name, l, b, c)


@pytest.mark.parametrize("name", ["filter_c1_corner_gauss", "filter_c3", "filter_lshape", "filter_ratio4"])
def test_filter_ghost_cells_match_oracle(gpu, name):
    check_ghost_cells(gpu, name)


MIDSIZE = [  # (filter_type, base_fgr, max_grid_size): ghost widths 1/2/4, 2/4/8, 3/6/12 (generic kernel), ragged 4 x 2 blocks
    (2, 2, 16), (2, 4, 16), (1, 6, 16), (1, 2, 10), (2, 4, 11), (4, 3, 7)]


def check_midsize(P, ftype, fgr, mgs):
    """hierarchies the fixtures do not hold (32^3 base, 3 levels, two variables) against the restatement"""
    from oracle import filter_oracle as FO
    pf = synth.config3(32, 16, names=("temp", "Y_CH4"))
    kw = dict(filter_type=ftype, base_fgr=fgr, max_grid_size=mgs)
    names, ref, _ = FO.filter_plotfile(pf, **kw)
    names2, got, _ = filterplt.filter_plotfile(P, pf, **kw)
    assert names == names2
    for (rb, rf), (gb, gf) in zip(ref, got):
        assert rb == gb
        for a, b in zip(rf, gf):
            assert bit_equal(b, a), (ftype, fgr, mgs)


@pytest.mark.parametrize("ftype,fgr,mgs", MIDSIZE)
def test_filter_midsize_matches_oracle(gpu, ftype, fgr, mgs):
    check_midsize(gpu, ftype, fgr, mgs)


def test_filter_large_matches_oracle(gpu):
    """128^3 base, 3 levels, 32^3 output boxes (the bench workload's box size and ghost widths 1/2/4; 6.3 M cells) against the
    restatement, bit for bit"""
    from oracle import filter_oracle as FO
    pf = synth.config3(128, 64)
    names, ref, _ = FO.filter_plotfile(pf, max_grid_size=32)
    _, got, _ = filterplt.filter_plotfile(gpu, pf, max_grid_size=32)
    for (rb, rf), (gb, gf) in zip(ref, got):
        assert rb == gb
        for a, b in zip(rf, gf):
            assert bit_equal(b, a)


def test_fill_patch_rejects_periodic_and_bad_nesting(gpu):
    P = gpu
    pf = synth.config1(16, 8)
    H = P.Hierarchy(pf.levels, is_per=(1, 1, 1))
    f = P.Field(H, 1, 2)
    with pytest.raises(P.PaError) as e:
        P.fill_patch(f, 0, 1, 0, 1)
    assert e.value.code == -4
    # fine boxes whose ghost region needs coarse cells that level 1 does not cover
    pf3 = synth.config3(16, 8)
    lv = pf3.levels
    bad = [lv[0], type(lv[1])(lv[1].domain_lo, lv[1].domain_hi, lv[1].dx, lv[1].boxes[:1], lv[1].fabs[:1]), lv[2]]
    H2 = P.Hierarchy(bad, is_per=(0, 0, 0))
    f2 = P.Field(H2, 1, 4)
    with pytest.raises(P.PaError) as e2:
        P.fill_patch(f2, 0, 1, 2, 4)
    assert e2.value.code == -1
