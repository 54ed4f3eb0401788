"""The NumPy restatement of the filterPlt path (oracle/filter_oracle.py) against the golden vectors produced by the compiled,
unmodified reference tool (tests/golden/make_golden_filter.py): bit for bit, every case, every filter type."""
import numpy as np
import pytest

from cases import FILTER_CASES, FILTER_TYPE_SWEEP
from helpers import bit_equal, load_golden
from oracle import filter_oracle as FO


def parse_opts(z):
    o = dict(s.split("=", 1) for s in (str(x) for x in z["opts"])) if "opts" in z.files else {}
    return dict(filter_type=int(o.get("filter_type", 1)), base_fgr=int(o.get("base_fgr", 2)),
                same_fgr_all_levels=bool(int(o.get("same_fgr_all_levels", 0))), max_grid_size=int(o.get("max_grid_size", 32)),
                interp_type=int(o.get("interp_type", 1)), variables=o["variables"].split() if "variables" in o else None,
                max_filter_level=int(o.get("max_filter_level", 1000)))


def out_boxes(z):
    res, o = [], 0
    for nb in z["out_nboxes"]:
        res.append([(tuple(int(v) for v in b[:3]), tuple(int(v) for v in b[3:])) for b in z["out_boxes"][o:o + int(nb)]])
        o += int(nb)
    return res


@pytest.mark.parametrize("name", sorted(FILTER_CASES))
def test_filter_oracle_matches_reference(name):
    pf, z = load_golden(name)
    names, levels, _ = FO.filter_plotfile(pf, **parse_opts(z))
    assert names == [str(n) for n in z["out_names"]]
    assert [b for b, _ in levels] == out_boxes(z)                       # BoxArray::maxSize, box for box and in order
    for ci, n in enumerate(names):
        got = np.concatenate([f[ci].ravel() for _, fabs in levels for f in fabs])
        assert bit_equal(got, z["out_" + n]), (name, n)


def test_filter_types_match_reference():
    name = FILTER_TYPE_SWEEP[0]
    pf, z = load_golden(name)
    for t, f in z["combos"]:
        _, levels, _ = FO.filter_plotfile(pf, filter_type=int(t), base_fgr=int(f), same_fgr_all_levels=True)
        got = np.concatenate([fab[0].ravel() for _, fabs in levels for fab in fabs])
        assert bit_equal(got, z["out_t%d_f%d" % (t, f)]), (int(t), int(f))


def test_weights_sum_and_symmetry():
    for t in range(0, 11):
        for f in (1, 2, 3, 4, 6, 8, 10, 12):
            if t in (1, 2) and f % 2 and f != 1:
                continue
            if t == 2 and f == 1:
                continue
            ng, w = FO.filter_weights(t, f)
            assert len(w) == 2 * ng + 1
            assert w == w[::-1]
            assert abs(sum(w) - 1.0) < 1e-12, (t, f, w)


def test_max_size_chops_like_boxlist():
    # 24 cells at chunk 8 -> 3 blocks; 20 at 8 -> (20 = 4*5, 8 = 4*2): 3 blocks of 2,2,1 coarse cells = 8, 8, 4
    assert FO.max_size([((0, 0, 0), (23, 7, 7))], 8) == [((0, 0, 0), (7, 7, 7)), ((8, 0, 0), (15, 7, 7)), ((16, 0, 0), (23, 7, 7))]
    assert [b[1][0] - b[0][0] + 1 for b in FO.max_size([((4, 0, 0), (23, 3, 3))], 8)] == [8, 8, 4]
    assert [b[1][0] - b[0][0] + 1 for b in FO.max_size([((0, 0, 0), (10, 3, 3))], 4)] == [4, 4, 3]
