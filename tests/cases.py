"""Named small test cases shared by the golden-fixture generator and the parity tests.

Each case: a synthetic hierarchy + the tool options (is_per, sym_dir) + which tools to run.  Sizes are chosen so
the reference runs in well under a second and every branch of SURVEY 3.3 is hit somewhere in the set."""
from peleanalysis_b200 import synth


def _np2_small():
    # non-cubic, non power-of-two dx (24 x 12 x 20 cells on 0.7 x 0.35 x 1.3)
    return synth.make_hierarchy((24, 12, 20), [[((12, 6, 10), (35, 17, 29))]], [2], 12, ("temp",), prob_hi=(0.7, 0.35, 1.3))


CASES = {
    # name: (builder, is_per, sym_dir, tools, curvature kwargs)
    "c1_periodic": (lambda: synth.config1(16, 8), (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "c1_walls": (lambda: synth.config1(16, 8), (0, 0, 0), (0, 0, 0), ("grad", "curvature"), {}),
    "c1_corner_sym": (lambda: synth.config1(16, 8, corner=True), (0, 1, 0), (1, 0, 0), ("grad", "curvature"), {}),
    "lshape": (lambda: synth.case_lshape(16, 8), (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "edge_walls": (lambda: synth.case_edge(16, 8), (0, 0, 0), (0, 0, 0), ("grad", "curvature"), {}),
    "edge_periodic": (lambda: synth.case_edge(16, 8), (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "np2": (_np2_small, (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "thin": (synth.case_thin, (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "mixed_boxes": (synth.case_mixed, (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "ratio4": (lambda: synth.case_ratio4(8, 16), (1, 1, 1), (0, 0, 0), ("grad",), {}),
    "c3_three_levels": (lambda: synth.config3(16, 8), (1, 1, 1), (0, 0, 0), ("grad", "curvature"), {}),
    "c3_threshold": (lambda: synth.config3(16, 8), (1, 1, 1), (0, 0, 0), ("curvature",),
                     dict(threshold_prog=1, threshold_value=0.01)),
    "c1_options": (lambda: synth.config1(16, 8, names=synth.FIELD_NAMES), (1, 1, 0), (0, 0, 0), ("curvature",),
                   dict(do_gaussCurv=1, do_strain=1, getStrainTensor=1, do_velnormal=1, threshold_prog=1, threshold_value=0.01)),
}

GRAD_OUT = ["gx", "gy", "gz", "mag"]
CURV_OUT = ["Progress", "MeanCurvature", "FlameNormalX", "FlameNormalY", "FlameNormalZ"]

# filterPlt cases: (builder, tool options).  The tool's geometry is always non-periodic; ghost widths follow from
# base_fgr x refinement ratios (1, 2, 4 cells on three ratio-2 levels with the default base_fgr = 2).
FILTER_CASES = {
    "filter_c1": (lambda: synth.config1(16, 8), {}),
    "filter_c1_corner_gauss": (lambda: synth.config1(16, 8, corner=True), dict(filter_type=2, base_fgr=4, max_grid_size=8)),
    "filter_c3": (lambda: synth.config3(16, 8, names=("temp", "x_velocity")), dict(max_grid_size=8)),
    "filter_c3_pc_samefgr": (lambda: synth.config3(16, 8), dict(interp_type=0, same_fgr_all_levels=1, base_fgr=4, filter_type=2)),
    "filter_c3_subset": (lambda: synth.config3(16, 8, names=("temp", "x_velocity", "Y_CH4")), dict(variables="Y_CH4 temp", max_filter_level=1, max_grid_size=12)),
    "filter_lshape": (lambda: synth.case_lshape(16, 8), dict(max_grid_size=8)),
    "filter_ratio4": (lambda: synth.case_ratio4(8, 16), {}),
    "filter_np2": (_np2_small, dict(filter_type=4, base_fgr=3)),
}
# every filter type at an even and an odd filter-to-grid ratio on one small file (level 0 and 1, same ratio on both)
FILTER_TYPE_SWEEP = ("filter_types", lambda: synth.config1(16, 8), [(t, f) for t in range(0, 11) for f in (3, 4) if not (t in (1, 2) and f % 2)] + [(6, 12), (9, 11)])
