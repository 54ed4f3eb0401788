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
