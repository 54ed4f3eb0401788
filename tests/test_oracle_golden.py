"""CPU: the C restatement (oracle/) against the golden vectors produced by the compiled, unmodified reference
(tests/golden/make_golden.py).  Bit-exact everywhere except GaussianCurvature (std::pow(x,4.0): libm on both
sides here, so it is bit-exact too, but the documented tolerance is 1e-12)."""
import numpy as np
import pytest

from cases import CASES
from helpers import bit_equal, load_golden, max_rel
from oracle import oracle as O


def _curv_kwargs(z):
    kw = dict(s.split("=") for s in z["curv_opts"]) if "curv_opts" in z and len(z["curv_opts"]) else {}
    return kw


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if "grad" in c[3]])
def test_oracle_grad_matches_reference(name):
    pf, z = load_golden(name)
    H = O.OracleHier(pf, tuple(z["is_per"]), tuple(z["sym_dir"]))
    g = H.grad(z["in_temp"])
    for c, k in enumerate(["gx", "gy", "gz", "mag"]):
        assert bit_equal(g[c], z["grad_" + k]), (name, k, max_rel(g[c], z["grad_" + k]))


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if "curvature" in c[3]])
def test_oracle_curvature_matches_reference(name):
    pf, z = load_golden(name)
    H = O.OracleHier(pf, tuple(z["is_per"]), tuple(z["sym_dir"]))
    kw = _curv_kwargs(z)
    thr = bool(int(kw.get("threshold_prog", 0)))
    tv = float(kw.get("threshold_value", 1e-4))
    pmin, pmax = float(z["prog_min"]), float(z["prog_max"])
    names = ["Progress", "MeanCurvature_temp", "FlameNormalX_temp", "FlameNormalY_temp", "FlameNormalZ_temp"]
    if int(kw.get("do_strain", 0)):
        U = np.stack([z["in_x_velocity"], z["in_y_velocity"], z["in_z_velocity"]])
        r = H.curvature_ex(z["in_temp"], U, pmin, pmax, thr, tv)
        core = r["core"]
        assert bit_equal(r["strain"], z["curv_StrainRate_temp"])
        assert bit_equal(r["veln"], z["curv_VelFlameNormal"])
        dirs = "xyz"
        for m in range(3):
            for n in range(3):
                assert bit_equal(r["rost"][3 * m + n], z["curv_ROST_dU%sd%s" % (dirs[m], dirs[n])]), (m, n)
        assert max_rel(r["gauss"], z["curv_GaussianCurvature_temp"]) <= 1e-12
    else:
        core = H.curvature(z["in_temp"], pmin, pmax, thr, tv)
    for c, n in enumerate(names):
        assert bit_equal(core[c], z["curv_" + n]), (name, n, max_rel(core[c], z["curv_" + n]))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (python oracle/build_ref.py)")
def test_live_reference_agrees_with_golden(tmp_path):
    """When the compiled reference is present, re-run one case end to end through plotfiles."""
    from peleanalysis_b200 import plotfile
    pf, z = load_golden("c1_periodic")
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    O.run_ref("grad", d, d + "_gt", gradVar="temp")
    r = plotfile.read_plotfile(d + "_gt")
    gx = np.concatenate([f[r.comp("temp_gx")].ravel() for l in r.levels for f in l.fabs])
    assert bit_equal(gx, z["grad_gx"])
