"""GPU (-m gpu): the C++ drop-in executables (peleanalysis_b200/host) end to end through plotfiles: same keys, same
output names, data bit-identical to the compiled reference's golden vectors; when oracle/_ref travelled to the box,
also compared live with AMReX's own fcompare against the reference executable's output."""
import os
import subprocess

import numpy as np
import pytest

from helpers import bit_equal, load_golden, max_rel
from oracle import oracle as O
from peleanalysis_b200 import plotfile

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "peleanalysis_b200", "host")


@pytest.fixture(scope="module")
def exes(gpu):
    # one make at a time: with pytest-xdist a second worker's relink would otherwise hit "Text file busy" on an
    # executable the first worker is already running
    import fcntl
    with open(os.path.join(HOST, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "grad3d.b200.ex"), os.path.join(HOST, "curvature3d.b200.ex")


def _flat(pf, name):
    c = pf.comp(name)
    return np.concatenate([f[c].ravel() for l in pf.levels for f in l.fabs])


def _run(exe, *args, cwd):
    p = subprocess.run([exe, *args], capture_output=True, text=True, cwd=cwd)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


@pytest.mark.parametrize("name", ["c1_periodic", "c1_corner_sym", "c3_three_levels", "ratio4"])
def test_grad_executable(exes, tmp_path, name):
    pf, z = load_golden(name)
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    per = " ".join(str(int(v)) for v in z["is_per"])
    sym = " ".join(str(int(v)) for v in z["sym_dir"])
    out = _run(exes[0], "infile=" + d, "gradVar=temp", "is_per=" + per, "sym_dir=" + sym, cwd=str(tmp_path))
    assert "Periodicity assumed for this case: " + per in out
    r = plotfile.read_plotfile(str(tmp_path / "plt_gt"))          # default outfile = <root>_gt in the cwd
    assert r.names == ["temp", "temp_gx", "temp_gy", "temp_gz", "||gradtemp||"]
    assert bit_equal(_flat(r, "temp"), z["in_temp"])
    for k, n in zip(["gx", "gy", "gz", "mag"], r.names[1:]):
        assert bit_equal(_flat(r, n), z["grad_" + k]), (name, n)
    if O.have_ref():
        O.run_ref("grad", d, str(tmp_path / "ref_gt"), gradVar="temp", is_per=list(z["is_per"]), sym_dir=list(z["sym_dir"]))
        p = subprocess.run([O.ref_exe("fcompare.ref.ex"), str(tmp_path / "plt_gt"), str(tmp_path / "ref_gt")], capture_output=True, text=True)
        assert "PLOTFILE AGREE" in p.stdout, p.stdout[-1500:]


def test_grad_executable_aux_and_inputs_file(exes, tmp_path):
    pf, z = load_golden("c1_options")
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    inp = tmp_path / "inputs.grad"
    inp.write_text("infile = %s\noutfile = %s   # comment\ngradVar = temp\nfinestLevel = 0\nis_per = 1 1 0\nAux_Variables = Y_CH4 x_velocity\n" % (d, tmp_path / "o"))
    _run(exes[0], str(inp), cwd=str(tmp_path))
    r = plotfile.read_plotfile(str(tmp_path / "o"))
    assert r.names == ["temp", "Y_CH4", "x_velocity", "temp_gx", "temp_gy", "temp_gz", "||gradtemp||"]
    assert len(r.levels) == 1
    n0 = pf.levels[0].ncells
    assert bit_equal(_flat(r, "Y_CH4"), z["in_Y_CH4"][:n0])
    # unknown aux variable aborts like the reference
    p = subprocess.run([exes[0], "infile=" + d, "Aux_Variables=nope"], capture_output=True, text=True, cwd=str(tmp_path))
    assert p.returncode != 0 and "Unknown auxiliary variable name: nope" in p.stderr


@pytest.mark.parametrize("name", ["c1_periodic", "c3_threshold", "c1_options"])
def test_curvature_executable(exes, tmp_path, name):
    pf, z = load_golden(name)
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    kw = [str(s) for s in z["curv_opts"]]
    per = " ".join(str(int(v)) for v in z["is_per"])
    _run(exes[1], "infile=" + d, "progressName=temp", "is_per=" + per, "outfile=" + str(tmp_path / "K"), *kw, cwd=str(tmp_path))
    r = plotfile.read_plotfile(str(tmp_path / "K"))
    for key in z.files:
        if not key.startswith("curv_") or key == "curv_opts":
            continue
        n = key[5:]
        got = _flat(r, n)
        if n.startswith("GaussianCurvature"):
            assert max_rel(got, z[key]) <= 1e-12
        else:
            assert bit_equal(got, z[key]), (name, n)
    assert "SmoothedProgress" in r.names and "GaussianCurvature_temp" in r.names


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs in one box (gpurun --gpus 2)")
@pytest.mark.parametrize("name", ["c3_three_levels", "mixed_boxes"])
def test_executables_on_two_gpus(exes, tmp_path, name):
    """ngpus=2: one host thread per GPU in one process (host/multi_gpu.hpp) -- peer links over NVLink by pointer, slab copies
    between the two GPUs, one Cell_D file per GPU; same bits as the golden vectors of the compiled reference."""
    pf, z = load_golden(name)
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    per = " ".join(str(int(v)) for v in z["is_per"])
    sym = " ".join(str(int(v)) for v in z["sym_dir"])
    _run(exes[0], "infile=" + d, "gradVar=temp", "is_per=" + per, "sym_dir=" + sym, "ngpus=2", cwd=str(tmp_path))
    r = plotfile.read_plotfile(str(tmp_path / "plt_gt"))
    for k, n in zip(["gx", "gy", "gz", "mag"], r.names[1:]):
        assert bit_equal(_flat(r, n), z["grad_" + k]), (name, n)
    _run(exes[1], "infile=" + d, "progressName=temp", "is_per=" + per, "sym_dir=" + sym, "ngpus=2", "outfile=" + str(tmp_path / "K"), cwd=str(tmp_path))
    r = plotfile.read_plotfile(str(tmp_path / "K"))
    for n in ["Progress", "MeanCurvature_temp", "FlameNormalX_temp", "FlameNormalY_temp", "FlameNormalZ_temp"]:
        assert bit_equal(_flat(r, n), z["curv_" + n]), (name, n)
