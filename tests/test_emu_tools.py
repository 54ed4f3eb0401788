"""CPU tier: the C++ host shells (peleanalysis_b200/host: ParmParse, plotfile reader / writer, sequencing of the C ABI)
linked against the EMULATED library of tests/emu instead of libpelestencil_b200.so, run end to end through plotfiles with
the same checks tests/test_gpu_tools.py makes on a B200.  Test infrastructure only: the executables built here live under
tests/emu/_build and are not the product's grad3d.b200.ex / curvature3d.b200.ex."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))

import test_gpu_tools as T  # noqa: E402  (test bodies reused as plain functions)


@pytest.fixture(scope="module")
def emu_exes():
    import build_emu
    lib = build_emu.build()
    out = os.path.dirname(lib)
    host = os.path.join(ROOT, "peleanalysis_b200", "host")
    exes = []
    import fcntl
    with open(os.path.join(out, ".exe_build.lock"), "w") as lk:      # one build at a time (pytest-xdist workers share the directory)
        fcntl.flock(lk, fcntl.LOCK_EX)
        for name, main in (("grad3d.emu.ex", "grad_main.cpp"), ("curvature3d.emu.ex", "curvature_main.cpp"), ("filterPlt3d.emu.ex", "filter_main.cpp")):
            exe = os.path.join(out, name)
            srcs = [os.path.join(host, main), os.path.join(host, "plotfile.cpp")]
            deps = srcs + [lib] + [os.path.join(host, f) for f in ("plotfile.hpp", "tool_common.hpp", "parmparse.hpp", "multi_gpu.hpp")]
            if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
                subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", *srcs, "-o", exe + ".tmp", "-L", out, "-lpelestencil_emu", "-Wl,-rpath," + out])
                os.replace(exe + ".tmp", exe)
            exes.append(exe)
    old = os.environ.get("PA_NORMAL_MATH")
    os.environ["PA_NORMAL_MATH"] = "fast"        # the emulator has no MUFU; see tests/test_emu_parity.py
    yield tuple(exes)
    if old is None:
        os.environ.pop("PA_NORMAL_MATH", None)
    else:
        os.environ["PA_NORMAL_MATH"] = old


@pytest.mark.parametrize("name", ["c1_periodic", "c1_corner_sym", "c3_three_levels", "ratio4", "mixed_boxes"])
def test_emulated_grad_executable(emu_exes, tmp_path, name):
    T.test_grad_executable(emu_exes, tmp_path, name)


def test_emulated_grad_executable_several_files_per_level(emu_exes, tmp_path, monkeypatch):
    T.test_grad_executable_several_files_per_level(emu_exes, tmp_path, monkeypatch)


def test_emulated_grad_executable_aux_and_inputs_file(emu_exes, tmp_path):
    T.test_grad_executable_aux_and_inputs_file(emu_exes, tmp_path)


@pytest.mark.parametrize("name", ["c1_periodic", "c3_threshold", "c1_options", "mixed_boxes"])
def test_emulated_curvature_executable(emu_exes, tmp_path, name):
    T.test_curvature_executable(emu_exes, tmp_path, name)


@pytest.mark.parametrize("name", ["filter_c1", "filter_c1_corner_gauss", "filter_c3", "filter_c3_subset", "filter_c3_pc_samefgr", "filter_ratio4"])
def test_emulated_filter_executable(emu_exes, tmp_path, name):
    T.test_filter_executable(emu_exes, tmp_path, name)


def test_emulated_filter_executable_errors(emu_exes, tmp_path):
    T.test_filter_executable_errors(emu_exes, tmp_path)


@pytest.mark.parametrize("name,ngpus", [("c1_periodic", 2), ("c3_three_levels", 2), ("mixed_boxes", 3), ("lshape", 4), ("c1_corner_sym", 2)])
def test_emulated_grad_executable_multi_gpu(emu_exes, tmp_path, name, ngpus):
    """grad3d ... ngpus=N: one host thread per (emulated) GPU in one process -- same-process peer links, slab copies between
    the threads' slabs, one Cell_D file per thread; the output equals the reference's golden vectors and AMReX's fcompare
    agrees with the reference executable's plotfile."""
    import numpy as np
    from helpers import bit_equal, load_golden
    from oracle import oracle as O
    from peleanalysis_b200 import plotfile
    pf, z = load_golden(name)
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    per = " ".join(str(int(v)) for v in z["is_per"])
    sym = " ".join(str(int(v)) for v in z["sym_dir"])
    env = dict(os.environ, CUEMU_DEVICES=str(ngpus), CUEMU_SEED="9")
    p = subprocess.run([emu_exes[0], "infile=" + d, "gradVar=temp", "is_per=" + per, "sym_dir=" + sym, "ngpus=%d" % ngpus],
                       capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    r = plotfile.read_plotfile(str(tmp_path / "plt_gt"))
    assert r.names == ["temp", "temp_gx", "temp_gy", "temp_gz", "||gradtemp||"]

    def flat(n):
        c = r.comp(n)
        return np.concatenate([f[c].ravel() for l in r.levels for f in l.fabs])
    assert bit_equal(flat("temp"), z["in_temp"])
    for k, n in zip(["gx", "gy", "gz", "mag"], r.names[1:]):
        assert bit_equal(flat(n), z["grad_" + k]), (name, n)
    files = {f for l in range(len(pf.levels)) for f in os.listdir(str(tmp_path / "plt_gt" / ("Level_%d" % l)))}
    assert len([f for f in files if f.startswith("Cell_D_")]) >= 2
    if O.have_ref():
        O.run_ref("grad", d, str(tmp_path / "ref_gt"), gradVar="temp", is_per=list(z["is_per"]), sym_dir=list(z["sym_dir"]))
        q = subprocess.run([O.ref_exe("fcompare.ref.ex"), str(tmp_path / "plt_gt"), str(tmp_path / "ref_gt")], capture_output=True, text=True)
        assert "PLOTFILE AGREE" in q.stdout, q.stdout[-1500:]


@pytest.mark.parametrize("name,ngpus", [("c1_periodic", 2), ("c3_threshold", 2), ("c1_options", 3), ("mixed_boxes", 2)])
def test_emulated_curvature_executable_multi_gpu(emu_exes, tmp_path, name, ngpus):
    """curvature3d ... ngpus=N (one host thread per emulated GPU): every option, incl. the per-level exchange of the flame
    normal under threshold_prog and the exchange of the internal gradient field for the Gaussian curvature."""
    env_old = {k: os.environ.get(k) for k in ("CUEMU_DEVICES", "CUEMU_SEED")}
    os.environ.update(CUEMU_DEVICES=str(ngpus), CUEMU_SEED="4")
    try:
        import numpy as np
        from helpers import bit_equal, load_golden, max_rel
        from peleanalysis_b200 import plotfile
        pf, z = load_golden(name)
        d = str(tmp_path / "plt")
        plotfile.write_plotfile(d, pf)
        kw = [str(s) for s in z["curv_opts"]]
        per = " ".join(str(int(v)) for v in z["is_per"])
        p = subprocess.run([emu_exes[1], "infile=" + d, "progressName=temp", "is_per=" + per, "outfile=" + str(tmp_path / "K"), "ngpus=%d" % ngpus, *kw],
                           capture_output=True, text=True, cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout + p.stderr
        r = plotfile.read_plotfile(str(tmp_path / "K"))
        for key in z.files:
            if not key.startswith("curv_") or key == "curv_opts":
                continue
            n = key[5:]
            c = r.comp(n)
            got = np.concatenate([f[c].ravel() for l in r.levels for f in l.fabs])
            if n.startswith("GaussianCurvature"):
                assert max_rel(got, z[key]) <= 1e-12
            else:
                assert bit_equal(got, z[key]), (name, n)
        assert "SmoothedProgress" in r.names and "GaussianCurvature_temp" in r.names
    finally:
        for k, v in env_old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ---- source-level drop-in: the reference tools with their operator block replaced by the C-ABI calls,
#      here linked against host AMReX (oracle/_ref/libamrex_ref.a) and the EMULATED library -------------------------------
@pytest.fixture(scope="module")
def emu_amrex_exes(emu_exes):
    from oracle import oracle as O
    ref = os.path.dirname(O.ref_exe("x"))
    srcdir, lib_a, inc = os.path.join(ref, "src"), os.path.join(ref, "libamrex_ref.a"), os.path.join(ref, "include")
    amrex = "/root/reference/Submodules/PelePhysics/Submodules/amrex"
    srcs = [os.path.join(srcdir, "grad_b200amrex.cpp"), os.path.join(srcdir, "curvature_b200amrex.cpp")]
    if not (os.path.isdir(amrex) and os.path.exists(lib_a) and all(os.path.exists(s) for s in srcs)):
        pytest.skip("needs /root/reference and oracle/_ref (python oracle/build_ref.py)")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import build_emu
    out = os.path.dirname(build_emu.build())
    glue = os.path.join(ROOT, "peleanalysis_b200", "host", "amrex_glue")
    flags = build_ref.cxx_flags(inc) + ["-O1", "-I" + glue, "-I" + os.path.join(ROOT, "include")]
    exes = []
    import fcntl
    with open(os.path.join(out, ".exe_build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        for src, name in zip(srcs, ("grad3d.emuamrex.ex", "curvature3d.emuamrex.ex")):
            exe = os.path.join(out, name)
            deps = [src, lib_a, os.path.join(out, "libpelestencil_emu.so")] + [os.path.join(glue, f) for f in os.listdir(glue)]
            if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
                subprocess.check_call(["g++", *flags, src, "-o", exe + ".tmp", lib_a, "-lgomp", "-lpthread", "-L", out, "-lpelestencil_emu", "-Wl,-rpath," + out])
                os.replace(exe + ".tmp", exe)
            exes.append(exe)
    return tuple(exes)


@pytest.mark.parametrize("name", ["c1_periodic", "c3_three_levels"])
def test_emulated_amrex_linked_grad(emu_amrex_exes, tmp_path, name):
    T.test_grad_executable(emu_amrex_exes, tmp_path, name)


@pytest.mark.parametrize("name", ["c1_periodic", "c1_options"])
def test_emulated_amrex_linked_curvature(emu_amrex_exes, tmp_path, name):
    T.check_curvature_executable(emu_amrex_exes, tmp_path, name)


def test_emulated_amrex_linked_curvature_do_smooth(emu_amrex_exes, tmp_path):
    T.test_amrex_linked_curvature_do_smooth(emu_amrex_exes, tmp_path)
