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
    for name, main in (("grad3d.emu.ex", "grad_main.cpp"), ("curvature3d.emu.ex", "curvature_main.cpp")):
        exe = os.path.join(out, name)
        srcs = [os.path.join(host, main), os.path.join(host, "plotfile.cpp")]
        deps = srcs + [lib] + [os.path.join(host, f) for f in ("plotfile.hpp", "tool_common.hpp", "parmparse.hpp")]
        if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", *srcs, "-o", exe, "-L", out, "-lpelestencil_emu", "-Wl,-rpath," + out])
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


def test_emulated_grad_executable_aux_and_inputs_file(emu_exes, tmp_path):
    T.test_grad_executable_aux_and_inputs_file(emu_exes, tmp_path)


@pytest.mark.parametrize("name", ["c1_periodic", "c3_threshold", "c1_options", "mixed_boxes"])
def test_emulated_curvature_executable(emu_exes, tmp_path, name):
    T.test_curvature_executable(emu_exes, tmp_path, name)
