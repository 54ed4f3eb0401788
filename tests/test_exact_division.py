"""CPU: the shared-reciprocal division of the MODE_NORMAL epilogue (stencil_tma.cu: div_by) equals IEEE a/b bit for bit
on random and adversarial operands -- checked with a small C program (hardware FMA) so that millions of cases run in a
second.  The GPU parity tests then pin the CUDA code itself against the reference's output."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not _has_fma(), reason="host CPU has no FMA instruction")
def test_markstein_division_is_exact(tmp_path):
    exe = str(tmp_path / "divcheck")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-o", exe, os.path.join(HERE, "divcheck", "markstein_div_check.c"), "-lm"])
    for seed in (1, 2, 3):
        out = subprocess.run([exe, str(seed), "30000000"], capture_output=True, text=True, check=True).stdout
        assert "bad=0" in out, out
