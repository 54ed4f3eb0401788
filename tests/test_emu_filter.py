"""CPU tier of the filterPlt path: the kernel sources of filter.cu under the CUDA execution-model emulator (tests/emu), same
C ABI, same golden vectors of the compiled reference, same checks as tests/test_gpu_filter.py.  Test infrastructure only."""
import pytest

import test_gpu_filter as GF
from cases import FILTER_CASES
from test_emu_parity import emu  # noqa: F401  (module-scoped fixture: the emulated library behind a private binding)


@pytest.mark.parametrize("name", sorted(FILTER_CASES))
def test_emulated_filter_matches_reference_golden(emu, name):  # noqa: F811
    GF.check_filter_case(emu, name)


def test_emulated_filter_types(emu):  # noqa: F811
    GF.check_filter_types(emu)


@pytest.mark.parametrize("name", ["filter_c1_corner_gauss", "filter_c3", "filter_lshape", "filter_ratio4"])
def test_emulated_filter_ghost_cells(emu, name):  # noqa: F811
    GF.check_ghost_cells(emu, name)


@pytest.mark.parametrize("ftype,fgr,mgs", GF.MIDSIZE)
def test_emulated_filter_midsize(emu, ftype, fgr, mgs):  # noqa: F811
    GF.check_midsize(emu, ftype, fgr, mgs)
