"""CPU: the C-ABI shared library loads and exports every symbol include/pele_stencil_b200.h declares; compute entry
points fail loudly (PA_ERR_CUDA) when no device is usable -- there is no CPU fallback to fall into."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pele_stencil_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(palib):
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(palib, n), n
        getattr(palib, n)


def test_python_binding_covers_header():
    from peleanalysis_b200 import capi
    assert sorted(capi.SYMBOLS) == _declared()


def test_product_does_not_import_oracle():
    """The product path may not import, link, dlopen or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "peleanalysis_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b|#include\s*[\"<][^\">]*oracle|libpa_oracle|oracle/_ref|\.ref\.ex)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), (dirpath, f, bad.search(txt).group(0))


def test_product_cannot_reach_the_emulator(palib):
    """tests/emu (the CPU emulation of the CUDA execution model used by tests/test_emu_parity.py) is test infrastructure:
    the product library is built by nvcc without PA_HOST_EMULATION and exports no emulator symbol, no product file names
    the emulated library or its directory, and the binding opens exactly lib/libpelestencil_b200.so."""
    import subprocess
    from peleanalysis_b200 import build, capi
    assert capi.LIB_PATH == build.LIB and capi.LIB_PATH.endswith(os.path.join("peleanalysis_b200", "lib", "libpelestencil_b200.so"))
    assert "PA_HOST_EMULATION" not in " ".join(build.FLAGS)
    syms = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True).stdout
    assert "cuemu" not in syms
    bad = re.compile(r"(libpelestencil_emu|tests/emu|build_emu|cuemu\.cpp)")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "peleanalysis_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                # the two kernel sources say, in comments, which test build defines PA_HOST_EMULATION
                txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("//") and "// tests/emu" not in l)
                assert not bad.search(txt), (dirpath, f, bad.search(txt).group(0))
    for f in ("bench.py", "__graft_entry__.py"):
        assert not bad.search(open(os.path.join(ROOT, f)).read()), f


def test_no_device_fails_loudly(palib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from peleanalysis_b200 import capi, synth
    with pytest.raises(capi.PaError) as e:
        capi.init(0)
    assert e.value.code == -2
    H = capi.Hierarchy(synth.config1(16, 8).levels)          # host-only: fine
    with pytest.raises(capi.PaError) as e:
        capi.Field(H, 1, 1)                                   # needs the device
    assert e.value.code == -2
