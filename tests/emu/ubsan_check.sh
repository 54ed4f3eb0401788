#!/bin/bash
# Alignment pass over the kernel sources: the emulated library built with -fsanitize=alignment (a misaligned 128-bit access --
# a fault on the GPU -- aborts here), golden cases in every stencil variant incl. the opt-in curvature kernels.   usage: tests/emu/ubsan_check.sh
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=$ROOT/tests/emu/_build/ubsan
mkdir -p "$OUT"
FLAGS="-std=c++17 -O1 -g -fPIC -fsanitize=alignment -fno-sanitize-recover=alignment -fno-omit-frame-pointer -ffp-contract=off -DPA_HOST_EMULATION=1 -I $ROOT/tests/emu"
for f in api.cu kernels.cu stencil_tma.cu curv_fused.cu curv_f3.cu normal_w.cu filter.cu hier.cpp; do g++ $FLAGS -x c++ -c "$ROOT/peleanalysis_b200/csrc/$f" -o "$OUT/${f%.*}.o" & done
g++ $FLAGS -c "$ROOT/tests/emu/cuemu.cpp" -o "$OUT/cuemu.o"
wait
g++ -shared -fsanitize=alignment -o "$OUT/libpelestencil_emu.so" "$OUT"/api.o "$OUT"/kernels.o "$OUT"/stencil_tma.o "$OUT"/curv_fused.o "$OUT"/curv_f3.o "$OUT"/normal_w.o "$OUT"/filter.o "$OUT"/hier.o "$OUT"/cuemu.o
cat > "$OUT/run.py" <<PY
import importlib.util, os, sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
from peleanalysis_b200 import capi as pc
import test_gpu_parity as G
from cases import CASES
spec = importlib.util.spec_from_file_location("capi_emulated", pc.__file__); emu = importlib.util.module_from_spec(spec); spec.loader.exec_module(emu)
emu.LIB_PATH = "$OUT/libpelestencil_emu.so"; os.environ["PA_NORMAL_MATH"] = "fast"; emu.init(0)
n = 0
for name in CASES:
    for st in ("tma", "tma_big", "simple"):
        for bc in ("0", "1"):
            os.environ["PA_BCFILL_V2"] = bc
            if "grad" in CASES[name][3]: G.test_grad_matches_reference_golden(emu, name, st, "links"); n += 1
            if "curvature" in CASES[name][3]: G.test_curvature_matches_reference_golden(emu, name, st, "links"); n += 1
    os.environ["PA_BCFILL_V2"] = "1"
    for st in ("tma_fused", "tma_fused3", "tma_n3", "tma_nw"):      # the opt-in curvature kernels, linked and materialised ghosts
        for links in ("links", "nolinks"):
            if "curvature" in CASES[name][3]: G.test_curvature_matches_reference_golden(emu, name, st, links); n += 1
    G.test_ghost_cells_match_oracle(emu, name); n += 1
for base, mgs in (((21, 14, 9), 7), ((25, 10, 12), 5), ((33, 12, 6), 11), ((48, 24, 24), 24)):   # odd box widths: every pitch / lead pad
    from peleanalysis_b200 import synth
    pf = synth.make_hierarchy(base, (), (), mgs, ("temp",))
    a, _, _ = G._gpu_grad(emu, pf, (1, 1, 1), (0, 0, 0), stencil="tma"); b, _, _ = G._gpu_grad(emu, pf, (1, 1, 1), (0, 0, 0), stencil="simple")
    assert all((a[c].view("u8") == b[c].view("u8")).all() for c in range(4)); n += 1
for args in ((32, 16, True), (72, 72, False)):                       # x strips of the later kernels, threshold clip
    G.test_fused3_strips_match_separate_kernels(emu, *args); n += 1
for ring in ("0", "1"):
    os.environ["PA_NW_RING"] = ring
    G.test_curvature_matches_reference_golden(emu, "c3_three_levels", "tma_nw", "links"); n += 1
os.environ["PA_NW_RING"] = "0"
print("ubsan (alignment) check: %d runs, no error" % n)
PY
LD_PRELOAD=$(gcc -print-file-name=libubsan.so) python "$OUT/run.py"
