"""Developer helper (CPU): run the curvature golden cases through the emulated library, fused vs unfused, and print whether
the fused kernel was the one that ran.  Usage: python scripts/emu_curv_check.py [case ...]"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests"), ROOT]
import build_emu  # noqa: E402
from peleanalysis_b200 import capi as pc  # noqa: E402

spec = importlib.util.spec_from_file_location("capi_emulated", pc.__file__)
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
m.LIB_PATH = build_emu.build()
os.environ["PA_NORMAL_MATH"] = "fast"
os.environ.setdefault("CUEMU_SEED", "0")
m.init(0)
import test_gpu_parity as G  # noqa: E402
from cases import CASES  # noqa: E402

names = sys.argv[1:] or [n for n, c in CASES.items() if "curvature" in c[3]]
bad = 0
for name in names:
    for st in ("tma", "tma_fused"):
        for links in ("links", "nolinks"):
            f0 = m.curv_fused_launches()
            try:
                G.test_curvature_matches_reference_golden(m, name, st, links)
                print(name, st, links, "OK   fused launches:", m.curv_fused_launches() - f0)
            except AssertionError as e:
                bad += 1
                print(name, st, links, "FAIL", str(e)[:300])
sys.exit(1 if bad else 0)
