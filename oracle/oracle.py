"""ctypes wrapper around oracle/libpa_oracle.so (the CPU restatement) and helpers to run the
compiled reference in oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (peleanalysis_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import shutil
import subprocess
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_LIB = None


def build() -> str:
    so = os.path.join(HERE, "libpa_oracle.so")
    src = os.path.join(HERE, "pa_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        import fcntl
        with open(os.path.join(HERE, ".build.lock"), "w") as lk:     # pytest-xdist workers: one make at a time
            fcntl.flock(lk, fcntl.LOCK_EX)
            if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
                subprocess.check_call(["make", "-C", HERE, "libpa_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.pao_hier_create.restype = C.c_void_p
        L.pao_hier_create.argtypes = [C.c_int] + [C.c_void_p] * 7
        L.pao_hier_destroy.argtypes = [C.c_void_p]
        L.pao_total_cells.restype = C.c_int64
        L.pao_total_cells.argtypes = [C.c_void_p]
        L.pao_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pao_curvature.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double,
                                    C.c_int, C.c_void_p, C.c_void_p]
        L.pao_curvature_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pao_filled_fabs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.pao_fb_source_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pao_mask.restype = C.c_int64
        L.pao_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleHier:
    """Hierarchy metadata of a Plotfile-like object (levels with domain, dx, boxes)."""

    def __init__(self, pf, is_per=(1, 1, 1), sym_dir=(0, 0, 0), ratios: Optional[Sequence[int]] = None):
        self.pf = pf
        nlev = len(pf.levels)
        dom = np.array([list(l.domain_lo) + list(l.domain_hi) for l in pf.levels], dtype=np.int32)
        dx = np.array([l.dx for l in pf.levels], dtype=np.float64)
        if ratios is None:   # from the level domains, as MLLinOp does (AMReX_MLLinOp.H:856-885)
            ratios = [(pf.levels[l + 1].domain_hi[0] - pf.levels[l + 1].domain_lo[0] + 1)
                      // (pf.levels[l].domain_hi[0] - pf.levels[l].domain_lo[0] + 1) for l in range(nlev - 1)]
        self.ratios = list(ratios)
        rat = np.array(self.ratios + [1], dtype=np.int32)
        nb = np.array([len(l.boxes) for l in pf.levels], dtype=np.int32)
        bx = np.array([list(lo) + list(hi) for l in pf.levels for lo, hi in l.boxes], dtype=np.int32)
        per = np.array(is_per, dtype=np.int32)
        bck = np.array(sym_dir, dtype=np.int32)
        self._keep = (dom, dx, rat, nb, bx, per, bck)
        self.h = lib().pao_hier_create(nlev, _p(dom), _p(dx), _p(rat), _p(nb), _p(bx), _p(per), _p(bck))
        self.total = lib().pao_total_cells(self.h)
        self.boxes = [(l, lo, hi) for l, lv in enumerate(pf.levels) for lo, hi in lv.boxes]

    def __del__(self):
        try:
            lib().pao_hier_destroy(self.h)
        except Exception:
            pass

    def flatten(self, comp: int) -> np.ndarray:
        return np.concatenate([f[comp].ravel() for l in self.pf.levels for f in l.fabs])

    def unflatten(self, flat: np.ndarray) -> List[List[np.ndarray]]:
        out, o = [], 0
        for lv in self.pf.levels:
            lst = []
            for lo, hi in lv.boxes:
                n = [hi[d] - lo[d] + 1 for d in range(3)]
                m = n[0] * n[1] * n[2]
                lst.append(flat[o:o + m].reshape(n[2], n[1], n[0]))
                o += m
            out.append(lst)
        return out

    def grad(self, s: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(s, dtype=np.float64)
        out = np.empty((4, self.total))
        lib().pao_grad(self.h, _p(s), _p(out))
        return out

    def curvature(self, S: np.ndarray, prog_min: float, prog_max: float, do_threshold=False, threshold=1e-4,
                  crse_ratio: int = 2, gauss: bool = False):
        S = np.ascontiguousarray(S, dtype=np.float64)
        out = np.empty((5, self.total))
        g = np.zeros(self.total) if gauss else None
        lib().pao_curvature(self.h, _p(S), prog_min, prog_max, int(do_threshold), threshold, crse_ratio,
                            _p(out), _p(g) if gauss else None)
        return (out, g) if gauss else out

    def curvature_ex(self, S, U, prog_min, prog_max, do_threshold=False, threshold=1e-4, crse_ratio=2):
        """All optional branches.  U: [3, total].  Returns dict of flat fields."""
        S = np.ascontiguousarray(S, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        T = self.total
        out = np.empty((5, T)); g = np.zeros(T); sr = np.zeros(T); rost = np.zeros((9, T)); vn = np.zeros(T)
        lib().pao_curvature_ex(self.h, _p(S), prog_min, prog_max, int(do_threshold), threshold, crse_ratio,
                               _p(out), _p(g), _p(U), _p(sr), _p(rost), _p(vn))
        return dict(core=out, gauss=g, strain=sr, rost=rost, veln=vn)

    def filled_fabs(self, lev: int, s: np.ndarray, ng: int = 1, ghost_init: float = 0.0, crse_ratio: int = 0):
        lv = self.pf.levels[lev]
        sizes = [np.prod([hi[d] - lo[d] + 1 + 2 * ng for d in range(3)]) for lo, hi in lv.boxes]
        out = np.empty(int(sum(sizes)))
        lib().pao_filled_fabs(self.h, lev, _p(np.ascontiguousarray(s)), ng, ghost_init, crse_ratio, _p(out))
        res, o = [], 0
        for (lo, hi), m in zip(lv.boxes, sizes):
            n = [hi[d] - lo[d] + 1 + 2 * ng for d in range(3)]
            res.append(out[o:o + int(m)].reshape(n[2], n[1], n[0]))
            o += int(m)
        return res

    def fb_source_map(self, lev: int, ng: int = 1):
        lv = self.pf.levels[lev]
        sizes = [int(np.prod([hi[d] - lo[d] + 1 + 2 * ng for d in range(3)])) for lo, hi in lv.boxes]
        out = np.empty(sum(sizes), dtype=np.int64)
        lib().pao_fb_source_map(self.h, lev, ng, _p(out))
        res, o = [], 0
        for (lo, hi), m in zip(lv.boxes, sizes):
            n = [hi[d] - lo[d] + 1 + 2 * ng for d in range(3)]
            res.append(out[o:o + m].reshape(n[2], n[1], n[0]))
            o += m
        return res

    def mask(self, lev: int, box: int, face: int, kind: int):
        pb = np.zeros(6, dtype=np.int32)
        n = lib().pao_mask(self.h, lev, box, face, kind, _p(pb), None)
        out = np.empty(n, dtype=np.int32)
        lib().pao_mask(self.h, lev, box, face, kind, _p(pb), _p(out))
        shp = [pb[3 + d] - pb[d] + 1 for d in range(3)]
        return pb, out.reshape(shp[2], shp[1], shp[0])


# ------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref)
# ------------------------------------------------------------------------------------------------

def ref_exe(name: str) -> str:
    return os.path.join(REF_DIR, name)


def have_ref() -> bool:
    return os.path.exists(ref_exe("grad3d.ref.ex")) and os.path.exists(ref_exe("curvature3d.ref.ex"))


def run_ref(tool: str, infile: str, outfile: str, threads: Optional[int] = None, timed: bool = False, variant: Optional[str] = None, **kv):
    """Run the reference `grad` or `curvature` executable.  Returns (stdout, hot_path_seconds or None).
    variant: "ref" (unmodified), "timed" (probes around the hot path), "cuda.timed" (the reference's own CUDA build, see
    build_ref_cuda.py)."""
    exe = ref_exe("%s3d.%s.ex" % (tool, variant or ("timed" if timed else "ref")))
    if os.path.lexists(outfile):
        shutil.rmtree(outfile)
    args = [exe, "infile=" + infile, "outfile=" + outfile]
    if (variant or "").startswith("cuda"):
        # the tools fill their MultiFabs from host code (AmrData::FillVar): on a GPU build that needs AMReX's managed arena
        # (a runtime ParmParse switch of AMReX, not a source change)
        args.append("amrex.the_arena_is_managed=1")
    for k, v in kv.items():
        if isinstance(v, (list, tuple)):
            v = " ".join(str(x) for x in v)
        args.append("%s=%s" % (k, v))
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    p = subprocess.run(args, capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.abspath(outfile)) or ".")
    if p.returncode != 0:
        raise RuntimeError("reference %s failed (%d):\n%s\n%s" % (tool, p.returncode, p.stdout[-2000:], p.stderr[-2000:]))
    m = re.search(r"hot_path_seconds\s+([0-9.eE+-]+)", p.stdout)
    return p.stdout, (float(m.group(1)) if m else None)


def run_ref_timed(tool: str, infile: str, outfile: str, threads: Optional[int] = None, reps: int = 1, variant: str = "timed", **kv):
    """The reference tool's timed build with its hot-path region repeated `reps` times inside ONE process (PA_TIMED_REPS,
    see build_ref.patch_timed).  Returns the list of hot-path seconds, one per repetition."""
    old = os.environ.get("PA_TIMED_REPS")
    os.environ["PA_TIMED_REPS"] = str(int(reps))
    try:
        out, _ = run_ref(tool, infile, outfile, threads=threads, variant=variant, **kv)
    finally:
        if old is None:
            os.environ.pop("PA_TIMED_REPS", None)
        else:
            os.environ["PA_TIMED_REPS"] = old
    hot = [float(x) for x in re.findall(r"hot_path_seconds\s+([0-9.eE+-]+)", out)]
    if len(hot) != int(reps):
        raise RuntimeError("reference %s: expected %d timed repetitions, saw %d" % (tool, reps, len(hot)))
    return hot
