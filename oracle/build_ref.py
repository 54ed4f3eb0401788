#!/usr/bin/env python3
"""Compile the UNMODIFIED reference (PeleAnalysis grad / curvature + the vendored AMReX 24.10
subset they need) into oracle/_ref/ with plain g++.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may execute it.

* Sources are compiled from where they lie under /root/reference; nothing is copied into the
  repo and outputs go only to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
* This is our own recipe: it does not run the reference's GNUmake / CMake build system.  The
  one generated header AMReX needs (AMReX_Config.H, normally emitted by
  Tools/libamrex/mkconfig.py from the make command line) is written here from the define
  list below -- the same list the reference's GNUmake uses for `BL_NO_FORT=TRUE USE_OMP=TRUE
  DIM=3 COMP=gnu` (Tools/GNUMake/Make.defs:83,841,902,931,936).
* Compile flags follow the reference build: g++ -O3 -finline-limit=43210 -std=c++17 -fopenmp,
  NO -march (so no FMA contraction) -- the parity anchor named in BASELINE.json.

Products (oracle/_ref/):
  grad3d.ref.ex, curvature3d.ref.ex      the reference tools, unmodified
  grad3d.timed.ex, curvature3d.timed.ex  same sources with amrex::second() probes inserted
                                         around the hot path (grad.cpp:151-236,
                                         curvature.cpp:283-791, FillVar excluded); the patched
                                         temporaries live only in oracle/_ref/src/
  fcompare.ref.ex                        AMReX's plotfile differ (Tools/Plotfile/fcompare.cpp)
  filterPlt3d.ref.ex, filterPlt3d.timed.ex  the neighbouring stencil tool (Src/filterPlt.cpp) with PelePhysics' Filter and
                                         PltFileManager and AMReX's FillPatchUtil; the timed variant probes
                                         filterPlt.cpp:166-221 (ghost-cell fill + filter, no plotfile I/O)
"""
import argparse
import concurrent.futures as cf
import glob
import os
import re
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PELE_ANALYSIS_HOME", "/root/reference")
AMREX = os.path.join(REF, "Submodules/PelePhysics/Submodules/amrex")
OUT = os.path.join(HERE, "_ref")

SRC_DIRS = ["Src/Base", "Src/Base/Parser", "Src/Boundary", "Src/Extern/amrdata",
            "Src/AmrCore", "Src/LinearSolvers/MLMG"]
# files guarded out by the reference's Make.package for this configuration
EXCLUDE = re.compile(r"(_fi\.cpp|_F\.cpp|CuptiTrace|MPMD|MemProfiler|Lazy\.cpp|_eb\.cpp|MLEB|EBABec|"
                     r"_hypre|Hypre|PETSc|HDF5|Conduit|SUNDIALS|AMReX_AmrMeshInSituBridge|"
                     r"AMReX_ErrorList|AMReX_TagBox|AMReX_AmrMesh|AMReX_AmrCore|AMReX_Cluster|"
                     r"AMReX_FluxRegister|AMReX_FillPatchUtil|AMReX_Interpolater|AMReX_MFInterpolater|"
                     r"AMReX_InterpBase|AMReX_FillPatcher|AMReX_AmrParGDB|AMReX_AmrParticles)")
# AmrCore is on the reference's include path (curvature.cpp includes AMReX_Interpolater.H) but no
# AmrCore object is referenced by either tool except through MLMG's YAFluxRegister.
KEEP_AMRCORE = re.compile(r"(YAFluxRegister|AMReX_InterpBase|AMReX_Interpolater|AMReX_MFInterpolater)")

DEFINES = {
    "BL_SPACEDIM": "3", "AMREX_SPACEDIM": "3", "BL_FORT_USE_UNDERSCORE": "1",
    "AMREX_FORT_USE_UNDERSCORE": "1", "BL_Linux": "1", "AMREX_Linux": "1", "NDEBUG": "1",
    "BL_NO_FORT": "1", "BL_USE_OMP": "1", "AMREX_USE_OMP": "1", "AMREX_GPU_MAX_THREADS": "0",
    "AMREX_DIMENSION_AGNOSTIC": "1", "AMREX_NO_PROBINIT": "1", "BL_USE_ARRAYVIEW": "1",
    "AMREX_USE_ARRAYVIEW": "1",
}


def write_headers(tiny_profile):
    inc = os.path.join(OUT, "include")
    os.makedirs(inc, exist_ok=True)
    d = dict(DEFINES)
    if tiny_profile:
        d["AMREX_TINY_PROFILING"] = "1"
    lines = ["#ifndef AMREX_HAVE_NO_CONFIG_H", "#define AMREX_HAVE_NO_CONFIG_H"]
    for k, v in d.items():
        lines += [f"#ifndef {k}", f"#define {k} {v}", "#endif"]
    lines += ["#endif"]
    _write_if_changed(os.path.join(inc, "AMReX_Config.H"), "\n".join(lines) + "\n")
    ver = ('#ifndef AMREX_VERSION_H_\n#define AMREX_VERSION_H_\n'
           '#define AMREX_GIT_VERSION "24.10"\n#define AMREX_RELEASE_NUMBER 241000\n#endif\n')
    _write_if_changed(os.path.join(inc, "AMReX_Version.H"), ver)
    # AMReX_buildInfo.cpp is referenced by Make.Analysis but by neither tool: not needed.
    return inc


def _write_if_changed(path, text):
    if os.path.exists(path) and open(path).read() == text:
        return
    with open(path, "w") as f:
        f.write(text)


def cxx_flags(inc):
    f = ["-O3", "-finline-limit=43210", "-std=c++17", "-fopenmp", "-pthread", "-w", "-I" + inc]
    for d in SRC_DIRS:
        f.append("-I" + os.path.join(AMREX, d))
    f.append("-I" + os.path.join(AMREX, "Tools/C_scripts"))
    return f


def compile_one(args):
    src, obj, flags = args
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return src, 0, ""
    p = subprocess.run(["g++", *flags, "-c", src, "-o", obj], capture_output=True, text=True)
    return src, p.returncode, p.stderr[-2000:]


def patch_timed(src_path, dst_path, begin_after, end_before, skip_regex, tag, cleanup=""):
    """Insert amrex::second() probes into a temporary copy of a reference tool (in oracle/_ref/src,
    never committed).  `begin_after` / `end_before` are regexes matching unique anchor lines;
    every line matching skip_regex (the FillVar disk read) is bracketed and its time subtracted.
    The probed region is repeated PA_TIMED_REPS times inside the process (a backward goto: the objects the
    region declares are destroyed and rebuilt each time, exactly as in a fresh run; `cleanup` frees what the
    region allocates with new) and prints one hot_path_seconds line per repetition, so a benchmark can time
    K steps of the full-size workload without K process start-ups and plotfile opens."""
    out = []
    began = ended = False
    for line in open(src_path):
        if not began and re.search(begin_after, line):
            out.append(line)
            out.append("    int pa_rep__ = 0; const int pa_reps__ = std::getenv(\"PA_TIMED_REPS\") ? std::atoi(std::getenv(\"PA_TIMED_REPS\")) : 1;\n")
            out.append("  pa_again__: ;\n")
            out.append("    double pa_t0__ = amrex::second(); double pa_skip__ = 0.0;\n")
            began = True
            continue
        if began and not ended and re.search(end_before, line):
            out.append('    { double pa_t1__ = amrex::second(); amrex::Print() << "' + tag +
                       ' hot_path_seconds " << std::setprecision(9) << (pa_t1__-pa_t0__-pa_skip__)'
                       ' << " skipped_io_seconds " << pa_skip__ << "\\n"; }\n')
            out.append("    if (++pa_rep__ < pa_reps__) { " + cleanup + " goto pa_again__; }\n")
            ended = True
        if began and not ended and re.search(skip_regex, line):
            out.append("    { double pa_s0__ = amrex::second();\n")
            out.append(line)
            out.append("      pa_skip__ += amrex::second() - pa_s0__; }\n")
            continue
        out.append(line)
    if not (began and ended):
        raise RuntimeError(f"anchors not found in {src_path}")
    os.makedirs(os.path.dirname(dst_path), exist_ok=True)
    _write_if_changed(dst_path, "#include <iomanip>\n#include <cstdlib>\n" + "".join(out))


def patch_splice(src_path, dst_path, begin_regex, end_regex, block_include):
    """A temporary copy of a reference tool (oracle/_ref/src, never committed) whose operator block -- the lines from the
    one matching begin_regex up to, not including, the one matching end_regex -- is replaced by `#include "<block>"`: the
    source-level integration INTEGRATION.md describes (AMReX keeps plotfile I/O, BoxArray / Geometry metadata and, for
    do_smooth, its own MLMG solve on the host; the stencil path goes through libpelestencil_b200's C ABI)."""
    out, state = ['#include "pa_amrex_glue.H"\n'], 0
    for line in open(src_path):
        if state == 0 and re.search(begin_regex, line):
            out.append('#include "%s"\n' % block_include)
            state = 1
            continue
        if state == 1:
            if re.search(end_regex, line):
                state = 2
            else:
                continue
        out.append(line)
    if state != 2:
        raise RuntimeError(f"splice anchors not found in {src_path}")
    os.makedirs(os.path.dirname(dst_path), exist_ok=True)
    _write_if_changed(dst_path, "".join(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--tiny-profile", action="store_true")
    ap.add_argument("--no-fcompare", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(AMREX):
        print(f"reference not present at {REF}; keeping prebuilt oracle/_ref", file=sys.stderr)
        return 0
    t0 = time.time()
    inc = write_headers(a.tiny_profile)
    flags = cxx_flags(inc)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for d in SRC_DIRS:
        for src in sorted(glob.glob(os.path.join(AMREX, d, "*.cpp"))):
            b = os.path.basename(src)
            if "AmrCore" in d:
                if not KEEP_AMRCORE.search(b):
                    continue
            elif EXCLUDE.search(b):
                continue
            jobs.append((src, os.path.join(objdir, b[:-4] + ".o"), flags))
    failed = []
    with cf.ThreadPoolExecutor(a.j) as ex:
        for src, rc, err in ex.map(compile_one, jobs):
            if rc != 0:
                failed.append((src, err))
    for src, err in failed:
        print(f"[build_ref] skipped (did not compile, not needed if link succeeds): {src}\n{err[-400:]}",
              file=sys.stderr)
    good = [o for (s, o, _) in jobs if os.path.exists(o)]
    lib = os.path.join(OUT, "libamrex_ref.a")
    if os.path.exists(lib):
        os.remove(lib)
    subprocess.check_call(["ar", "rcs", lib, *good])

    srcdir = os.path.join(OUT, "src")
    patch_timed(os.path.join(REF, "Src/grad.cpp"), os.path.join(srcdir, "grad_timed.cpp"),
                r"const int nGrow = 1;", r"Write the results", r"amrData\.FillVar\(", "[grad]")
    patch_timed(os.path.join(REF, "Src/curvature.cpp"), os.path.join(srcdir, "curvature_timed.cpp"),
                r"const int nGrow = 2 ;", r"Set-up the output", r"amrData\.FillVar\(", "[curvature]",
                cleanup="for (int pa_l__ = 0; pa_l__ < Nlev; ++pa_l__) { delete state[pa_l__]; delete flame_normal[pa_l__]; "
                        "delete cell_normal[pa_l__]; delete geoms[pa_l__]; }")
    # source-level drop-in: the reference tools with their operator blocks replaced by the C-ABI calls, linked against the
    # host-only AMReX built above and the product library (skipped when the library has not been built yet)
    repo = os.path.dirname(HERE)
    glue = os.path.join(repo, "peleanalysis_b200", "host", "amrex_glue")
    palib = os.path.join(repo, "peleanalysis_b200", "lib", "libpelestencil_b200.so")
    shells = []
    if os.path.exists(palib) and os.path.isdir(glue):
        patch_splice(os.path.join(REF, "Src/grad.cpp"), os.path.join(srcdir, "grad_b200amrex.cpp"),
                     r"Get face-centered gradients from MLMG", r"Write the results", "grad_block.inc")
        patch_splice(os.path.join(REF, "Src/curvature.cpp"), os.path.join(srcdir, "curvature_b200amrex.cpp"),
                     r"Compute curvature using LinearOperators", r"Set-up the output", "curvature_block.inc")
        shells = [(os.path.join(srcdir, "grad_b200amrex.cpp"), "grad3d.b200amrex.ex"),
                  (os.path.join(srcdir, "curvature_b200amrex.cpp"), "curvature3d.b200amrex.ex")]
    shell_flags = ["-I" + glue, "-I" + os.path.join(repo, "include")]
    shell_libs = ["-L" + os.path.dirname(palib), "-lpelestencil_b200", "-Wl,-rpath," + os.path.dirname(palib),
                  "-Wl,-rpath,$ORIGIN/../../peleanalysis_b200/lib"]
    tools = [(os.path.join(REF, "Src/grad.cpp"), "grad3d.ref.ex"),
             (os.path.join(REF, "Src/curvature.cpp"), "curvature3d.ref.ex"),
             (os.path.join(srcdir, "grad_timed.cpp"), "grad3d.timed.ex"),
             (os.path.join(srcdir, "curvature_timed.cpp"), "curvature3d.timed.ex")]
    if not a.no_fcompare:
        tools.append((os.path.join(AMREX, "Tools/Plotfile/fcompare.cpp"), "fcompare.ref.ex"))
    tools += shells

    # filterPlt: three more objects (AMReX's FillPatchUtil is excluded from the library above because neither grad nor
    # curvature references it; Filter and PltFileManager live in PelePhysics)
    pp = os.path.join(REF, "Submodules/PelePhysics/Source/Utility")
    filt_flags = ["-I" + os.path.join(pp, "Filter"), "-I" + os.path.join(pp, "PltFileManager")]
    filt_objs = []
    filt_jobs = [(os.path.join(AMREX, "Src/AmrCore/AMReX_FillPatchUtil.cpp"), os.path.join(objdir, "AMReX_FillPatchUtil.o"), flags),
                 (os.path.join(pp, "Filter/Filter.cpp"), os.path.join(objdir, "PP_Filter.o"), flags + filt_flags),
                 (os.path.join(pp, "PltFileManager/PltFileManager.cpp"), os.path.join(objdir, "PP_PltFileManager.o"), flags + filt_flags)]
    if all(os.path.exists(j[0]) for j in filt_jobs):
        with cf.ThreadPoolExecutor(a.j) as ex:
            for src, rc, err in ex.map(compile_one, filt_jobs):
                if rc != 0:
                    print(f"[build_ref] FAILED {src}\n{err}", file=sys.stderr)
        filt_objs = [j[1] for j in filt_jobs if os.path.exists(j[1])]
    filt_tools = []
    if len(filt_objs) == len(filt_jobs):
        patch_timed(os.path.join(REF, "Src/filterPlt.cpp"), os.path.join(srcdir, "filterPlt_timed.cpp"),
                    r"fillPatchFromPlt doesn.t fill ghost cells", r"Saving filtered data", r"PA_NEVER_MATCHES", "[filterPlt]")
        filt_tools = [(os.path.join(REF, "Src/filterPlt.cpp"), "filterPlt3d.ref.ex"),
                      (os.path.join(srcdir, "filterPlt_timed.cpp"), "filterPlt3d.timed.ex")]
        tools += filt_tools

    def link(t):
        src, exe = t
        exe = os.path.join(OUT, exe)
        if os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(src), os.path.getmtime(lib)):
            return exe, 0, ""
        extra_f = shell_flags if t in shells else (filt_flags if t in filt_tools else [])
        extra_l = shell_libs if t in shells else []
        extra_o = filt_objs if t in filt_tools else []
        p = subprocess.run(["g++", *flags, *extra_f, src, *extra_o, "-o", exe, lib, "-lgomp", "-lpthread", *extra_l],
                           capture_output=True, text=True)
        return exe, p.returncode, p.stderr[-3000:]
    rc_all = 0
    with cf.ThreadPoolExecutor(a.j) as ex:
        for exe, rc, err in ex.map(link, tools):
            if rc != 0:
                rc_all = 1
                print(f"[build_ref] FAILED {exe}\n{err}", file=sys.stderr)
            else:
                print(f"[build_ref] built {exe}")
    print(f"[build_ref] done in {time.time()-t0:.0f}s")
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
