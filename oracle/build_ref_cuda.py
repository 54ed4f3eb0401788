#!/usr/bin/env python3
"""Secondary baseline: the UNMODIFIED reference built the way its own GNUmake builds it for GPUs (USE_CUDA=TRUE,
CUDA_ARCH=100: AMReX's generic ParallelFor backend, Tools/GNUMake/comps/nvcc.mak) into oracle/_ref/ -- so the time of the
reference's hot path on the same B200 can be reported next to ours (bench.py: ref_gpu_baseline).  TEST / BENCH INFRASTRUCTURE
ONLY, same rules as build_ref.py: our own recipe (no GNUmake / CMake run), sources compiled from where they lie under
/root/reference, nothing copied, outputs only under oracle/_ref/.

Flags follow nvcc.mak:103-165 and Make.defs:824-835,1155 for `USE_CUDA=TRUE USE_OMP=FALSE USE_GPU_RDC=FALSE`:
  nvcc -ccbin=g++ -Xcompiler='-O3 -std=c++17 -pthread' --std=c++17 -m64 -maxrregcount=255 --expt-relaxed-constexpr
       --expt-extended-lambda --forward-unknown-to-host-compiler -lineinfo --ptxas-options=-O3 --use_fast_math
       -gencode arch=compute_100,code=sm_100 -x cu -c     (+ -DAMREX_USE_CUDA -DAMREX_USE_GPU -DBL_COALESCE_FABS
       -DAMREX_GPU_MAX_THREADS=256), link with -lcuda -lcurand.
(--use_fast_math is the reference's default for CUDA builds: its GPU results are not bit-identical to its CPU results.)

Products: grad3d.cuda.timed.ex, curvature3d.cuda.timed.ex (the timed variants of build_ref.py, PA_TIMED_REPS loop included).
"""
import argparse
import concurrent.futures as cf
import glob
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_ref as B  # noqa: E402

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
OUT = os.path.join(B.OUT, "cuda")

DEFINES = dict(B.DEFINES)
for k in ("BL_USE_OMP", "AMREX_USE_OMP"):
    DEFINES.pop(k)
DEFINES.update({"AMREX_USE_CUDA": "1", "AMREX_USE_GPU": "1", "BL_COALESCE_FABS": "1", "AMREX_GPU_MAX_THREADS": "256"})


def write_headers():
    inc = os.path.join(OUT, "include")
    os.makedirs(inc, exist_ok=True)
    lines = ["#ifndef AMREX_HAVE_NO_CONFIG_H", "#define AMREX_HAVE_NO_CONFIG_H"]
    for k, v in DEFINES.items():
        lines += [f"#ifndef {k}", f"#define {k} {v}", "#endif"]
    lines += ["#endif"]
    B._write_if_changed(os.path.join(inc, "AMReX_Config.H"), "\n".join(lines) + "\n")
    B._write_if_changed(os.path.join(inc, "AMReX_Version.H"),
                        '#ifndef AMREX_VERSION_H_\n#define AMREX_VERSION_H_\n#define AMREX_GIT_VERSION "24.10"\n#define AMREX_RELEASE_NUMBER 241000\n#endif\n')
    return inc


def flags(inc):
    f = ["-ccbin=g++", "-Xcompiler=-O3 -std=c++17 -pthread -w", "--std=c++17", "-Wno-deprecated-gpu-targets", "-m64", "-maxrregcount=255",
         "--expt-relaxed-constexpr", "--expt-extended-lambda", "--forward-unknown-to-host-compiler", "-lineinfo", "--ptxas-options=-O3",
         "--use_fast_math", "-w", "-gencode", "arch=compute_100,code=sm_100", "-I" + inc]
    for d in B.SRC_DIRS:
        f.append("-I" + os.path.join(B.AMREX, d))
    f.append("-I" + os.path.join(B.AMREX, "Tools/C_scripts"))
    return f


def compile_one(args):
    src, obj, fl = args
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return src, 0, ""
    p = subprocess.run([NVCC, *fl, "-x", "cu", "-c", src, "-o", obj], capture_output=True, text=True)
    return src, p.returncode, p.stderr[-2000:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    a = ap.parse_args()
    if not os.path.isdir(B.AMREX):
        print("reference not present; keeping prebuilt oracle/_ref/cuda", file=sys.stderr)
        return 0
    t0 = time.time()
    inc = write_headers()
    fl = flags(inc)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for d in B.SRC_DIRS:
        for src in sorted(glob.glob(os.path.join(B.AMREX, d, "*.cpp"))):
            b = os.path.basename(src)
            if "AmrCore" in d:
                if not B.KEEP_AMRCORE.search(b):
                    continue
            elif B.EXCLUDE.search(b):
                continue
            jobs.append((src, os.path.join(objdir, b[:-4] + ".o"), fl))
    failed = []
    with cf.ThreadPoolExecutor(a.j) as ex:
        for src, rc, err in ex.map(compile_one, jobs):
            if rc != 0:
                failed.append((src, err))
    for src, err in failed:
        print(f"[build_ref_cuda] skipped (did not compile, not needed if link succeeds): {src}\n{err[-600:]}", file=sys.stderr)
    good = [o for (s, o, _) in jobs if os.path.exists(o)]
    lib = os.path.join(OUT, "libamrex_cuda.a")
    if os.path.exists(lib):
        os.remove(lib)
    subprocess.check_call(["ar", "rcs", lib, *good])
    srcdir = os.path.join(B.OUT, "src")          # the timed temporaries build_ref.py wrote
    rc_all = 0
    for src, exe in ((os.path.join(srcdir, "grad_timed.cpp"), "grad3d.cuda.timed.ex"), (os.path.join(srcdir, "curvature_timed.cpp"), "curvature3d.cuda.timed.ex")):
        if not os.path.exists(src):
            print(f"[build_ref_cuda] {src} missing: run oracle/build_ref.py first", file=sys.stderr)
            return 1
        obj = os.path.join(objdir, os.path.basename(src)[:-4] + ".o")
        p = subprocess.run([NVCC, *fl, "-x", "cu", "-c", src, "-o", obj], capture_output=True, text=True)
        if p.returncode == 0:
            p = subprocess.run([NVCC, "-ccbin=g++", "-m64", "-gencode", "arch=compute_100,code=sm_100", obj, "-o", os.path.join(B.OUT, exe), lib,
                                "-L/usr/local/cuda/lib64/stubs", "-lcuda", "-lcurand", "-lpthread"], capture_output=True, text=True)
        if p.returncode != 0:
            rc_all = 1
            print(f"[build_ref_cuda] FAILED {exe}\n{p.stderr[-3000:]}", file=sys.stderr)
        else:
            print(f"[build_ref_cuda] built {os.path.join(B.OUT, exe)}")
    print(f"[build_ref_cuda] done in {time.time()-t0:.0f}s")
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
