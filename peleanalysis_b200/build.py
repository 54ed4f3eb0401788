"""Compile the sm_100a CUDA library in-tree: peleanalysis_b200/lib/libpelestencil_b200.so.

nvcc cross-compiles without a GPU.  -fmad=false keeps separate IEEE multiplies and adds (bit parity with the
reference's FMA-free CPU build); -lineinfo lets ncu map SASS to source lines.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpelestencil_b200.so")
SOURCES = ["api.cu", "kernels.cu", "stencil_tma.cu", "curv_fused.cu", "curv_f3.cu", "normal_w.cu", "filter.cu", "hier.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--use_fast_math=false".replace("--use_fast_math=false", "-Xptxas=-v")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pele_stencil_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lk:       # several processes (pytest-xdist workers, torchrun ranks) may get here at once
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        objs = []
        procs = []
        for s in SOURCES:
            o = os.path.join(LIBDIR, s.rsplit(".", 1)[0] + ".o")
            objs.append(o)
            cmd = [NVCC, *FLAGS, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        log = []
        for s, p in procs:
            out, _ = p.communicate()
            log.append("== %s\n%s" % (s, out))
            if p.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        subprocess.check_call([NVCC, "-shared", "-o", LIB + ".tmp", *objs, "-lcudart"])
        os.replace(LIB + ".tmp", LIB)
        with open(os.path.join(LIBDIR, "build.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
