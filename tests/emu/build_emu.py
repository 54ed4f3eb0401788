"""Build tests/emu/_build/libpelestencil_emu.so: the kernel sources of peleanalysis_b200/csrc compiled by g++ against the
CUDA execution-model emulator (TEST INFRASTRUCTURE; see tests/emu/cuda_runtime.h).  Same floating-point rules as the GPU
build: no FMA contraction (-ffp-contract=off; explicit fma() calls go to libm's correctly rounded fma)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "peleanalysis_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libpelestencil_emu.so")
SOURCES = [os.path.join(CSRC, f) for f in ("api.cu", "kernels.cu", "stencil_tma.cu", "curv_fused.cu", "curv_f3.cu", "normal_w.cu", "filter.cu", "hier.cpp")] + [os.path.join(HERE, "cuemu.cpp")]
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DPA_HOST_EMULATION=1", "-I", HERE,
         "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-unused-but-set-variable"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cuemu.cpp", "cuda_runtime.h", "build_emu.py")]
    deps.append(os.path.join(ROOT, "include", "pele_stencil_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    import fcntl
    with open(os.path.join(OUT, ".build.lock"), "w") as lk:          # pytest-xdist workers share this directory: one build at a time
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not force and not needs_build():                          # another worker built it while this one waited
            return LIB
        procs = []
        for s in SOURCES:
            o = os.path.join(OUT, os.path.basename(s).rsplit(".", 1)[0] + ".o")
            procs.append((s, o, subprocess.Popen(["g++", *FLAGS, "-x", "c++", "-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for s, o, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError("g++ failed on %s:\n%s" % (s, out))
        subprocess.check_call(["g++", "-shared", "-o", LIB + ".tmp"] + [o for _, o, _ in procs])
        os.replace(LIB + ".tmp", LIB)                                # a process that already mapped the old file keeps its copy
    return LIB


if __name__ == "__main__":
    print(build(force=True))
