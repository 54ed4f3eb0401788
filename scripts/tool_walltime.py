#!/usr/bin/env python3
"""Tool-level wall time on the GPU box: the reference executables on the host cores against the drop-in
executables on the GPU, whole process from argv to a finished output plotfile on the same /dev/shm plotfile.
python scripts/tool_walltime.py [base=256]  ->  one JSON line per tool"""
import json, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from peleanalysis_b200 import plotfile, synth

base = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = sys.argv[2].split(",") if len(sys.argv) > 2 else ["grad", "curvature", "filterPlt"]
HOST = os.path.join(ROOT, "peleanalysis_b200", "host")
subprocess.check_call(["make", "-s", "-C", HOST])
tmp = tempfile.mkdtemp(prefix="pa_wall_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))


def run(cmd, cwd):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, env=env)
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-600:] + p.stderr[-600:])
    return dt, p.stdout


try:
    # grad / curvature: 3 levels, base^3 cells each (the configs[2] shape), 64^3 boxes, temp + velocities
    pf = synth.config3(base, 64, names=("temp", "x_velocity", "y_velocity", "z_velocity"))
    d = os.path.join(tmp, "plt")
    plotfile.write_plotfile(d, pf, clean="remove")
    cells = sum(l.ncells for l in pf.levels)
    del pf
    for tool, ref, mine, args in (
            ("grad", "grad3d.ref.ex", "grad3d.b200.ex", ["gradVar=temp"]),
            ("curvature", "curvature3d.ref.ex", "curvature3d.b200.ex", ["progressName=temp"]),
            ("filterPlt", "filterPlt3d.ref.ex", "filterPlt3d.b200.ex", ["variables=temp", "max_filter_level=1"])):
        if tool not in only:
            continue
        rec = {"tool": tool, "cells": cells if tool != "filterPlt" else cells * 2 // 3, "args": args, "host_cores": os.cpu_count()}
        for who, exe in (("reference_s", O.ref_exe(ref)), ("b200_s", os.path.join(HOST, mine)), ("b200_second_run_s", os.path.join(HOST, mine))):
            wd = os.path.join(tmp, who + "_" + tool)
            os.makedirs(wd, exist_ok=True)
            extra = [] if tool == "filterPlt" else ["outfile=" + os.path.join(wd, "out")]
            if who != "reference_s":
                extra.append("verbose=1")
            dt, out = run([exe, "infile=" + d, *args, *extra], wd)
            rec[who] = round(dt, 3)
            if who == "b200_s":
                rec["b200_breakdown"] = [ln.strip() for ln in out.splitlines() if ln.startswith("[b200]")]
            shutil.rmtree(wd, ignore_errors=True)
        rec["speedup_whole_process"] = round(rec["reference_s"] / rec["b200_second_run_s"], 2)
        print(json.dumps(rec), flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
