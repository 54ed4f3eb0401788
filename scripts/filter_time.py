#!/usr/bin/env python3
"""Per-level timing of the filterPlt path on the device (CUDA events on the library's stream = torch's current stream).
python scripts/filter_time.py [base] [plotfile max_grid_size] [output max_grid_size] [filter_type] [base_fgr] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peleanalysis_b200 import capi as P, filterplt, synth

base = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mgs_in = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mgs = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ftype = int(sys.argv[4]) if len(sys.argv) > 4 else 1
fgr = int(sys.argv[5]) if len(sys.argv) > 5 else 2
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
P.init(0)
pf = synth.config3(base, mgs_in, fill=False)
run = filterplt.FilterRun(P, pf, filter_type=ftype, base_fgr=fgr, max_grid_size=mgs, upload=False)
run.fin.set_val(1.0)
P.sync()
ev = lambda: torch.cuda.Event(enable_timing=True)
res = {"base": base, "mgs": mgs, "filter_type": ftype, "base_fgr": fgr, "ngrow": run.ngrow, "cells": run.cells(), "levels": []}
for _ in range(2):
    run.step()
P.sync()
tot = 0.0
for l in range(run.nlev):
    tf = tk = 0.0
    for _ in range(reps):
        a, b, c = ev(), ev(), ev()
        a.record()
        P.fill_patch(run.fin, 0, run.ncomp, l, run.ngrow[l], run.interp_type)
        b.record()
        P.filter_level(run.fin, 0, run.fout, 0, run.ncomp, l, run.filter_type, run.fgr[l])
        c.record()
        torch.cuda.synchronize()
        tf += a.elapsed_time(b); tk += b.elapsed_time(c)
    cells = run.levels[l].ncells
    g = run.ngrow[l]
    flops = 2.0 * (2 * g + 1) ** 3 * cells
    res["levels"].append({"lev": l, "ngrow": g, "boxes": len(run.levels[l].boxes), "cells": cells, "fill_ms": tf / reps, "filter_ms": tk / reps,
                          "filter_Gcells_s": cells / (tk / reps) / 1e6, "filter_TFLOPs_fp64": flops / (tk / reps) / 1e9})
    tot += (tf + tk) / reps
a, b = ev(), ev()
a.record()
for _ in range(reps):
    run.step()
b.record()
torch.cuda.synchronize()
res["step_ms"] = a.elapsed_time(b) / reps
res["Gcells_s"] = run.cells() / res["step_ms"] / 1e6
print(json.dumps(res))
