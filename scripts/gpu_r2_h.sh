#!/bin/bash
# L2 cache-hint A/B for the TMA pipeline; why the reference's CUDA build did not run; N=1 sanity of peer-less changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for h in 0 1 2 3; do
  for ex in config2 target_grad target_curv grad5 curvature3; do
    PA_TMA_L2HINT=$h timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2h_${ex}_hint$h.log 2>&1
  done
done
el hints
PA_TMA_L2HINT=3 timeout -s KILL 200 ncu --set full --clock-control none -k regex:k_stencil_tma -s 4 -c 1 -o $O/r2h_grad_config2_hint3 -f python bench.py --only-extra config2 --steps 2 --warmup 3 > $O/r2h_ncu.log 2>&1
el ncu
# the reference's own CUDA build on this GPU
python - > $O/r2h_refcuda.log 2>&1 <<'PY'
import os, sys, tempfile, shutil, subprocess
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
from peleanalysis_b200 import plotfile, synth
pf = synth.make_hierarchy(256, [], [], 128, ("temp",))
tmp = tempfile.mkdtemp(prefix="pa_rc_", dir="/dev/shm")
d = os.path.join(tmp, "plt")
plotfile.write_plotfile(d, pf)
exe = O.ref_exe("grad3d.cuda.timed.ex")
env = dict(os.environ, PA_TIMED_REPS="3")
p = subprocess.run([exe, "infile=" + d, "gradVar=temp", "outfile=" + os.path.join(tmp, "o")], capture_output=True, text=True, env=env, cwd=tmp)
print("rc", p.returncode); print(p.stdout[-3000:]); print(p.stderr[-3000:])
shutil.rmtree(tmp, ignore_errors=True)
PY
tail -n 15 $O/r2h_refcuda.log
el refcuda
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_*_hint*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {a:(round(d[a],4) if not isinstance(d[a],dict) else d[a].get('ok')) for a in ('value','ms_per_step','roofline_frac','output_hash') if a in d})
PY
