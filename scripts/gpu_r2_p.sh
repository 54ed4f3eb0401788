#!/bin/bash
# block-per-tag same-level copy in the filterPlt fill; phase breakdown of the grad tool's wall time
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 400 python -m pytest tests/test_gpu_filter.py -q -m gpu -n 4 --timeout 380 -p no:cacheprovider > $O/r2p_pytest.log 2>&1; echo "rc=$?" >> $O/r2p_pytest.log
el pytest; tail -3 $O/r2p_pytest.log
timeout -s KILL 200 python bench.py --only-extra filter3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2p_bench_filter3.log 2> $O/r2p_bench_filter3.err; echo "rc=$?" >> $O/r2p_bench_filter3.err
python - <<'PY'
import json
for line in open('gpurun_out/r2p_bench_filter3.log'):
    if line.startswith('{'):
        d=json.loads(line); print('filter3', d['value'], d['ms_per_step'], d['roofline']['frac'], d['output_hash'])
        for L in d['levels']: print('   ', L)
PY
timeout -s KILL 120 python scripts/filter_time.py 512 128 128 1 2 3 > $O/r2p_time_512_128_box.log 2>&1; tail -n1 $O/r2p_time_512_128_box.log | cut -c1-900
el filter
python - > $O/r2p_grad_phases.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, shutil, time
sys.path.insert(0, os.getcwd())
from peleanalysis_b200 import synth, plotfile
from oracle import oracle as O
subprocess.check_call(["make", "-s", "-C", "peleanalysis_b200/host"])
tmp = tempfile.mkdtemp(prefix="pa_ph_", dir="/dev/shm")
d = os.path.join(tmp, "plt")
plotfile.write_plotfile(d, synth.config3(256, 64, names=("temp", "x_velocity", "y_velocity", "z_velocity")), clean="remove")
env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))
for rep in range(2):
    for name, exe in (("b200", os.path.abspath("peleanalysis_b200/host/grad3d.b200.ex")), ("ref", O.ref_exe("grad3d.ref.ex"))):
        t0 = time.time()
        p = subprocess.run([exe, "infile=" + d, "gradVar=temp", "outfile=" + os.path.join(tmp, "o_" + name), "verbose=1"], capture_output=True, text=True, cwd=tmp, env=env)
        print(name, "wall %.3f s rc %d" % (time.time() - t0, p.returncode))
        for ln in p.stdout.splitlines():
            if ln.startswith("[b200]"): print("   ", ln)
        shutil.rmtree(os.path.join(tmp, "o_" + name), ignore_errors=True)
# raw costs on this box: CUDA context, pinned allocation, /dev/shm read and write
import ctypes, numpy as np
t0 = time.time(); import torch; torch.cuda.init(); torch.zeros(1, device="cuda"); print("import torch + context %.3f s" % (time.time() - t0))
t0 = time.time(); x = torch.empty(2 * 1024**3 // 8, dtype=torch.float64, pin_memory=True); print("pin 2 GB %.3f s" % (time.time() - t0))
a = np.ones(400 * 1024**2 // 8)
t0 = time.time(); a.tofile(os.path.join(tmp, "w.bin")); print("write 400 MB to /dev/shm %.3f s" % (time.time() - t0))
t0 = time.time(); b = np.fromfile(os.path.join(tmp, "w.bin")); print("read 400 MB from /dev/shm %.3f s" % (time.time() - t0))
shutil.rmtree(tmp, ignore_errors=True)
PY
el phases; cat $O/r2p_grad_phases.log
