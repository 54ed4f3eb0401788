#!/bin/bash
# validation snapshot: all GPU tests, smoke, default bench (+ filter3 extra), reference arm, reference-CUDA diagnostic
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2l_gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu -n 8 --timeout 600 -p no:cacheprovider > $O/r2l_pytest.log 2>&1; echo "rc=$?" >> $O/r2l_pytest.log
el pytest; tail -4 $O/r2l_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/r2l_smoke.log 2>&1; echo "rc=$?" >> $O/r2l_smoke.log
el smoke; tail -2 $O/r2l_smoke.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > $O/r2l_bench.log 2> $O/r2l_bench.err; echo "rc=$?" >> $O/r2l_bench.err
el bench; tail -c 400 $O/r2l_bench.err
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 3 > $O/r2l_bench_ref.log 2> $O/r2l_bench_ref.err; echo "rc=$?" >> $O/r2l_bench_ref.err
el refarm; cut -c1-300 $O/r2l_bench_ref.log
# the reference's own CUDA build: why does bench.py report ref_gpu_baseline = null?
python - > $O/r2l_refcuda.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, shutil
sys.path.insert(0, os.getcwd())
from peleanalysis_b200 import synth, plotfile
exe = os.path.join("oracle", "_ref", "grad3d.cuda.timed.ex")
print("exists", os.path.exists(exe))
tmp = tempfile.mkdtemp(prefix="refcuda_")
d = os.path.join(tmp, "plt")
plotfile.write_plotfile(d, synth.config1(64, 32), clean="remove")
for extra in ([], ["amrex.the_arena_is_managed=1"], ["amrex.the_arena_is_managed=1", "amrex.use_gpu_aware_mpi=0", "amrex.abort_on_out_of_gpu_memory=1"]):
    p = subprocess.run([os.path.abspath(exe), "infile=" + d, "outfile=" + d + "_gt", "gradVar=temp", *extra], capture_output=True, text=True, cwd=tmp)
    print("== args", extra, "rc", p.returncode)
    print(p.stdout[-1500:]); print(p.stderr[-1500:])
    bt = os.path.join(tmp, "Backtrace.0")
    if os.path.exists(bt):
        print(open(bt).read()[:3000]); os.remove(bt)
shutil.rmtree(tmp, ignore_errors=True)
PY
el refcuda; tail -5 $O/r2l_refcuda.log
