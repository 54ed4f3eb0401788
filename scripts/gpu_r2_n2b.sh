#!/bin/bash
# two GPUs of one box: the C++ ngpus=2 thread host and mgtools.py (one process per GPU) on real hardware, then a short N=2 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 300 python -m pytest tests/test_gpu_tools.py -q -m gpu -k "two_gpus" --timeout 280 -p no:cacheprovider > $O/r2n_pytest_two_gpus.log 2>&1; echo "rc=$?" >> $O/r2n_pytest_two_gpus.log
el pytest; tail -3 $O/r2n_pytest_two_gpus.log
python - > $O/r2n_mgtools.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from helpers import load_golden
from oracle import oracle as O
from peleanalysis_b200 import plotfile
tmp = tempfile.mkdtemp(prefix="mgtools_")
ok = True
for name, tool, extra in (("c3_three_levels", "grad", ["gradVar=temp"]), ("mixed_boxes", "grad", ["gradVar=temp"]),
                          ("c3_threshold", "curvature", ["progressName=temp"]), ("c1_options", "curvature", ["progressName=temp"])):
    pf, z = load_golden(name)
    d = os.path.join(tmp, "plt_" + name + "_" + tool)
    plotfile.write_plotfile(d, pf, clean="remove")
    per = " ".join(str(int(v)) for v in z["is_per"]); sym = " ".join(str(int(v)) for v in z["sym_dir"])
    kw = [str(s) for s in z["curv_opts"]] if tool == "curvature" and "curv_opts" in z.files else []
    out = d + "_out"
    for transport in ("peer", "slab"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29531",
               "-m", "peleanalysis_b200.mgtools", tool, "infile=" + d, "outfile=" + out, "is_per=" + per, "sym_dir=" + sym, "transport=" + transport, *extra, *kw]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            ok = False; print("FAILED", name, tool, transport, p.stdout[-800:], p.stderr[-1500:]); continue
        ref = d + "_ref"
        kv = dict(s.split("=", 1) for s in extra + kw)
        O.run_ref(tool, d, ref, is_per=list(z["is_per"]), sym_dir=list(z["sym_dir"]), **kv)
        q = subprocess.run([O.ref_exe("fcompare.ref.ex"), out, ref], capture_output=True, text=True)
        agree = "PLOTFILE AGREE" in q.stdout
        # the reference never writes SmoothedProgress / (without do_gaussCurv) GaussianCurvature: uninitialised there
        if not agree:
            bad = [ln for ln in q.stdout.splitlines() if ln.strip() and ln.split()[0] not in ("variable", "level", "SmoothedProgress", "GaussianCurvature_temp", "Level") and len(ln.split()) >= 3 and ln.split()[-1] not in ("0", "0.0")]
            agree = all(("SmoothedProgress" in b or "GaussianCurvature" in b or "----" in b or "name" in b) for b in bad)
        print("mgtools", name, tool, transport, "fcompare:", "AGREE" if agree else "DIFFERS")
        if not agree:
            ok = False; print(q.stdout[-1500:])
print("MGTOOLS OK" if ok else "MGTOOLS FAILED")
PY
el mgtools; grep -E "mgtools|MGTOOLS|FAILED" $O/r2n_mgtools.log | head -12
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 1 > $O/r2n_bench_n2.log 2> $O/r2n_bench_n2.err; echo "rc=$?" >> $O/r2n_bench_n2.err
el bench; tail -c 300 $O/r2n_bench_n2.err
python - <<'PY'
import json
for line in open('gpurun_out/r2n_bench_n2.log'):
    if line.startswith('{'):
        d=json.loads(line); print('N=2 value %.1f ms %.4f frac %.4f hash %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['output_hash']))
        for k,x in (d.get('extras') or {}).items(): print('   ',k,{a:(round(x[a],4) if isinstance(x[a],float) else x[a]) for a in ('value','ms_per_step','roofline_frac') if a in x}, x.get('output_hash',{}).get('ok'), x.get('error'))
PY
