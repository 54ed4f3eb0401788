#!/bin/bash
# session 7: the one remaining GPU call of round 1 (about a quarter of an hour of box time).  Most important first; every
# step writes its own file under gpurun_out/ so that a cut-off call still leaves what finished.
#  1 smoke                                   (does the default configuration run at all; which flame-normal arithmetic)
#  2 pytest -m gpu, 4 workers                (full parity picture quickly)
#  3 bench.py default                        (headline line + extras + cpu baseline)
#  4 A/B: 8-consumer-warp shape for the flame-normal modes on the two curvature extras
#  5 ncu launch list of the default bench + curvature extra
#  6 ncu --set full of the curvature kernels
#  7 pytest -m gpu serial, as the driver runs it (whatever time is left)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/x_gpu.txt 2>&1
nproc >> $O/x_gpu.txt; free -g >> $O/x_gpu.txt

timeout -s KILL 240 python -c "import __graft_entry__ as g; g.smoke()" > $O/x_smoke.log 2>&1; echo "rc=$?" >> $O/x_smoke.log
el smoke; tail -n 3 $O/x_smoke.log
ENVFIX=""
if ! grep -q "rc=0" $O/x_smoke.log; then
  # the default configuration failed: try the 8-warp shape, then the plain operators too, and carry on with what works
  for fix in "PA_TMA_CW16=0" "PA_TMA_CW16=0 PA_NORMAL_MATH=plain"; do
    env $fix timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/x_smoke_fix.log 2>&1; rc=$?
    echo "fix '$fix' rc=$rc" >> $O/x_smoke.log
    if [ $rc -eq 0 ]; then ENVFIX="$fix"; break; fi
  done
  el "smoke fallback: ENVFIX='$ENVFIX'"
fi
echo "ENVFIX='$ENVFIX'" > $O/x_envfix.txt

env $ENVFIX timeout -s KILL 420 python -m pytest tests -q -m gpu -n 4 --timeout 200 --timeout-method=thread -p no:cacheprovider > $O/x_pytest_par.log 2>&1; echo "rc=$?" >> $O/x_pytest_par.log
el pytest-par; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/x_pytest_par.log | head -30

env $ENVFIX timeout -s KILL 420 python bench.py > $O/x_bench_n1.log 2>&1; echo "rc=$?" >> $O/x_bench_n1.log
el bench; tail -c 600 $O/x_bench_n1.log

for ex in curvature3 target_curv; do
  env $ENVFIX PA_TMA_CW16=0 timeout -s KILL 150 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/x_${ex}_cw8.log 2>&1
  env $ENVFIX PA_NORMAL_MATH=plain timeout -s KILL 150 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/x_${ex}_plain.log 2>&1
done
el ab

env $ENVFIX timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file $O/x_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/x_ncu_bench.log 2>&1; echo "rc=$?" >> $O/x_ncu_bench.log
env $ENVFIX timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file $O/x_launches_curv.csv \
    python bench.py --only-extra target_curv --steps 3 --warmup 3 > $O/x_ncu_curvl.log 2>&1; echo "rc=$?" >> $O/x_ncu_curvl.log
el launchlists

env $ENVFIX timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 8 -c 4 -o $O/x_curv -f \
    python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/x_ncu_curv.log 2>&1; echo "rc=$?" >> $O/x_ncu_curv.log
ncu -i $O/x_curv.ncu-rep --page raw --csv > $O/x_curv_raw.csv 2>/dev/null
el ncufull

python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/x_bench_*.log'))+sorted(glob.glob('gpurun_out/x_*_cw8.log'))+sorted(glob.glob('gpurun_out/x_*_plain.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            if 'roofline' in d:
                print(f, 'value %.1f ms %.3f frac %.3f e2e %.3f launches %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('gpu_launches')))
                print('   cpu', d.get('cpu_baseline')); print('   clocks', d.get('clocks'))
                for k,v in (d.get('extras') or {}).items(): print('   ',k, {a:v[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in v} or v)
            else:
                print(f, {a:d[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step','unavailable') if a in d})
PY

env $ENVFIX timeout -s KILL 600 python -m pytest tests -x -q -m gpu --timeout 200 --timeout-method=thread -p no:cacheprovider > $O/x_pytest_serial.log 2>&1; echo "rc=$?" >> $O/x_pytest_serial.log
el pytest-serial; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/x_pytest_serial.log | head -10
