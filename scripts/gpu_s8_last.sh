#!/bin/bash
# session 8: the last GPU call of round 1 (about 8 minutes of command time).  Most important first; every step writes its
# own file under gpurun_out/ so that a cut-off call still leaves what finished.
#  1 smoke (default configuration; fallbacks if it fails)     2 pytest -m gpu, 8 workers
#  3 bench.py default (headline + extras + cpu baseline)      4 ncu launch lists (bench, curvature)
#  5 A/B of the flame-normal shapes / arithmetic              6 ncu --set full of the curvature kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/y_gpu.txt 2>&1
nproc >> $O/y_gpu.txt; free -g >> $O/y_gpu.txt

timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/y_smoke.log 2>&1; echo "rc=$?" >> $O/y_smoke.log
el smoke; tail -n 3 $O/y_smoke.log
ENVFIX=""
if ! grep -q "rc=0" $O/y_smoke.log; then
  for fix in "PA_TMA_CW16=0" "PA_TMA_CW16=0 PA_NORMAL_MATH=plain"; do
    env $fix timeout -s KILL 90 python -c "import __graft_entry__ as g; g.smoke()" > $O/y_smoke_fix.log 2>&1; rc=$?
    echo "fix '$fix' rc=$rc" >> $O/y_smoke.log
    if [ $rc -eq 0 ]; then ENVFIX="$fix"; break; fi
  done
  el "smoke fallback: ENVFIX='$ENVFIX'"
fi
echo "ENVFIX='$ENVFIX'" > $O/y_envfix.txt

env $ENVFIX timeout -s KILL 240 python -m pytest tests -q -m gpu -n 8 --timeout 150 --timeout-method=thread -p no:cacheprovider > $O/y_pytest_par.log 2>&1; echo "rc=$?" >> $O/y_pytest_par.log
el pytest-par; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/y_pytest_par.log | head -30

env $ENVFIX timeout -s KILL 300 python bench.py > $O/y_bench_n1.log 2>&1; echo "rc=$?" >> $O/y_bench_n1.log
el bench; tail -c 400 $O/y_bench_n1.log

summ() {
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/y_bench_*.log'))+sorted(glob.glob('gpurun_out/y_ab_*.log')):
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
}
summ

env $ENVFIX timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 120 --csv --log-file $O/y_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/y_ncu_bench.log 2>&1; echo "rc=$?" >> $O/y_ncu_bench.log
el launchlist-bench
env $ENVFIX timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 120 --csv --log-file $O/y_launches_curv.csv \
    python bench.py --only-extra target_curv --steps 3 --warmup 3 > $O/y_ncu_curvl.log 2>&1; echo "rc=$?" >> $O/y_ncu_curvl.log
el launchlist-curv

for ex in target_curv curvature3; do
  env $ENVFIX PA_TMA_CW16=0 timeout -s KILL 100 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/y_ab_${ex}_cw8.log 2>&1
  env $ENVFIX PA_NORMAL_MATH=plain timeout -s KILL 100 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/y_ab_${ex}_plain.log 2>&1
  env $ENVFIX PA_TMA_CW16=0 PA_NORMAL_MATH=plain timeout -s KILL 100 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/y_ab_${ex}_cw8plain.log 2>&1
done
el ab
summ

env $ENVFIX timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 8 -c 4 -o $O/y_curv -f \
    python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/y_ncu_curv.log 2>&1; echo "rc=$?" >> $O/y_ncu_curv.log
ncu -i $O/y_curv.ncu-rep --page raw --csv > $O/y_curv_raw.csv 2>/dev/null
el ncufull
