#!/bin/bash
# session 8, last call of round 1: parity of the final tree (side-stream overlap, class-major tiles), final default bench,
# A/B of the overlap on the multi-level extras
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python -m pytest tests -q -m gpu -n 8 --timeout 100 --timeout-method=thread -p no:cacheprovider > $O/f_pytest.log 2>&1; echo "rc=$?" >> $O/f_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/f_pytest.log | head -20
timeout -s KILL 120 python bench.py > $O/f_bench_n1.log 2>&1; echo "rc=$?" >> $O/f_bench_n1.log
el bench
for ex in target_curv grad5 target_grad curvature3; do
  PA_STREAM_OVERLAP=0 timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/f_ab_${ex}_nooverlap.log 2>&1
  el $ex
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_bench_*.log'))+sorted(glob.glob('gpurun_out/f_ab_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            if 'roofline' in d:
                print(f, 'value %.1f ms %.3f frac %.3f e2e %.3f launches %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('gpu_launches')))
                print('   cpu', d.get('cpu_baseline')); print('   clocks', d.get('clocks'))
                for k,v in (d.get('extras') or {}).items(): print('   ',k, {a:v[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in v} or v)
            else:
                print(f, {a:d[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in d})
PY
