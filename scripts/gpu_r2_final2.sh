#!/bin/bash
# final single-GPU validation of the tree after the later curvature kernels: all GPU tests, smoke, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 600 python -m pytest tests -q -m gpu -n 8 --timeout 400 -p no:cacheprovider > $O/fin2_pytest.log 2>&1; echo "rc=$?" >> $O/fin2_pytest.log
el pytest; tail -4 $O/fin2_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/fin2_smoke.log 2>&1; echo "rc=$?" >> $O/fin2_smoke.log
el smoke; tail -2 $O/fin2_smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $O/fin2_bench.log 2> $O/fin2_bench.err; echo "rc=$?" >> $O/fin2_bench.err
el bench; tail -c 300 $O/fin2_bench.err
python - <<'PY'
import json
for line in open('gpurun_out/fin2_bench.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],2),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],3),'cpu',round(d['cpu_baseline']['value'],4),'hash',d['output_hash']['ok'], 'clocks', d['clocks'])
        for k,v in d['extras'].items():
            if 'error' in v: print('  ',k,v); continue
            if k=='filter3': print('   filter3', round(v['value'],2), round(v['ms_per_step'],4), round(v['roofline']['frac'],4), v['output_hash']['ok'])
            else: print('  ',k, round(v['value'],2), round(v['ms_per_step'],4), round(v['roofline_frac'],4), v['output_hash']['ok'], v['launches_per_step'])
PY
