#!/bin/bash
# final single-GPU validation of the tree: all GPU tests, smoke, default bench, reference arm, tool wall times, launch list + ncu of the headline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/fin_gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu -n 8 --timeout 600 -p no:cacheprovider > $O/fin_pytest.log 2>&1; echo "rc=$?" >> $O/fin_pytest.log
el pytest; tail -4 $O/fin_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/fin_smoke.log 2>&1; echo "rc=$?" >> $O/fin_smoke.log
el smoke; tail -2 $O/fin_smoke.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > $O/fin_bench.log 2> $O/fin_bench.err; echo "rc=$?" >> $O/fin_bench.err
el bench; tail -c 300 $O/fin_bench.err
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 3 > $O/fin_bench_ref.log 2> $O/fin_bench_ref.err; echo "rc=$?" >> $O/fin_bench_ref.err
el refarm
timeout -s KILL 400 python scripts/tool_walltime.py 256 grad,curvature > $O/fin_tool_walltime.log 2>&1
el walltime; cut -c1-500 $O/fin_tool_walltime.log
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/fin_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/fin_ncu_bench.log 2>&1
el launches
python - <<'PY'
import json
for line in open('gpurun_out/fin_bench.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],2),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],3),'cpu',round(d['cpu_baseline']['value'],4),'hash',d['output_hash']['ok'], 'clocks', d['clocks'])
        for k,v in d['extras'].items():
            if 'error' in v: print('  ',k,v); continue
            if k=='filter3': print('   filter3', round(v['value'],2), round(v['ms_per_step'],4), round(v['roofline']['frac'],4), v['output_hash']['ok'], v['cpu_baseline'].get('value'))
            else: print('  ',k, round(v['value'],2), round(v['ms_per_step'],4), round(v['roofline_frac'],4), v['output_hash']['ok'], v['launches_per_step'])
PY
