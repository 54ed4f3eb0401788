#!/bin/bash
# session 6, call 1: validate the restored tree (GPU tests, smoke, both bench arms), collect the launch list of the
# default bench and full ncu captures of the curvature kernels, and sweep the ring depth for the curvature passes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/v_gpu.txt 2>&1
nproc >> $O/v_gpu.txt; free -g >> $O/v_gpu.txt
timeout -s KILL 700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $O/v_pytest.log 2>&1; echo "rc=$?" >> $O/v_pytest.log
grep -E "passed|failed|^FAILED|rc=" $O/v_pytest.log | head -20
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/v_smoke.log 2>&1; echo "rc=$?" >> $O/v_smoke.log; tail -n 2 $O/v_smoke.log
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/v_bench_ref.log 2>&1; echo "rc=$?" >> $O/v_bench_ref.log
timeout -s KILL 600 python bench.py > $O/v_bench_n1.log 2>&1; echo "rc=$?" >> $O/v_bench_n1.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file $O/v_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/v_ncu_bench.log 2>&1; echo "rc=$?" >> $O/v_ncu_bench.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 8 -c 4 -o $O/v_curv -f \
    python bench.py --only-extra curvature3 --steps 2 --warmup 3 > $O/v_ncu_curv.log 2>&1; echo "rc=$?" >> $O/v_ncu_curv.log
ncu -i $O/v_curv.ncu-rep --page raw --csv > $O/v_curv_raw.csv 2>/dev/null
for kb in 12 32 48; do
  PA_TMA_INFLIGHT_KB=$kb timeout -s KILL 200 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/v_tcurv_kb$kb.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/v_bench_*.log'))+sorted(glob.glob('gpurun_out/v_tcurv_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            if 'roofline' in d:
                print(f, 'value %.1f ms %.3f frac %.3f e2e %.3f launches %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('gpu_launches')))
                print('   cpu', d.get('cpu_baseline'))
                for k,v in (d.get('extras') or {}).items(): print('   ',k, {a:v[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in v} or v)
            else:
                print(f, {a:d[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step','unavailable') if a in d})
PY
tail -n 3 $O/v_ncu_curv.log $O/v_ncu_bench.log
