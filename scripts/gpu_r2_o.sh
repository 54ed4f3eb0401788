#!/bin/bash
# filter path after the per-level halo table, tool-level wall times, larger filter parity case
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 400 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tools.py -q -m gpu -k "filter" -n 4 --timeout 380 -p no:cacheprovider > $O/r2o_pytest.log 2>&1; echo "rc=$?" >> $O/r2o_pytest.log
el pytest; tail -3 $O/r2o_pytest.log
timeout -s KILL 200 python bench.py --only-extra filter3 --steps 10 --warmup 3 > $O/r2o_bench_filter3.log 2> $O/r2o_bench_filter3.err; echo "rc=$?" >> $O/r2o_bench_filter3.err
el bench; python - <<'PY'
import json
for line in open('gpurun_out/r2o_bench_filter3.log'):
    if line.startswith('{'):
        d=json.loads(line); print('filter3', d['value'], d['ms_per_step'], d['roofline']['frac'], d['output_hash'])
        for L in d['levels']: print('   ', L)
PY
timeout -s KILL 120 python scripts/filter_time.py 512 128 128 1 2 3 > $O/r2o_time_512_128_box.log 2>&1
timeout -s KILL 600 python scripts/tool_walltime.py 256 > $O/r2o_tool_walltime.log 2>&1
el walltime; cat $O/r2o_tool_walltime.log | cut -c1-700
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2o_launches.csv python scripts/filter_time.py 256 64 32 1 2 1 > /dev/null 2>&1
el launches
