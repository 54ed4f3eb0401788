#!/bin/bash
# single-GPU check: GPU tests, bench (headline + extras), ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $O/o_pytest.log 2>&1; echo "rc=$?" >> $O/o_pytest.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/o_bench_n1.log 2>&1; echo "rc=$?" >> $O/o_bench_n1.log
PA_CURV_UNFUSED=1 timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/o_bench_unfused.log 2>&1; echo "rc=$?" >> $O/o_bench_unfused.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file $O/o_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/o_ncu_bench.log 2>&1; echo "rc=$?" >> $O/o_ncu_bench.log
tail -n 5 $O/o_pytest.log; tail -n 2 $O/o_bench_n1.log $O/o_bench_unfused.log
