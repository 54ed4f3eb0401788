#!/bin/bash
# filter kernel v2 (4x2 blocks): parity on the GPU incl. the tool tests, timing, FP64 rate, bench extra, ncu of the new kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 400 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tools.py -q -m gpu -k "filter" --timeout 380 -p no:cacheprovider > $O/r2k_pytest.log 2>&1; echo "rc=$?" >> $O/r2k_pytest.log
el pytest; tail -4 $O/r2k_pytest.log
timeout -s KILL 120 python scripts/filter_time.py 256 64 32 1 2 > $O/r2k_time_256_32_box.log 2>&1
timeout -s KILL 120 python scripts/filter_time.py 256 64 32 2 4 > $O/r2k_time_256_32_gauss4.log 2>&1
timeout -s KILL 120 python scripts/filter_time.py 512 128 128 1 2 3 > $O/r2k_time_512_128_box.log 2>&1
el timing; tail -qn1 $O/r2k_time_*.log | cut -c1-1500
timeout -s KILL 300 python bench.py --only-extra filter3 --steps 10 --warmup 3 > $O/r2k_bench_filter3.log 2> $O/r2k_bench_filter3.err; echo "rc=$?" >> $O/r2k_bench_filter3.err
el bench; tail -c 2500 $O/r2k_bench_filter3.log; tail -3 $O/r2k_bench_filter3.err
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_filter -c 3 -o $O/r2k_filter python scripts/filter_time.py 256 64 32 1 2 1 > $O/r2k_ncu.log 2>&1
el ncu; tail -2 $O/r2k_ncu.log
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2k_launches.csv python scripts/filter_time.py 256 64 32 1 2 1 > /dev/null 2>&1
el launches
