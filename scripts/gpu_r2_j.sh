#!/bin/bash
# filterPlt path: GPU parity, per-level timing, one ncu --set full capture of the filter kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
echo skip-pytest > $O/r2j_pytest2.log
el pytest; tail -5 $O/r2j_pytest.log
timeout -s KILL 120 python scripts/filter_time.py 256 64 32 1 2 > $O/r2j_time_256_32_box.log 2>&1
timeout -s KILL 120 python scripts/filter_time.py 256 64 64 1 2 > $O/r2j_time_256_64_box.log 2>&1
timeout -s KILL 120 python scripts/filter_time.py 256 64 32 2 4 > $O/r2j_time_256_32_gauss4.log 2>&1
timeout -s KILL 120 python scripts/filter_time.py 512 128 128 1 2 3 > $O/r2j_time_512_128_box.log 2>&1
el timing; tail -n1 $O/r2j_time_*.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_filter -c 3 -o $O/r2j_filter python scripts/filter_time.py 256 64 32 1 2 1 > $O/r2j_ncu.log 2>&1
el ncu; tail -2 $O/r2j_ncu.log
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2j_launches.csv python scripts/filter_time.py 256 64 32 1 2 1 > /dev/null 2>&1
el launches
