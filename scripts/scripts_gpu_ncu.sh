#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# launch list (cold-cache, serialised: shares only)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_bench.log
# full capture of the stencil kernel and the two ghost-fill kernels
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:^k_stencil_tma -s 3 -c 2 -o gpurun_out/prof_stencil -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:^k_(halo|bcfill)" -s 6 -c 2 -o gpurun_out/prof_ghost -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_ghost.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_ghost.log
ls -la gpurun_out
