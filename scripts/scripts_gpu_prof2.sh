#!/bin/bash
# full ncu captures: NORMAL_S + DIV (curvature, config 3), bcfill, small-box grad (config 5), linked grad (config 2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 8 -c 4 -o $O/p_curv -f \
    python bench.py --only-extra curvature3 --steps 2 --warmup 3 > $O/p_curv.log 2>&1; echo "rc=$?" >> $O/p_curv.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 4 -c 2 -o $O/p_grad5 -f \
    python bench.py --only-extra grad5 --steps 2 --warmup 3 > $O/p_grad5.log 2>&1; echo "rc=$?" >> $O/p_grad5.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:^k_stencil_tma -s 3 -c 1 -o $O/p_grad2 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/p_grad2.log 2>&1; echo "rc=$?" >> $O/p_grad2.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/p_bench_e2e.log 2>&1; echo "rc=$?" >> $O/p_bench_e2e.log
tail -n 3 $O/p_curv.log $O/p_grad5.log $O/p_grad2.log $O/p_bench_e2e.log
