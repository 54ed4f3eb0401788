#!/bin/bash
# profiles of the kernels that ship (ncu --set full, launch lists), smoke(), the default bench line incl. ref_gpu_baseline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; echo "rc=$?" >> $O/r2g_smoke.log; tail -n 2 $O/r2g_smoke.log
el smoke
timeout -s KILL 700 python bench.py --steps 20 --warmup 5 > $O/r2g_bench.log 2> $O/r2g_bench.err; echo "rc=$?" >> $O/r2g_bench.err
el bench; tail -c 300 $O/r2g_bench.err
# launch lists (cold-cache, serialised: shares only)
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_stencil|k_curv|k_div|k_bcfill|k_halo|k_clip" -c 40 --csv --log-file $O/r2g_launches_bench_config2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/r2g_ncu_l1.log 2>&1
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_stencil|k_curv|k_div|k_bcfill|k_halo|k_clip" -c 40 --csv --log-file $O/r2g_launches_target_curv.csv python bench.py --only-extra target_curv --steps 3 --warmup 3 > $O/r2g_ncu_l2.log 2>&1
PA_CURV_FUSED=1 timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_stencil|k_curv|k_div|k_bcfill|k_halo|k_clip" -c 40 --csv --log-file $O/r2g_launches_target_curv_fused.csv python bench.py --only-extra target_curv --steps 3 --warmup 3 > $O/r2g_ncu_l3.log 2>&1
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_stencil|k_curv|k_div|k_bcfill|k_halo|k_clip" -c 40 --csv --log-file $O/r2g_launches_target_grad.csv python bench.py --only-extra target_grad --steps 3 --warmup 3 > $O/r2g_ncu_l4.log 2>&1
el launchlists
# full captures
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_stencil_tma -s 4 -c 1 -o $O/r2g_grad_config2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > $O/r2g_ncu_f1.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 12 -c 4 -o $O/r2g_curv_default -f python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2g_ncu_f2.log 2>&1
PA_CURV_FUSED=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_fused|k_div_shell" -s 6 -c 2 -o $O/r2g_curv_fused -f python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2g_ncu_f3.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 6 -c 2 -o $O/r2g_grad_target -f python bench.py --only-extra target_grad --steps 2 --warmup 3 > $O/r2g_ncu_f4.log 2>&1
el ncu
