#!/bin/bash
# links / peer-links validation: GPU tests, 2-rank dist check (slab + peer), benches, ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $O/l_pytest.log 2>&1; echo "rc=$?" >> $O/l_pytest.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/l_dist.log 2>&1; echo "rc=$?" >> $O/l_dist.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/l_bench_n1.log 2>&1; echo "rc=$?" >> $O/l_bench_n1.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-links --no-cpu-baseline --e2e-steps 1 > $O/l_bench_n1_nolinks.log 2>&1; echo "rc=$?" >> $O/l_bench_n1_nolinks.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/l_bench_n2_peer.log 2>&1; echo "rc=$?" >> $O/l_bench_n2_peer.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --transport slab --e2e-steps 1 > $O/l_bench_n2_slab.log 2>&1; echo "rc=$?" >> $O/l_bench_n2_slab.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file $O/l_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/l_ncu_bench.log 2>&1; echo "rc=$?" >> $O/l_ncu_bench.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:^k_stencil_tma -s 3 -c 1 -o $O/l_prof_stencil -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/l_ncu_full.log 2>&1; echo "rc=$?" >> $O/l_ncu_full.log
tail -n 5 $O/l_pytest.log $O/l_dist.log; tail -n 2 $O/l_bench_n1.log $O/l_bench_n1_nolinks.log $O/l_bench_n2_peer.log $O/l_bench_n2_slab.log
