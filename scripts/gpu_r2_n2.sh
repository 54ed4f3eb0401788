#!/bin/bash
# Round 2, 2-GPU call (~2 min x 2 GPUs): the distributed checks that ran only under the emulator / gloo in round 1
# (two-pass curvature and every curvature option over NVLink, peer and slab transport), then the N=2 bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/r2b_dist.log 2>&1; echo "rc=$?" >> $O/r2b_dist.log
grep -E "dist_check|DIST_CHECK|rc=" $O/r2b_dist.log | tail -40
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/r2b_bench_n2.log 2>&1; echo "rc=$?" >> $O/r2b_bench_n2.log
tail -c 700 $O/r2b_bench_n2.log
