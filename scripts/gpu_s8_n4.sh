#!/bin/bash
# session 8, call 4 (4 GPUs): the default bench at N=4 as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 3 > $O/n4_bench.log 2>&1; echo "rc=$?" >> $O/n4_bench.log
tail -c 1800 $O/n4_bench.log
