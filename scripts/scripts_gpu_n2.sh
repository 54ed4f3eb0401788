#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_smi.txt
nvidia-smi topo -m >> gpurun_out/n2_smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/n2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/n2_pytest.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/n2_dist.log 2>&1; echo "rc=$?" >> gpurun_out/n2_dist.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.log 2>&1; echo "rc=$?" >> gpurun_out/n2_bench.log
timeout -s KILL 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/n1_bench.log 2>&1; echo "rc=$?" >> gpurun_out/n1_bench.log
tail -n 6 gpurun_out/n2_pytest.log gpurun_out/n2_dist.log gpurun_out/n2_bench.log gpurun_out/n1_bench.log
