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
# the plotfile tools with one process per GPU (first run over NCCL), checked against the single-GPU C++ executable with fcompare
python - <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
from peleanalysis_b200 import plotfile
pf, z = load_golden("c3_three_levels"); plotfile.write_plotfile("gpurun_out/r2b_plt", pf, clean="remove")
PY
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m peleanalysis_b200.mgtools grad infile=gpurun_out/r2b_plt outfile=gpurun_out/r2b_gt_n2 > $O/r2b_mgtools.log 2>&1; echo "rc=$?" >> $O/r2b_mgtools.log
peleanalysis_b200/host/grad3d.b200.ex infile=gpurun_out/r2b_plt outfile=gpurun_out/r2b_gt_n1 >> $O/r2b_mgtools.log 2>&1
oracle/_ref/fcompare.ref.ex gpurun_out/r2b_gt_n2 gpurun_out/r2b_gt_n1 2>&1 | tail -3 | tee -a $O/r2b_mgtools.log
# the C++ executable with one host thread per GPU (first run on hardware)
peleanalysis_b200/host/grad3d.b200.ex infile=gpurun_out/r2b_plt outfile=gpurun_out/r2b_gt_threads ngpus=2 >> $O/r2b_mgtools.log 2>&1; echo "threads rc=$?" >> $O/r2b_mgtools.log
oracle/_ref/fcompare.ref.ex gpurun_out/r2b_gt_threads gpurun_out/r2b_gt_n1 2>&1 | tail -3 | tee -a $O/r2b_mgtools.log
peleanalysis_b200/host/curvature3d.b200.ex infile=gpurun_out/r2b_plt outfile=gpurun_out/r2b_K_threads ngpus=2 threshold_prog=1 threshold_value=0.01 >> $O/r2b_mgtools.log 2>&1; echo "curv threads rc=$?" >> $O/r2b_mgtools.log
peleanalysis_b200/host/curvature3d.b200.ex infile=gpurun_out/r2b_plt outfile=gpurun_out/r2b_K_n1 threshold_prog=1 threshold_value=0.01 >> $O/r2b_mgtools.log 2>&1
oracle/_ref/fcompare.ref.ex gpurun_out/r2b_K_threads gpurun_out/r2b_K_n1 2>&1 | tail -3 | tee -a $O/r2b_mgtools.log
# curvature on 2 GPUs (first multi-GPU curvature timing)
for tr in peer slab; do
  timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --only-extra target_curv --steps 10 --warmup 3 --transport $tr > $O/r2b_curv_n2_$tr.log 2>&1; tail -c 400 $O/r2b_curv_n2_$tr.log
done
