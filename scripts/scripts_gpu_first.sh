#!/bin/bash
# first GPU run: smoke (simple + TMA), memcheck, parity tests, short bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== smoke simple" > gpurun_out/smoke.log
PA_STENCIL=simple timeout -s KILL 300 python __graft_entry__.py --smoke >> gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
echo "== smoke tma" >> gpurun_out/smoke.log
timeout -s KILL 200 python __graft_entry__.py --smoke >> gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
echo "== memcheck" > gpurun_out/memcheck.log
timeout -s KILL 400 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke >> gpurun_out/memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck.log
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
PA_STENCIL=simple timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_simple.log 2>&1; echo "rc=$?" >> gpurun_out/bench_simple.log
tail -3 gpurun_out/smoke.log gpurun_out/pytest.log gpurun_out/bench.log gpurun_out/bench_simple.log
