#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k degenerate --timeout 300 -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest.log
PA_CURV_UNFUSED=1 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k degenerate --timeout 300 -p no:cacheprovider > gpurun_out/d_pytest_unfused.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest_unfused.log
grep -E "AssertionError|passed|failed" gpurun_out/d_pytest.log gpurun_out/d_pytest_unfused.log | cut -c1-900
