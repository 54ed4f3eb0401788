#!/bin/bash
# whole bench flow on one GPU (new bench.py), fingerprints recorded, full GPU test suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --write-hashes > $O/r2f_bench.log 2> $O/r2f_bench.err; echo "rc=$?" >> $O/r2f_bench.err
cp tests/golden/bench_hashes.json $O/r2f_bench_hashes.json 2>/dev/null
el bench; tail -c 600 $O/r2f_bench.err
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/r2f_bench_ref.log 2> $O/r2f_bench_ref.err; echo "rc=$?" >> $O/r2f_bench_ref.err
el refarm; tail -c 300 $O/r2f_bench_ref.err
timeout -s KILL 900 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider --durations=15 > $O/r2f_pytest.log 2>&1; echo "rc=$?" >> $O/r2f_pytest.log
el pytest; tail -n 30 $O/r2f_pytest.log
