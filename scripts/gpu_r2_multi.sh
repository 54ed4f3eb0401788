#!/bin/bash
# N GPUs of one box (gpurun --gpus N -- 'N=<n> bash scripts/gpu_r2_multi.sh'): correctness of every multi-rank path against
# the single-rank oracle (tests/dist_check.py), then the benchmark line of the default run (headline + the extras every N
# carries, each with its output fingerprint), then the A/B of the box -> rank curve.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
N=${N:-2}
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
  timeout -s KILL 400 $TR --master-port 29511 tests/dist_check.py > $O/r2m_dist_check_n$N.log 2>&1; echo "rc=$?" >> $O/r2m_dist_check_n$N.log
  el dist_check; grep -E "DIST_CHECK|rc=|mismatching boxes=[1-9]|Error|error" $O/r2m_dist_check_n$N.log | head -20
fi
timeout -s KILL 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2m_bench_n$N.log 2> $O/r2m_bench_n$N.err; echo "rc=$?" >> $O/r2m_bench_n$N.err
el bench; tail -c 400 $O/r2m_bench_n$N.err
PA_DISTRIBUTE=morton timeout -s KILL 200 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --e2e-steps 1 > $O/r2m_bench_n${N}_morton.log 2>&1
timeout -s KILL 200 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --e2e-steps 1 --transport slab > $O/r2m_bench_n${N}_slab.log 2>&1
el ab
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_bench_n${N}*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value %.1f ms %.3f kernel_ms %.3f hash %s e2e %.2f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['output_hash'], (d.get('e2e') or {}).get('value', 0)))
            for k, x in (d.get('extras') or {}).items():
                print('    ', k, {a: (round(x[a], 4) if isinstance(x[a], float) else x[a]) for a in ('value', 'ms_per_step', 'roofline_frac', 'output_hash') if a in x} if 'error' not in x else x)
PY
