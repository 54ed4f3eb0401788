#!/bin/bash
# N GPUs of one box (gpurun --gpus N -- 'N=<n> bash scripts/gpu_r2_scale.sh'): the default bench line with its extras; every
# workload's output fingerprint must equal the single-GPU one (bench.py exits 3 otherwise)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_bench_n$N.log 2> $O/scale_bench_n$N.err; echo "rc=$?" >> $O/scale_bench_n$N.err
tail -c 300 $O/scale_bench_n$N.err
python - <<PY
import json
for line in open('gpurun_out/scale_bench_n$N.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('N=$N value %.1f ms %.4f kernel_ms %.4f frac %.4f hash %s e2e %.2f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['output_hash'], (d.get('e2e') or {}).get('value', 0)))
        for k, x in (d.get('extras') or {}).items():
            print('    ', k, {a: (round(x[a], 4) if isinstance(x[a], float) else x[a]) for a in ('value', 'ms_per_step', 'roofline_frac') if a in x}, x.get('output_hash', {}).get('ok'), x.get('error'))
PY
