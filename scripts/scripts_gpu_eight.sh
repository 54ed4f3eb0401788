#!/bin/bash
# 8-GPU validation: dist check at 8 ranks (slab + peer), strong-scaling benches N=8,4 (peer), N=8 slab
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/e_topo.txt 2>&1
run() { # nproc port out args...
  n=$1; port=$2; out=$3; shift 3
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@" > $out 2>&1; echo "rc=$?" >> $out
}
run 8 29511 $O/e_dist8.log tests/dist_check.py
tail -n 3 $O/e_dist8.log
run 8 29512 $O/e_bench_n8_peer.log bench.py --gpus 8 --steps 20 --warmup 3
run 4 29513 $O/e_bench_n4_peer.log bench.py --gpus 4 --steps 20 --warmup 3
run 8 29514 $O/e_bench_n8_slab.log bench.py --gpus 8 --steps 20 --warmup 3 --transport slab --e2e-steps 1
run 8 29515 $O/e_bench_n8_c4.log bench.py --gpus 8 --steps 20 --warmup 3 --workload config4 --e2e-steps 1
run 2 29516 $O/e_bench_n2_c4.log bench.py --gpus 2 --steps 20 --warmup 3 --workload config4 --e2e-steps 1
timeout -s KILL 300 python bench.py --gpus 1 --steps 10 --warmup 3 --workload config4 --e2e-steps 1 --no-extras --no-cpu-baseline > $O/e_bench_n1_c4.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e_bench_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); ok=True
            print(f, 'N=%d value %.1f ms %.3f kernel_ms %.3f frac %.3f e2e %.3f launches %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches']))
    if not ok: print(f,'NO JSON'); print(open(f).read()[-1500:])
PY
