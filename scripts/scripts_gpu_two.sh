#!/bin/bash
# 2-GPU validation after the persistent-kernel rewrite: GPU tests, dist check (slab + peer), N=1/N=2 benches, full ncu of the curvature kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $O/t_pytest.log 2>&1; echo "rc=$?" >> $O/t_pytest.log
tail -n 4 $O/t_pytest.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/t_dist.log 2>&1; echo "rc=$?" >> $O/t_dist.log
tail -n 3 $O/t_dist.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/t_bench_n1.log 2>&1; echo "rc=$?" >> $O/t_bench_n1.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/t_bench_n2_peer.log 2>&1; echo "rc=$?" >> $O/t_bench_n2_peer.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --transport slab --e2e-steps 1 > $O/t_bench_n2_slab.log 2>&1; echo "rc=$?" >> $O/t_bench_n2_slab.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma|k_bcfill" -s 8 -c 4 -o $O/t_curv -f \
    python bench.py --only-extra curvature3 --steps 2 --warmup 3 > $O/t_curv.log 2>&1; echo "rc=$?" >> $O/t_curv.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/t_bench_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'N=%d value %.1f ms %.3f frac %.3f e2e %.3f launches %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches']))
            if d.get('extras'):
                for k,v in d['extras'].items(): print('   ',k, {a:v[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in v} or v)
PY
