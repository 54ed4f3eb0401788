#!/bin/bash
# adaptive work-item depth on the per-rank load of N=8 strong scaling (config2_small = 8 boxes x 5 variables on one GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for z in 0 1 2 auto; do
  if [ "$z" = auto ]; then unset PA_TMA_ZDIV; else export PA_TMA_ZDIV=$z; fi
  timeout -s KILL 60 python bench.py --only-extra config2_small --steps 30 --warmup 5 > $O/r2m_zdiv_config2_small_${z}_$rep.log 2>&1
  timeout -s KILL 60 python bench.py --only-extra curvature3_small --steps 30 --warmup 5 > $O/r2m_zdiv_curvature3_small_${z}_$rep.log 2>&1
done; done
unset PA_TMA_ZDIV
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_zdiv_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'))
PY
