#!/bin/bash
# planes per work item for the curvature modes: PA_TMA_ZC = 32 (default) / 64 / 128, current tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for zc in 32 64 128; do
for ex in target_curv curvature3; do
  PA_TMA_ZC=$zc timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2r_${ex}_zc${zc}_$rep.log 2>&1
done; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'))
PY
