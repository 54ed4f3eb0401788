#!/bin/bash
# rows aligned to 32 bytes (PA_ROW_ALIGN default): one timing of the north-star curvature workload (the last seconds of the round's GPU budget)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 25 python bench.py --only-extra target_curv --steps 10 --warmup 3 > gpurun_out/r2al_target_curv_align32.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/r2al_target_curv_align32.log'):
    if line.startswith('{'):
        d=json.loads(line); print('align32 target_curv', round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
