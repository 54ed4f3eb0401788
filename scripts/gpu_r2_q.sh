#!/bin/bash
# coarse-fine fill with component groups (k_bcfill_v2 reworked): A/B against the morning snapshot; tool wall times after the host-side changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 280 -p no:cacheprovider -k "golden or ghost or midsize" > $O/r2q_pytest.log 2>&1; echo "rc=$?" >> $O/r2q_pytest.log
el pytest; tail -3 $O/r2q_pytest.log
for rep in 1 2; do
for ex in target_grad target_curv curvature3 grad5; do
  timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2q_${ex}_$rep.log 2>&1
  PA_BCFILL_V2=0 timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2q_${ex}_v1_$rep.log 2>&1
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_*_[12].log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'))
PY
el ab
timeout -s KILL 400 python scripts/tool_walltime.py 256 grad,curvature > $O/r2q_tool_walltime.log 2>&1
el walltime; cut -c1-600 $O/r2q_tool_walltime.log
