#!/bin/bash
# second fused curvature kernel, iteration 2 (no integer division): fingerprints, timing, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python scripts/gpu_hash_check.py > $O/r2t_hash.log 2>&1; echo "rc=$?" >> $O/r2t_hash.log
el hash; tail -2 $O/r2t_hash.log | cut -c1-200
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 180 -p no:cacheprovider -k "fused2" > $O/r2t_pytest.log 2>&1; echo "rc=$?" >> $O/r2t_pytest.log
el pytest; tail -2 $O/r2t_pytest.log
for rep in 1 2; do
  for ex in target_curv curvature3; do
    PA_CURV_FUSED=2 timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2t_${ex}_fused2_$rep.log 2>&1
  done
done
PA_CURV_FUSED=0 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2t_target_curv_fused0_1.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_*fused*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
el timing
PA_CURV_FUSED=2 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_curv_f2 -c 1 -o $O/r2t_curv_f2 python bench.py --only-extra target_curv --steps 1 --warmup 3 > $O/r2t_ncu.log 2>&1
el ncu; tail -1 $O/r2t_ncu.log
