#!/bin/bash
# flame-normal-only form of curv_f3.cu in front of MODE_DIV (PA_NORMAL_F3=1): fingerprints, parity tests, timing, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python scripts/gpu_hash_check.py > $O/r2w_hash.log 2>&1; echo "rc=$?" >> $O/r2w_hash.log
el hash; tail -2 $O/r2w_hash.log | cut -c1-200
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 180 -p no:cacheprovider -k "fused3 or n3" > $O/r2w_pytest.log 2>&1; echo "rc=$?" >> $O/r2w_pytest.log
el pytest; tail -2 $O/r2w_pytest.log
for ex in target_curv curvature3; do
  PA_NORMAL_F3=1 timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2w_${ex}_fusedn3.log 2>&1
  PA_NORMAL_F3=1 PA_NF3_CTAS=3 timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2w_${ex}_fusedn3_ctas3.log 2>&1
done
PA_NORMAL_F3=1 PA_NF3_ZC=32 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2w_target_curv_fusedn3_zc32.log 2>&1
PA_NORMAL_F3=1 PA_NF3_ZC=128 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2w_target_curv_fusedn3_zc128.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w_*fused*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
el timing
PA_NORMAL_F3=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_f3" -c 1 -o $O/r2w_normal_f3 python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2w_ncu.log 2>&1
el ncu; tail -1 $O/r2w_ncu.log
