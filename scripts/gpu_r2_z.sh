#!/bin/bash
# barrier-free flame-normal kernel (normal_w.cu, PA_NORMAL_W=1): fingerprints, parity tests, timing, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python scripts/gpu_hash_check.py > $O/r2z_hash.log 2>&1; echo "rc=$?" >> $O/r2z_hash.log
el hash; tail -2 $O/r2z_hash.log | cut -c1-200
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 180 -p no:cacheprovider -k "nw or fused3_strips" > $O/r2z_pytest.log 2>&1; echo "rc=$?" >> $O/r2z_pytest.log
el pytest; tail -2 $O/r2z_pytest.log
for c in 3 2 4; do
  PA_NORMAL_W=1 PA_NW_CTAS=$c timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2z_target_curv_fusednw_ctas$c.log 2>&1
done
for zc in 16 64 128; do
  PA_NORMAL_W=1 PA_NW_ZC=$zc timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2z_target_curv_fusednw_zc$zc.log 2>&1
done
PA_NORMAL_W=1 timeout -s KILL 90 python bench.py --only-extra curvature3 --steps 20 --warmup 5 > $O/r2z_curvature3_fusednw.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_*fused*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
el timing
PA_NORMAL_W=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_normal_w" -c 1 -o $O/r2z_normal_w python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2z_ncu.log 2>&1
el ncu; tail -1 $O/r2z_ncu.log
