#!/bin/bash
# second fused curvature kernel on hardware: fingerprints against the separate kernels, golden parity, timing, one ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python scripts/gpu_hash_check.py > $O/r2s_hash.log 2>&1; echo "rc=$?" >> $O/r2s_hash.log
el hash; tail -4 $O/r2s_hash.log | cut -c1-300
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 280 -p no:cacheprovider -k "curvature" > $O/r2s_pytest.log 2>&1; echo "rc=$?" >> $O/r2s_pytest.log
el pytest; tail -3 $O/r2s_pytest.log
for rep in 1 2; do
for f in 0 2 1; do
  for ex in target_curv curvature3; do
    PA_CURV_FUSED=$f timeout -s KILL 90 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/r2s_${ex}_fused${f}_$rep.log 2>&1
  done
done; done
for zc in 31 126; do
  PA_CF2_ZC=$zc PA_CURV_FUSED=2 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2s_target_curv_fused2_zc$zc.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_*fused*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
el timing
PA_CURV_FUSED=2 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_curv_f2 -c 1 -o $O/r2s_curv_f2 python bench.py --only-extra target_curv --steps 1 --warmup 3 > $O/r2s_ncu.log 2>&1
el ncu; tail -2 $O/r2s_ncu.log
