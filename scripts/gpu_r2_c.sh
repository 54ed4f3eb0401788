#!/bin/bash
# Round 2, third GPU call: fused curvature kernel v2 (warp-to-warp sync, n-row = warp) -- parity, A/B and ablations.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 200 python scripts/gpu_hash_check.py > $O/r2c_hash.log 2>&1; tail -n 2 $O/r2c_hash.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 200 -p no:cacheprovider -k "curvature or midsize or extreme or degenerate or flat or selftest" > $O/r2c_pytest.log 2>&1; echo "rc=$?" >> $O/r2c_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2c_pytest.log | head -20
for cw in 15 19; do
  for abl in 0 1 2 4 8 3 15; do
    PA_CF_CW=$cw PA_CF_ABLATE=$abl timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2c_target_curv_cw${cw}_abl${abl}.log 2>&1
  done
  PA_CF_CW=$cw PA_CF_STAGES=6 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2c_target_curv_cw${cw}_st6.log 2>&1
  PA_CF_CW=$cw PA_CF_ZC=126 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2c_target_curv_cw${cw}_zc126.log 2>&1
  PA_CF_CW=$cw timeout -s KILL 60 python bench.py --only-extra curvature3 --steps 10 --warmup 3 > $O/r2c_curvature3_cw${cw}.log 2>&1
  el cw$cw
done
PA_CURV_FUSED=0 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2c_target_curv_unfused.log 2>&1
for cw in 15 19; do
  PA_CF_CW=$cw timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_fused" -s 2 -c 1 -o $O/r2c_curv_fused_cw$cw -f \
      python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2c_ncu_full_cw$cw.log 2>&1
  el ncu$cw
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {a:(round(d[a],4) if not isinstance(d[a],dict) else d[a].get('value')) for a in ('value','ms_per_step','roofline_frac','launches_per_step','output_hash') if a in d})
PY
