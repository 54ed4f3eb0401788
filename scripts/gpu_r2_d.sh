#!/bin/bash
# fused curvature v3 (TMA bulk stores of result rows): parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 200 python scripts/gpu_hash_check.py > $O/r2d_hash.log 2>&1; tail -n 2 $O/r2d_hash.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 200 -p no:cacheprovider -k "curvature or midsize or extreme or degenerate or flat or selftest" > $O/r2d_pytest.log 2>&1; echo "rc=$?" >> $O/r2d_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2d_pytest.log | head -20
for cw in 15 19; do
  for abl in 0 16 2 1; do
    PA_CF_CW=$cw PA_CF_ABLATE=$abl timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2d_target_curv_cw${cw}_abl${abl}.log 2>&1
  done
  PA_CF_CW=$cw PA_CF_STAGES=6 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2d_target_curv_cw${cw}_st6.log 2>&1
  PA_CF_CW=$cw timeout -s KILL 60 python bench.py --only-extra curvature3 --steps 10 --warmup 3 > $O/r2d_curvature3_cw${cw}.log 2>&1
  el cw$cw
done
PA_CF_CW=15 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_fused" -s 2 -c 1 -o $O/r2d_curv_fused_cw15 -f \
      python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2d_ncu_full_cw15.log 2>&1
el ncu
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {a:(round(d[a],4) if not isinstance(d[a],dict) else d[a].get('value')) for a in ('value','ms_per_step','roofline_frac','launches_per_step','output_hash') if a in d})
PY
