#!/bin/bash
# fused curvature v4 (reader-normalised, reordered step, producer back-off): parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 90 python scripts/gpu_hash_check.py > $O/r2e_hash.log 2>&1; tail -n 2 $O/r2e_hash.log
PA_CF_BULK=1 timeout -s KILL 90 python scripts/gpu_hash_check.py > $O/r2e_hash_bulk.log 2>&1; tail -n 1 $O/r2e_hash_bulk.log
timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 200 -p no:cacheprovider -k "curvature or midsize or extreme or degenerate or flat or selftest" > $O/r2e_pytest.log 2>&1; echo "rc=$?" >> $O/r2e_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2e_pytest.log | head -20
for ps in 0 100 200 500; do
  for bulk in 0 1; do
    PA_CF_CW=15 PA_CF_PSLEEP=$ps PA_CF_BULK=$bulk timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2e_target_curv_ps${ps}_bulk${bulk}.log 2>&1
  done
done
for abl in 1 2 4 8 15; do
  PA_CF_CW=15 PA_CF_ABLATE=$abl timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2e_target_curv_abl${abl}.log 2>&1
done
PA_CF_CW=19 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2e_target_curv_cw19.log 2>&1
PA_CF_CW=15 timeout -s KILL 60 python bench.py --only-extra curvature3 --steps 10 --warmup 3 > $O/r2e_curvature3_cw15.log 2>&1
el fused
for ps in 0 200; do
  PA_CURV_FUSED=0 PA_TMA_PSLEEP=$ps timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2e_target_curv_unfused_ps${ps}.log 2>&1
  PA_TMA_PSLEEP=$ps timeout -s KILL 60 python bench.py --only-extra target_grad --steps 10 --warmup 3 > $O/r2e_target_grad_ps${ps}.log 2>&1
  PA_TMA_PSLEEP=$ps timeout -s KILL 60 python bench.py --only-extra config2 --steps 10 --warmup 3 > $O/r2e_config2_ps${ps}.log 2>&1
  PA_TMA_PSLEEP=$ps timeout -s KILL 60 python bench.py --only-extra grad5 --steps 10 --warmup 3 > $O/r2e_grad5_ps${ps}.log 2>&1
done
el unfused
PA_CF_CW=15 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_fused|k_div_shell" -s 4 -c 2 -o $O/r2e_curv_fused_cw15 -f \
      python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2e_ncu_full_cw15.log 2>&1
el ncu
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_*_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {a:(round(d[a],4) if not isinstance(d[a],dict) else d[a].get('value')) for a in ('value','ms_per_step','roofline_frac','launches_per_step','output_hash') if a in d})
PY
