#!/bin/bash
# Round 2, second GPU call (1 GPU): the fused curvature kernel on hardware for the first time.
#  1 parity: the curvature / golden GPU tests (fused and unfused routes)
#  2 A/B: fused (CW 15 / 19, PA_CF_ZC 32/64/126, stages 3/4/5) vs unfused on target_curv and curvature3
#  3 ncu: launch list of one target_curv step, --set full of the fused kernel and the shell kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 200 -p no:cacheprovider -k "curvature or midsize or full_size or extreme or wide or degenerate or flat or selftest" > $O/r2b_pytest.log 2>&1; echo "rc=$?" >> $O/r2b_pytest.log
timeout -s KILL 200 python scripts/gpu_hash_check.py > $O/r2b_hash.log 2>&1; tail -n 3 $O/r2b_hash.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2b_pytest.log | head -20
for ex in target_curv curvature3; do
  PA_CURV_FUSED=0 timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2b_${ex}_unfused.log 2>&1
  for cw in 15 19; do
    for zc in 32 64 126; do
      PA_CF_CW=$cw PA_CF_ZC=$zc timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2b_${ex}_cw${cw}_zc${zc}.log 2>&1
    done
    for st in 3 5; do
      PA_CF_CW=$cw PA_CF_STAGES=$st timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2b_${ex}_cw${cw}_zc64_st${st}.log 2>&1
    done
  done
  el $ex
done
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2b_launches_target_curv.csv python bench.py --only-extra target_curv --steps 2 --warmup 1 > $O/r2b_ncu_launches.log 2>&1
el launches
for cw in 15 19; do
  PA_CF_CW=$cw timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_curv_fused|k_div_shell|k_bcfill" -s 8 -c 4 -o $O/r2b_curv_fused_cw$cw -f \
      python bench.py --only-extra target_curv --steps 2 --warmup 2 > $O/r2b_ncu_full_cw$cw.log 2>&1
  el ncu$cw
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            ok=True; d=json.loads(line)
            print(f, {a:round(d[a],4) for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in d})
    if not ok and 'ncu' not in f and 'pytest' not in f: print(f, "NO JSON", open(f).read()[-400:])
PY
