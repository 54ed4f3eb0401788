#!/bin/bash
# adaptive work-item depth: parity + A/B on one GPU; the reference's CUDA build with the managed arena; default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
PA_TMA_ZDIV=2 timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -n 8 --timeout 200 -p no:cacheprovider -k "golden or midsize or full_size" > $O/r2i_pytest_zdiv2.log 2>&1; echo "rc=$?" >> $O/r2i_pytest_zdiv2.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2i_pytest_zdiv2.log | head
for z in 0 1 2; do
  for ex in config2 target_grad target_curv grad5 curvature3; do
    PA_TMA_ZDIV=$z timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2i_${ex}_zdiv$z.log 2>&1
  done
done
for ex in config2 target_grad target_curv grad5 curvature3; do
  timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2i_${ex}_auto.log 2>&1
done
el zdiv
timeout -s KILL 700 python bench.py --steps 20 --warmup 5 > $O/r2i_bench.log 2> $O/r2i_bench.err; echo "rc=$?" >> $O/r2i_bench.err
el bench; tail -c 300 $O/r2i_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_*_zdiv*.log'))+sorted(glob.glob('gpurun_out/r2i_*_auto.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {a:(round(d[a],4) if not isinstance(d[a],dict) else d[a].get('ok')) for a in ('value','ms_per_step','roofline_frac','output_hash') if a in d})
for line in open('gpurun_out/r2i_bench.log'):
    if line.startswith('{'):
        d=json.loads(line); print('bench', d['value'], d['roofline']['frac'], 'cpu', (d['cpu_baseline'] or {}).get('value'), 'refgpu', d.get('ref_gpu_baseline'))
PY
