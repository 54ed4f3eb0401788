#!/bin/bash
# session 6, call 2: branch-free flame-normal epilogue + 16-consumer-warp CTA shape: parity, A/B of the shapes per mode, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $O/w_pytest.log 2>&1; echo "rc=$?" >> $O/w_pytest.log
grep -E "passed|failed|^FAILED|^E  |rc=" $O/w_pytest.log | head -20
for mask in 0 20 28; do
  for ex in curvature3 target_curv; do
    PA_TMA_CW16=$mask timeout -s KILL 200 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/w_${ex}_m$mask.log 2>&1
  done
done
for mask in 0 1; do
  PA_TMA_CW16=$mask timeout -s KILL 200 python bench.py --only-extra target_grad --steps 10 --warmup 3 > $O/w_target_grad_m$mask.log 2>&1
  PA_TMA_CW16=$mask timeout -s KILL 200 python bench.py --only-extra grad5 --steps 10 --warmup 3 > $O/w_grad5_m$mask.log 2>&1
  PA_TMA_CW16=$mask timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > $O/w_bench_m$mask.log 2>&1
done
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_stencil_tma" -s 4 -c 2 -o $O/w_curv -f \
    python bench.py --only-extra curvature3 --steps 2 --warmup 3 > $O/w_ncu_curv.log 2>&1; echo "rc=$?" >> $O/w_ncu_curv.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/w_*_m*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            ok=True
            d=json.loads(line)
            if 'roofline' in d:
                print(f, 'value %.1f ms %.3f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))
            else:
                print(f, {a:round(d[a],4) for a in ('value','ms_per_step','roofline_frac') if a in d})
    if not ok: print(f, "NO JSON", open(f).read()[-600:])
PY
tail -n 2 $O/w_ncu_curv.log
