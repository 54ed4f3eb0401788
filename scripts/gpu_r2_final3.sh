#!/bin/bash
# validation after the L2 hints became the default of the divergence kernel: all GPU tests, smoke, the two curvature workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests -q -m gpu -n 8 --timeout 300 -p no:cacheprovider > $O/fin3_pytest.log 2>&1; echo "rc=$?" >> $O/fin3_pytest.log
tail -3 $O/fin3_pytest.log
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/fin3_smoke.log 2>&1; echo "rc=$?" >> $O/fin3_smoke.log
tail -2 $O/fin3_smoke.log
for ex in target_curv curvature3; do
  timeout -s KILL 60 python bench.py --only-extra $ex --steps 20 --warmup 5 > $O/fin3_$ex.log 2>&1
  python - "$O/fin3_$ex.log" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1].split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
done
