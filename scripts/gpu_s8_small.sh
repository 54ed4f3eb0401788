#!/bin/bash
# session 8, call 2: small-tile CTA shapes (2 / 4 consumer warps): parity, then A/B on the 16^3-box hierarchy
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 150 python -m pytest tests -q -m gpu -n 8 --timeout 120 --timeout-method=thread -p no:cacheprovider > $O/z_pytest.log 2>&1; echo "rc=$?" >> $O/z_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/z_pytest.log | head -20
for sm in 1 0; do
  PA_TMA_SMALL=$sm timeout -s KILL 90 python bench.py --only-extra grad5 --steps 20 --warmup 3 > $O/z_grad5_small$sm.log 2>&1
done
for kb in 10 40; do
  PA_TMA_INFLIGHT_KB=$kb timeout -s KILL 90 python bench.py --only-extra grad5 --steps 20 --warmup 3 > $O/z_grad5_small1_kb$kb.log 2>&1
done
el grad5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/z_grad5*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            ok=True; d=json.loads(line)
            print(f, {a:round(d[a],4) for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in d})
    if not ok: print(f, "NO JSON", open(f).read()[-400:])
PY
