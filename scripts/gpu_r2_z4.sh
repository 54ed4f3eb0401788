#!/bin/bash
# barrier-free flame-normal kernel with the warp-private cp.async ring: fingerprint, timing, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 120 python scripts/gpu_hash_check.py > $O/r2z4_hash.log 2>&1; echo "rc=$?" >> $O/r2z4_hash.log
el hash; tail -2 $O/r2z4_hash.log | cut -c1-200
for pf in 4 2; do
  PA_NORMAL_W=1 PA_NW_PF=$pf timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2z4_target_curv_fusednw_ring_d$pf.log 2>&1
done
PA_NORMAL_W=1 PA_NW_CTAS=4 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2z4_target_curv_fusednw_ring_ctas4.log 2>&1
PA_NORMAL_W=1 PA_NW_ZC=64 timeout -s KILL 90 python bench.py --only-extra target_curv --steps 20 --warmup 5 > $O/r2z4_target_curv_fusednw_ring_zc64.log 2>&1
PA_NORMAL_W=1 timeout -s KILL 90 python bench.py --only-extra curvature3 --steps 20 --warmup 5 > $O/r2z4_curvature3_fusednw_ring.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z4_*fused*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline_frac'],4), d['output_hash'].get('ok'), d['launches_per_step'])
PY
el timing
PA_NORMAL_W=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:"k_normal_w" -c 1 -o $O/r2z4_normal_w python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2z4_ncu.log 2>&1
el ncu; tail -1 $O/r2z4_ncu.log
