#!/bin/bash
# Round 2, first GPU call (1 GPU, ~6 min): run the variants that round 1 built and verified only under the emulator, and
# collect the A/Bs that decide their defaults.  Every step writes its own file under gpurun_out/.
#  1 pytest -m gpu incl. PA_TEST_EXPERIMENTAL=1 (descriptor-prefetch variant on hardware for the first time)
#  2 A/B per extra: PA_TMA_PREFETCH 0/1  x  PA_TMA_ZC 32/64/128            (ms per step, roofline fraction)
#  3 default bench with PA_TMA_PREFETCH=1 (does the headline move?)
#  4 ncu --set full of the 16^3-box grad kernel and of NORMAL_S with / without the prefetch warp (stall reasons)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
PA_TEST_EXPERIMENTAL=1 timeout -s KILL 200 python -m pytest tests -q -m gpu -n 8 --timeout 120 --timeout-method=thread -p no:cacheprovider > $O/r2a_pytest.log 2>&1; echo "rc=$?" >> $O/r2a_pytest.log
el pytest; grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r2a_pytest.log | head -20
for ex in target_curv grad5 target_grad curvature3; do
  for pf in 0 1; do
    for zc in 32 64 128; do
      PA_TMA_PREFETCH=$pf PA_TMA_ZC=$zc timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2a_${ex}_pf${pf}_zc${zc}.log 2>&1
    done
  done
  el $ex
done
# rows per tile pick the CTA shape class (128-wide boxes: TY=2 -> 128 pairs -> 2 consumer warps x 5-6 CTAs/SM, TY=4 -> 4 warps x 3-4)
for ty in 2 4; do
  for pf in 0 1; do
    PA_TMA_PREFETCH=$pf PA_TMA_TY=$ty timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2a_target_curv_pf${pf}_zc32_ty$ty.log 2>&1
  done
done
el shapes
# DIV waits on its full barrier 39 % of the time with a 3-stage ring (3 components per stage): the 16-warp shape (mask bit 8)
# has room for 4-5 stages
for mask in 20 28; do
  for kb in 20 40; do
    PA_TMA_CW16=$mask PA_TMA_INFLIGHT_KB=$kb timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2a_target_curv_pf0_zc32_mask${mask}_kb$kb.log 2>&1
  done
done
el ring
# PF kernels stage DIV "lean" (no y-halo rows for n_x / n_z, no z-halo planes for n_x / n_y); PA_DIV_LEAN=0 = prefetch warp alone
PA_TMA_PREFETCH=1 PA_DIV_LEAN=0 timeout -s KILL 60 python bench.py --only-extra target_curv --steps 10 --warmup 3 > $O/r2a_target_curv_pf1_zc32_nolean.log 2>&1
el lean
# staged coarse-fine fill (coarse register cells through shared memory, each loaded once)
for ex in grad5 target_curv curvature3 target_grad; do
  PA_BCFILL_V2=1 timeout -s KILL 60 python bench.py --only-extra $ex --steps 10 --warmup 3 > $O/r2a_${ex}_pf0_zc32_bcfillv2.log 2>&1
done
PA_BCFILL_V2=1 timeout -s KILL 200 python -m pytest tests -q -m gpu -n 8 --timeout 120 -p no:cacheprovider -k "golden or ghost_cells or midsize" > $O/r2a_pytest_bcfillv2.log 2>&1; tail -n 2 $O/r2a_pytest_bcfillv2.log
el bcfill
PA_TMA_PREFETCH=1 timeout -s KILL 150 python bench.py --no-extras > $O/r2a_bench_pf1.log 2>&1
timeout -s KILL 150 python bench.py --no-extras > $O/r2a_bench_pf0.log 2>&1
el bench
for pf in 0 1; do
  PA_TMA_PREFETCH=$pf timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:k_stencil_tma -s 6 -c 2 -o $O/r2a_grad5_pf$pf -f \
      python bench.py --only-extra grad5 --steps 2 --warmup 3 > $O/r2a_ncu_grad5_pf$pf.log 2>&1
  PA_TMA_PREFETCH=$pf timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:k_stencil_tma -s 6 -c 2 -o $O/r2a_curv_pf$pf -f \
      python bench.py --only-extra target_curv --steps 2 --warmup 3 > $O/r2a_ncu_curv_pf$pf.log 2>&1
done
el ncu
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_*_pf*_zc*.log'))+sorted(glob.glob('gpurun_out/r2a_bench_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            ok=True; d=json.loads(line)
            if 'roofline' in d: print(f, 'value %.1f ms %.3f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))
            else: print(f, {a:round(d[a],4) for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in d})
    if not ok: print(f, "NO JSON", open(f).read()[-300:])
PY
