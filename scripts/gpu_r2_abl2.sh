#!/bin/bash
# K-less plane-staged kernel: streaming stores (ablate 4: right results; 5: + no FP64 chain); the default kernels with the
# L2 hints forced on; write-only / copy bandwidth of the device -- kernel times from ncu launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 60 python scripts/gpu_write_bw.py > $O/r2abl_write_bw.log 2>&1; cat $O/r2abl_write_bw.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__cycles_active.avg.pct_of_peak_sustained_elapsed
for a in 4 5; do
  PA_NORMAL_F3=1 PA_NF3_ABLATE=$a timeout -s KILL 120 ncu --metrics $M --clock-control none -k regex:"k_curv_f3" -c 1 --csv --log-file $O/r2abl_$a.csv python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2abl_$a.log 2>&1
done
PA_TMA_L2HINT=3 timeout -s KILL 120 ncu --metrics $M --clock-control none -k regex:"k_stencil_tma" -c 2 --csv --log-file $O/r2abl_tma_hint3.csv python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2abl_tma_hint3.log 2>&1
python - <<'PY'
import csv
for a in ('4','5','tma_hint3'):
    rows=[r for r in csv.reader(open('gpurun_out/r2abl_%s.csv'%a)) if len(r)>10]
    h=rows[0]; im=h.index('Metric Name'); iv=h.index('Metric Value'); ii=h.index('ID'); ik=h.index('Kernel Name')
    d={}
    for r in rows[1:]: d.setdefault((r[ii],r[ik][:40]),{})[r[im]]=r[iv]
    print('case',a,d)
PY
