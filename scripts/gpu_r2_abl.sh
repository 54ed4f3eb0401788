#!/bin/bash
# where does the time of the flame-normal pass go?  K-less plane-staged kernel (PA_NORMAL_F3=1) with its FP64 chain and / or its
# global stores removed (wrong results, timing only) -- kernel times from an ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for a in 0 1 2 3; do
  PA_NORMAL_F3=1 PA_NF3_ABLATE=$a timeout -s KILL 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_curv_f3" -c 1 --csv --log-file $O/r2abl_$a.csv python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2abl_$a.log 2>&1
done
python - <<'PY'
import csv
for a in range(4):
    rows=[r for r in csv.reader(open('gpurun_out/r2abl_%d.csv'%a)) if len(r)>10]
    h=rows[0]; im=h.index('Metric Name'); iv=h.index('Metric Value')
    print('ablate',a,{r[im]:r[iv] for r in rows[1:]})
PY
