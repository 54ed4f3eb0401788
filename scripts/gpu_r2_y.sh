#!/bin/bash
# how fast is the barrier-free plain-load stencil kernel on the flame-normal pass? (launch list of the PA_STENCIL=simple route)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
PA_STENCIL=simple timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_stencil_simple|k_progress|k_bcfill" -c 6 --csv --log-file $O/r2y_launches_simple.csv python bench.py --only-extra target_curv --steps 1 --warmup 0 > $O/r2y_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2y_launches_simple.csv')) if len(r)>10]
h=rows[0]
ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ik][:60]),{})[r[im]]=r[iv]
for k,v in d.items(): print(k, v)
PY
