#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/o_bench_*.log
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $O/o_pytest.log 2>&1; echo "rc=$?" >> $O/o_pytest.log
grep -E "passed|failed|^FAILED" $O/o_pytest.log | head -20
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/o_bench_n1.log 2>&1; echo "rc=$?" >> $O/o_bench_n1.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $O/o_launches_curv.csv \
    python bench.py --only-extra curvature3 --steps 3 --warmup 3 > $O/o_ncu_curv.log 2>&1
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob('gpurun_out/o_bench_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value %.1f ms %.3f frac %.3f e2e %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
            if d.get('extras'):
                for k,v in d['extras'].items(): print('   ',k, {a:v[a] for a in ('value','ms_per_step','roofline_frac','launches_per_step') if a in v} or v)
for f in ['gpurun_out/o_launches_curv.csv']:
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    seq=[(r[ki][:50], float(r[vi].replace(',',''))/1000) for r in rows[1:] if 'valid_copy' not in r[ki]]
    print(f); 
    for n,v in seq[-4:]: print('   %-52s %9.1f us'%(n,v))
PY
