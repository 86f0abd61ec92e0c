#!/bin/bash
# round 2, call 5: full GPU suite (PreComp over edge records, packed tables, full-size #4/#5, multi-replica) + default bench with extras
mkdir -p gpurun_out
P=gpurun_out/r2c5
timeout 2400 python -m pytest tests -m gpu -q -x > ${P}_t_all.log 2>&1; echo "gpu suite: $(tail -1 ${P}_t_all.log)"
( time python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err ) 2> ${P}_bench_time.txt; tail -3 ${P}_bench_time.txt
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c5_bench_default.json'))
    def show(n,x):
        r=x.get('roofline') or {}
        print(n, round(x['value']/1e9,3),'Gsteps/s', r.get('kernel'), 'frac',round(r.get('frac',0),3),'ms',round(x['ms_per_step'],3),'passes',x.get('timed_passes'),'e2e',x['e2e'] and round(x['e2e']['value']/1e9,3),'cpu',x['cpu_baseline'] and round(x['cpu_baseline']['value']/1e6,2),'clk',x['clocks'], 'chk', x.get('checksum'))
    show('headline',d)
    for k,v in d.get('extra',{}).items():
        if 'error' in v: print(k,'ERROR',v['error'])
        else: show(k,v)
except Exception as e:
    print('bench FAILED', e); print(open('gpurun_out/r2c5_bench_default.err').read()[-1500:])
PY
