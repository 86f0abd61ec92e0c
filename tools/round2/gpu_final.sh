#!/bin/bash
# round 2, final validation on one B200 (what the GPU budget had left: ~10 min): full GPU suite, default bench, smoke(),
# launch list of the default bench, one full ncu capture of the headline kernel (full bench launch)
mkdir -p gpurun_out
P=gpurun_out/r2final
timeout 330 python -m pytest tests -m gpu -q > ${P}_t_all.log 2>&1; echo "gpu suite: $(tail -1 ${P}_t_all.log)"
( time python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err ) 2> ${P}_bench_time.txt; tail -3 ${P}_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2final_bench_default.json') if l.startswith('{')][-1])
    def show(n,x):
        r=x.get('roofline') or {}
        print(n, round(x['value']/1e9,3),'Gsteps/s', r.get('kernel'), 'frac',round(r.get('frac',0),3),'dram_frac',r.get('dram_frac') and round(r['dram_frac'],3),'ms',round(x['ms_per_step'],3),'e2e',x['e2e'] and round(x['e2e']['value']/1e9,3),'cpu',x['cpu_baseline'] and round(x['cpu_baseline']['value']/1e6,2),'clk',x['clocks']['samples'], x['clocks']['reasons'])
    show('headline',d)
    for k,v in d.get('extra',{}).items():
        if 'error' in v: print(k,'ERROR',v['error'])
        else: show(k,v)
except Exception as e:
    print('bench FAILED', e); print(open('gpurun_out/r2final_bench_default.err').read()[-1500:])
PY
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > ${P}_ncu_launches.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw_edge -o ${P}_uw_edge_pl_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra > ${P}_ncu_full.log 2>&1
ls -la gpurun_out | grep r2final
