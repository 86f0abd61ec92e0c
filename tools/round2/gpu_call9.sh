#!/bin/bash
# round 2, call 9: weighted edge index (tests, bench), launch-size experiment
mkdir -p gpurun_out
P=gpurun_out/r2c9
timeout 900 python -m pytest tests/test_gpu_wedge.py -q -x > ${P}_t_wedge.log 2>&1; echo "wedge tests: $(tail -1 ${P}_t_wedge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_index.py -q -x > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "weighted or sparse" > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
python tools/round2/launch_size.py > ${P}_launch_size.txt 2>&1; cat ${P}_launch_size.txt | tail -8
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 3 --warmup 3 --no-extra $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']/1e6,2), 'widx_ms', d.get('weighted_index_build_ms'), d.get('weighted_index_bytes'), d.get('weighted_index_counts'), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
}
run plw "--workload powerlaw-1M-10M-sparseotf-weighted"
run plx "--workload powerlaw-1M-10M-sparseotf-n2vplus"
