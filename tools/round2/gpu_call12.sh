#!/bin/bash
# round 2, call 12: warp-converged loops + cooperative off-edge steps
mkdir -p gpurun_out
P=gpurun_out/r2c12
timeout 900 python -m pytest tests/test_gpu_edge_index.py tests/test_gpu_wedge.py -q -x > ${P}_t_edge.log 2>&1; echo "edge+wedge tests: $(tail -1 ${P}_t_edge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or power_law or dropin" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "sparse" > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
}
run pl ""
run pl_mb6 "--flags $((6*65536))"
run er "--workload er-100k-1M-sparseotf"
run plw "--workload powerlaw-1M-10M-sparseotf-weighted"
run plx "--workload powerlaw-1M-10M-sparseotf-n2vplus"
python tools/round2/launch_size.py > ${P}_launch_size.txt 2>&1; cat ${P}_launch_size.txt | tail -8
