#!/bin/bash
# round 2, call 2: first run of the edge-index kernel: parity, bench, ncu
mkdir -p gpurun_out
P=gpurun_out/r2c2
timeout 900 python -m pytest tests/test_gpu_edge_index.py -q -x > ${P}_t_edge.log 2>&1; echo "edge tests: $(tail -1 ${P}_t_edge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin or power_law or sharding" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 5 --warmup 3 --no-cpu $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,3), 'idx_ms', d.get('edge_index_build_ms'), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-600:])
PY
}
run pl ""
run er "--workload er-100k-1M-sparseotf"
run pl_mb4 "--flags $((4*65536)) --no-e2e"
run pl_mb6 "--flags $((6*65536)) --no-e2e"
run pl_mb8 "--flags $((8*65536)) --no-e2e"
run pl_old "--flags 64 --no-e2e"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
timeout 600 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw_edge -o ${P}_edge_pl_nw1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --num-walks 1 > ${P}_ncu.log 2>&1
ls -la gpurun_out | grep r2c2
