#!/bin/bash
# round 2, call 3: edge kernel v2 (branch-light solve, integer replay arithmetic, sector-aligned row stores, L2 fetch 32 B)
mkdir -p gpurun_out
P=gpurun_out/r2c3
timeout 900 python -m pytest tests/test_gpu_edge_index.py -q -x > ${P}_t_edge.log 2>&1; echo "edge tests: $(tail -1 ${P}_t_edge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin or power_law or sharding" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
run() { # name, env, extra args
  local out=${P}_$1.json
  env $2 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e $3 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), 'idx_ms', d.get('edge_index_build_ms'), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-600:])
PY
}
run pl "A=1" ""
run pl_f64 "B2W_L2_FETCH=64" ""
run pl_f128 "B2W_L2_FETCH=128" ""
run pl_mb4 "A=1" "--flags $((4*65536))"
run pl_mb6 "A=1" "--flags $((6*65536))"
run er "A=1" "--workload er-100k-1M-sparseotf"
run er_f64 "B2W_L2_FETCH=64" "--workload er-100k-1M-sparseotf"
run pc "A=1" "--workload er-50k-1M-precomp"
run pc_f64 "B2W_L2_FETCH=64" "--workload er-50k-1M-precomp"
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
timeout 600 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw_edge -o ${P}_edge_pl_nw1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --num-walks 1 > ${P}_ncu.log 2>&1
