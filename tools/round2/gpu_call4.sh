#!/bin/bash
# round 2, call 4: edge kernel v3 (empty-range fix, integer replay) -- full GPU suite + bench + profile
mkdir -p gpurun_out
P=gpurun_out/r2c4
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_t_all.log 2>&1; echo "gpu suite: $(tail -1 ${P}_t_all.log)"
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
run pl_mb4 "A=1" "--flags $((4*65536))"
run pl_mb6 "A=1" "--flags $((6*65536))"
run er "A=1" "--workload er-100k-1M-sparseotf"
timeout 600 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw_edge -o ${P}_edge_pl_nw1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --num-walks 1 > ${P}_ncu.log 2>&1
