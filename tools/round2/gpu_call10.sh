#!/bin/bash
# round 2, call 10: exact-cdf checkpoints for the unweighted edge kernel
mkdir -p gpurun_out
P=gpurun_out/r2c10
timeout 900 python -m pytest tests/test_gpu_edge_index.py -q -x > ${P}_t_edge.log 2>&1; echo "edge tests: $(tail -1 ${P}_t_edge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or power_law or dropin" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "test_full_size_sparse_otf" > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), 'ckpt_ms', d.get('edge_ckpt_build_ms'), d.get('edge_ckpt_bytes'), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
}
run pl ""
run pl_nock "--flags 128"
run pl_mb6 "--flags $((6*65536))"
run er "--workload er-100k-1M-sparseotf"
python tools/round2/launch_size.py > ${P}_launch_size.txt 2>&1; cat ${P}_launch_size.txt | tail -8
timeout 600 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw_edge -o ${P}_edge_pl_nw1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra --num-walks 1 > ${P}_ncu.log 2>&1
