#!/bin/bash
# round 2, call 14: 32-byte edge records with inline lists
mkdir -p gpurun_out
P=gpurun_out/r2c14
timeout 900 python -m pytest tests/test_gpu_edge_index.py -q -x > ${P}_t_edge.log 2>&1; echo "edge tests: $(tail -1 ${P}_t_edge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or power_law or dropin or precomp" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "test_full_size_sparse_otf or precomp" > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],3), 'idx_ms', d.get('edge_index_build_ms'), d.get('edge_index_bytes'), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
}
run pl ""
run er "--workload er-100k-1M-sparseotf"
run pc "--workload er-50k-1M-precomp"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:walk_uw_edge -c 1 --csv --log-file ${P}_traffic_pl.csv python bench.py --steps 1 --warmup 0 --no-extra --no-cpu --no-e2e > ${P}_traffic_pl.log 2>&1
grep walk_ ${P}_traffic_pl.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
