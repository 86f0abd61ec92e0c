#!/bin/bash
# round 2, call 16: interpolation search in the weighted kernel; L2 persistence window for the PreComp records
mkdir -p gpurun_out
P=gpurun_out/r2c16
timeout 900 python -m pytest tests/test_gpu_wedge.py -q -x > ${P}_t_wedge.log 2>&1; echo "wedge tests: $(tail -1 ${P}_t_wedge.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or power_law or precomp" > ${P}_t_parity.log 2>&1; echo "parity: $(tail -1 ${P}_t_parity.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "weighted" > ${P}_t_full.log 2>&1; echo "fullsize: $(tail -1 ${P}_t_full.log)"
run() { # name, extra args
  local out=${P}_$1.json
  python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e $2 > $out 2>${P}_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s', d['roofline']['kernel'], 'ms', round(d['ms_per_step'],3), d['walk_stats_rank0'], flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
}
run plw "--workload powerlaw-1M-10M-sparseotf-weighted"
run plx "--workload powerlaw-1M-10M-sparseotf-n2vplus"
run pc "--workload er-50k-1M-precomp"
run plw_mb5 "--workload powerlaw-1M-10M-sparseotf-weighted --flags $((5*65536))"
run pc_mb5 "--workload er-50k-1M-precomp --flags $((5*65536))"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:walk_wedge -c 1 --csv --log-file ${P}_traffic_plw.csv python bench.py --workload powerlaw-1M-10M-sparseotf-weighted --steps 1 --warmup 0 --no-extra --no-cpu --no-e2e > ${P}_traffic_plw.log 2>&1
grep walk_ ${P}_traffic_plw.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
