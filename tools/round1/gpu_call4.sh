#!/bin/bash
mkdir -p gpurun_out
V=pecanpy_b200/lib/variants
run() { # name, lib, extra args
  local out=gpurun_out/ab_$1.json
  B2W_LIBRARY=$2 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $3 > $out 2>gpurun_out/ab_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
}
for v in default dual; do
  lib=""; [ $v != default ] && lib=$PWD/$V/libb2w_$v.so
  B2W_LIBRARY=$lib python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin" > gpurun_out/t_$v.log 2>&1; echo "tests $v: $(tail -1 gpurun_out/t_$v.log)"
  run pl_$v "$lib" ""
  run pl_${v}_l2 "$lib" "--flags 64"
  run er_$v "$lib" "--workload er-100k-1M-sparseotf"
done
run er_default_l2 "" "--workload er-100k-1M-sparseotf --flags 64"
run pc_default "" "--workload er-50k-1M-precomp"
run plw_default "" "--workload powerlaw-1M-10M-sparseotf-weighted"
B2W_LIBRARY=$PWD/$V/libb2w_dual.so python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/t_full_dual.log 2>&1; echo "fullsize dual: $(tail -1 gpurun_out/t_full_dual.log)"
