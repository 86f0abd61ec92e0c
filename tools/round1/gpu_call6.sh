#!/bin/bash
mkdir -p gpurun_out
V=pecanpy_b200/lib/variants
run() { # name, lib, extra args
  local out=gpurun_out/r6_$1.json
  B2W_LIBRARY=$2 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $3 > $out 2>gpurun_out/r6_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
}
for v in default c3 c4; do
  lib=""; [ $v != default ] && lib=$PWD/$V/libb2w_$v.so
  B2W_LIBRARY=$lib python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin" > gpurun_out/t6_$v.log 2>&1; echo "tests $v: $(tail -1 gpurun_out/t6_$v.log)"
  run pl_$v "$lib" ""
  run er_$v "$lib" "--workload er-100k-1M-sparseotf"
done
run plw_c4 "$PWD/$V/libb2w_c4.so" "--workload powerlaw-1M-10M-sparseotf-weighted"
