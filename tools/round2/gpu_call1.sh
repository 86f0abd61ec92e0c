#!/bin/bash
# round 2, call 1: A/B of the two variants written at the end of round 1 (in-register membership, sub-warp REDUX)
mkdir -p gpurun_out
V=pecanpy_b200/lib/variants
run() { # name, lib, extra args
  local out=gpurun_out/r2c1_$1.json
  B2W_LIBRARY=$2 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $3 > $out 2>gpurun_out/r2c1_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
}
for v in default inreg redux; do
  lib=""; [ $v != default ] && lib=$PWD/$V/libb2w_$v.so
  B2W_LIBRARY=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin" > gpurun_out/r2c1_t_$v.log 2>&1; echo "tests $v: $(tail -1 gpurun_out/r2c1_t_$v.log)"
done
run pl_default "" ""
run pl_inreg "$PWD/$V/libb2w_inreg.so" ""
run er_default "" "--workload er-100k-1M-sparseotf"
run er_redux "$PWD/$V/libb2w_redux.so" "--workload er-100k-1M-sparseotf"
run pc_default "" "--workload er-50k-1M-precomp"
run plw_default "" "--workload powerlaw-1M-10M-sparseotf-weighted"
run dense_default "" "--workload dense-20k-denseotf-n2vplus --steps 2 --warmup 3"
