#!/bin/bash
# A/B of the uw kernel structures (tools/build_variants.py) + occupancy variants
mkdir -p gpurun_out
V=pecanpy_b200/lib/variants
run() { # name, lib, extra args
  local out=gpurun_out/ab_$1.json
  B2W_LIBRARY=$2 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $3 > $out 2>gpurun_out/ab_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
}
for v in default flat nested_inl; do
  lib=""; [ $v != default ] && lib=$PWD/$V/libb2w_$v.so
  B2W_LIBRARY=$lib python -m pytest tests/test_gpu_parity.py -q -x -k "sparse_otf or dropin" > gpurun_out/t_$v.log 2>&1; echo "tests $v: $(tail -1 gpurun_out/t_$v.log)"
  run pl_$v "$lib" ""
  run er_$v "$lib" "--workload er-100k-1M-sparseotf"
done
run pl_default_mb6 "" "--flags $((6<<16))"
run pl_default_mb4 "" "--flags $((4<<16))"
run pl_flat_mb6 "$PWD/$V/libb2w_flat.so" "--flags $((6<<16))"
run er_default_g16 "" "--workload er-100k-1M-sparseotf --flags $((16<<8))"
for mb in 0 5 6 8; do run pc_mb$mb "" "--workload er-50k-1M-precomp --flags $((mb<<16))"; done
run plw_default "" "--workload powerlaw-1M-10M-sparseotf-weighted"
