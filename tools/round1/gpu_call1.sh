#!/bin/bash
# GPU call: new threshold tests + source-level ncu captures of the walk kernels (1 walk per node)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "threshold or dropin_extend" > gpurun_out/t_thr.log 2>&1
tail -3 gpurun_out/t_thr.log
NCU="ncu --set full --import-source on --clock-control none -c 1 -f"
timeout 400 $NCU -k regex:walk_uw -o gpurun_out/uw_pl_nw1 python bench.py --num-walks 1 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_pl.log 2>&1
timeout 200 $NCU -k regex:walk_uw -o gpurun_out/uw_er_nw10 python bench.py --workload er-100k-1M-sparseotf --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_er.log 2>&1
timeout 200 $NCU -k regex:walk_thread -o gpurun_out/precomp_er50k python bench.py --workload er-50k-1M-precomp --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_pc.log 2>&1
ls -la gpurun_out
