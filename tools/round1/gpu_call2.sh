#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/t_parity.log 2>&1; tail -3 gpurun_out/t_parity.log
python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/t_full.log 2>&1; tail -3 gpurun_out/t_full.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_pl.json 2> gpurun_out/bench_pl.err; cat gpurun_out/bench_pl.json | cut -c1-400
python bench.py --workload er-100k-1M-sparseotf --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_er.json 2> gpurun_out/bench_er.err; cat gpurun_out/bench_er.json | cut -c1-400
python bench.py --workload powerlaw-1M-10M-sparseotf-weighted --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_plw.json 2> gpurun_out/bench_plw.err; cat gpurun_out/bench_plw.json | cut -c1-400
NCU="ncu --set full --import-source on --clock-control none -c 1 -f"
timeout 300 $NCU -k regex:walk_uw -o gpurun_out/uw_pl_nw1_v3 python bench.py --num-walks 1 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_pl.log 2>&1
timeout 200 $NCU -k regex:walk_uw -o gpurun_out/uw_er_v3 python bench.py --workload er-100k-1M-sparseotf --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_er.log 2>&1
timeout 300 $NCU -k regex:walk_sparse_warp -o gpurun_out/generic_plw_nw1 python bench.py --workload powerlaw-1M-10M-sparseotf-weighted --num-walks 1 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_plw.log 2>&1
ls -la gpurun_out | head -30
