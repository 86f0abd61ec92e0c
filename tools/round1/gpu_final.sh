#!/bin/bash
# final round-1 validation: full GPU suite, bench lines, ncu launch list + one full capture of the default launch
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "gpu tests: $(tail -1 gpurun_out/final_tests.log)"
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; cut -c1-300 gpurun_out/final_bench_default.json
python bench.py --workload er-100k-1M-sparseotf > gpurun_out/final_bench_er.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_er.json
python bench.py --workload er-50k-1M-precomp > gpurun_out/final_bench_precomp.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_precomp.json
python bench.py --workload powerlaw-1M-10M-sparseotf-weighted --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/final_bench_plw.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_plw.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/final_ncu_launches.log 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -c 1 -f -k regex:walk_uw -o gpurun_out/final_uw_pl_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out | grep final
