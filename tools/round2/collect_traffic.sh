#!/bin/bash
# DRAM traffic of the walk kernel of every bench workload, FULL bench launch (final tree): ncu with three metrics only
# (one replay pass), first launch of the kernel.  tools/round2/make_traffic_json.py turns the CSVs into profiles/traffic.json.
mkdir -p gpurun_out
P=gpurun_out/traffic
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for W in powerlaw-1M-10M-sparseotf er-100k-1M-sparseotf er-50k-1M-precomp powerlaw-1M-10M-sparseotf-weighted powerlaw-1M-10M-sparseotf-n2vplus; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:'walk_(uw_edge|precomp_edge|wedge|uw_kernel|sparse_warp|thread)' -c 1 --csv --log-file ${P}_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 0 --no-extra --no-cpu --no-e2e > ${P}_$W.log 2>&1
  echo "$W: $(grep -c walk_ ${P}_$W.csv) rows"
done
timeout 900 ncu --metrics $M --clock-control none -k regex:walk_dense -c 1 --csv --log-file ${P}_dense-20k-denseotf-n2vplus.csv \
  python bench.py --workload dense-20k-denseotf-n2vplus --steps 1 --warmup 0 --no-extra --no-cpu --no-e2e --num-walks 1 > ${P}_dense.log 2>&1
echo "dense: $(grep -c walk_ ${P}_dense-20k-denseotf-n2vplus.csv) rows"
