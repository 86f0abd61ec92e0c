#!/bin/bash
# round 2, call 23 (2 GPUs): mirrored stores, coalesced by the warp (WarpRowTile)
mkdir -p gpurun_out
P=gpurun_out/r2c23
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "mirror" > ${P}_t_multi.log 2>&1; echo "mirror tests: $(tail -1 ${P}_t_multi.log)"
for cfg in "1 mirror"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --no-e2e --batches $1 --gather $2 > ${P}_n2_b$1_$2.json 2> ${P}_n2_b$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c23_n2_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r2c23_')[1], round(d['value']/1e9,3),'Gsteps/s ms',round(d['ms_per_step'],3),'kernel_ms',round(d['kernel_ms_max_over_ranks'],3), d['checksum']['fnv_like_u64'], d['config']['parallelism'][-70:])
    except Exception as e:
        print(f,'FAILED',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
tail -5 ${P}_t_multi.log
