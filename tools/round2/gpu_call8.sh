#!/bin/bash
# round 2, call 8 (2 GPUs): copy-engine all-gather (CUDA IPC) vs NCCL, batches
mkdir -p gpurun_out
P=gpurun_out/r2c8
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -x > ${P}_t_multi.log 2>&1; echo "multi tests: $(tail -1 ${P}_t_multi.log)"; tail -30 ${P}_t_multi.log | grep -E "Error|error|FAILED" | head
for NB in 1 0 4 8; do
E2E="--no-e2e"; [ $NB = 0 ] && E2E=""
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra $E2E --batches $NB > ${P}_n2_b$NB.json 2> ${P}_n2_b$NB.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --batches 1 --gather nccl --no-e2e > ${P}_n2_nccl.json 2> ${P}_n2_nccl.err
python - <<'PY'
import json
for f in ['n2_b1','n2_b0','n2_b4','n2_b8','n2_nccl']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2c8_{f}.json') if l.startswith('{')][-1])
        print(f, round(d['value']/1e9,3),'Gsteps/s ms',round(d['ms_per_step'],3),'kernel_ms',round(d['kernel_ms_max_over_ranks'],3),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,3), d['checksum']['fnv_like_u64'], d['config']['parallelism'][40:])
    except Exception as e:
        print(f,'FAILED',e); print(open(f'gpurun_out/r2c8_{f}.err').read()[-1200:])
PY
