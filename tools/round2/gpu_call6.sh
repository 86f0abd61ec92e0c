#!/bin/bash
# round 2, call 6 (2 GPUs): multi-GPU tests over NCCL, scaling bench N=1,2
mkdir -p gpurun_out
P=gpurun_out/r2c6
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -x > ${P}_t_multi.log 2>&1; echo "multi tests: $(tail -1 ${P}_t_multi.log)"; tail -30 ${P}_t_multi.log | grep -E "Error|error|FAILED" | head
python bench.py --no-extra --no-cpu --steps 10 --warmup 3 > ${P}_n1.json 2> ${P}_n1.err
for NB in 1 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --batches $NB > ${P}_n2_b$NB.json 2> ${P}_n2_b$NB.err
done
python - <<'PY'
import json
for f in ['n1','n2_b1','n2_b4','n2_b8']:
    try:
        d=json.load(open(f'gpurun_out/r2c6_{f}.json'))
        print(f, round(d['value']/1e9,3),'Gsteps/s ms',round(d['ms_per_step'],3),'kernel_ms',round(d['kernel_ms_max_over_ranks'],3),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,3), d['checksum'], d['clocks'])
    except Exception as e:
        print(f,'FAILED',e); print(open(f'gpurun_out/r2c6_{f}.err').read()[-1200:])
PY
