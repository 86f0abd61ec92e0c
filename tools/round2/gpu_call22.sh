#!/bin/bash
# round 2, call 22 (2 GPUs): what is the GPU-to-GPU link of these boxes?
mkdir -p gpurun_out
P=gpurun_out/r2c22
nvidia-smi topo -m > ${P}_topo.txt 2>&1
nvidia-smi nvlink --status > ${P}_nvlink.txt 2>&1
python tools/round2/p2p_diag.py > ${P}_p2p.txt 2>&1
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/round2/nccl_diag.py > ${P}_nccl.txt 2>&1
cat ${P}_topo.txt | head -12; head -8 ${P}_nvlink.txt; cat ${P}_p2p.txt; grep -E "via|all_gather world|NVLS|Channel 00" ${P}_nccl.txt | head -12
