"""NCCL all-gather of the walk matrix alone (no walk kernel): time per call, 20 calls."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
rows = 10_000_000 // world * world
full = torch.zeros((rows, 82), dtype=torch.int32, device="cuda")
mine = full[rank * (rows // world):(rank + 1) * (rows // world)]
for _ in range(3):
    dist.all_gather_into_tensor(full.view(-1), mine.view(-1))
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_gather_into_tensor(full.view(-1), mine.view(-1))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
if rank == 0:
    recv = full.numel() * 4 * (world - 1) / world
    print(f"nccl all_gather world {world}: {ms:.3f} ms per call; {recv / ms / 1e6:.1f} GB/s received per GPU")
dist.destroy_process_group()
